from dataclasses import dataclass

import torch


@dataclass
class HunyuanVideoPipelineOutput:
    frames: torch.Tensor
