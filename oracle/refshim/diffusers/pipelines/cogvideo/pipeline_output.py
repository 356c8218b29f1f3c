from dataclasses import dataclass

import torch


@dataclass
class CogVideoXPipelineOutput:
    frames: torch.Tensor
