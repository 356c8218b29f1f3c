from typing import List, Union

import numpy as np
import PIL.Image
import torch

PipelineImageInput = Union[PIL.Image.Image, np.ndarray, torch.Tensor, List[PIL.Image.Image], List[np.ndarray], List[torch.Tensor]]
