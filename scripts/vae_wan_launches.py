"""One native Wan VAE decode of a 5-latent-frame 480 x 832 clip (17 frames) for `ncu --metrics gpu__time_duration.sum`."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from alg_b200.vae_wan import AutoencoderKLWan  # noqa: E402

vae = AutoencoderKLWan.from_synthetic(seed=0, device="cuda")
z = torch.randn(1, 16, 5, 60, 104, device="cuda")
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
v = vae.decode(z).sample
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print(v.shape, bool(torch.isfinite(v).all()))
