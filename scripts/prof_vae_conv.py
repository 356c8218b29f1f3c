"""One implicit 3x3x3 convolution of the Wan VAE's top level (96 -> 96 channels, 3 frames of 480 x 832, float32 via the bf16 3-term
split) for `ncu --set full -k regex:gemm_kernel`: tap-mode GEMM, M = 5 x 482 x 834 padded-raster rows, N = 96, K = 27 x 320."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from alg_b200.vae_wan import AutoencoderKLWan, _Act  # noqa: E402

vae = AutoencoderKLWan.from_synthetic(seed=0, device="cuda")
T, H, W, C = 3, 480, 832, 96
x = _Act(torch.randn(T * H * W, C, device="cuda"), T, H, W, C)
name = "decoder.up_blocks.3.resnets.1"
for _ in range(2):
    s3p = vae._split_pad(vae._to_padded(x), vae._w[name + ".norm1.gamma"], True)
    y = vae._conv_taps(s3p, (T, H, W), name + ".conv1")
torch.cuda.synchronize()
rows, cs = s3p.shape
print("rows", rows, "Cs", cs, "algorithmic bytes: operand", rows * cs * 2, "x 3 temporal taps if the 9 spatial taps hit L2; output", rows * 96 * 4,
      "flops", 2 * rows * 96 * 27 * cs)
