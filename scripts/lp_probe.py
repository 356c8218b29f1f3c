"""Two launches of each low-pass kernel at saturating sizes, for ncu (see scripts/hbm_bench.py for the timed numbers)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lp_utils

x = torch.randn(1, 8192, 1, 60, 104, device="cuda")
y = torch.randn(1, 4096, 1, 90, 160, device="cuda").bfloat16()
z = torch.randn(1, 192, 480, 720, device="cuda").bfloat16()
for _ in range(2):
    lp_utils.apply_low_pass_filter(x, "down_up", 0.0, 0.0, 0.4)
    lp_utils.apply_low_pass_filter(y, "down_up", 0.0, 0.0, 0.625)
    lp_utils.apply_low_pass_filter(z, "gaussian_blur", 15.0, 0.02734375, 0.25)
torch.cuda.synchronize()
