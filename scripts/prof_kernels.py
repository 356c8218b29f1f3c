"""Launch the hot kernels once at the Wan-I2V-14B shapes so `ncu --set full -k regex:...` can capture them."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from alg_b200 import ops, _lib

which = sys.argv[1] if len(sys.argv) > 1 else "attn"
torch.manual_seed(0)
if which == "attn":
    B, H, D, N = 1, int(os.environ.get("HEADS", "8")), 128, 32760
    q = torch.randn(B, N, H, D, device="cuda").bfloat16()
    k = torch.randn(B, N, H, D, device="cuda").bfloat16()
    vt = torch.randn(B, H, D, N, device="cuda").bfloat16()
    for _ in range(2):
        ops.attention(q, k, vt)
elif which == "gemm":
    M = 65520
    for (N, K, epi) in ((5120, 5120, 0), (13824, 5120, 1), (5120, 13824, 2)):
        a = torch.randn(M, K, device="cuda").bfloat16()
        w = (torch.randn(N, K, device="cuda") * 0.02).bfloat16()
        b = torch.randn(N, device="cuda").bfloat16()
        res = torch.randn(M, N, device="cuda").bfloat16() if epi == 2 else None
        gate = torch.randn(1, N, device="cuda") if epi == 2 else None
        ops.gemm(a, w, b, epilogue=epi, residual=res, gate=gate)
elif which == "lowpass":
    import lp_utils
    x = torch.randn(1, 8192, 1, 60, 104, device="cuda")
    for _ in range(2):
        lp_utils.apply_low_pass_filter(x, "down_up", 0.0, 0.0, 0.4)
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(10):
        lp_utils.apply_low_pass_filter(x, "down_up", 0.0, 0.0, 0.4)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print("down_up 8192 planes ms", ms, "GB/s", 2 * x.numel() * 4 / ms / 1e6)
    x = torch.randn(1, 20, 21, 60, 104, device="cuda")
    for _ in range(3):
        lp_utils.apply_low_pass_filter(x, "down_up", 0.0, 0.0, 0.4)
    e0.record()
    for _ in range(20):
        lp_utils.apply_low_pass_filter(x, "down_up", 0.0, 0.0, 0.4)
    e1.record(); torch.cuda.synchronize()
    print("down_up wan shape us/call", e0.elapsed_time(e1) * 1000 / 20)
torch.cuda.synchronize()
