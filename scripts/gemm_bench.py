"""GEMM timing at the Wan shapes.  Env ALG_GEMM_BN forces the N tile."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from alg_b200 import ops
torch.manual_seed(0)
M = int(os.environ.get("GEMM_M", "65520"))
for (N, K, epi) in ((5120, 5120, 0), (13824, 5120, 1), (5120, 13824, 2)):
    a = torch.randn(M, K, device="cuda").bfloat16()
    w = (torch.randn(N, K, device="cuda") * 0.02).bfloat16()
    b = torch.randn(N, device="cuda").bfloat16()
    res = torch.randn(M, N, device="cuda").bfloat16() if epi == 2 else None
    gate = torch.randn(1, N, device="cuda") if epi == 2 else None
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    for _ in range(2):
        ops.gemm(a, w, b, epilogue=epi, residual=res, gate=gate, out=out)
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(5):
        ops.gemm(a, w, b, epilogue=epi, residual=res, gate=gate, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    sl = slice(M - 512, M)  # the last rows: an odd M-tile count exercises the ragged / dummy tile of a CTA pair
    ref = torch.nn.functional.linear(a[sl], w, b).float()
    if epi == 1: ref = torch.nn.functional.gelu(ref.bfloat16().float(), approximate="tanh")
    if epi == 2: ref = res[sl].float() + ref.bfloat16().float() * gate
    err = float((out[sl].float() - ref).norm() / ref.norm())
    print(f"BN={os.environ.get('ALG_GEMM_BN','auto')} CL={os.environ.get('ALG_GEMM_CLUSTER','default')} gemm M{M} N{N} K{K} epi{epi}: {ms:.3f} ms {2*M*N*K/ms/1e9:.1f} TFLOP/s  rel {err:.2e}", flush=True)
