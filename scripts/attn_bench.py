"""Attention kernel check + timing at the Wan self-attention shape.  Usage: python scripts/attn_bench.py [reps] [heads]
Env ALG_ATTN_VARIANT selects a kernel variant (see attention.cu)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.nn.functional as F
from alg_b200 import ops


def rel(a, b):
    a, b = a.float(), b.float()
    return float((a - b).norm() / b.norm())


def check(B, H, D, Nq, Nkv, scale_q=1.0):
    torch.manual_seed(0)
    q = (torch.randn(B, Nq, H, D, device="cuda") * scale_q).bfloat16()
    k = torch.randn(B, Nkv, H, D, device="cuda").bfloat16()
    v = torch.randn(B, Nkv, H, D, device="cuda").bfloat16()
    pad = (Nkv + 7) // 8 * 8
    vt = torch.zeros(B, H, D, pad, device="cuda", dtype=torch.bfloat16)
    vt[..., :Nkv] = v.permute(0, 2, 3, 1)
    o = ops.attention(q, k, vt, n_kv=Nkv)
    torch.cuda.synchronize()
    ref = F.scaled_dot_product_attention(q.float().transpose(1, 2), k.float().transpose(1, 2), v.float().transpose(1, 2)).transpose(1, 2)
    return rel(o, ref)


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
    H = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    var = "poly=" + os.environ.get("ALG_ATTN_POLY", "default") + " split=" + os.environ.get("ALG_ATTN_SPLIT", "default")
    errs = [check(1, 2, 128, 1000, 2000, 4.0), check(2, 3, 128, 512, 1024), check(1, 2, 128, 300, 257), check(1, 2, 64, 700, 1000),
            check(1, 4, 128, 4096, 4096, 3.0)]
    B, D, N = 1, 128, 32760
    q = torch.randn(B, N, H, D, device="cuda").bfloat16()
    k = torch.randn(B, N, H, D, device="cuda").bfloat16()
    vt = torch.randn(B, H, D, N, device="cuda").bfloat16()
    o = torch.empty_like(q)
    for _ in range(2):
        ops.attention(q, k, vt, out=o)
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        ops.attention(q, k, vt, out=o)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    fl = 4 * B * H * N * N * D
    print(f"variant {var}: errs {' '.join(f'{e:.2e}' for e in errs)} | wan shape H={H} reps={reps}: {ms:.2f} ms {fl/ms/1e9:.1f} TFLOP/s", flush=True)


if __name__ == "__main__":
    main()
