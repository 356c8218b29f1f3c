"""Attention timing at the three models' self-attention shapes.  Usage: python scripts/attn_bench_shapes.py [reps]
(env ALG_ATTN_POLY / ALG_ATTN_SPLIT select the kernel variant)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from alg_b200 import ops

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
var = f"poly={os.environ.get('ALG_ATTN_POLY', 'default')} split={os.environ.get('ALG_ATTN_SPLIT', 'default')}"
SHAPES = (("wan", 1, 40, 128, 32760, 32760), ("cog", 2, 48, 64, 17776, 17776), ("hunyuan", 1, 24, 128, 118980, 118980),
          ("wan_cross_text", 3, 40, 128, 32760, 512), ("wan_cross_image", 3, 40, 128, 32760, 257))
only = os.environ.get("SHAPES")
for name, B, H, D, Nq, N in SHAPES:
    if only and name not in only.split(","):
        continue
    q = torch.randn(B, Nq, H, D, device="cuda").bfloat16()
    k = torch.randn(B, N, H, D, device="cuda").bfloat16()
    vt = torch.randn(B, H, D, (N + 7) // 8 * 8, device="cuda").bfloat16()
    o = torch.empty_like(q)
    r = reps if name != "hunyuan" else max(1, reps // 3)
    ops.attention(q, k, vt, n_kv=N, out=o)
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(r):
        ops.attention(q, k, vt, n_kv=N, out=o)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / r
    print(f"{var} {name}: B={B} H={H} D={D} Nq={Nq} Nkv={N}: {ms:.3f} ms {4 * B * H * Nq * N * D / ms / 1e9:.1f} TFLOP/s", flush=True)
    del q, k, vt, o
