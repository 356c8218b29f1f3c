"""Attention at the full HunyuanVideo shape (118 980 tokens: an ODD grid of 465 CTAs, padded to whole multicast clusters): three 300-row
slices, first / middle / last, against fp32 SDPA.  python scripts/attn_fullshape_check.py"""
import sys, torch, torch.nn.functional as F
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from alg_b200 import ops
torch.manual_seed(0)
B, H, D, N = 1, 2, 128, 118980
q = torch.randn(B, N, H, D, device='cuda').bfloat16(); k = torch.randn(B, N, H, D, device='cuda').bfloat16()
v = torch.randn(B, N, H, D, device='cuda').bfloat16()
vt = torch.zeros(B, H, D, (N + 7) // 8 * 8, device='cuda', dtype=torch.bfloat16); vt[..., :N] = v.permute(0, 2, 3, 1)
o = ops.attention(q, k, vt, n_kv=N)
torch.cuda.synchronize()
for lo in (0, 59000, N - 300):
    sl = slice(lo, lo + 300)
    ref = F.scaled_dot_product_attention(q[:, sl].float().transpose(1, 2), k.float().transpose(1, 2), v.float().transpose(1, 2)).transpose(1, 2)
    e = float((o[:, sl].float() - ref).norm() / ref.norm())
    print('rows', lo, 'rel', e); assert e < 2 ** -8
print('grid', (N + 255) // 256, 'OK')
