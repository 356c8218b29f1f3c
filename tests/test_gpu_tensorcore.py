"""tcgen05 GEMM and attention (through the C ABI) vs fp32 PyTorch references of the same op.

Tolerances: outputs are bf16, so one rounding = 2^-9 relative per element; GEMM/attention results are required
within 2^-8 relative L2 of the fp32 reference."""
import pytest
import torch
import torch.nn.functional as F

from conftest import rel_l2

pytestmark = pytest.mark.gpu
TOL = 2 ** -8


def _gemm_ref(a, w, b, epi, res, gate, per_row):
    ref = a.float() @ w.float().t()
    if b is not None:
        ref = ref + (b.float()[:, None] if per_row else b.float()[None, :])
    if epi == 1:
        ref = F.gelu(ref.bfloat16().float(), approximate="tanh")
    elif epi == 4:
        ref = F.gelu(ref.bfloat16().float())
    elif epi == 5:
        ref = F.silu(ref.bfloat16().float())
    elif epi == 2:
        ref = res.float() + ref.bfloat16().float() * gate
    elif epi == 3:
        ref = res.float() + ref.bfloat16().float()
    return ref


@pytest.mark.parametrize("M,N,K,epi,per_row,f32", [
    (128, 256, 64, 0, False, True), (300, 256, 512, 0, False, False), (128, 128, 128, 0, False, False),
    (128, 64, 128, 0, False, False), (1000, 1280, 1280, 4, False, False), (1, 30720, 5120, 0, False, False),
    (512, 5120, 4096, 1, False, False), (777, 5120, 144, 0, False, False), (640, 72, 512, 0, False, False),
    (5120, 3000, 5120, 0, True, False), (5120, 257, 5120, 0, True, False), (4096, 5120, 5120, 2, False, False),
    (4096, 5120, 5120, 3, False, False), (4096, 64, 5120, 0, False, False), (333, 13824, 5120, 1, False, False),
    (2049, 5120, 13824, 2, False, False), (64, 512, 8, 5, False, False)])
def test_gemm(M, N, K, epi, per_row, f32):
    from alg_b200 import ops
    torch.manual_seed(M + N + K)
    a = torch.randn(M, K, device="cuda").bfloat16()
    w = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
    b = torch.randn(M if per_row else N, device="cuda").bfloat16()
    res = torch.randn(M, N, device="cuda").bfloat16() if epi in (2, 3) else None
    gate = torch.randn(1, N, device="cuda") if epi == 2 else None
    out = ops.gemm(a, w, b, epilogue=epi, residual=res, gate=gate, bias_per_row=per_row,
                   out_dtype=torch.float32 if f32 else torch.bfloat16)
    assert rel_l2(out, _gemm_ref(a, w, b, epi, res, gate, per_row)) < TOL


def test_gemm_exact_small_integers():
    """Known-answer test: small-integer operands make every product and sum exact, so the result is bit-exact."""
    from alg_b200 import ops
    torch.manual_seed(0)
    a = torch.randint(-4, 5, (384, 320), device="cuda").bfloat16()
    w = torch.randint(-4, 5, (512, 320), device="cuda").bfloat16()
    out = ops.gemm(a, w, None, out_dtype=torch.float32)
    assert torch.equal(out, a.float() @ w.float().t())


def test_gemm_in_place_residual_and_linearity():
    from alg_b200 import ops
    torch.manual_seed(1)
    a = torch.randn(1024, 512, device="cuda").bfloat16()
    w = (torch.randn(768, 512, device="cuda") * 0.05).bfloat16()
    x = torch.randn(1024, 768, device="cuda").bfloat16()
    want = ops.gemm(a, w, None, epilogue=3, residual=x)
    x2 = x.clone()
    ops.gemm(a, w, None, epilogue=3, residual=x2, out=x2)  # D aliases R, as the engine's residual stream does
    assert torch.equal(x2, want)
    y1 = ops.gemm(a, w, None, out_dtype=torch.float32)
    y2 = ops.gemm((2 * a.float()).bfloat16(), w, None, out_dtype=torch.float32)
    assert torch.equal(y2, 2 * y1)  # scaling by a power of two is exact


def _attn_ref(q, k, v):
    return F.scaled_dot_product_attention(q.float().transpose(1, 2), k.float().transpose(1, 2), v.float().transpose(1, 2)).transpose(1, 2)


def _vt(v, pad_to=8):
    B, N, H, D = v.shape
    pad = (N + pad_to - 1) // pad_to * pad_to
    vt = torch.zeros(B, H, D, pad, device=v.device, dtype=v.dtype)
    vt[..., :N] = v.permute(0, 2, 3, 1)
    return vt


@pytest.mark.parametrize("B,H,D,Nq,Nkv,sq", [(1, 1, 128, 256, 128, 1.0), (2, 3, 128, 512, 1024, 1.0), (1, 2, 128, 300, 257, 1.0),
                                             (1, 2, 128, 100, 77, 1.0), (1, 2, 128, 1000, 2000, 4.0), (1, 2, 64, 256, 256, 1.0),
                                             (2, 3, 64, 700, 1000, 1.0), (1, 4, 128, 4096, 4096, 3.0), (3, 2, 128, 1, 512, 1.0),
                                             (1, 48, 64, 520, 1226, 2.0), (2, 2, 128, 600, 1500, 1.0), (1, 3, 64, 1300, 1100, 1.0)])
def test_attention(B, H, D, Nq, Nkv, sq):
    from alg_b200 import ops
    torch.manual_seed(Nq + Nkv)
    q = (torch.randn(B, Nq, H, D, device="cuda") * sq).bfloat16()
    k = torch.randn(B, Nkv, H, D, device="cuda").bfloat16()
    v = torch.randn(B, Nkv, H, D, device="cuda").bfloat16()
    o = ops.attention(q, k, _vt(v), n_kv=Nkv)
    assert rel_l2(o, _attn_ref(q, k, v)) < TOL


def test_attention_accumulate_is_bf16_sum():
    from alg_b200 import ops
    torch.manual_seed(5)
    q, k, v = (torch.randn(1, 520, 2, 128, device="cuda").bfloat16() for _ in range(3))
    base = torch.randn(1, 520, 2, 128, device="cuda").bfloat16()
    plain = ops.attention(q, k, _vt(v))
    acc = ops.attention(q, k, _vt(v), out=base.clone(), accumulate=True)
    assert torch.equal(acc, (plain.float() + base.float()).bfloat16())  # hidden_states + hidden_states_img


def test_attention_rescale_path_large_logit_growth():
    """Keys sorted so the running max keeps growing by > 2^8: exercises the lazy O-rescale in TMEM."""
    from alg_b200 import ops
    torch.manual_seed(6)
    B, H, D, N = 1, 2, 128, 2048
    q = torch.randn(B, 256, H, D, device="cuda").bfloat16()
    k = (torch.randn(B, N, H, D, device="cuda") * torch.linspace(0.1, 6.0, N, device="cuda")[None, :, None, None]).bfloat16()
    v = torch.randn(B, N, H, D, device="cuda").bfloat16()
    assert rel_l2(ops.attention(q, k, _vt(v)), _attn_ref(q, k, v)) < TOL


def test_attention_full_size_properties():
    """Wan config size (N = 32 760, d_head 128): softmax rows sum to one (V = const => O = const) and the result
    is invariant under a permutation of the keys."""
    from alg_b200 import ops
    torch.manual_seed(7)
    B, H, D, N = 1, 2, 128, 32760
    q = torch.randn(B, N, H, D, device="cuda").bfloat16()
    k = torch.randn(B, N, H, D, device="cuda").bfloat16()
    ones = torch.full((B, H, D, N), 0.5, device="cuda", dtype=torch.bfloat16)
    o = ops.attention(q, k, ones)
    assert float((o.float() - 0.5).abs().max()) < 2 ** -8
    v = torch.randn(B, N, H, D, device="cuda").bfloat16()
    perm = torch.randperm(N, device="cuda")
    o1 = ops.attention(q[:, :1024], k, _vt(v))
    o2 = ops.attention(q[:, :1024], k[:, perm].contiguous(), _vt(v[:, perm].contiguous()))
    assert rel_l2(o1, o2) < TOL
    ref = _attn_ref(q[:, :512], k, v)
    assert rel_l2(o1[:, :512], ref) < TOL


@pytest.mark.parametrize("N", [72, 96, 160, 192, 320, 384])
def test_gemm_narrow_tiles_96_and_192(N):
    """The 96- and 192-wide tiles (N of the float32 VAE levels): plain and fp32-output GEMMs with fp32 bias / residual."""
    import torch
    from alg_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(N)
    M, K = 1000, 448
    a = torch.randn(M, K, generator=g, device="cuda").bfloat16()
    w = (torch.randn(N, K, generator=g, device="cuda") * K ** -0.5).bfloat16()
    b = torch.randn(N, generator=g, device="cuda")
    r = torch.randn(M, N, generator=g, device="cuda")
    ref = a.float() @ w.float().t()
    out = ops.gemm(a, w, None, out_dtype=torch.float32)
    assert float((out - ref).abs().max()) < 1e-3 * float(ref.abs().max())
    out = torch.empty(M, N, device="cuda")
    ops.gemm(a, w, None, out=out, out_dtype=torch.float32, bias_f32=b, residual_f32=r)
    assert float((out - (ref + b + r)).abs().max()) < 1e-3 * float(ref.abs().max())
    out16 = ops.gemm(a, w, b.bfloat16())
    assert float((out16.float() - (ref + b.bfloat16().float())).abs().max()) < 2 ** -7 * float(ref.abs().max())


@pytest.mark.parametrize("env", [dict(ALG_ATTN_S128="1"), dict(ALG_ATTN_PS="1"), dict(ALG_ATTN_PS="2"), dict(ALG_ATTN_PAIR="1"),
                                 dict(ALG_ATTN_MC="0"), dict(ALG_ATTN_MC="2"), dict(ALG_ATTN_FX="1", ALG_ATTN_MC="0")])
def test_attention_experimental_variants_keep_parity(env):
    """The schedules kept behind knobs (profiles/r02_attention_s128.md, r02_pair_mma.md) stay correct: each runs in a fresh process
    (the knobs are read once per process) on a long ragged head_dim-128 problem against fp32 SDPA."""
    import os
    import subprocess
    import sys
    code = (
        "import torch, torch.nn.functional as F\n"
        "from alg_b200 import ops\n"
        "torch.manual_seed(0)\n"
        "B, H, D, Nq, Nkv = 1, 3, 128, 700, 2500\n"
        "q = (torch.randn(B, Nq, H, D, device='cuda') * 3).bfloat16(); k = torch.randn(B, Nkv, H, D, device='cuda').bfloat16()\n"
        "v = torch.randn(B, Nkv, H, D, device='cuda').bfloat16()\n"
        "pad = (Nkv + 7) // 8 * 8\n"
        "vt = torch.zeros(B, H, D, pad, device='cuda', dtype=torch.bfloat16); vt[..., :Nkv] = v.permute(0, 2, 3, 1)\n"
        "o = ops.attention(q, k, vt, n_kv=Nkv)\n"
        "ref = F.scaled_dot_product_attention(q.float().transpose(1, 2), k.float().transpose(1, 2), v.float().transpose(1, 2)).transpose(1, 2)\n"
        "e = float((o.float() - ref).norm() / ref.norm()); print('REL', e); assert e < 2 ** -8, e\n")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", code], cwd=root, env=dict(os.environ, **env), capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "REL" in r.stdout, (env, r.stdout[-500:], r.stderr[-1500:])


@pytest.mark.parametrize("amp", [1.5, 2.0, 3.0])
def test_attention_fixed_reference_range_guard(amp):
    """ALG_ATTN_FX=1 (fixed softmax reference): rows whose later scores tower over the first 64 keys by 100-400 log2 units must hit
    the range guard (reference raised by exact powers of two, step recomputed) and still match fp32 SDPA; in a fresh process
    because the knob is read once."""
    import os
    import subprocess
    import sys
    code = (
        "import torch, torch.nn.functional as F\n"
        "from alg_b200 import ops\n"
        "torch.manual_seed(1)\n"
        f"B, H, D, Nq, Nkv, amp = 1, 2, 128, 300, 2000, {amp}\n"
        "q = (4 + 0.25 * torch.randn(B, Nq, H, D, device='cuda')).bfloat16()\n"
        "k = (0.25 * torch.randn(B, Nkv, H, D, device='cuda')); k[:, :64] -= amp; k[:, 64:1000] += 0.2 * amp; k[:, 1000:] += amp\n"
        "k = k.bfloat16(); v = torch.randn(B, Nkv, H, D, device='cuda').bfloat16()\n"
        "pad = (Nkv + 7) // 8 * 8\n"
        "vt = torch.zeros(B, H, D, pad, device='cuda', dtype=torch.bfloat16); vt[..., :Nkv] = v.permute(0, 2, 3, 1)\n"
        "o = ops.attention(q, k, vt, n_kv=Nkv)\n"
        "ref = F.scaled_dot_product_attention(q.float().transpose(1, 2), k.float().transpose(1, 2), v.float().transpose(1, 2)).transpose(1, 2)\n"
        "assert torch.isfinite(o.float()).all()\n"
        "e = float((o.float() - ref).norm() / ref.norm()); print('REL', e); assert e < 2 ** -7, e\n")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, ALG_ATTN_FX="1", ALG_ATTN_MC="0")
    r = subprocess.run([sys.executable, "-c", code], cwd=root, env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "REL" in r.stdout, (r.stdout[-500:], r.stderr[-1500:])
