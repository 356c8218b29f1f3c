"""Per-op parity of the HBM-bound DiT building blocks (C ABI: alg_layer_norm, alg_head_norm_rope, alg_patch_gather,
alg_unpatchify, alg_timestep_embedding, ...) and of the ABI-2 GEMM gating variants against their eager restatements in
oracle/ops_emulation.py.  bf16 outputs: bit-exact or within one bf16 ulp of the emulation (reductions are ordered
differently from ATen's), so the bar is rel-L2 <= 2^-9 with at most a small fraction of elements off by one ulp."""
import pytest
import torch

from conftest import rel_l2

pytestmark = pytest.mark.gpu


def _close(a, b, tol=2e-3, frac=0.02):
    a, b = a.float(), b.float()
    assert rel_l2(a, b) < tol, rel_l2(a, b)
    mism = (a != b).float().mean().item()
    assert mism <= frac, mism


@pytest.mark.parametrize("rows,d", [(37, 3072), (300, 5120), (5, 64), (64, 8192)])
@pytest.mark.parametrize("mode", ["plain", "affine_f32", "affine_bf16_chain", "mod_f32", "mod_bf16_chain_split"])
def test_layer_norm(rows, d, mode):
    from alg_b200 import ops
    from oracle import ops_emulation as emu
    g = torch.Generator(device="cuda").manual_seed(rows * 7 + d)
    x = (torch.randn(rows, d, generator=g, device="cuda") * 1.7 + 0.3).bfloat16()
    kw = dict(eps=1e-5)
    rn = lambda *s, dt=torch.float32: (torch.randn(*s, generator=g, device="cuda") * 0.3).to(dt)  # noqa: E731
    if mode == "affine_f32":
        kw.update(weight=1 + rn(d), bias=rn(d))
    elif mode == "affine_bf16_chain":
        kw.update(weight=(1 + rn(d)).bfloat16(), bias=rn(d, dt=torch.bfloat16), scale=rn(d, dt=torch.bfloat16),
                  shift=rn(d, dt=torch.bfloat16), chain_bf16=True)
    elif mode == "mod_f32":
        kw.update(scale=rn(d), shift=rn(d))
    elif mode == "mod_bf16_chain_split":
        rpb = max(rows // 2, 1)
        mods = rn(2, 4, d, dt=torch.bfloat16)
        kw.update(scale=mods[:, 0], shift=mods[:, 1], scale_alt=mods[:, 2], shift_alt=mods[:, 3], rows_per_batch=rpb,
                  split_row=rpb // 3 + 1, chain_bf16=True)
        x = x[: 2 * rpb].contiguous()
    got = ops.layer_norm(x, **kw)
    ref = emu.layer_norm(x, **kw)
    _close(got, ref)


@pytest.mark.parametrize("hd,heads", [(64, 48), (128, 24), (64, 3), (128, 1)])
@pytest.mark.parametrize("kind", ["rms", "layer", "none"])
def test_head_norm_rope(hd, heads, kind):
    from alg_b200 import _lib, ops
    from oracle import ops_emulation as emu
    g = torch.Generator(device="cuda").manual_seed(hd + heads)
    rpb, n_txt = 45, 7
    rows = 2 * rpb
    x = torch.randn(rows, heads * hd, generator=g, device="cuda").bfloat16()
    w = (1 + 0.2 * torch.randn(hd, generator=g, device="cuda")).bfloat16()
    b = (0.2 * torch.randn(hd, generator=g, device="cuda")).bfloat16()
    ang = torch.rand(rpb - n_txt, hd // 2, generator=g, device="cuda") * 6.28
    cos, sin = ang.cos().repeat_interleave(2, dim=1).contiguous(), ang.sin().repeat_interleave(2, dim=1).contiguous()
    kw = dict(norm_kind={"rms": _lib.NORM_RMS, "layer": _lib.NORM_LAYER, "none": _lib.NORM_NONE}[kind], weight=w,
              bias=b if kind == "layer" else None, eps=1e-6, cos=cos, sin=sin, rows_per_batch=rpb, rope_row0=n_txt,
              rope_rows=rpb - n_txt)
    ref = emu.head_norm_rope(x.clone(), heads, hd, **kw)
    got = ops.head_norm_rope(x.clone(), heads, hd, **kw)
    _close(got, ref)
    # text rows (no RoPE) of the un-normalised variant must be untouched
    if kind == "none":
        assert torch.equal(got[:n_txt], x[:n_txt])


def test_patch_gather_and_unpatchify_layouts():
    from alg_b200 import ops
    from oracle import ops_emulation as emu
    g = torch.Generator(device="cuda").manual_seed(0)
    T, H, W = 3, 8, 12
    # Wan layout: fp32 [C, T, H, W] sources; Cog layout: bf16 [F, C, H, W] transposed views; Hy: frame-0 override
    lat = torch.randn(16, T, H, W, generator=g, device="cuda")
    cond = [torch.randn(20, T, H, W, generator=g, device="cuda") for _ in range(3)]
    passes = [[lat, c] for c in cond]
    N = T * (H // 2) * (W // 2)
    out = torch.zeros(3 * N, 144, device="cuda", dtype=torch.bfloat16)
    assert torch.equal(ops.patch_gather(passes, out), emu.patch_gather(passes, torch.zeros_like(out)))
    lat_c = torch.randn(T, 16, H, W, generator=g, device="cuda").bfloat16()
    img_c = torch.randn(T, 16, H, W, generator=g, device="cuda").bfloat16()
    passes = [[lat_c.transpose(0, 1), img_c.transpose(0, 1)]] * 2
    out = torch.zeros(2 * N, 128, device="cuda", dtype=torch.bfloat16)
    assert torch.equal(ops.patch_gather(passes, out), emu.patch_gather(passes, torch.zeros_like(out)))
    first = torch.randn(16, 1, H, W, generator=g, device="cuda")
    passes = [[(lat, first)]]
    out = torch.zeros(N, 64, device="cuda", dtype=torch.bfloat16)
    assert torch.equal(ops.patch_gather(passes, out), emu.patch_gather(passes, torch.zeros_like(out)))
    # unpatchify: both column orders, contiguous [P, C, T, H, W] and Cog's [P, F, C, H, W] storage
    proj = torch.randn(2 * N, 64, generator=g, device="cuda").bfloat16()
    for cm in (False, True):
        o1 = torch.empty(2, 16, T, H, W, device="cuda", dtype=torch.bfloat16)
        assert torch.equal(ops.unpatchify(proj, o1, cm), emu.unpatchify(proj, torch.empty_like(o1), cm))
    store = torch.empty(2, T, 16, H, W, device="cuda", dtype=torch.bfloat16)
    ops.unpatchify(proj, store.transpose(1, 2), True)
    assert torch.equal(store.transpose(1, 2), emu.unpatchify(proj, torch.empty(2, 16, T, H, W, device="cuda", dtype=torch.bfloat16), True))


def test_small_vector_ops():
    from alg_b200 import ops
    from oracle import ops_emulation as emu
    for t, dim in ((999.0, 3072), (0.0, 256), (6000.0, 256), (17.0, 64)):
        for dt in (torch.float32, torch.bfloat16):
            got, ref = ops.timestep_embedding(t, dim, dt, "cuda"), emu.timestep_embedding(t, dim, dt, "cuda")
            assert (got.float() - ref.float()).abs().max() < (2e-3 if dt == torch.float32 else 1e-2)
    g = torch.Generator(device="cuda").manual_seed(1)
    a, b = (torch.randn(3072, generator=g, device="cuda").bfloat16() for _ in range(2))
    assert torch.equal(ops.add(a, b), emu.add(a, b))
    _close(ops.silu(a), emu.silu(a))
    x = torch.randn(77, 4096, generator=g, device="cuda").bfloat16()
    _close(ops.mean_rows(x), emu.mean_rows(x))
    dst = torch.zeros(77, 5000, device="cuda", dtype=torch.bfloat16)
    ops.copy_rows(x, dst[:, 8:4104])
    assert torch.equal(dst[:, 8:4104], x) and not dst[:, :8].any() and not dst[:, 4104:].any()


@pytest.mark.parametrize("gate_dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("gate_round", [False, True])
def test_gemm_gate_variants(gate_dtype, gate_round):
    from alg_b200 import _lib, ops
    from oracle import ops_emulation as emu
    g = torch.Generator(device="cuda").manual_seed(3)
    rpb, K, N = 200, 256, 384
    M = 2 * rpb
    a = torch.randn(M, K, generator=g, device="cuda").bfloat16()
    w = (torch.randn(N, K, generator=g, device="cuda") * 0.05).bfloat16()
    bias = torch.randn(N, generator=g, device="cuda").bfloat16()
    res = torch.randn(M, N, generator=g, device="cuda").bfloat16()
    gate = torch.randn(2, N, generator=g, device="cuda").to(gate_dtype)
    gate_alt = torch.randn(2, N, generator=g, device="cuda").to(gate_dtype)
    kw = dict(epilogue=_lib.EPI_GATE_RESIDUAL, residual=res, gate=gate, gate_alt=gate_alt, gate_split_row=37,
              gate_round=gate_round, rows_per_batch=rpb)
    _close(ops.gemm(a, w, bias, **kw), emu.gemm(a, w, bias, **kw), frac=0.05)
    # column-offset output inside a wider buffer (HunyuanVideo single-stream concat) and V^T placement
    wide = torch.zeros(M, 1024, device="cuda", dtype=torch.bfloat16)
    ops.gemm(a, w, bias, epilogue=_lib.EPI_GELU_TANH, out=wide[:, 512:512 + N])
    _close(wide[:, 512:512 + N], emu.gemm(a, w, bias, epilogue=_lib.EPI_GELU_TANH), tol=4e-3, frac=0.2)
    assert not wide[:, :512].any() and not wide[:, 512 + N:].any()
