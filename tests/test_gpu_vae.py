"""Native CogVideoX VAE encoder (alg_b200/vae_cogvideox.py) on the GPU: per-op parity of alg_im2col_bf16 /
alg_group_norm_bf16 against their emulations and torch, and the encoder against oracle/vae_oracle.py (fp32 ground truth
of the same bf16 weights; eager bf16 = what the reference runs through diffusers + cuDNN)."""
import pytest
import torch
import torch.nn.functional as F

from conftest import rel_l2

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("T,H,W,C,kernel,stride,pads,out_hw", [
    (1, 12, 20, 16, (3, 3, 3), (1, 1, 1), (2, 1, 1), None),
    (3, 7, 9, 8, (3, 3, 3), (1, 1, 1), (2, 1, 1), None),
    (1, 16, 24, 32, (1, 3, 3), (1, 2, 2), (0, 0, 0), (8, 12)),
    (1, 15, 23, 8, (1, 3, 3), (1, 2, 2), (0, 0, 0), (7, 11)),
    (1, 5, 6, 64, (1, 1, 1), (1, 1, 1), (0, 0, 0), None),
])
def test_im2col_bit_exact(T, H, W, C, kernel, stride, pads, out_hw):
    from alg_b200 import ops
    from oracle import ops_emulation as emu
    x = torch.randn(T * H * W, C, device="cuda").bfloat16()
    kw = dict(kernel=kernel, stride=stride, pad_t=pads[0], pad_top=pads[1], pad_left=pads[2], out_hw=out_hw)
    got = ops.im2col(x, T, H, W, **kw)
    ref = emu.im2col(x.cpu(), T, H, W, **kw)
    assert got.shape == ref.shape and torch.equal(got.cpu(), ref)
    ws = torch.full((got.numel() + 64,), 7.0, device="cuda").bfloat16()  # workspace variant writes only its head
    got2 = ops.im2col(x, T, H, W, out=ws, **kw)
    assert torch.equal(got2, got) and torch.all(ws[got.numel():] == 7.0)


def test_conv_as_im2col_gemm_matches_conv3d():
    from alg_b200 import ops
    H, W, Ci, Co = 24, 40, 64, 96
    x = torch.randn(H * W, Ci, device="cuda").bfloat16()
    w = (torch.randn(Co, Ci, 3, 3, 3, device="cuda") * (27 * Ci) ** -0.5).bfloat16()
    b = (0.1 * torch.randn(Co, device="cuda")).bfloat16()
    cols = ops.im2col(x, 1, H, W, kernel=(3, 3, 3), pad_t=2, pad_top=1, pad_left=1)
    got = ops.gemm(cols, w.movedim(1, -1).reshape(Co, -1).contiguous(), b)
    xc = x.view(1, H, W, Ci).permute(3, 0, 1, 2)[None].float()
    ref = F.conv3d(torch.cat([xc] * 3, dim=2), w.float(), b.float(), padding=(0, 1, 1))[0, :, 0].permute(1, 2, 0).reshape(-1, Co)
    assert rel_l2(got, ref) < 4e-3  # one bf16 rounding of the output


@pytest.mark.parametrize("rows,C,groups,silu", [(24 * 40, 128, 32, True), (1000, 256, 32, False), (333, 512, 32, True),
                                                (77, 32, 32, True), (4096, 64, 32, True)])
def test_group_norm_matches_torch(rows, C, groups, silu):
    from alg_b200 import ops
    from oracle import ops_emulation as emu
    x = (torch.randn(rows, C, device="cuda") * 2 + 0.7).bfloat16()
    w = (1 + 0.1 * torch.randn(C, device="cuda")).bfloat16()
    b = (0.1 * torch.randn(C, device="cuda")).bfloat16()
    got = ops.group_norm(x, groups, w, b, eps=1e-6, silu=silu)
    ref = emu.group_norm(x, groups, w, b, eps=1e-6, silu=silu)
    d = (got.float() - ref.float()).abs()
    # fp32 statistics reduced in a different order: a handful of 1-ulp bf16 flips, nothing else
    assert (d > 0).float().mean() < 2e-3 and d.max() <= 2 ** -6 * max(1.0, ref.float().abs().max().item())
    # torch's own GroupNorm on [1, C, rows] (fp32 math on the bf16 values) as the independent reference
    y = F.group_norm(x.float().t()[None], groups, w.float(), b.float(), eps=1e-6)[0].t()
    y = y.bfloat16().float()
    if silu:
        y = F.silu(y).bfloat16().float()
    assert rel_l2(got, y) < 3e-3
    # in place
    x2 = x.clone()
    ops.group_norm(x2, groups, w, b, eps=1e-6, silu=silu, out=x2)
    assert torch.equal(x2, got)


def _run_encoder(cfg_over, H, W, seed):
    from alg_b200 import vae_cogvideox as V
    from oracle import vae_oracle as Vo
    m = V.AutoencoderKLCogVideoX.from_synthetic(seed=seed, **cfg_over)
    sd = V.synthetic_state_dict(m._cfg, seed=seed)
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = torch.randn(1, 3, 1, H, W, generator=g, device="cuda").clamp(-1, 1).bfloat16()
    post = m.encode(x).latent_dist
    ref = Vo.encode_moments(x, sd, m._cfg, dtype=torch.float32)
    eager = Vo.encode_moments(x, sd, m._cfg, dtype=torch.bfloat16)
    return post, ref, eager


@pytest.mark.parametrize("over,H,W", [(dict(block_out_channels=(32, 64, 64, 64), layers_per_block=2, latent_channels=4), 64, 96),
                                      (dict(layers_per_block=1), 64, 96)])
def test_encoder_matches_oracle(over, H, W):
    post, ref, eager = _run_encoder(over, H, W, seed=2)
    assert post.parameters.shape == ref.shape
    e_native, e_eager = rel_l2(post.parameters, ref), rel_l2(eager, ref)
    assert e_native < 1.5 * e_eager + 1e-3, (e_native, e_eager)


def test_encoder_full_architecture_480x720():
    """CogVideoX-5b VAE encoder at the config-3 image size: the call the pixel-space ALG loop makes every step."""
    post, ref, eager = _run_encoder({}, 480, 720, seed=4)
    assert post.parameters.shape == (1, 32, 1, 60, 90)
    e_native, e_eager = rel_l2(post.parameters, ref), rel_l2(eager, ref)
    assert e_native < 1.5 * e_eager + 1e-3, (e_native, e_eager)
    g = torch.Generator(device="cuda").manual_seed(0)
    z = post.sample(g)
    assert z.shape == (1, 16, 1, 60, 90) and torch.isfinite(z).all()


def test_gemm_a_k_period_equals_repeated_operand():
    """alg_gemm_t.a_k_period: A [M, period] walked K / period times == the GEMM on A.repeat(1, K / period), bit for bit."""
    from alg_b200 import _lib, ops
    M, N, period, reps = 1000, 96, 192, 3
    a = torch.randn(M, period, device="cuda").bfloat16()
    w = (torch.randn(N, period * reps, device="cuda") * 0.05).bfloat16()
    b = torch.randn(N, device="cuda").bfloat16()
    r = torch.randn(M, N, device="cuda").bfloat16()
    got = ops.gemm(a, w, b, epilogue=_lib.EPI_RESIDUAL, residual=r, a_k_period=period)
    ref = ops.gemm(a.repeat(1, reps).contiguous(), w, b, epilogue=_lib.EPI_RESIDUAL, residual=r)
    assert torch.equal(got, ref)
    with pytest.raises(RuntimeError):
        ops.gemm(a[:, :100].contiguous(), w, a_k_period=100)  # not a multiple of 64
