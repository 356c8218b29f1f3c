"""CPU checks of the native CogVideoX VAE encoder's host side: alg_b200/vae_cogvideox.py only orders C-ABI calls, so with
``ops`` swapped for the eager emulation (oracle/ops_emulation.py, test infrastructure) it must reproduce the encoder oracle
(oracle/vae_oracle.py: F.conv3d / F.group_norm on the diffusers-named weights) -- weight re-layout, causal temporal taps,
the asymmetric down-sampling pad, residual / shortcut wiring.  The product path has no CPU mode."""
import pytest
import torch

from conftest import rel_l2

TINY = dict(block_out_channels=(32, 64, 64, 64), layers_per_block=2, latent_channels=4)


def test_im2col_emulation_is_the_conv():
    """cols @ W_perm^T == F.conv3d with the causal front replication (and the stride-2 (0,1,0,1)-padded conv2d)."""
    import torch.nn.functional as F
    from oracle import ops_emulation as emu
    g = torch.Generator().manual_seed(0)
    T, H, W, Ci, Co = 3, 6, 10, 8, 5
    x = torch.randn(T * H * W, Ci, generator=g)
    w = torch.randn(Co, Ci, 3, 3, 3, generator=g)
    cols = emu.im2col(x, T, H, W, kernel=(3, 3, 3), pad_t=2, pad_top=1, pad_left=1)
    got = cols @ w.movedim(1, -1).reshape(Co, -1).t()
    xc = x.view(T, H, W, Ci).permute(3, 0, 1, 2)[None]
    ref = F.conv3d(torch.cat([xc[:, :, :1]] * 2 + [xc], dim=2), w, padding=(0, 1, 1))[0].permute(1, 2, 3, 0).reshape(-1, Co)
    assert torch.allclose(got, ref, atol=1e-4)
    w2 = torch.randn(Co, Ci, 3, 3, generator=g)
    cols2 = emu.im2col(x[: H * W], 1, H, W, kernel=(1, 3, 3), stride=(1, 2, 2), out_hw=(H // 2, W // 2))
    got2 = cols2 @ w2.movedim(1, -1).reshape(Co, -1).t()
    x2 = x[: H * W].view(H, W, Ci).permute(2, 0, 1)[None]
    ref2 = F.conv2d(F.pad(x2, (0, 1, 0, 1)), w2, stride=2)[0].permute(1, 2, 0).reshape(-1, Co)
    assert torch.allclose(got2, ref2, atol=1e-4)


def test_vae_encoder_sequencer_matches_oracle(monkeypatch):
    from alg_b200 import vae_cogvideox as V
    from oracle import ops_emulation as emu, vae_oracle as Vo
    monkeypatch.setattr(V, "ops", emu)
    m = V.AutoencoderKLCogVideoX(**TINY)
    sd = V.synthetic_state_dict(m._cfg, seed=3, device="cpu")
    m.load_state_dict(sd)
    g = torch.Generator().manual_seed(1)
    x = torch.randn(2, 3, 1, 32, 48, generator=g).bfloat16()
    post = m.encode(x).latent_dist
    ref = Vo.encode_moments(x, sd, m._cfg, dtype=torch.float32)
    assert post.parameters.shape == (2, 8, 1, 4, 6) and post.parameters.dtype == torch.bfloat16
    # protocol of the DiT tests: no further from the fp32 evaluation of the same bf16 weights than eager bf16 is (x1.5)
    eager = Vo.encode_moments(x, sd, m._cfg, dtype=torch.bfloat16)
    e_native, e_eager = rel_l2(post.parameters, ref), rel_l2(eager, ref)
    assert e_native < 1.5 * e_eager + 1e-3, (e_native, e_eager)
    # latent_dist surface: mode, clamp, seeded sample == mean + std * draw on the caller's generator
    assert torch.equal(post.mode(), post.parameters[:, :4])
    g1, g2 = torch.Generator().manual_seed(5), torch.Generator().manual_seed(5)
    z = post.sample(g1)
    noise = torch.randn(post.mean.shape, generator=g2, dtype=post.parameters.dtype)
    assert torch.equal(z, Vo.sample(post.parameters, noise))


def test_vae_surface_and_errors():
    from alg_b200 import vae_cogvideox as V
    m = V.AutoencoderKLCogVideoX()
    assert m.config.scaling_factor == 0.7 and m.config.temporal_compression_ratio == 4 and not m.config.invert_scale_latents
    assert len(m.config.block_out_channels) == 4
    shapes = V.encoder_parameter_shapes(m._cfg)
    assert shapes["encoder.conv_in.conv.weight"] == (128, 3, 3, 3, 3)
    assert shapes["encoder.down_blocks.1.resnets.0.conv_shortcut.weight"] == (256, 128, 1, 1, 1)
    assert shapes["encoder.down_blocks.2.downsamplers.0.conv.weight"] == (256, 256, 3, 3)
    assert "encoder.down_blocks.3.downsamplers.0.conv.weight" not in shapes
    assert shapes["encoder.conv_out.conv.weight"] == (32, 512, 3, 3, 3)
    with pytest.raises(RuntimeError):
        m.encode(torch.zeros(1, 3, 1, 16, 16))
    t = V.AutoencoderKLCogVideoX(block_out_channels=(32, 32, 32, 32), layers_per_block=1)
    t.load_state_dict(V.synthetic_state_dict(t._cfg, device="cpu"))
    with pytest.raises(NotImplementedError):
        t.encode(torch.zeros(1, 3, 5, 16, 16))
    with pytest.raises(ValueError):
        t.encode(torch.zeros(1, 3, 1, 20, 16))
    with pytest.raises(NotImplementedError):
        t.decode(torch.zeros(1, 16, 1, 2, 2))
    with pytest.raises(RuntimeError):  # the product path has no CPU mode
        t.encode(torch.zeros(1, 3, 1, 16, 16))


def test_patch_gather_and_group_norm_properties():
    """Seeded sweep over random geometries (CPU): the patch-gather + matmul formulation IS conv3d / strided conv2d with
    CogVideoX's padding rules, and the GroupNorm emulation IS torch's group_norm (then one bf16 rounding)."""
    import random

    import torch.nn.functional as F
    from oracle import ops_emulation as emu
    rnd = random.Random(7)
    g = torch.Generator().manual_seed(7)
    for _ in range(12):
        T, H, W = rnd.randint(1, 4), rnd.randint(3, 9), rnd.randint(3, 11)
        Ci, Co = 8 * rnd.randint(1, 3), rnd.randint(1, 6)
        x = torch.randn(T * H * W, Ci, generator=g)
        w = torch.randn(Co, Ci, 3, 3, 3, generator=g)
        cols = emu.im2col(x, T, H, W, kernel=(3, 3, 3), pad_t=2, pad_top=1, pad_left=1)
        xc = x.view(T, H, W, Ci).permute(3, 0, 1, 2)[None]
        ref = F.conv3d(torch.cat([xc[:, :, :1]] * 2 + [xc], dim=2), w, padding=(0, 1, 1))[0].permute(1, 2, 3, 0)
        assert torch.allclose(cols @ w.movedim(1, -1).reshape(Co, -1).t(), ref.reshape(-1, Co), atol=2e-4)
        # CogVideoXDownsample3D: F.pad (0, 1, 0, 1) then Conv2d(3, stride 2, padding 0), frame by frame
        w2 = torch.randn(Co, Ci, 3, 3, generator=g)
        Ho, Wo = (H + 1 - 3) // 2 + 1, (W + 1 - 3) // 2 + 1
        cols2 = emu.im2col(x, T, H, W, kernel=(1, 3, 3), stride=(1, 2, 2), out_hw=(Ho, Wo))
        x2 = x.view(T, H, W, Ci).permute(0, 3, 1, 2)
        ref2 = F.conv2d(F.pad(x2, (0, 1, 0, 1)), w2, stride=2).permute(0, 2, 3, 1).reshape(-1, Co)
        assert torch.allclose(cols2 @ w2.movedim(1, -1).reshape(Co, -1).t(), ref2, atol=2e-4)
    for _ in range(8):
        rows, groups = rnd.randint(5, 300), rnd.choice([1, 2, 4, 8])
        Cc = groups * rnd.choice([1, 2, 4, 8, 16])
        if Cc % 8:
            Cc *= 8
        x = (torch.randn(rows, Cc, generator=g) * 3 + 1).bfloat16()
        w, b = (1 + 0.2 * torch.randn(Cc, generator=g)).bfloat16(), (0.3 * torch.randn(Cc, generator=g)).bfloat16()
        got = emu.group_norm(x, groups, w, b, eps=1e-6, silu=False).float()
        ref = F.group_norm(x.float().t()[None], groups, w.float(), b.float(), eps=1e-6)[0].t()
        assert (got - ref).abs().max() <= 2 ** -7 * max(1.0, ref.abs().max().item())
