"""oracle/wan_vae_oracle.py on CPU in float64: diffusers' CHUNKED feature-cache evaluation of AutoencoderKLWan (encoder 1 + 4 + 4 ...
frames, decoder one latent frame at a time, the "Rep" marker of upsample3d) equals the closed whole-clip form the product
implements (alg_b200/vae_wan.py): zero-front-padded causal convolutions, "downsample3d" / "upsample3d" pass frame 0 through, and
upsample3d's temporal windows read frame 0 as zero.  Also the HunyuanVideo VAE oracle's shape contract."""
import pytest
import torch


def _cfg():
    from oracle import wan_vae_oracle as V
    return dict(V.WAN21_VAE, base_dim=8, z_dim=4)


@pytest.mark.parametrize("T", [1, 5, 9])
def test_wan_encode_chunked_equals_closed_form(T):
    from oracle import wan_vae_oracle as V
    cfg = _cfg()
    sd = V.make_weights(cfg, seed=1, dtype=torch.float64)
    x = torch.randn(1, 3, T, 16, 24, generator=torch.Generator().manual_seed(T), dtype=torch.float64)
    a, b = V.encode_moments(x, sd, cfg, torch.float64), V.encode_closed_form(x, sd, cfg, torch.float64)
    assert a.shape == (1, 8, 1 + (T - 1) // 4, 2, 3) and float((a - b).abs().max()) < 1e-12 * max(1.0, float(a.abs().max()))


@pytest.mark.parametrize("T", [1, 2, 4])
def test_wan_decode_chunked_equals_closed_form(T):
    from oracle import wan_vae_oracle as V
    cfg = _cfg()
    sd = V.make_weights(cfg, seed=2, dtype=torch.float64)
    z = torch.randn(1, 4, T, 2, 3, generator=torch.Generator().manual_seed(10 + T), dtype=torch.float64)
    a, b = V.decode(z, sd, cfg, torch.float64), V.decode_closed_form(z, sd, cfg, torch.float64)
    assert a.shape == (1, 3, 4 * T - 3, 16, 24) and float((a - b).abs().max()) < 1e-12
    assert float(a.abs().max()) <= 1.0


def test_upsample3d_never_sees_frame_zero():
    """The quirk the closed form encodes: perturbing latent frame 0 changes decoded frame 0 and, through the 3x3x3 convolutions,
    later frames -- but NOT through upsample3d's temporal convolution; with every other path cut (1-frame clip vs its first frame
    inside a longer clip) the first decoded frame is identical."""
    from oracle import wan_vae_oracle as V
    cfg = _cfg()
    sd = V.make_weights(cfg, seed=3, dtype=torch.float64)
    z = torch.randn(1, 4, 3, 2, 3, generator=torch.Generator().manual_seed(5), dtype=torch.float64)
    full = V.decode(z, sd, cfg, torch.float64)
    first = V.decode(z[:, :, :1], sd, cfg, torch.float64)
    assert float((full[:, :, :1] - first).abs().max()) < 1e-12  # causal: frame 0 does not depend on later latent frames


def test_hunyuan_vae_oracle_shapes():
    from oracle import hunyuan_vae_oracle as V
    cfg = dict(V.HUNYUAN_VAE, block_out_channels=(32, 32, 32, 32), latent_channels=4)
    sd = V.make_weights(cfg, seed=1)
    x = torch.randn(1, 3, 5, 16, 16, generator=torch.Generator().manual_seed(0))
    m = V.encode_moments(x, sd, cfg)
    assert m.shape == (1, 8, 2, 2, 2)
    d = V.decode(m[:, :4], sd, cfg)
    assert d.shape == (1, 3, 5, 16, 16) and bool(torch.isfinite(d).all())
