"""Native ``AutoencoderKLHunyuanVideo`` (alg_b200/vae_hunyuan.py, float32 split-GEMM path) against ``oracle/hunyuan_vae_oracle.py``
(diffusers' network restated with F.conv3d on replicate-padded clips, F.group_norm, masked SDPA; parity unpinned, diffusers absent).
Oracle in float64 on the GPU; tolerance 1e-4 relative L2."""
import pytest
import torch

from conftest import rel_l2

pytestmark = pytest.mark.gpu

TINY = dict(block_out_channels=(32, 64, 64, 64), latent_channels=4)


def _pair(cfg=TINY, seed=1):
    from alg_b200.vae_hunyuan import AutoencoderKLHunyuanVideo
    from oracle import hunyuan_vae_oracle as V
    ocfg = dict(V.HUNYUAN_VAE, **cfg)
    sd = V.make_weights(ocfg, seed=seed, device="cuda")
    return AutoencoderKLHunyuanVideo(**cfg).load_state_dict({k: v.clone() for k, v in sd.items()}), sd, ocfg


def test_group_norm_and_causal_softmax_kernels():
    import torch.nn.functional as F
    from alg_b200 import _lib
    g = torch.Generator(device="cuda").manual_seed(0)
    for rows, C, groups in ((1000, 64, 32), (77, 96, 32), (5000, 512, 32)):
        x = torch.randn(rows, C, generator=g, device="cuda") * 3 + 1
        w, b = torch.randn(C, generator=g, device="cuda"), torch.randn(C, generator=g, device="cuda")
        stats = torch.empty(2 * groups, device="cuda", dtype=torch.float64)
        for silu in (0, 1):
            y = torch.empty_like(x)
            _lib.check(_lib.lib().alg_group_norm_f32(x.data_ptr(), y.data_ptr(), rows, C, groups, 1e-6, w.data_ptr(), b.data_ptr(), silu,
                                                     stats.data_ptr(), _lib.stream_ptr(x.device)))
            ref = F.group_norm(x.t()[None].double(), groups, w.double(), b.double(), 1e-6)[0].t()
            ref = F.silu(ref) if silu else ref
            assert rel_l2(y, ref) < 2e-6, (rows, C, silu, rel_l2(y, ref))
    N, block = 96, 24
    s = torch.randn(N, N, generator=g, device="cuda") * 4
    ref = s.double() * 0.3
    frame = torch.arange(N, device="cuda") // block
    ref = ref.masked_fill(frame[None, :] > frame[:, None], float("-inf")).softmax(-1)
    _lib.check(_lib.lib().alg_softmax_rows_f32(s.data_ptr(), N, N, N, 0.3, block, _lib.stream_ptr(s.device)))
    assert rel_l2(s, ref) < 2e-6 and bool((s[0, block:] == 0).all())


def test_im2col_replicate_and_temporal_duplication():
    """HunyuanVideoCausalConv3d padding (replicate in T, H, W) and HunyuanVideoUpsampleCausal3D's resize fused into the gather."""
    import ctypes as C
    import torch.nn.functional as F
    from alg_b200 import _lib
    g = torch.Generator(device="cuda").manual_seed(1)

    def run(T, H, W, Cc, stride, up, tdup):
        x = torch.randn(T, H, W, Cc, generator=g, device="cuda")
        clip = x.permute(3, 0, 1, 2)[None]  # [1, C, T, H, W]
        if up > 1 or tdup > 1:
            first = F.interpolate(clip[:, :, 0], scale_factor=(float(up), float(up)), mode="nearest").unsqueeze(2)
            if T > 1:
                other = F.interpolate(clip[:, :, 1:].contiguous(), scale_factor=(float(tdup), float(up), float(up)), mode="nearest")
                clip = torch.cat((first, other), 2)
            else:
                clip = first
        padded = F.pad(clip, (1, 1, 1, 1, 2, 0), mode="replicate")
        TL, HL, WL = clip.shape[2:]
        To, Ho, Wo = (TL - 1) // stride[0] + 1, (HL - 1) // stride[1] + 1, (WL - 1) // stride[2] + 1
        want = padded.unfold(2, 3, stride[0]).unfold(3, 3, stride[1]).unfold(4, 3, stride[2])  # [1, C, To, Ho, Wo, 3, 3, 3]
        want = want[0].permute(1, 2, 3, 4, 5, 6, 0).reshape(To * Ho * Wo, 27 * Cc)
        K = 27 * Cc
        ld = (K + 7) // 8 * 8
        cols = torch.empty(To * Ho * Wo, 3 * ld, device="cuda", dtype=torch.bfloat16)
        p = _lib.Im2colF32()
        p.x, p.cols, p.T, p.H, p.W, p.C = x.data_ptr(), cols.data_ptr(), T, H, W, Cc
        p.kt, p.kh, p.kw, p.st, p.sh, p.sw = 3, 3, 3, *stride
        p.pad_t, p.pad_top, p.pad_left, p.To, p.Ho, p.Wo, p.to0, p.up, p.t_min, p.ld = 2, 1, 1, To, Ho, Wo, 0, up, 0, ld
        p.replicate, p.tdup = 1, tdup
        _lib.check(_lib.lib().alg_im2col_split3_f32(C.byref(p), _lib.stream_ptr(x.device)))
        hi, lo = cols[:, :K].float(), cols[:, 2 * ld:2 * ld + K].float()
        assert torch.equal(hi, want.bfloat16().float()), (T, H, W, Cc, stride, up, tdup)
        assert (hi + lo - want).abs().max() <= want.abs().max() * 2 ** -15

    run(4, 6, 10, 8, (1, 1, 1), 1, 1)
    run(5, 6, 10, 8, (2, 2, 2), 1, 1)
    run(3, 7, 9, 3, (1, 2, 2), 1, 1)   # RGB scalar path, odd sizes
    run(3, 4, 6, 8, (1, 1, 1), 2, 2)   # spatial + temporal upsample fused
    run(3, 4, 6, 8, (1, 1, 1), 2, 1)   # spatial only
    run(1, 4, 6, 8, (1, 1, 1), 2, 1)   # a single frame


@pytest.mark.parametrize("T,H,W", [(1, 32, 48), (5, 32, 32), (3, 48, 32)])
def test_encode_matches_oracle(T, H, W):
    from oracle import hunyuan_vae_oracle as V
    vae, sd, ocfg = _pair()
    x = torch.rand(1, 3, T, H, W, generator=torch.Generator(device="cuda").manual_seed(4), device="cuda") * 2 - 1
    dist = vae.encode(x).latent_dist
    ref = V.encode_moments(x.double(), {k: v.double() for k, v in sd.items()}, ocfg, torch.float64)
    assert dist.parameters.shape == ref.shape and rel_l2(dist.parameters, ref) < 1e-4, rel_l2(dist.parameters, ref)


@pytest.mark.parametrize("T,h,w", [(1, 4, 6), (3, 4, 4), (2, 6, 4)])
def test_decode_matches_oracle(T, h, w):
    from oracle import hunyuan_vae_oracle as V
    vae, sd, ocfg = _pair(seed=2)
    z = torch.randn(1, 4, T, h, w, generator=torch.Generator(device="cuda").manual_seed(5), device="cuda")
    out = vae.decode(z).sample
    ref = V.decode(z.double(), {k: v.double() for k, v in sd.items()}, ocfg, torch.float64)
    assert out.shape == ref.shape == (1, 3, 4 * (T - 1) + 1, 8 * h, 8 * w) and rel_l2(out, ref) < 1e-4, rel_l2(out, ref)


def test_pipeline_image_latents_and_attention_budget():
    """hy:550-592: ``prepare_latents`` encodes the single conditioning frame (argmax) and scales it; with the native VAE the
    result equals the oracle's.  A clip whose N x N attention scores cannot exist is refused with the reason."""
    from types import SimpleNamespace
    from alg_b200.schedulers import FlowMatchEulerDiscreteScheduler
    from oracle import hunyuan_vae_oracle as V
    from pipeline_hunyuan_video_image2video_lowpass import HunyuanVideoImageToVideoPipeline
    cfg = dict(TINY, latent_channels=16)
    vae, sd, ocfg = _pair(cfg, seed=3)

    class OracleVAE:
        dtype = torch.float32
        config = SimpleNamespace(**ocfg)
        temporal_compression_ratio, spatial_compression_ratio = 4, 8

        def encode(self, x):
            m = V.encode_moments(x.double(), {k: v.double() for k, v in sd.items()}, ocfg, torch.float64).float()
            return SimpleNamespace(latent_dist=SimpleNamespace(mode=lambda: m[:, :16], sample=lambda generator=None: m[:, :16]))

    dummy = SimpleNamespace(config=SimpleNamespace(in_channels=16), dtype=torch.bfloat16, to=lambda *a, **k: None)
    res = []
    for v in (vae, OracleVAE()):
        pipe = HunyuanVideoImageToVideoPipeline(text_encoder=None, tokenizer=None, transformer=dummy, vae=v,
                                                scheduler=FlowMatchEulerDiscreteScheduler(shift=7.0), text_encoder_2=None,
                                                tokenizer_2=None, image_processor=None)
        image = (torch.rand(1, 3, 32, 48, generator=torch.Generator().manual_seed(8)) * 2 - 1).cuda()
        lat, img_lat = pipe.prepare_latents(image, 1, 16, 32, 48, 9, torch.float32, torch.device("cuda"),
                                            torch.Generator(device="cuda").manual_seed(9), None, "token_replace", False)
        res.append((lat, img_lat))
    assert torch.equal(res[0][0], res[1][0]) and res[0][1].shape == res[1][1].shape == (1, 16, 1, 4, 6)
    assert rel_l2(res[0][1], res[1][1]) < 1e-4
    vae._attn_budget = 1 << 20
    with pytest.raises(NotImplementedError, match="score matrix"):
        vae.decode(torch.randn(1, 16, 3, 16, 16, device="cuda"))


@pytest.mark.parametrize("T", [5, 9, 8])
def test_temporal_tiled_decode_matches_oracle(T):
    """hy:1292 on clips longer than one temporal tile: diffusers' default ``use_framewise_decoding`` decodes overlapping tiles of 5
    latent frames (stride 3) and cross-fades 4 frames (``_temporal_tiled_decode`` / ``blend_t``); T = 5 is the first tiled length,
    T = 8 ends on a short last tile."""
    from oracle import hunyuan_vae_oracle as V
    vae, sd, ocfg = _pair(seed=6)
    z = torch.randn(1, 4, T, 4, 4, generator=torch.Generator(device="cuda").manual_seed(T), device="cuda")
    out = vae.decode(z).sample
    ref = V.decode(z.double(), {k: v.double() for k, v in sd.items()}, ocfg, torch.float64)
    assert out.shape == ref.shape == (1, 3, 4 * (T - 1) + 1, 32, 32) and rel_l2(out, ref) < 1e-4, rel_l2(out, ref)
    whole = V.decode(z.double(), {k: v.double() for k, v in sd.items()}, ocfg, torch.float64, framewise=False)
    assert rel_l2(ref, whole) > 1e-3  # tiling is not a no-op: the tiles lose their causal context, which is why it must be reproduced
    with pytest.raises(NotImplementedError, match="spatial"):
        vae.enable_tiling()
