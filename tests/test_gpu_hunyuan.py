"""HunyuanVideo-I2V: native DiT sequencing (alg_b200/hunyuan.py over the C ABI) + ALG loop vs the oracle restatement.

Tolerance protocol as for Wan.  The oracle keeps diffusers' padded text tokens and boolean masks; the
engine drops the padded tokens, so these tests also pin that shortcut on the GPU.  The latents live in an fp32 container
with bf16-rounded values on frames 1.. (quirk q12): per-step bar = one bf16 rounding of the state, rel-L2 <= 2^-8."""
import pytest
import torch

from conftest import rel_l2

pytestmark = pytest.mark.gpu

TINY = dict(num_attention_heads=2, attention_head_dim=128, num_layers=2, num_single_layers=2, num_refiner_layers=1,
            text_embed_dim=64, pooled_projection_dim=32)
ALG = dict(use_low_pass_guidance=True, lp_filter_type="down_up", lp_filter_in_latent=True, lp_blur_sigma=15.0,
           lp_blur_kernel_size=0.02734375, lp_resize_factor=0.625, lp_strength_schedule_type="interval",
           schedule_blur_kernel_size=False, schedule_interval_start_time=0.0, schedule_interval_end_time=0.3,
           schedule_linear_start_weight=1.0, schedule_linear_end_weight=0.0, schedule_linear_end_time=0.5,
           schedule_exp_decay_rate=10.0)


def _model(seed=0, **over):
    from alg_b200 import hunyuan
    cfg = dict(TINY, **over)
    return cfg, hunyuan.HunyuanVideoTransformer3DModel.from_synthetic(seed=seed, device="cuda", **cfg)


def _ocfg(cfg):
    from oracle import hunyuan_oracle as Ho
    keys = Ho.HunyuanConfig.__dataclass_fields__
    return Ho.HunyuanConfig(**{k: v for k, v in cfg.items() if k in keys})


def _inputs(B, T, H, W, L, valid, seed=0, text_dim=64, pooled_dim=32):
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = torch.randn(B, 16, T, H, W, generator=g, device="cuda")
    text = torch.randn(B, L, text_dim, generator=g, device="cuda").bfloat16()
    mask = torch.zeros(B, L, device="cuda")
    for b, v in enumerate(valid):
        mask[b, :v] = 1
    pooled = torch.randn(B, pooled_dim, generator=g, device="cuda").bfloat16()
    t = torch.tensor([873.0] * B, device="cuda").bfloat16()
    gd = torch.tensor([6.0] * B, device="cuda").bfloat16() * 1000.0
    return x, text, mask, pooled, t, gd


@pytest.mark.parametrize("over", [{}, {"num_attention_heads": 3, "rope_axes_dim": (16, 56, 56), "num_layers": 1, "num_single_layers": 3,
                                       "num_refiner_layers": 2}])
def test_forward_matches_oracle(over):
    from oracle import hunyuan_oracle as Ho
    cfg, model = _model(1, **over)
    ocfg = _ocfg(cfg)
    x, text, mask, pooled, t, gd = _inputs(2, 3, 8, 16, 24, (9, 24))
    out = model(x, t, text, mask, pooled, gd, return_dict=False)[0]
    sd = model.state_dict()
    ref16 = Ho.forward(sd, ocfg, x.bfloat16(), t, text, mask, pooled, gd)
    ref32 = Ho.forward({k: v.float() for k, v in sd.items()}, ocfg, x.bfloat16().float(), t.float(), text.float(), mask,
                       pooled.float(), gd.float())
    e_eng, e_torch = rel_l2(out, ref32), rel_l2(ref16, ref32)
    assert e_eng < max(1.5 * e_torch, 3e-3), (e_eng, e_torch)
    assert rel_l2(out, ref16) < 2e-2


def test_forward_full_width_one_block_each():
    """True HunyuanVideo width (24 heads x 128, mlp 12288, text 4096, pooled 768), 1 dual + 1 single + 1 refiner block."""
    from alg_b200 import hunyuan
    from oracle import hunyuan_oracle as Ho
    over = dict(num_layers=1, num_single_layers=1, num_refiner_layers=1)
    model = hunyuan.HunyuanVideoTransformer3DModel.from_synthetic(seed=4, device="cuda", **over)
    ocfg = Ho.HunyuanConfig(**over)
    x, text, mask, pooled, t, gd = _inputs(1, 2, 32, 48, 64, (37,), text_dim=4096, pooled_dim=768)
    out = model(x, t, text, mask, pooled, gd, return_dict=False)[0]
    sd = model.state_dict()
    ref16 = Ho.forward(sd, ocfg, x.bfloat16(), t, text, mask, pooled, gd)
    ref32 = Ho.forward({k: v.float() for k, v in sd.items()}, ocfg, x.bfloat16().float(), t.float(), text.float(), mask,
                       pooled.float(), gd.float())
    e_eng, e_torch = rel_l2(out, ref32), rel_l2(ref16, ref32)
    assert e_eng < max(1.5 * e_torch, 3e-3), (e_eng, e_torch)


def _pipe(model):
    from pipeline_hunyuan_video_image2video_lowpass import HunyuanVideoImageToVideoPipeline
    pipe = HunyuanVideoImageToVideoPipeline.from_pretrained("synthetic", transformer=model, synthetic=True).to("cuda")
    pipe.set_progress_bar_config(disable=True)
    return pipe


@pytest.mark.parametrize("true_cfg", [1.0, 4.0])
def test_loop_teacher_forced_per_step_latents(true_cfg):
    """hy:1126-1270 against oracle/hunyuan_oracle.denoise_loop: single-pass ALG branch (the shipped yaml) and the
    true-CFG 3-pass / 2-pass branch, both sides consuming the oracle's x_i."""
    from alg_b200.schedulers import FlowMatchEulerDiscreteScheduler
    from oracle import hunyuan_oracle as Ho, lp_oracle, prepare_lp_oracle as P, sched_oracle
    cfg, model = _model(2)
    ocfg = _ocfg(cfg)
    pipe = _pipe(model)
    pipe.scheduler = FlowMatchEulerDiscreteScheduler.from_config(pipe.scheduler.config, flow_shift=7.0, invert_sigmas=False)
    steps, T, H, W, L = 10, 3, 8, 16, 20
    x, text, mask, pooled, _, _ = _inputs(2, T, H, W, L, (11, 6), seed=5)
    lat0 = x[:1].contiguous()
    image_latents = x[1:2, :, :1].contiguous()
    pos = (text[:1], pooled[:1], mask[:1])
    neg = (text[1:2], pooled[1:2], mask[1:2]) if true_cfg > 1 else None
    sd = model.state_dict()

    def transformer(xx, timestep, emb, m, pl, guidance):
        return Ho.forward(sd, ocfg, xx, timestep, emb, m, pl, guidance)

    def lp_filter(img, kind, sigma, k, f):
        return P.hunyuan_prepare_lp(2, kind, sigma, k, f, True, True, img)  # pinned to hy:650-792 by the loop fixtures

    per_step = []
    sched = sched_oracle.FlowEulerOracle(shift=7.0)
    Ho.denoise_loop(transformer, sched, lat0, image_latents, pos, neg, steps, 6.0, true_cfg, ALG, lp_filter,
                    lp_oracle.get_lp_strength, on_step=lambda i, t, lat, npred: per_step.append((lat, npred)))
    xs = [lat0] + [p[0] for p in per_step]
    import numpy as np
    pipe.scheduler.set_timesteps(sigmas=np.linspace(1.0, 0.0, steps + 1)[:-1], device="cuda")
    guidance = float((torch.tensor([6.0], dtype=torch.bfloat16) * 1000.0)[0])
    n_pass_seen = []
    ts = pipe.scheduler.timesteps.float().cpu()
    for i in range(steps):
        x_next, npred = pipe.denoise_step(i, ts[i], xs[i], image_latents, pos, neg, guidance, 9, steps, ALG, true_cfg)
        n_pass_seen.append(npred.shape[0])
        assert npred.shape == per_step[i][1].shape
        assert torch.equal(x_next[:, :, :1], image_latents)  # hy:1270: the image frame is re-prepended untouched
        assert rel_l2(x_next, xs[i + 1]) < 2 ** -8, (i, rel_l2(x_next, xs[i + 1]))
    assert n_pass_seen == ([3] * 3 + [2] * 7 if true_cfg > 1 else [1] * 10)


def test_pipeline_call_surface():
    import inspect
    from PIL import Image
    cfg, model = _model(3)
    pipe = _pipe(model)
    names = list(inspect.signature(pipe.__call__).parameters)
    assert names[:5] == ["image", "prompt", "prompt_2", "negative_prompt", "negative_prompt_2"]
    assert names[-3:] == ["lp_on_noisy_latent", "enable_lp_img_embeds", "i2v_stable"] and len(names) == 46
    img = Image.new("RGB", (140, 70), (10, 200, 90))
    kw = dict(image=img, prompt="a cat walks", height=64, width=128, num_frames=9, num_inference_steps=3, guidance_scale=6.0,
              max_sequence_length=16)
    seen = []
    out = pipe(**kw, output_type="latent", generator=torch.Generator("cuda").manual_seed(42),
               callback_on_step_end=lambda p, i, t, k: (seen.append(i), {})[1], **ALG)
    assert out.frames.shape == (1, 16, 3, 8, 16) and out.frames.dtype == torch.float32 and seen == [0, 1, 2]
    assert torch.isfinite(out.frames).all()
    frames = pipe(**kw, output_type="np", true_cfg_scale=3.0, **ALG).frames
    assert frames.shape == (1, 9, 64, 128, 3)
    with pytest.raises(ValueError, match="divisible by 16"):
        pipe(**dict(kw, height=72), **ALG)
    with pytest.raises(AssertionError, match="image embeds is not supported"):
        pipe(**kw, enable_lp_img_embeds=True, **ALG)
    with pytest.raises(NotImplementedError, match="latent space only"):
        pipe(**kw, **dict(ALG, lp_filter_in_latent=False))
