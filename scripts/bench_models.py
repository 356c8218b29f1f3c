"""Full-size step timing of the other two BASELINE.json configs through the pipelines' denoise_step (device-timed):

    python scripts/bench_models.py cog      # configs[2]: CogVideoX-5b-I2V 720x480, 49 frames, gaussian_blur in pixel space
    python scripts/bench_models.py hunyuan  # configs[3]: HunyuanVideo-I2V 720p, 129 frames, interval schedule (single pass)

Prints one JSON line per model: ms per 3-pass / 2-pass (Cog) or 1-pass (Hy) step, frames/s of the full schedule derived
from them, and the model FLOP rate.  Synthetic weights at the true architecture (no checkpoints offline)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch


def timed(fn, reps):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def cog(reps):
    from pipeline_cogvideox_image2video_lowpass import CogVideoXImageToVideoPipeline
    pipe = CogVideoXImageToVideoPipeline.from_pretrained("synthetic", synthetic=True).to("cuda")
    steps, frames, H, W = 50, 49, 480, 720
    alg = dict(use_low_pass_guidance=True, lp_filter_type="gaussian_blur", lp_filter_in_latent=False, lp_blur_sigma=15.0,
               lp_blur_kernel_size=0.02734375, lp_resize_factor=0.25, lp_strength_schedule_type="interval",
               schedule_blur_kernel_size=False, schedule_interval_start_time=0.0, schedule_interval_end_time=0.04,
               schedule_linear_start_weight=1.0, schedule_linear_end_weight=0.0, schedule_linear_end_time=0.5,
               schedule_exp_decay_rate=10.0)
    g = torch.Generator(device="cuda").manual_seed(42)
    lat = torch.randn(1, 13, 16, 60, 90, generator=g, device="cuda").bfloat16()
    img_lat = torch.cat([torch.randn(1, 1, 16, 60, 90, generator=g, device="cuda"), torch.zeros(1, 12, 16, 60, 90, device="cuda")], 1).bfloat16()
    rgb = (torch.rand(1, 3, H, W, generator=g, device="cuda") * 2 - 1).bfloat16()
    pos, neg = (torch.randn(1, 226, 4096, generator=g, device="cuda").bfloat16() for _ in range(2))
    rope = pipe._prepare_rotary_positional_embeddings(H, W, 13, "cuda")
    pipe.scheduler.set_timesteps(steps, device="cuda")
    ts = pipe.scheduler.timesteps.tolist()
    with torch.no_grad():
        ms3 = timed(lambda: pipe.denoise_step(0, ts[0], lat, img_lat, rgb, pos, neg, rope, g, frames, steps, alg, 6.0), reps)
        ms2 = timed(lambda: pipe.denoise_step(10, ts[10], lat, img_lat, rgb, pos, neg, rope, g, frames, steps, alg, 6.0), reps)
    sec_video = (2 * ms3 + 48 * ms2) / 1e3
    fwd = 3.32e14  # FLOP per sample-forward at config 3
    print(json.dumps({"model": "CogVideoX-5b-I2V 720x480 49f 50 steps, gaussian_blur pixel-space ALG (BASELINE configs[2])",
                      "ms_step_3pass": ms3, "ms_step_2pass": ms2, "frames_per_s": frames / sec_video,
                      "model_tflops_2pass": 2 * fwd / (ms2 / 1e3) / 1e12, "tokens": 17776}), flush=True)


def hunyuan(reps):
    from pipeline_hunyuan_video_image2video_lowpass import HunyuanVideoImageToVideoPipeline
    pipe = HunyuanVideoImageToVideoPipeline.from_pretrained("synthetic", synthetic=True).to("cuda")
    steps, frames = 30, 129
    T, H, W = 33, 90, 160
    alg = dict(use_low_pass_guidance=True, lp_filter_type="down_up", lp_filter_in_latent=True, lp_blur_sigma=15.0,
               lp_blur_kernel_size=0.02734375, lp_resize_factor=0.625, lp_strength_schedule_type="interval",
               schedule_blur_kernel_size=False, schedule_interval_start_time=0.0, schedule_interval_end_time=0.04,
               schedule_linear_start_weight=1.0, schedule_linear_end_weight=0.0, schedule_linear_end_time=0.5,
               schedule_exp_decay_rate=10.0)
    g = torch.Generator(device="cuda").manual_seed(42)
    lat = torch.randn(1, 16, T, H, W, generator=g, device="cuda")
    img_lat = torch.randn(1, 16, 1, H, W, generator=g, device="cuda")
    pos = (torch.randn(1, 400, 4096, generator=g, device="cuda").bfloat16(), torch.randn(1, 768, generator=g, device="cuda").bfloat16(), 180)
    pipe.scheduler.set_timesteps(sigmas=np.linspace(1.0, 0.0, steps + 1)[:-1], device="cuda")
    ts = pipe.scheduler.timesteps.float().cpu()
    with torch.no_grad():
        ms1 = timed(lambda: (setattr(pipe.scheduler, "_step_index", 0),
                             pipe.denoise_step(0, ts[0], lat, img_lat, pos, None, 6016.0, frames, steps, alg, 1.0))[1], reps)
    fwd = 1.21e16  # FLOP per sample-forward at config 4
    print(json.dumps({"model": "HunyuanVideo-I2V 720x1280 129f 30 steps, down_up ALG, embedded guidance (BASELINE configs[3])",
                      "ms_step_1pass": ms1, "frames_per_s": frames / (steps * ms1 / 1e3),
                      "model_tflops": fwd / (ms1 / 1e3) / 1e12, "tokens": T * (H // 2) * (W // 2) + 180}), flush=True)


def vae(reps):
    """The per-step VAE call of pixel-space ALG (cog:645): AutoencoderKLCogVideoX.encode of one 480x720 frame."""
    from alg_b200 import _lib, vae_cogvideox as V
    m = V.AutoencoderKLCogVideoX.from_synthetic(seed=0)
    x = torch.randn(1, 3, 1, 480, 720, device="cuda").clamp(-1, 1).bfloat16()
    n0 = _lib.lib().alg_launch_count()
    m.encode(x)
    launches = _lib.lib().alg_launch_count() - n0
    ms = timed(lambda: m.encode(x), reps)
    flop = 0
    H, W = 480, 720
    for name, shape in V.encoder_parameter_shapes(m._cfg).items():
        if name.endswith(".weight") and len(shape) >= 4:
            lvl = int(name.split("down_blocks.")[1][0]) if "down_blocks" in name else 3
            px = (H >> lvl) * (W >> lvl)
            if "downsamplers" in name:
                px //= 4
            k = 1
            for d in shape[1:]:
                k *= d
            flop += 2 * px * shape[0] * k
    print(json.dumps({"model": "AutoencoderKLCogVideoX.encode, one 480x720 frame (cog:645, every step of pixel-space ALG)",
                      "ms_per_encode": ms, "launches": int(launches), "conv_tflop": flop / 1e12,
                      "conv_tflops_rate": flop / ms / 1e9}), flush=True)


def vae_decode(reps):
    """AutoencoderKLCogVideoX.decode of the full clip (cog:428-433): 13 latent frames of 60 x 90 -> 49 frames of 480 x 720."""
    from alg_b200 import _lib, vae_cogvideox as V
    m = V.AutoencoderKLCogVideoX.from_synthetic(seed=0, with_decoder=True)
    z = torch.randn(1, 16, 13, 60, 90, device="cuda").bfloat16()
    n0 = _lib.lib().alg_launch_count()
    out = m.decode(z).sample
    launches = _lib.lib().alg_launch_count() - n0
    torch.cuda.synchronize()
    ms = timed(lambda: m.decode(z), reps)
    flop = 0
    for name, shape in V.decoder_parameter_shapes(m._cfg).items():
        if name.endswith(".weight") and len(shape) >= 4 and "conv_y" not in name and "conv_b" not in name:
            lvl = int(name.split("up_blocks.")[1][0]) if "up_blocks" in name else (0 if "mid_block" in name or "conv_in" in name else 3)
            if "upsamplers" in name:
                lvl += 1
            frames = {0: 13, 1: 25, 2: 49, 3: 49}[min(lvl, 3)]
            px = frames * (60 << min(lvl, 3)) * (90 << min(lvl, 3))
            k = 1
            for d in shape[1:]:
                k *= d
            flop += 2 * px * shape[0] * k
    print(json.dumps({"model": "AutoencoderKLCogVideoX.decode, 13 x 60 x 90 latents -> 49 frames of 480 x 720 (cog:428-433, once per video)",
                      "ms_per_decode": ms, "launches": int(launches), "conv_tflop": flop / 1e12, "conv_tflops_rate": flop / ms / 1e9,
                      "finite": bool(torch.isfinite(out).all()), "shape": list(out.shape),
                      "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30}), flush=True)


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "cog"
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    {"cog": cog, "hunyuan": hunyuan, "vae": vae, "vae_decode": vae_decode}[which](reps)
