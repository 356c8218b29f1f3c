"""bench.py -- video frames/sec of the ALG denoise loop (BASELINE.json metric), one harness for every config.

    python bench.py --gpus N --steps K --warmup W                      # wan480 = BASELINE.json configs[1], the headline
    python bench.py --config {wan480,wan720,cog,hunyuan} ...           # configs[1] / its 720p variant / configs[2] / configs[3]
    python bench.py --impl reference --gpus N --steps K ...            # the reference's CPU path (oracle port) on host cores
    python bench.py --full-video                                       # pipe.__call__ for ALL steps, free-running (wall clock)

A "step" is one iteration of the denoise loop (wan:844-927 / cog:1005-1123 / hy:1126-1270): low-pass filter of the
conditioning image, the 1/2/3-pass DiT forward, the CFG combine and the scheduler update, for ONE video sample per GPU
(independent samples shard across GPUs: weak scaling, no data-path collective; NCCL only broadcasts the weights at init).
K timed steps sample the real schedule uniformly (index floor(k * steps / K)); frames/sec comes from the schedule-weighted
step time (e.g. Wan: 10 three-pass + 40 two-pass steps), so the number does not depend on K.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

FAST = bool(os.environ.get("ALG_BENCH_FAST"))  # profiling runs (ncu): device-timed region only, no e2e / CPU / parity legs
GUIDANCE = 5.0
FLOW_SHIFT = 5.0  # run.py:63 compares int 480 to '480' => always 5.0 (quirk q1): measure the as-shipped behaviour
# kept as module globals for scripts/parity_fullsize.py (Wan geometry of the selected config)
NUM_FRAMES, STEPS_PER_VIDEO, HEIGHT, WIDTH = 81, 50, 480, 832
T_LAT, H_LAT, W_LAT = 21, 60, 104


def alg_kwargs(**over):
    base = dict(use_low_pass_guidance=True, lp_filter_type="down_up", lp_filter_in_latent=True, lp_blur_sigma=15.0,
                lp_blur_kernel_size=0.02734375, lp_resize_factor=0.4, lp_strength_schedule_type="interval",
                schedule_blur_kernel_size=False, schedule_interval_start_time=0.0, schedule_interval_end_time=0.20,
                schedule_linear_start_weight=1.0, schedule_linear_end_weight=0.0, schedule_linear_end_time=0.5,
                schedule_exp_decay_rate=10.0)
    base.update(over)
    return base


ALG = alg_kwargs()  # Wan (configs/wan_alg.yaml)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops", 1590.0), d.get("bf16_tflops_sustained", 1400.0), "measured"
    return 6650.0, 1590.0, 1400.0, "fallback"


_RESULT_FD = None


def emit(line):
    """The one JSON line, on the process's ORIGINAL stdout (see main)."""
    data = (json.dumps(line) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_RESULT_FD, data)


def strength(idx, total, alg):
    import lp_utils
    return lp_utils.get_lp_strength(idx, total, alg["lp_strength_schedule_type"], alg["schedule_interval_start_time"],
                                    alg["schedule_interval_end_time"], alg["schedule_linear_start_weight"],
                                    alg["schedule_linear_end_weight"], alg["schedule_linear_end_time"], alg["schedule_exp_decay_rate"])


def n_pass_of(idx):  # Wan (used by scripts/parity_fullsize.py too)
    return 3 if strength(idx, STEPS_PER_VIDEO, ALG) != 0.0 else 2


def forward_flops(n_tok, d=5120, ffn=13824, layers=40, ctx=769, text_dim=4096):
    """Algorithmic FLOPs of one Wan sample-forward: (whole forward, self-attention core)."""
    self_attn = 4 * n_tok * n_tok * d
    per_layer = 8 * n_tok * d * d + self_attn + 4 * n_tok * d * d + 4 * ctx * d * d + 4 * n_tok * ctx * d + 4 * n_tok * d * ffn
    return layers * per_layer, layers * self_attn


def joint_dit_flops(n_tok, d, layers):
    """CogVideoX / HunyuanVideo blocks: 24 N d^2 of linears + 4 N^2 d of joint attention per block."""
    attn = 4 * n_tok * n_tok * d
    return layers * (24 * n_tok * d * d + attn), layers * attn


class ClockSampler(threading.Thread):
    """nvidia-smi-equivalent (NVML) SM clock + throttle reasons during the timed region."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None

    def run(self):
        try:
            import pynvml as N
            N.nvmlInit()
            h = N.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = N.nvmlDeviceGetMaxClockInfo(h, N.NVML_CLOCK_SM)
            names = {N.nvmlClocksEventReasonHwSlowdown: "hw_slowdown", N.nvmlClocksEventReasonHwThermalSlowdown: "hw_thermal_slowdown",
                     N.nvmlClocksEventReasonSwThermalSlowdown: "sw_thermal_slowdown", N.nvmlClocksEventReasonSwPowerCap: "sw_power_cap",
                     N.nvmlClocksEventReasonHwPowerBrakeSlowdown: "hw_power_brake"}
            while not self.stop_flag:
                self.samples.append(N.nvmlDeviceGetClockInfo(h, N.NVML_CLOCK_SM))
                r = N.nvmlDeviceGetCurrentClocksEventReasons(h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
                time.sleep(0.1)
        except Exception as ex:  # pragma: no cover
            self.reasons.add(f"nvml_unavailable:{type(ex).__name__}")

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ======================================================================================================
# workloads
# ======================================================================================================
class WanWorkload:
    key = "wan"
    frames, steps_per_video = 81, 50

    def __init__(self, resolution):
        global HEIGHT, WIDTH, H_LAT, W_LAT
        if resolution == "720p":
            HEIGHT, WIDTH, H_LAT, W_LAT = 720, 1280, 90, 160
        self.res, self.alg = resolution, ALG
        self.n_tok = T_LAT * (H_LAT // 2) * (W_LAT // 2)
        self.metric = f"video frames/sec (Wan-I2V-14B {HEIGHT}p, 81 frames, 50 steps, ALG down_up)"
        base = (f"Wan-I2V-14B {HEIGHT}x{WIDTH}, 81 frames, 50 steps, ALG down_up f=0.4 interval[0,0.2], gs 5, UniPC flow_shift 5.0")
        self.workload = base + (" (BASELINE.json configs[1]); one sample per GPU" if HEIGHT == 480 else
                                " (720p variant of BASELINE.json configs[1]); one sample per GPU")
        self.kernel = f"attention_kernel<128> (DiT self-attention, N={self.n_tok}, 40 heads x 128)"
        self.weights = "seeded random init at the true 16.4 B-parameter architecture (no checkpoints offline)"
        self.l2 = "inputs_exceed_l2 (32.8 GB of weights stream through the 126 MB L2 every step)"
        self.dtype = "bf16"

    def n_pass(self, idx):
        return 3 if strength(idx, self.steps_per_video, self.alg) != 0.0 else 2

    def forward_flops(self):
        return forward_flops(self.n_tok)

    def attn_flops(self, n_pass):  # one self-attention launch covers all passes of a step x 40 heads
        return 4 * n_pass * self.n_tok * self.n_tok * 5120

    def setup(self, device, rank, world):
        from alg_b200 import distributed as D, wan
        from alg_b200.schedulers import UniPCMultistepScheduler
        from pipeline_wan_image2video_lowpass import WanImageToVideoPipeline
        cfg = dict(wan.WAN_I2V_14B)
        shapes = {k: (s, torch.float32 if any(f in k for f in wan.FP32_KEYS) else torch.bfloat16) for k, s in wan.parameter_shapes(cfg).items()}
        sd, arena = D.arena_state_dict(shapes, device)  # one flat arena: ONE NCCL broadcast for the 16.4 B parameters
        if rank == 0:
            for k, v in wan.synthetic_state_dict(cfg, seed=0, device=device).items():
                sd[k].copy_(v)
        D.broadcast_arena(arena)  # NCCL over NVLink: rank 0 owns the seeded weights (init only, outside the timed region)
        self.transformer = wan.WanTransformer3DModel(**cfg).load_state_dict(sd)
        self.pipe = WanImageToVideoPipeline.from_pretrained("synthetic", transformer=self.transformer, synthetic=True)
        self.pipe.scheduler = UniPCMultistepScheduler.from_config(self.pipe.scheduler.config, flow_shift=FLOW_SHIFT)
        self.pipe.to(device)
        self.pipe.set_progress_bar_config(disable=True)
        self.pipe._guidance_scale = GUIDANCE
        self.transformer.context_cache(True)  # like pipe.__call__: the conditioning is fixed over the steps of a video
        self.sched = self.pipe.scheduler
        self.sched.set_timesteps(self.steps_per_video, device=device)
        self.ts = self.sched.timesteps.tolist()
        self.lat0, self.cond, self.pos, self.neg, self.img = synthetic_inputs(device, D.sample_seed(rank))
        self.image_rgb = torch.zeros(1, 3, HEIGHT, WIDTH, device=device)
        self.device = device
        self.host_in = self.host_out = None

    def _jump(self, idx):
        # jump to schedule position idx (the multistep history buffers hold the previous timed step's values: identical
        # work, the numbers only matter to the parity tests)
        s = self.sched
        s._step_index, s.lower_order_nums, s.this_order = idx, min(idx, 2), (min(idx, 2) or None)
        s._have_last = idx > 0 and s._state is not None

    def step(self, idx, inputs=None):
        lat, cond, pos, neg, img = inputs or (self.lat0, self.cond, self.pos, self.neg, self.img)
        self._jump(idx)
        out, _ = self.pipe.denoise_step(idx, self.ts[idx], lat, cond, self.image_rgb, pos, neg, img, None, self.frames,
                                        self.steps_per_video, self.alg)
        return out

    def host_inputs(self):
        return [self.lat0, self.cond, self.pos, self.neg, self.img]

    def profile_begin(self):
        self.transformer.profile(True)

    def profile_end(self):
        prof = self.transformer.profile_read()
        self.transformer.profile(False)
        return prof

    def full_video(self, generator):
        """pipe.__call__ for all 50 steps, free-running, from HOST inputs (embeddings / latents on the host, image tensor)."""
        h = [t.cpu() for t in (self.pos, self.neg, self.img, self.lat0)]
        image = torch.rand(1, 3, HEIGHT, WIDTH, generator=torch.Generator().manual_seed(1))

        def run():
            return self.pipe(image=image, prompt_embeds=h[0].to(self.device), negative_prompt_embeds=h[1].to(self.device),
                             latents=h[3].to(self.device), height=HEIGHT, width=WIDTH, num_frames=self.frames,
                             num_inference_steps=self.steps_per_video, guidance_scale=GUIDANCE, generator=generator,
                             output_type="latent", **self.alg).frames.cpu()
        return run, sum(t.numel() * t.element_size() for t in h) + image.numel() * 4

    def parity(self, log):
        sys.path.insert(0, os.path.join(ROOT, "scripts"))
        import parity_fullsize as PF
        res = PF.wan_case(layers=40, fp32_layers=0, eager_backends=("cudnn",), resolution=self.res, log=log)
        n3 = sum(1 for i in range(self.steps_per_video) if self.n_pass(i) == 3)
        ms = {r["n_pass"]: r.get("eager_cudnn_ms") for r in res["steps"]}
        eager = None
        if ms.get(2) and ms.get(3):
            sec = (n3 * ms[3] + (self.steps_per_video - n3) * ms[2]) / 1e3
            eager = {"value": self.frames / sec, "unit": "frames/s", "ms_step_3pass": ms[3], "ms_step_2pass": ms[2],
                     "what": "eager PyTorch bf16 on this GPU: oracle/wan_oracle.forward (cuBLAS linears, cuDNN SDPA) + torch CFG / "
                             "UniPC, same tensors, same 40-layer weights; schedule-weighted 10 x 3-pass + 40 x 2-pass"}
        return res, eager


def synthetic_inputs(device, seed):
    """Wan conditioning of one sample: latents ~ N(0,1); condition = [mask | N(0,1)]; zero-padded prompts."""
    g = torch.Generator(device=device).manual_seed(seed)
    lat = torch.randn(1, 16, T_LAT, H_LAT, W_LAT, generator=g, device=device)
    mask = torch.zeros(1, 4, T_LAT, H_LAT, W_LAT, device=device)
    mask[:, :, 0] = 1  # wan:436-447: ones on latent frame 0
    cond = torch.cat([mask, torch.randn(1, 16, T_LAT, H_LAT, W_LAT, generator=g, device=device)], dim=1)
    n_tok = 64 + seed % 200

    def text():
        e = torch.randn(1, 512, 4096, generator=g, device=device)
        e[:, n_tok:] = 0  # Wan zero-pads beyond the prompt length (wan:214-217)
        return e.bfloat16()

    return lat, cond, text(), text(), torch.randn(1, 257, 1280, generator=g, device=device).bfloat16()


class _SequencedWorkload:
    """CogVideoX / HunyuanVideo: Python sequencers over the C ABI; per-class timing from alg_b200.ops' event spans."""
    dtype = "bf16"

    def profile_begin(self):
        from alg_b200 import ops
        ops.profile_start()

    def profile_end(self):
        from alg_b200 import ops
        p = ops.profile_stop()
        out = {k: {"ms": p.get(k, {}).get("ms", 0.0), "launches": p.get(k, {}).get("launches", 0)}
               for k in ("self_attention", "short_attention", "gemm")}
        return out

    def _broadcast(self, model, device):
        from alg_b200 import distributed as D
        D.broadcast_state_dict(model.state_dict())  # identical seeded weights on every rank; the broadcast is the init-time collective


class CogWorkload(_SequencedWorkload):
    key = "cog"
    frames, steps_per_video = 49, 50
    H, W, F_LAT = 480, 720, 13

    def __init__(self):
        self.alg = alg_kwargs(lp_filter_type="gaussian_blur", lp_filter_in_latent=False, lp_resize_factor=0.25,
                              schedule_interval_end_time=0.04)  # configs/cogvideox_alg.yaml: pixel-space Gaussian, interval [0, 0.04]
        self.n_tok = self.F_LAT * 30 * 45 + 226
        self.metric = "video frames/sec (CogVideoX-5b-I2V 480x720, 49 frames, 50 steps, ALG gaussian_blur in pixel space)"
        self.workload = ("CogVideoX-5b-I2V 480x720, 49 frames, 50 steps, ALG gaussian_blur k=13 sigma=15 in pixel space + native VAE "
                         "encode every step, interval[0,0.04], gs 6, DDIM (BASELINE.json configs[2]); one sample per GPU")
        self.kernel = f"attention_kernel<64> (joint text+video attention, N={self.n_tok}, 48 heads x 64)"
        self.weights = "seeded random init at the true CogVideoX-5b-I2V architecture (42 blocks, d 3072; no checkpoints offline)"
        self.l2 = "inputs_exceed_l2 (11 GB of weights stream through the 126 MB L2 every step)"
        self.guidance = 6.0

    def n_pass(self, idx):
        return 3 if strength(idx, self.steps_per_video, self.alg) != 0.0 else 2

    def forward_flops(self):
        return joint_dit_flops(self.n_tok, 3072, 42)

    def attn_flops(self, n_pass):
        return 4 * n_pass * self.n_tok * self.n_tok * 3072

    def setup(self, device, rank, world):
        from alg_b200 import distributed as D
        from pipeline_cogvideox_image2video_lowpass import CogVideoXImageToVideoPipeline
        self.pipe = CogVideoXImageToVideoPipeline.from_pretrained("synthetic", synthetic=True, device=device).to(device)
        self.pipe.set_progress_bar_config(disable=True)
        self._broadcast(self.pipe.transformer, device)
        g = torch.Generator(device=device).manual_seed(D.sample_seed(rank))
        F_, h, w = self.F_LAT, 60, 90
        self.lat = torch.randn(1, F_, 16, h, w, generator=g, device=device).bfloat16()
        self.img_lat = torch.cat([torch.randn(1, 1, 16, h, w, generator=g, device=device), torch.zeros(1, F_ - 1, 16, h, w, device=device)], 1).bfloat16()
        self.rgb = (torch.rand(1, 3, self.H, self.W, generator=g, device=device) * 2 - 1).bfloat16()
        self.pos, self.neg = (torch.randn(1, 226, 4096, generator=g, device=device).bfloat16() for _ in range(2))
        self.rope = self.pipe._prepare_rotary_positional_embeddings(self.H, self.W, F_, device)
        self.pipe.scheduler.set_timesteps(self.steps_per_video, device=device)
        self.ts = self.pipe.scheduler.timesteps.tolist()
        self.gen = g
        self.device = device

    def step(self, idx, inputs=None):
        lat, img_lat, rgb, pos, neg = inputs or (self.lat, self.img_lat, self.rgb, self.pos, self.neg)
        out, _ = self.pipe.denoise_step(idx, self.ts[idx], lat, img_lat, rgb, pos, neg, self.rope, self.gen, self.frames,
                                        self.steps_per_video, self.alg, self.guidance)
        return out

    def host_inputs(self):
        return [self.lat, self.img_lat, self.rgb, self.pos, self.neg]

    def full_video(self, generator):
        image = torch.rand(1, 3, self.H, self.W, generator=torch.Generator().manual_seed(1))
        h = [self.pos.cpu(), self.neg.cpu(), self.lat.cpu()]

        def run():
            return self.pipe(image=image, prompt_embeds=h[0].to(self.device), negative_prompt_embeds=h[1].to(self.device),
                             latents=h[2].to(self.device), height=self.H, width=self.W, num_frames=self.frames,
                             num_inference_steps=self.steps_per_video, guidance_scale=self.guidance, generator=generator,
                             output_type="latent", **self.alg).frames.cpu()
        return run, sum(t.numel() * t.element_size() for t in h) + image.numel() * 4

    def parity(self, log):
        sys.path.insert(0, os.path.join(ROOT, "scripts"))
        import parity_fullsize as PF
        res = PF.cog_case(layers=42, fp32_layers=0, eager_backends=("cudnn",), log=log)
        ms = {r["n_pass"]: r.get("eager_cudnn_ms") for r in res["steps"]}
        eager = None
        if ms.get(2) and ms.get(3):
            sec = (2 * ms[3] + 48 * ms[2]) / 1e3
            eager = {"value": self.frames / sec, "unit": "frames/s", "ms_step_3pass": ms[3], "ms_step_2pass": ms[2],
                     "what": "eager PyTorch bf16 DiT forward on this GPU (oracle/cog_oracle.forward, cuBLAS + cuDNN SDPA), "
                             "schedule-weighted 2 x 3-pass + 48 x 2-pass; the per-step VAE encode is not included"}
        return res, eager


class HunyuanWorkload(_SequencedWorkload):
    key = "hunyuan"
    frames, steps_per_video = 129, 30
    T, H, W = 33, 90, 160

    def __init__(self):
        self.alg = alg_kwargs(lp_resize_factor=0.625, schedule_interval_end_time=0.04)  # configs/hunyuan_video_alg.yaml
        self.text_valid = 180
        self.n_tok = self.T * 45 * 80 + self.text_valid
        self.metric = "video frames/sec (HunyuanVideo-I2V 720p, 129 frames, 30 steps, ALG down_up, single-pass branch)"
        self.workload = ("HunyuanVideo-I2V 720x1280, 129 frames, 30 steps, ALG down_up f=0.625 interval[0,0.04] on the first-frame "
                         "latent, embedded guidance 6.0, true_cfg_scale 1.0 (single-pass branch hy:1196-1235), FlowMatchEuler shift 7 "
                         "(BASELINE.json configs[3]); one sample per GPU")
        self.kernel = f"attention_kernel<128> (joint latent+text attention, N={self.n_tok}, 24 heads x 128)"
        self.weights = "seeded random init at the true HunyuanVideo-I2V architecture (20 dual + 40 single blocks, d 3072)"
        self.l2 = "inputs_exceed_l2 (26 GB of weights stream through the 126 MB L2 every step)"

    def n_pass(self, idx):
        return 1

    def forward_flops(self):
        return joint_dit_flops(self.n_tok, 3072, 60)

    def attn_flops(self, n_pass):
        return 4 * n_pass * self.n_tok * self.n_tok * 3072

    def setup(self, device, rank, world):
        import numpy as np
        from alg_b200 import distributed as D
        from pipeline_hunyuan_video_image2video_lowpass import HunyuanVideoImageToVideoPipeline
        self.pipe = HunyuanVideoImageToVideoPipeline.from_pretrained("synthetic", synthetic=True, device=device).to(device)
        self.pipe.set_progress_bar_config(disable=True)
        self._broadcast(self.pipe.transformer, device)
        g = torch.Generator(device=device).manual_seed(D.sample_seed(rank))
        self.lat = torch.randn(1, 16, self.T, self.H, self.W, generator=g, device=device)
        self.img_lat = torch.randn(1, 16, 1, self.H, self.W, generator=g, device=device)
        self.text = torch.randn(1, 256, 4096, generator=g, device=device).bfloat16()
        self.pooled = torch.randn(1, 768, generator=g, device=device).bfloat16()
        self.pipe.scheduler.set_timesteps(sigmas=np.linspace(1.0, 0.0, self.steps_per_video + 1)[:-1], device=device)
        self.ts = self.pipe.scheduler.timesteps.float().cpu()
        self.device = device

    def step(self, idx, inputs=None):
        lat, img_lat, text, pooled = inputs or (self.lat, self.img_lat, self.text, self.pooled)
        self.pipe.scheduler._step_index = idx
        out, _ = self.pipe.denoise_step(idx, self.ts[idx], lat, img_lat, (text, pooled, self.text_valid), None, 6016.0, self.frames,
                                        self.steps_per_video, self.alg, 1.0)
        return out

    def host_inputs(self):
        return [self.lat, self.img_lat, self.text, self.pooled]

    def full_video(self, generator):
        image = torch.rand(1, 3, 720, 1280, generator=torch.Generator().manual_seed(1))
        mask = torch.zeros(1, 256, dtype=torch.int64)
        mask[:, :self.text_valid] = 1
        h = [self.text.cpu(), self.pooled.cpu(), self.lat.cpu()]

        def run():
            return self.pipe(image=image, prompt_embeds=h[0].to(self.device), pooled_prompt_embeds=h[1].to(self.device),
                             prompt_attention_mask=mask.to(self.device), latents=h[2].to(self.device), height=720, width=1280,
                             num_frames=self.frames, num_inference_steps=self.steps_per_video, guidance_scale=6.0,
                             true_cfg_scale=1.0, generator=generator, output_type="latent", **self.alg).frames.cpu()
        return run, sum(t.numel() * t.element_size() for t in h) + image.numel() * 4

    def parity(self, log):
        sys.path.insert(0, os.path.join(ROOT, "scripts"))
        import parity_fullsize as PF
        res = PF.hunyuan_case(eager_backends=("cudnn",), log=log)
        ms = res["steps"][0].get("eager_cudnn_ms")
        eager = None if not ms else {"value": self.frames / (self.steps_per_video * ms / 1e3), "unit": "frames/s", "ms_step_1pass": ms,
                                     "what": "eager PyTorch bf16 DiT forward on this GPU (oracle/hunyuan_oracle.forward, cuBLAS + cuDNN SDPA "
                                             "with the key-padding mask), 30 single-pass steps"}
        return res, eager


def make_workload(name):
    if name == "wan480":
        return WanWorkload("480p")
    if name == "wan720":
        return WanWorkload("720p")
    if name == "cog":
        return CogWorkload()
    return HunyuanWorkload()


# ======================================================================================================
# reference arm / cpu_baseline: the oracle port of the loop on host cores, bounded sample + FLOP extrapolation
# ======================================================================================================
def _host_tflops(threads):
    a = torch.randn(2048, 2048)
    a @ a
    t0 = time.perf_counter()
    for _ in range(3):
        a @ a
    return 3 * 2 * 2048 ** 3 / (time.perf_counter() - t0) / 1e12


def cpu_sample(wl, steps, warmup, budget_s=None, threads=None):
    """Time the CPU restatement (oracle/) of a denoise step on a BOUNDED sample of the config: ONE transformer block (of
    40 / 42 / 60) at full width and -- when the budget allows, as BASELINE.md section 4 plans -- the FULL token count (else
    the largest whole number of latent frames that fits), the real ATen low-pass filter (the reference's own two calls,
    oracle/prepare_lp_oracle.low_pass), CFG and the scheduler update; fp32, all host threads; scaled to the video by
    algorithmic FLOPs.  Returns (frames/s extrapolated, seconds per sample step, description, threads)."""
    from oracle import prepare_lp_oracle as P, sched_oracle

    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    rate = _host_tflops(threads) * 1e12
    g = torch.Generator().manual_seed(42)
    n_steps_total = max(steps + warmup, 1)
    budget = (budget_s if budget_s is not None else 150.0) / n_steps_total

    if wl.key == "wan":
        from oracle import wan_oracle as W
        full_t = T_LAT
        per_frame = (H_LAT // 2) * (W_LAT // 2)
        cand = [t for t in (full_t, 14, 10, 7, 5, 3, 1) if t <= full_t]
        passes = 2
        sel = next((t for t in cand if passes * forward_flops(t * per_frame, layers=1)[0] / rate <= budget), 1)
        cfg = W.WanConfig(num_layers=1)
        sd = W.make_weights(cfg, seed=0, device="cpu", dtype=torch.float32)
        lat = torch.randn(1, 16, sel, H_LAT, W_LAT, generator=g)
        cond = torch.randn(1, 20, sel, H_LAT, W_LAT, generator=g)
        pos, neg = torch.randn(1, 512, 4096, generator=g), torch.randn(1, 512, 4096, generator=g)
        img = torch.randn(1, 257, 1280, generator=g)
        sched = sched_oracle.UniPCOracle(flow_shift=FLOW_SHIFT)
        sched.set_timesteps(STEPS_PER_VIDEO)

        def one_step(i):
            lp = P.low_pass(cond, "down_up", 0.0, 0.0, 0.4)
            x = torch.cat([torch.cat([lat] * 2), torch.cat([cond, lp])], dim=1)
            noise = W.forward(sd, cfg, x, sched.timesteps[i].expand(2), torch.cat([neg, pos]), img.repeat(2, 1, 1))
            sched.step_index, sched.model_outputs, sched.last_sample = i, [lat, lat], lat
            sched.this_order, sched.lower_order_nums = 2, 2
            return sched.step(sched_oracle.cfg_combine(noise.bfloat16(), GUIDANCE), lat)

        n_s = sel * per_frame
        f_sample = passes * forward_flops(n_s, layers=1)[0]
        f_video = sum(wl.n_pass(i) for i in range(wl.steps_per_video)) * wl.forward_flops()[0]
        what = f"one 2-pass denoise step on {sel}/{full_t} latent frames ({n_s} of {wl.n_tok} tokens) x 1/40 blocks"
    elif wl.key == "cog":
        from oracle import cog_oracle as Co
        F_full, per_frame = wl.F_LAT, 30 * 45
        passes = 2
        sel = next((f for f in (F_full, 10, 7, 5, 3, 1) if passes * joint_dit_flops(f * per_frame + 226, 3072, 1)[0] / rate <= budget), 1)
        cfg = Co.CogConfig(num_layers=1, sample_frames=(sel - 1) * 4 + 1)
        sd = Co.make_weights(cfg, seed=0, device="cpu", dtype=torch.float32)
        lat = torch.randn(1, sel, 16, 60, 90, generator=g)
        img_lat = torch.randn(1, sel, 16, 60, 90, generator=g)
        rgb = torch.rand(1, 3, 480, 720, generator=g) * 2 - 1
        pos, neg = torch.randn(1, 226, 4096, generator=g), torch.randn(1, 226, 4096, generator=g)
        rope = Co.rotary_tables(cfg, 30, 45, sel)
        sched = sched_oracle.CogDDIMOracle()
        sched.set_timesteps(50)

        def one_step(i):
            P.low_pass(rgb, "gaussian_blur", 15.0, 0.02734375, 0.25)  # the per-step pixel-space filter (the VAE encode is not timed)
            x = torch.cat([torch.cat([lat] * 2), torch.cat([img_lat] * 2)], dim=2)
            t = sched.timesteps[i]
            noise = Co.forward(sd, cfg, x, torch.cat([neg, pos]), t.expand(2), rope).float()
            return sched.step(sched_oracle.cfg_combine(noise, 6.0, fp32=True), int(t), lat)

        n_s = sel * per_frame + 226
        f_sample = passes * joint_dit_flops(n_s, 3072, 1)[0]
        f_video = sum(wl.n_pass(i) for i in range(wl.steps_per_video)) * wl.forward_flops()[0]
        what = f"one 2-pass denoise step on {sel}/{F_full} latent frames ({n_s} of {wl.n_tok} tokens) x 1/42 blocks"
    else:
        from oracle import hunyuan_oracle as Ho
        per_frame = 45 * 80
        sel = next((t for t in (33, 17, 9, 5, 3, 2) if joint_dit_flops(t * per_frame + 256, 3072, 2)[0] / rate <= budget), 2)
        cfg = Ho.HunyuanConfig(num_layers=1, num_single_layers=1, num_refiner_layers=2)
        sd = Ho.make_weights(cfg, seed=0, device="cpu", dtype=torch.float32)
        lat = torch.randn(1, 16, sel, 90, 160, generator=g)
        first = torch.randn(1, 16, 1, 90, 160, generator=g)
        text, pooled = torch.randn(1, 256, 4096, generator=g), torch.randn(1, 768, generator=g)
        mask = torch.zeros(1, 256)
        mask[:, :wl.text_valid] = 1
        sched = sched_oracle.FlowEulerOracle(shift=7.0)
        import numpy as np
        sched.set_timesteps(30, sigmas=np.linspace(1.0, 0.0, 31)[:-1])

        def one_step(i):
            lp = P.low_pass(first, "down_up", 0.0, 0.0, 0.625)
            x = torch.cat([lp, lat[:, :, 1:]], dim=2)
            noise = Ho.forward(sd, cfg, x, sched.timesteps[i].reshape(1), text, mask, pooled, torch.tensor([6000.0]))
            sched.step_index = i
            return sched.step(noise[:, :, 1:], lat[:, :, 1:])

        n_s = sel * per_frame + 256
        f_sample = joint_dit_flops(n_s, 3072, 2)[0]
        f_video = wl.steps_per_video * wl.forward_flops()[0]
        what = f"one single-pass denoise step on {sel}/33 latent frames ({n_s} of {wl.n_tok} tokens) x (1 dual + 1 single)/60 blocks"

    with torch.no_grad():
        for _ in range(warmup):
            one_step(1)
        t0 = time.perf_counter()
        for k in range(max(steps, 1)):
            one_step(1 + k % 20)
        dt = (time.perf_counter() - t0) / max(steps, 1)
    sec_video = dt * f_video / f_sample
    desc = (f"oracle port (PyTorch fp32 restatement of the reference loop + the reference's own ATen low-pass calls, {threads} "
            f"threads): {what} = {f_sample:.3e} FLOP in {dt:.2f} s; extrapolated by FLOPs to the {f_video:.3e} FLOP video")
    return wl.frames / sec_video, dt, desc, threads


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = make_workload(args.config)
    fps, dt, desc, threads = cpu_sample(wl, args.steps, args.warmup)
    line = {
        "metric": wl.metric, "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "impl": "reference",
        "config": {"workload": wl.workload + "; CPU bounded sample, FLOP-extrapolated"},
        # one host, whatever --gpus says: the CPU arm does not scale with the GPU count (rank 0 alone runs it)
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": threads, "kind": "port", "sample": desc},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ======================================================================================================
# this repo's arm
# ======================================================================================================
def vae_sample(wl, device, fps):
    """AutoencoderKLWan at the workload's size (alg_b200/vae_wan.py, float32 through the bf16 3-term split): decode of one clip of
    latents and encode of the [image, zeros...] condition clip, CUDA-event timed, outside the timed regions."""
    from alg_b200.vae_wan import AutoencoderKLWan
    if hasattr(wl, "pipe"):
        del wl.pipe
    if hasattr(wl, "transformer"):
        del wl.transformer
    torch.cuda.empty_cache()
    vae = AutoencoderKLWan.from_synthetic(seed=0, device=device)
    z = torch.randn(1, 16, *wl.lat0.shape[2:], generator=torch.Generator(device=device).manual_seed(3), device=device)

    def timed(fn):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        out = fn()
        b.record()
        torch.cuda.synchronize()
        return out, a.elapsed_time(b)
    with torch.no_grad():
        vae.decode(z[:, :, :2])  # warm-up (tensor maps, allocator)
        video, ms_dec = timed(lambda: vae.decode(z).sample)
        clip = torch.zeros(1, 3, video.shape[2], video.shape[3], video.shape[4], device=device)
        clip[:, :, 0] = video[:, :, 0]
        finite, shape = bool(torch.isfinite(video).all()), list(video.shape)
        del video
        lat, ms_enc = timed(lambda: vae.encode(clip).latent_dist.mode())
    loop_s = wl.frames / fps
    return {"what": "native AutoencoderKLWan, seeded weights at the Wan2.1 VAE architecture, float32 via the bf16 3-term split; "
                    "NOT inside the metric's timed region (the denoise loop)",
            "decode_ms": ms_dec, "decoded": shape, "decode_finite": finite, "encode_condition_ms": ms_enc, "latent": list(lat.shape),
            "encode_finite": bool(torch.isfinite(lat).all()),
            "frames_per_s_loop_plus_encode_decode": wl.frames / (loop_s + (ms_dec + ms_enc) / 1e3),
            "peak_mem_gb": torch.cuda.max_memory_allocated(device) / 2 ** 30}


def run_ours(args):
    import torch.distributed as dist
    from alg_b200 import _lib
    from alg_b200 import distributed as D

    rank, world, local = D.env_rank()
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the engine has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    _lib.check(_lib.lib().alg_check_device())
    log = lambda s: print(s, file=sys.stderr, flush=True)  # noqa: E731

    wl = make_workload(args.config)
    wl.setup(device, rank, world)
    S = wl.steps_per_video

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    full = None
    if args.full_video:
        args.steps = S
    idxs = [int(i * S / args.steps) for i in range(args.steps)]
    kinds_present = sorted({wl.n_pass(i) for i in range(S)})
    with torch.no_grad():
        warm = [0, min(S - 1, S // 5), min(S - 1, S // 5 + 1)]
        for w in range(args.warmup):
            wl.step(warm[w % 3])
        # ---------------- device-timed region: inputs resident in HBM ------------------------------------
        wl.profile_begin()
        sampler = ClockSampler(local)
        barrier()
        sampler.start()
        launches0 = _lib.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if FAST:
            torch.cuda.cudart().cudaProfilerStart()  # ncu --profile-from-start off: capture the timed region only
        marks = [torch.cuda.Event(enable_timing=True) for _ in idxs]
        e0.record()
        out = None
        for k, idx in enumerate(idxs):
            out = wl.step(idx)
            marks[k].record()
        e1.record()
        barrier()
        if FAST:
            torch.cuda.cudart().cudaProfilerStop()
        sampler.stop_flag = True
        launches = _lib.launch_count() - launches0
        ms_total = e0.elapsed_time(e1)
        step_ms = [(e0 if k == 0 else marks[k - 1]).elapsed_time(marks[k]) for k in range(len(idxs))]
        prof = wl.profile_end()
        finite = bool(torch.isfinite(out).all())
        # ---------------- end-to-end region: host buffers in, host result out, every step ----------------
        ms_e2e, h2d, d2h = 0.0, 0, 0
        if not FAST:
            host_in = [t.cpu().pin_memory() for t in wl.host_inputs()]
            host_out = torch.empty_like(out, device="cpu").pin_memory()
            h2d = sum(t.numel() * t.element_size() for t in host_in)
            d2h = host_out.numel() * host_out.element_size()
            barrier()
            f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            f0.record()
            for idx in idxs:
                dev_in = tuple(t.to(device, non_blocking=True) for t in host_in)
                host_out.copy_(wl.step(idx, dev_in), non_blocking=True)
            f1.record()
            barrier()
            ms_e2e = f0.elapsed_time(f1)
        # ---------------- optional: the public __call__ for the whole schedule, free-running ------------------
        if args.full_video:
            run, in_bytes = wl.full_video(torch.Generator(device=device).manual_seed(D.sample_seed(rank)))
            barrier()
            t0 = time.perf_counter()
            frames = run()
            torch.cuda.synchronize()
            wall = time.perf_counter() - t0
            full = {"what": "pipe.__call__(output_type='latent') for all steps, free-running, host inputs in / host latents out",
                    "steps": S, "wall_s": wall, "frames_per_s": wl.frames / wall, "finite": bool(torch.isfinite(frames).all()),
                    "h2d_bytes": in_bytes, "d2h_bytes": frames.numel() * frames.element_size()}

    ms_total, ms_e2e = D.max_over_ranks([ms_total, ms_e2e], device)
    ms_step = ms_total / args.steps
    # The video mixes step kinds (Wan: 10 three-pass + 40 two-pass).  K uniformly sampled steps reproduce that mix only for
    # some K, so the rate is taken from the schedule-weighted step time whenever every kind was sampled; `ms_per_step` stays
    # the raw mean of the K timed steps.  mix = weighted / raw.
    kinds = {p_: [m for m, i in zip(step_ms, idxs) if wl.n_pass(i) == p_] for p_ in kinds_present}
    mix = 1.0
    if all(kinds[p_] for p_ in kinds_present):
        weighted = sum(sum(1 for i in range(S) if wl.n_pass(i) == p_) * sum(kinds[p_]) / len(kinds[p_]) for p_ in kinds_present) / S
        mix = weighted / (sum(step_ms) / len(step_ms))
    fps = D.aggregate_rate(wl.frames / S, world, ms_step * mix)
    fps_e2e = D.aggregate_rate(wl.frames / S, world, ms_e2e / args.steps * mix) if ms_e2e > 0 else None

    if rank == 0:
        hbm, tf_burst, tf_sust, src = peaks()
        sa = prof["self_attention"]
        passes = sum(wl.n_pass(i) for i in idxs)
        flops_total = sum(wl.attn_flops(wl.n_pass(i)) for i in idxs) * (sa["launches"] / max(len(idxs), 1))  # launches per step = layers
        achieved = flops_total / (sa["ms"] / 1e3) / 1e12 if sa["ms"] > 0 else None
        total_flops = passes * wl.forward_flops()[0]
        traffic = None  # DRAM bytes per launch from the committed ncu --set full capture of this kernel (per head x heads)
        tpath = os.path.join(ROOT, "profiles", "r02_attention_traffic.json")
        if wl.key == "wan" and os.path.exists(tpath) and sa["launches"]:
            heads_per_launch = 40 * passes / len(idxs)
            traffic = json.load(open(tpath))["dram_bytes_per_head"] * heads_per_launch * wl.n_tok / 32760  # linear in tokens
        roofline = {"kernel": wl.kernel, "bound": "tensor", "achieved": achieved, "peak": tf_sust, "unit": "TFLOP/s",
                    "frac": achieved / tf_sust if achieved else None,
                    "peak_kind": f"bf16_tflops_sustained ({src}); kernel timed inside a long step",
                    "frac_of_burst_peak": achieved / tf_burst if achieved else None, "traffic": traffic,
                    "traffic_unit": "bytes per launch (ncu dram__bytes_read + write per head x mean heads per launch)",
                    "launches": sa["launches"], "avg_launch_ms": sa["ms"] / max(sa["launches"], 1),
                    "share_of_step": sa["ms"] / ms_total,
                    "per_class_ms": {k: round(v["ms"], 3) for k, v in prof.items()},
                    "whole_step_tflops": total_flops / (ms_total / 1e3) / 1e12,
                    "ms_per_step_by_passes": {str(p_): round(sum(kinds[p_]) / len(kinds[p_]), 3) for p_ in kinds_present if kinds[p_]}}
        if "gemm" in prof and prof["gemm"]["ms"] > 0 and wl.key == "wan":
            gemm_flops = passes * (wl.forward_flops()[0] - wl.forward_flops()[1] - 40 * 4 * wl.n_tok * 769 * 5120)
            roofline["gemm_tflops_in_step"] = gemm_flops / (prof["gemm"]["ms"] / 1e3) / 1e12
        cpu = parity = eager = None
        if not FAST:
            if world == 1:  # the CPU baseline is an N = 1 leg (rank 0's host cores); N > 1 lines carry null
                try:
                    c_fps, c_dt, c_desc, c_thr = cpu_sample(wl, steps=1, warmup=0, budget_s=45.0)
                    cpu = {"value": c_fps, "unit": "frames/s", "cores": c_thr, "kind": "port", "sample": c_desc}
                except Exception as ex:  # pragma: no cover
                    cpu = {"value": None, "unit": "frames/s", "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {ex}"}
            if args.parity and world == 1:
                # checker leg, OUTSIDE every timed region: a teacher-forced step of each kind at the config size and FULL depth,
                # engine vs the eager bf16 oracle on this GPU (scripts/parity_fullsize.py); the eager step time is the GPU bar
                try:
                    del wl.pipe
                    if hasattr(wl, "transformer"):
                        del wl.transformer
                    torch.cuda.empty_cache()
                    with torch.no_grad():
                        res, eager = wl.parity(log)
                    parity = {"protocol": "teacher-forced, scheduler history synced; noise_pred / stepped-latent relative L2 vs "
                                          "the eager bf16 oracle (itself ~1.2e-2 / ~1.2e-3 away from its own flash-SDPA variant: "
                                          "profiles/r02_parity_fullsize.json)",
                              "layers": res.get("layers"), "tokens": res.get("tokens"),
                              "steps": [{k: r.get(k) for k in ("schedule_index", "n_pass", "noise_rel_l2", "latent_rel_l2",
                                                               "noise_finite", "sched_kernel_on_oracle_noise_bitexact")}
                                        for r in res.get("steps", [])],
                              "max_latent_rel_l2": res.get("max_latent_rel_l2")}
                except Exception as ex:  # pragma: no cover
                    parity = {"error": f"{type(ex).__name__}: {ex}"[:300]}
        vae_leg = None
        if not FAST and args.vae and world == 1 and wl.key == "wan":
            # informational leg, OUTSIDE every timed region: what turns the loop's latents into the frames the metric counts
            # (wan:959) and what built the condition (wan:429-434), on the native float32 VAE at the config size
            try:
                vae_leg = vae_sample(wl, device, fps)
            except Exception as ex:  # pragma: no cover
                vae_leg = {"error": f"{type(ex).__name__}: {ex}"[:300]}
        line = {
            "metric": wl.metric, "value": fps, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": wl.dtype,
            "data": "synthetic",
            "config": {"workload": wl.workload,
                       "step_sampling": f"schedule indices {idxs} ({passes} sample-forwards in {args.steps} steps); value = frames per "
                                        f"step / (ms_per_step x {mix:.4f}), the schedule-weighted step time",
                       "weights": wl.weights, "l2": wl.l2,
                       "parallelism": f"dp{world} (independent samples; one flat-arena NCCL weight broadcast at init only)"},
            "roofline": roofline, "cpu_baseline": cpu,
            "e2e": None if FAST else {"value": fps_e2e, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches), "clocks": sampler.summary(), "output_finite": finite,
        }
        if eager is not None:
            line["eager_gpu"] = eager
        if parity is not None:
            line["parity_at_config"] = parity
        if full is not None:
            line["full_video"] = full
        if vae_leg is not None:
            line["vae"] = vae_leg
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default=None, choices=["wan480", "wan720", "cog", "hunyuan"],
                    help="wan480 = BASELINE.json configs[1] (the headline workload, default); wan720 = the same loop on 720x1280 "
                         "latents (75 600 tokens); cog = configs[2]; hunyuan = configs[3]")
    ap.add_argument("--resolution", default="480p", choices=["480p", "720p"], help="(older spelling) 720p = --config wan720")
    ap.add_argument("--full-video", action="store_true", help="also run pipe.__call__ for the whole schedule, free-running")
    ap.add_argument("--parity", type=int, default=int(os.environ.get("ALG_BENCH_PARITY", "1")),
                    help="1 (default, N = 1 only): after the timed regions, run one teacher-forced step of each kind at full depth "
                         "against the eager bf16 oracle on the GPU and report parity_at_config + eager_gpu")
    ap.add_argument("--vae", type=int, default=int(os.environ.get("ALG_BENCH_VAE", "1")),
                    help="1 (default, N = 1, Wan configs): after the timed regions, decode one clip of latents and encode the condition "
                         "clip on the native AutoencoderKLWan (seeded weights at the true architecture) and report the times")
    args = ap.parse_args()
    if args.config is None:
        args.config = "wan720" if args.resolution == "720p" else "wan480"
    # stdout carries exactly ONE line, the JSON: everything a library prints there while the run is in flight (NCCL's
    # "NCCL version ..." banner at init, for one) is routed to stderr, and the result goes to the saved descriptor.
    global _RESULT_FD
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
