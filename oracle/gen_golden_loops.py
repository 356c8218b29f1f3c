"""Golden vectors of the three ALG denoise loops, produced by the UNMODIFIED reference pipelines.

TEST INFRASTRUCTURE.  Run in the build container (it needs /root/reference):

    python oracle/gen_golden_loops.py            # writes tests/golden/loop_*.npz

``/root/reference/pipeline_{wan,cogvideox,hunyuan_video}_image2video_lowpass.py`` import ``diffusers``, which is absent;
``oracle/refshim`` provides a fake ``diffusers`` with exactly the imported names, so the first-party files import as they
are and their real ``__call__`` runs here: ``check_inputs``, ``prepare_latents``, the per-step ``get_lp_strength`` /
parameter modulation / ``prepare_lp`` / model-input assembly / CFG combine / ``scheduler.step`` / callback handling.  The
objects handed to the constructor are the stand-ins the third-party pieces need: the DiT = ``oracle/*_oracle.forward`` on
seeded tiny weights (bf16, CPU), the scheduler = ``oracle/sched_oracle`` behind the diffusers call surface, the VAE =
``oracle/stub_vae.ArithVAE``.  Every transformer call (inputs and output) and every post-step latent is recorded.

What the fixtures pin: rows a2, a3-a6, a7, a9, a11 of SURVEY 8 (first-party code) bit for bit; the DiT / scheduler
arithmetic stays "parity unpinned" (third-party, absent) -- the fixtures carry the DiT outputs so that tests can REPLAY
them and compare everything around the DiT exactly.
"""
from __future__ import annotations

import json
import os
import sys
from types import SimpleNamespace

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference"
sys.path[:0] = [os.path.join(HERE, "refshim"), REF]
sys.path.append(ROOT)  # for `oracle.*` only: the reference directory shadows the repo's same-named pipeline modules

import numpy as np  # noqa: E402
import torch  # noqa: E402

import lp_utils  # noqa: E402  (the reference's)
import pipeline_cogvideox_image2video_lowpass as ref_cog  # noqa: E402
import pipeline_hunyuan_video_image2video_lowpass as ref_hy  # noqa: E402
import pipeline_wan_image2video_lowpass as ref_wan  # noqa: E402
from diffusers import schedulers as shim_sched  # noqa: E402
from oracle import cog_oracle as Co, hunyuan_oracle as Ho, sched_oracle, wan_oracle as W  # noqa: E402
from oracle.stub_vae import ArithVAE, StubImageEncoder, StubImageProcessor, StubTextEncoder, StubTokenizer  # noqa: E402

for m in (lp_utils, ref_wan, ref_cog, ref_hy):
    assert m.__file__.startswith(REF), m.__file__

OUT = os.path.join(ROOT, "tests", "golden")
DT = torch.bfloat16


def pack(t: torch.Tensor):
    """bf16 -> its bit pattern (numpy has no bf16); everything else as is."""
    t = t.detach().cpu().contiguous()
    if t.dtype == torch.bfloat16:
        return t.view(torch.int16).numpy(), "bf16"
    return t.numpy(), str(t.dtype).replace("torch.", "")


class Recorder:
    def __init__(self):
        self.arrays, self.dtypes = {}, {}

    def put(self, name, t):
        a, d = pack(t)
        self.arrays[name], self.dtypes[name] = a, d

    def save(self, path, meta):
        meta = dict(meta, dtypes=self.dtypes)
        np.savez_compressed(path, meta=np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8), **self.arrays)
        print(f"{os.path.basename(path)}: {os.path.getsize(path) / 1024:.0f} KB, {len(self.arrays)} arrays")


def which_rows(batch, named):
    """Describe a stacked conditioning batch as a string of names ('nnp', 'np', ...) by exact comparison."""
    out = ""
    for row in batch:
        hit = [k for k, v in named.items() if v.shape == row.shape and torch.equal(v, row)]
        assert len(hit) >= 1, "conditioning row matches none of the known tensors"
        out += hit[0]
    return out


# ----------------------------------------------------------------------------------------------------------------
# scheduler adapters: oracle/sched_oracle.py behind the diffusers call surface
# ----------------------------------------------------------------------------------------------------------------
class UniPCAdapter(shim_sched.UniPCMultistepScheduler):
    def __init__(self, flow_shift):
        self.o = sched_oracle.UniPCOracle(flow_shift=flow_shift)
        self.config = SimpleNamespace(flow_shift=flow_shift)

    def set_timesteps(self, num_inference_steps, device=None):
        self.o.set_timesteps(num_inference_steps)
        self.timesteps = self.o.timesteps.to(device)

    def step(self, model_output, timestep, sample, return_dict=True):
        return (self.o.step(model_output, sample),)


class CogDDIMAdapter(shim_sched.CogVideoXDDIMScheduler):
    def __init__(self):
        self.o = sched_oracle.CogDDIMOracle()

    def set_timesteps(self, num_inference_steps, device=None):
        self.o.set_timesteps(num_inference_steps)
        self.timesteps = self.o.timesteps.to(device)

    def scale_model_input(self, sample, timestep=None):
        return sample

    def step(self, model_output, timestep, sample, eta: float = 0.0, return_dict: bool = True):
        return (self.o.step(model_output, int(timestep), sample),)


class CogDPMAdapter(shim_sched.CogVideoXDPMScheduler):
    def __init__(self):
        self.o = sched_oracle.CogDPMOracle()

    def set_timesteps(self, num_inference_steps, device=None):
        self.o.set_timesteps(num_inference_steps)
        self.timesteps = self.o.timesteps.to(device)

    def scale_model_input(self, sample, timestep=None):
        return sample

    def step(self, model_output, old_pred_original_sample, timestep, timestep_back, sample, eta: float = 0.0,
             generator=None, return_dict: bool = False):
        from diffusers.utils.torch_utils import randn_tensor
        draw = lambda: randn_tensor(sample.shape, generator=generator, device=sample.device, dtype=sample.dtype)  # noqa: E731
        return self.o.step(model_output, old_pred_original_sample, int(timestep),
                           None if timestep_back is None else int(timestep_back), sample, draw)


class EulerAdapter(shim_sched.FlowMatchEulerDiscreteScheduler):
    def __init__(self, shift):
        self.o = sched_oracle.FlowEulerOracle(shift=shift)

    def set_timesteps(self, num_inference_steps=None, device=None, sigmas=None):
        self.o.set_timesteps(num_inference_steps, sigmas=sigmas)
        self.timesteps = self.o.timesteps.to(device)

    def step(self, model_output, timestep, sample, return_dict=True):
        return (self.o.step(model_output, sample),)


# ----------------------------------------------------------------------------------------------------------------
# Wan
# ----------------------------------------------------------------------------------------------------------------
WAN_TINY = dict(num_attention_heads=2, attention_head_dim=128, text_dim=64, freq_dim=256, ffn_dim=512, num_layers=2,
                image_dim=64, text_len=32)


class WanDiT:
    def __init__(self, rec, named):
        self.cfg = W.WanConfig(**WAN_TINY)
        self.sd = W.make_weights(self.cfg, seed=11, device="cpu", dtype=DT)
        self.config = SimpleNamespace(patch_size=self.cfg.patch_size)
        self.dtype = DT
        self.rec, self.named, self.n = rec, named, 0

    def __call__(self, hidden_states, timestep, encoder_hidden_states, encoder_hidden_states_image=None,
                 attention_kwargs=None, return_dict=True):
        out = W.forward(self.sd, self.cfg, hidden_states, timestep, encoder_hidden_states, encoder_hidden_states_image)
        i = self.n
        self.rec.put(f"hs_{i}", hidden_states)
        self.rec.put(f"noise_{i}", out)
        self.rec.put(f"t_{i}", timestep)
        self.rec.arrays[f"text_{i}"] = np.frombuffer(which_rows(encoder_hidden_states, self.named).encode(), dtype=np.uint8)
        self.n += 1
        return (out,)


def wan_case(name, alg, steps=5, prompts=None, n_videos=1, last=False, guidance=5.0, num_frames=5, height=96, width=128, seed=5,
             expect_error=None):
    """``prompts`` = (prompt, negative_prompt) strings / lists: conditioning through the (stub) tokenizer + text encoder,
    the only way ``num_videos_per_prompt > 1`` works in the reference; otherwise seeded ``prompt_embeds``."""
    g = torch.Generator().manual_seed(seed)
    pos = torch.randn(1, 32, 64, generator=g).to(DT)
    neg = torch.randn(1, 32, 64, generator=g).to(DT)
    img_table = torch.randn(2, 9, 64, generator=g).to(DT)  # CLIP penultimate hidden states of (image, last_image)
    image = torch.rand(1, 3, height, width, generator=g)
    last_image = torch.rand(1, 3, height, width, generator=g) if last else None
    rec = Recorder()
    named = {"p": pos[0], "n": neg[0]}
    dit = WanDiT(rec, named)
    pipe = ref_wan.WanImageToVideoPipeline(tokenizer=StubTokenizer(), text_encoder=StubTextEncoder(),
                                           image_encoder=StubImageEncoder(img_table), image_processor=StubImageProcessor(),
                                           transformer=dit, vae=ArithVAE("wan"), scheduler=UniPCAdapter(5.0))
    text_kw = dict(prompt_embeds=pos, negative_prompt_embeds=neg)
    batch = 1
    if prompts is not None:
        text_kw = dict(prompt=prompts[0], negative_prompt=prompts[1], max_sequence_length=32)
        batch = (len(prompts[0]) if isinstance(prompts[0], list) else 1) * n_videos
        pe, ne = pipe.encode_prompt(prompts[0], prompts[1], True, 1, max_sequence_length=32, device=torch.device("cpu"))
        named.clear()
        for b in range(pe.shape[0]):
            named[f"p{b}"], named[f"n{b}"] = pe[b].to(DT), ne[b].to(DT)
    lat_after = []

    def cb(p, i, t, kw):
        lat_after.append(kw["latents"].clone())
        return {}

    gen = [torch.Generator().manual_seed(100 + b) for b in range(batch)] if batch > 1 else torch.Generator().manual_seed(100)
    call = dict(image=image, last_image=last_image, height=height, width=width, num_frames=num_frames, num_inference_steps=steps,
                guidance_scale=guidance, num_videos_per_prompt=n_videos, generator=gen, output_type="latent",
                callback_on_step_end=cb, **text_kw, **alg)
    if expect_error is not None:
        try:
            pipe(**call)
        except expect_error as ex:
            return {"case": name, "raises": type(ex).__name__, "message": str(ex).splitlines()[0][:200],
                    "transformer_calls_before_error": dit.n}
        raise AssertionError(f"{name}: the reference was expected to raise {expect_error}")
    out = pipe(**call)
    assert torch.equal(out.frames, lat_after[-1])
    for i, x in enumerate(lat_after):
        rec.put(f"lat_{i}", x)
    # what prepare_latents produced (replayed with the same generator seeds)
    gen2 = [torch.Generator().manual_seed(100 + b) for b in range(batch)] if batch > 1 else torch.Generator().manual_seed(100)
    img_t = pipe.video_processor.preprocess(image, height=height, width=width).to(torch.float32)
    last_t = pipe.video_processor.preprocess(last_image, height=height, width=width).to(torch.float32) if last else None
    lat0, cond = pipe.prepare_latents(img_t, batch, 16, height, width, num_frames, torch.float32, torch.device("cpu"), gen2,
                                      None, last_t)
    for k, v in dict(pos=pos, neg=neg, image_table=img_table, image=image, lat0=lat0, condition=cond).items():
        rec.put(k, v)
    if last:
        rec.put("last_image", last_image)
    rec.save(os.path.join(OUT, f"loop_wan_{name}.npz"),
             dict(model="wan", cfg=WAN_TINY, weights_seed=11, alg=alg, steps=steps, batch=batch, n_videos=n_videos,
                  prompts=prompts, guidance=guidance, num_frames=num_frames, height=height, width=width, flow_shift=5.0,
                  generator_seed=100, n_calls=dit.n, text_encoder_seed=99,
                  reference="pipeline_wan_image2video_lowpass.WanImageToVideoPipeline.__call__ (unmodified, via oracle/refshim)"))


# ----------------------------------------------------------------------------------------------------------------
# CogVideoX
# ----------------------------------------------------------------------------------------------------------------
COG_TINY = dict(num_attention_heads=2, attention_head_dim=64, time_embed_dim=64, text_embed_dim=64, num_layers=2,
                sample_width=12, sample_height=8, sample_frames=9, max_text_seq_length=16)


class CogDiT:
    def __init__(self, rec, named, dt):
        self.cfg = Co.CogConfig(**COG_TINY)
        self.sd = Co.make_weights(self.cfg, seed=12, device="cpu", dtype=dt)
        c = self.cfg
        self.config = SimpleNamespace(patch_size=c.patch_size, patch_size_t=None, sample_width=c.sample_width,
                                      sample_height=c.sample_height, sample_frames=c.sample_frames, in_channels=c.in_channels,
                                      attention_head_dim=c.attention_head_dim, use_rotary_positional_embeddings=True,
                                      ofs_embed_dim=None)
        self.dtype = dt
        self.rec, self.named, self.n = rec, named, 0

    def __call__(self, hidden_states, encoder_hidden_states, timestep, ofs=None, image_rotary_emb=None, attention_kwargs=None,
                 return_dict=True):
        out = Co.forward(self.sd, self.cfg, hidden_states, encoder_hidden_states, timestep, image_rotary_emb)
        i = self.n
        self.rec.put(f"hs_{i}", hidden_states)
        self.rec.put(f"noise_{i}", out)
        self.rec.put(f"t_{i}", timestep)
        self.rec.arrays[f"text_{i}"] = np.frombuffer(which_rows(encoder_hidden_states, self.named).encode(), dtype=np.uint8)
        if i == 0:
            self.rec.put("rope_cos", image_rotary_emb[0])
            self.rec.put("rope_sin", image_rotary_emb[1])
        self.n += 1
        return (out,)


def cog_case(name, alg, steps=5, guidance=6.0, dpm=False, use_dynamic_cfg=False, num_frames=9, seed=6, dt=torch.float32):
    """fp32 by default = BASELINE.json configs[0] (CPU ATen has no bf16 antialiased bilinear, quirk q16); the Gaussian
    pixel-space case runs in bf16 like configs[2]."""
    g = torch.Generator().manual_seed(seed)
    pos = torch.randn(1, 16, 64, generator=g).to(dt)
    neg = torch.randn(1, 16, 64, generator=g).to(dt)
    height, width = 64, 96
    image = torch.rand(1, 3, height, width, generator=g)
    rec = Recorder()
    dit = CogDiT(rec, {"p": pos[0], "n": neg[0]}, dt)
    pipe = ref_cog.CogVideoXImageToVideoPipeline(tokenizer=None, text_encoder=None, vae=ArithVAE("cog", dtype=dt), transformer=dit,
                                                 scheduler=CogDPMAdapter() if dpm else CogDDIMAdapter())
    lat_after = []

    def cb(p, i, t, kw):
        lat_after.append(kw["latents"].clone())
        return {}

    out = pipe(image=image, prompt_embeds=pos, negative_prompt_embeds=neg, height=height, width=width, num_frames=num_frames,
               num_inference_steps=steps, guidance_scale=guidance, use_dynamic_cfg=use_dynamic_cfg,
               generator=torch.Generator().manual_seed(200), output_type="latent", callback_on_step_end=cb, **alg)
    assert torch.equal(out.frames, lat_after[-1])
    for i, x in enumerate(lat_after):
        rec.put(f"lat_{i}", x)
    img_t = pipe.video_processor.preprocess(image, height=height, width=width).to(torch.device("cpu"), dtype=dt)
    lat0, img_lat = pipe.prepare_latents(img_t, 1, 16, num_frames, height, width, dt, torch.device("cpu"),
                                         torch.Generator().manual_seed(200), None)
    for k, v in dict(pos=pos, neg=neg, image=image, lat0=lat0, image_latents=img_lat).items():
        rec.put(k, v)
    rec.save(os.path.join(OUT, f"loop_cog_{name}.npz"),
             dict(model="cog", cfg=COG_TINY, weights_seed=12, alg=alg, steps=steps, guidance=guidance, dpm=dpm,
                  use_dynamic_cfg=use_dynamic_cfg, num_frames=num_frames, height=height, width=width, generator_seed=200,
                  dtype=str(dt).replace("torch.", ""),
                  n_calls=dit.n,
                  reference="pipeline_cogvideox_image2video_lowpass.CogVideoXImageToVideoPipeline.__call__ (unmodified, via oracle/refshim)"))


# ----------------------------------------------------------------------------------------------------------------
# HunyuanVideo
# ----------------------------------------------------------------------------------------------------------------
HY_TINY = dict(num_attention_heads=2, attention_head_dim=128, num_layers=2, num_single_layers=2, num_refiner_layers=1,
               text_embed_dim=64, pooled_projection_dim=32)


class HyDiT:
    def __init__(self, rec, named_text, named_pooled):
        self.cfg = Ho.HunyuanConfig(**HY_TINY)
        self.sd = Ho.make_weights(self.cfg, seed=13, device="cpu", dtype=DT)
        self.config = SimpleNamespace(in_channels=16, guidance_embeds=True, patch_size=2, patch_size_t=1,
                                      image_condition_type="token_replace")
        self.dtype = DT
        self.rec, self.named_text, self.named_pooled, self.n = rec, named_text, named_pooled, 0

    def __call__(self, hidden_states, timestep, encoder_hidden_states, encoder_attention_mask, pooled_projections,
                 guidance=None, attention_kwargs=None, return_dict=True):
        out = Ho.forward(self.sd, self.cfg, hidden_states, timestep, encoder_hidden_states, encoder_attention_mask,
                         pooled_projections, guidance)
        i = self.n
        self.rec.put(f"hs_{i}", hidden_states)
        self.rec.put(f"noise_{i}", out)
        self.rec.put(f"t_{i}", timestep)
        self.rec.put(f"guidance_{i}", guidance)
        self.rec.put(f"mask_{i}", encoder_attention_mask)
        self.rec.arrays[f"text_{i}"] = np.frombuffer(which_rows(encoder_hidden_states, self.named_text).encode(), dtype=np.uint8)
        self.rec.arrays[f"pooled_{i}"] = np.frombuffer(which_rows(pooled_projections, self.named_pooled).encode(), dtype=np.uint8)
        self.n += 1
        return (out,)


def hy_case(name, alg, steps=5, guidance=6.0, true_cfg=1.0, num_frames=9, height=64, width=128, seed=7, **extra):
    g = torch.Generator().manual_seed(seed)
    L = 24
    pos, neg = torch.randn(1, L, 64, generator=g).to(DT), torch.randn(1, L, 64, generator=g).to(DT)
    ppos, pneg = torch.randn(1, 32, generator=g).to(DT), torch.randn(1, 32, generator=g).to(DT)
    mpos, mneg = torch.zeros(1, L, dtype=torch.int64), torch.zeros(1, L, dtype=torch.int64)
    mpos[:, :17] = 1
    mneg[:, :9] = 1
    image = torch.rand(1, 3, height, width, generator=g)
    rec = Recorder()
    dit = HyDiT(rec, {"p": pos[0], "n": neg[0]}, {"p": ppos[0], "n": pneg[0]})
    pipe = ref_hy.HunyuanVideoImageToVideoPipeline(text_encoder=None, tokenizer=None, transformer=dit, vae=ArithVAE("hunyuan"),
                                                   scheduler=EulerAdapter(7.0), text_encoder_2=None, tokenizer_2=None,
                                                   image_processor=None)
    lat_after = []

    def cb(p, i, t, kw):
        lat_after.append(kw["latents"].clone())
        return {}

    kw = dict(prompt_embeds=pos, pooled_prompt_embeds=ppos, prompt_attention_mask=mpos)
    if true_cfg > 1:
        kw.update(negative_prompt_embeds=neg, negative_pooled_prompt_embeds=pneg, negative_prompt_attention_mask=mneg)
    out = pipe(image=image, height=height, width=width, num_frames=num_frames, num_inference_steps=steps,
               guidance_scale=guidance, true_cfg_scale=true_cfg, generator=torch.Generator().manual_seed(300),
               output_type="latent", callback_on_step_end=cb, **kw, **alg, **extra)
    assert torch.equal(out.frames, lat_after[-1])
    for i, x in enumerate(lat_after):
        rec.put(f"lat_{i}", x)
    img_t = pipe.video_processor.preprocess(image, height, width).to(torch.device("cpu"), torch.float32)
    lat0, img_lat = pipe.prepare_latents(img_t, 1, 16, height, width, num_frames, torch.float32, torch.device("cpu"),
                                         torch.Generator().manual_seed(300), None, "token_replace", extra.get("i2v_stable", False))
    for k, v in dict(pos=pos, neg=neg, pooled_pos=ppos, pooled_neg=pneg, mask_pos=mpos, mask_neg=mneg, image=image, lat0=lat0,
                     image_latents=img_lat).items():
        rec.put(k, v)
    rec.save(os.path.join(OUT, f"loop_hunyuan_{name}.npz"),
             dict(model="hunyuan", cfg=HY_TINY, weights_seed=13, alg=alg, steps=steps, guidance=guidance, true_cfg=true_cfg,
                  num_frames=num_frames, height=height, width=width, shift=7.0, generator_seed=300, n_calls=dit.n, extra=extra,
                  reference="pipeline_hunyuan_video_image2video_lowpass.HunyuanVideoImageToVideoPipeline.__call__ (unmodified, via oracle/refshim)"))


def alg_kwargs(**over):
    base = dict(use_low_pass_guidance=True, lp_filter_type="down_up", lp_filter_in_latent=True, lp_blur_sigma=15.0,
                lp_blur_kernel_size=0.02734375, lp_resize_factor=0.4, lp_strength_schedule_type="interval",
                schedule_blur_kernel_size=False, schedule_interval_start_time=0.0, schedule_interval_end_time=0.3,
                schedule_linear_start_weight=1.0, schedule_linear_end_weight=0.0, schedule_linear_end_time=0.5,
                schedule_exp_decay_rate=10.0)
    base.update(over)
    return base


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(1)  # one summation order
    with torch.no_grad():
        wan_case("latent_down_up", alg_kwargs())
        wan_case("pixel_gaussian_linear", alg_kwargs(lp_filter_type="gaussian_blur", lp_filter_in_latent=False, lp_blur_sigma=3.0,
                                                    lp_blur_kernel_size=0.15, lp_strength_schedule_type="linear",
                                                    schedule_blur_kernel_size=True, schedule_linear_end_time=0.6), steps=4)
        # batch > 1 in the reference: works for vanilla CFG with string prompts + num_videos_per_prompt; a 3-pass ALG step
        # breaks (wan:919 tests shape[0] == 3, so 6 rows are chunked in two), and a LIST of prompts breaks at wan:905-908
        # (image_embeds repeated batch x rows times)
        wan_case("two_videos_vanilla", dict(use_low_pass_guidance=False), steps=3,
                 prompts=("a red bus turning a corner", "blurry, low quality"), n_videos=2)
        quirks = [wan_case("two_videos_alg", alg_kwargs(), steps=2, prompts=("a red bus turning a corner", "blurry, low quality"),
                           n_videos=2, expect_error=RuntimeError),
                  wan_case("two_prompts", dict(use_low_pass_guidance=False), steps=2,
                           prompts=(["a red bus", "a green tram"], ["blurry", "dark"]), expect_error=RuntimeError)]
        json.dump({"note": "behaviour of the UNMODIFIED reference observed by oracle/gen_golden_loops.py", "cases": quirks},
                  open(os.path.join(OUT, "loop_quirks.json"), "w"), indent=1)
        wan_case("last_image", alg_kwargs(lp_strength_schedule_type="exponential", schedule_exp_decay_rate=4.0), steps=3, last=True)
        wan_case("vanilla", dict(use_low_pass_guidance=False), steps=3)
        cog_case("latent_down_up", alg_kwargs(lp_resize_factor=0.25))
        cog_case("pixel_gaussian", alg_kwargs(lp_filter_type="gaussian_blur", lp_filter_in_latent=False, lp_blur_sigma=3.0,
                                              lp_blur_kernel_size=0.2, schedule_interval_end_time=0.45), steps=4, dt=DT)
        cog_case("exponential_q11", alg_kwargs(lp_resize_factor=0.5, lp_strength_schedule_type="exponential",
                                               schedule_exp_decay_rate=6.0), steps=4)
        cog_case("dpm_latent", alg_kwargs(lp_resize_factor=0.25), steps=4, dpm=True)
        cog_case("vanilla_dynamic_cfg", dict(use_low_pass_guidance=False), steps=3, use_dynamic_cfg=True)
        hy_case("single_pass_down_up", alg_kwargs(lp_resize_factor=0.625))
        hy_case("true_cfg_down_up", alg_kwargs(lp_resize_factor=0.625), true_cfg=4.0, steps=4)
        hy_case("true_cfg_noisy_latent", alg_kwargs(lp_resize_factor=0.625), true_cfg=4.0, steps=3, lp_on_noisy_latent=True)
        hy_case("vanilla_stable", dict(use_low_pass_guidance=False), steps=3, i2v_stable=True)


if __name__ == "__main__":
    main()
