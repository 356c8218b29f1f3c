"""The three drop-in pipelines' ``__call__`` on the GPU against vectors produced by the UNMODIFIED reference pipelines
(tests/golden/loop_*.npz; generator oracle/gen_golden_loops.py).

The DiT is REPLAYED (a stand-in ``transformer`` object returns the noise prediction the reference's transformer produced
and checks the model input it is handed), so everything the product does AROUND the DiT -- ``prepare_latents``, the
strength schedule, ``prepare_lp`` through the CUDA low-pass kernels (+ VAE sample on the caller's generator in pixel
mode), the never-materialised model input, the fused CFG + scheduler kernels, callbacks -- is compared with the
reference's own per-step latents.  Tolerances: model input within 1 bf16 ulp on < 0.5 % of the elements (the CUDA filters
are within 2e-6 of ATen's, which can flip a bf16 rounding); per-step latents <= 2e-6 relative L2 for fp32 state, 1 bf16 ulp
on < 0.5 % of the elements for bf16 state."""
import pytest
import torch

import golden_loops as GL
from conftest import rel_l2

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _ulp_close(a, b, what, frac=5e-3):
    assert a.shape == b.shape, (what, a.shape, b.shape)
    if b.dtype == torch.bfloat16:
        a = a.to(torch.bfloat16)
        diff = (a.contiguous().view(torch.int16).int() - b.contiguous().view(torch.int16).int()).abs()
        # sign-magnitude bit patterns: +-0 and tiny values of opposite sign differ by a large integer but not in value
        bad = (diff > 1) & ((a.float() - b.float()).abs() > 1e-30)
        assert not bool(bad.any()) and float((diff > 0).float().mean()) < frac, (what, int(diff.max()), float((diff > 0).float().mean()))
    else:
        assert rel_l2(a.float(), b.float()) < 2e-6, (what, rel_l2(a.float(), b.float()))


def _check_latents(got, want, what):
    if want.dtype == torch.bfloat16:
        _ulp_close(got, want, what)
    else:
        assert got.dtype == want.dtype
        assert rel_l2(got, want) < 2e-6, (what, rel_l2(got, want))


class _Replay:
    dtype = torch.bfloat16

    def __init__(self, d, named, cfg):
        from types import SimpleNamespace
        self.d, self.named, self.calls, self.config = d, named, 0, SimpleNamespace(**cfg)
        self.device = torch.device(DEV)

    def to(self, *a, **k):
        return self

    def _texts(self, i, texts):
        want = GL.split_names(self.d[f"text_{i}"])
        assert len(want) == len(texts), (i, want, len(texts))
        for r, w in enumerate(want):
            ref = self.named[w]
            assert torch.equal(texts[r].reshape(-1, texts[r].shape[-1])[: ref.shape[0]].to(ref.dtype), ref), (i, r, w)


# ----------------------------------------------------------------------------------------------------------------
class ReplayWan(_Replay):
    def forward_passes(self, latents, cond, text, image, timestep, out=None):
        i, d = self.calls, self.d
        hs = d[f"hs_{i}"]
        n_pass = len(cond)
        n_samples = hs.shape[0] // n_pass
        b = self.sample  # rows of the reference batch are pass-major: row = pass * n_samples + sample
        rows = [p * n_samples + b for p in range(n_pass)]
        x = torch.stack([torch.cat([latents[p].reshape(16, *cond[p].shape[-3:]), cond[p]], dim=0) for p in range(n_pass)])
        _ulp_close(x, hs[rows], f"model input of call {i} sample {b}")
        assert int(timestep) == int(d[f"t_{i}"][0])
        want = GL.split_names(d[f"text_{i}"])
        for p in range(n_pass):
            ref = self.named[want[rows[p]]]
            assert torch.equal(text[p].to(ref.dtype), ref), (i, p, want)
        noise = d[f"noise_{i}"][rows].contiguous()
        self.sample += 1
        if self.sample == n_samples:
            self.sample, self.calls = 0, i + 1
        return noise


@pytest.mark.parametrize("name", GL.names("wan"))
def test_wan_pipeline_matches_reference(name):
    from alg_b200.schedulers import UniPCMultistepScheduler
    from oracle.stub_vae import ArithVAE, StubImageEncoder, StubImageProcessor, StubTextEncoder, StubTokenizer
    from pipeline_wan_image2video_lowpass import WanImageToVideoPipeline
    meta, d = GL.load("wan", name, DEV)
    enc = StubTextEncoder(seed=meta["text_encoder_seed"])
    named = {"p": d["pos"][0], "n": d["neg"][0]}
    kw = dict(prompt_embeds=d["pos"], negative_prompt_embeds=d["neg"])
    if meta["prompts"]:
        kw = dict(prompt=meta["prompts"][0], negative_prompt=meta["prompts"][1], max_sequence_length=32)
        tok = StubTokenizer()

        def embed(p):
            t = tok(p, max_length=32)
            h = enc(t.input_ids, t.attention_mask).last_hidden_state
            h[0, int(t.attention_mask.sum()):] = 0
            return h[0].to(DEV)
        named = {"p0": embed(meta["prompts"][0]), "n0": embed(meta["prompts"][1])}
    replay = ReplayWan(d, named, dict(patch_size=(1, 2, 2), text_dim=64, image_dim=64))
    replay.sample = 0
    pipe = WanImageToVideoPipeline(tokenizer=StubTokenizer(), text_encoder=enc, image_encoder=StubImageEncoder(d["image_table"]),
                                   image_processor=StubImageProcessor(), transformer=replay, vae=ArithVAE("wan"),
                                   scheduler=UniPCMultistepScheduler(flow_shift=meta["flow_shift"])).to(DEV)
    pipe.set_progress_bar_config(disable=True)
    B = meta["batch"]
    gen = [torch.Generator().manual_seed(meta["generator_seed"] + b) for b in range(B)] if B > 1 else \
        torch.Generator().manual_seed(meta["generator_seed"])
    seen = []
    out = pipe(image=d["image"], last_image=d.get("last_image"), height=meta["height"], width=meta["width"],
               num_frames=meta["num_frames"], num_inference_steps=meta["steps"], guidance_scale=meta["guidance"],
               num_videos_per_prompt=meta["n_videos"], generator=gen, output_type="latent",
               callback_on_step_end=lambda p, i, t, k: seen.append(k["latents"].clone()) or {}, **kw, **meta["alg"])
    assert replay.calls == meta["n_calls"] and len(seen) == meta["steps"]
    for i, lat in enumerate(seen):
        _check_latents(lat, d[f"lat_{i}"], f"latents after step {i}")
    assert torch.equal(out.frames, seen[-1])


def test_wan_pipeline_batch_quirks():
    """What the reference cannot do, the drop-in refuses with a message (tests/golden/loop_quirks.json)."""
    from alg_b200.schedulers import UniPCMultistepScheduler
    from oracle.stub_vae import ArithVAE, StubImageEncoder, StubImageProcessor, StubTextEncoder, StubTokenizer
    from pipeline_wan_image2video_lowpass import WanImageToVideoPipeline
    meta, d = GL.load("wan", "latent_down_up", DEV)
    replay = ReplayWan(d, {}, dict(patch_size=(1, 2, 2), text_dim=64, image_dim=64))
    replay.sample = 0
    pipe = WanImageToVideoPipeline(tokenizer=StubTokenizer(), text_encoder=StubTextEncoder(), image_encoder=StubImageEncoder(d["image_table"]),
                                   image_processor=StubImageProcessor(), transformer=replay, vae=ArithVAE("wan"),
                                   scheduler=UniPCMultistepScheduler(flow_shift=5.0)).to(DEV)
    pipe.set_progress_bar_config(disable=True)
    common = dict(image=d["image"], height=meta["height"], width=meta["width"], num_frames=meta["num_frames"], num_inference_steps=2,
                  output_type="latent", max_sequence_length=32)
    with pytest.raises(ValueError, match="list of 2 prompts"):
        pipe(prompt=["a", "b"], negative_prompt=["c", "d"], **common, **meta["alg"])
    with pytest.raises(RuntimeError, match="three-pass ALG steps support one sample"):
        pipe(prompt="a red bus", negative_prompt="blurry", num_videos_per_prompt=2, **common, **meta["alg"])


# ----------------------------------------------------------------------------------------------------------------
class ReplayCog(_Replay):
    def forward_passes(self, latents, cond, text, timestep, rope):
        i, d = self.calls, self.d
        x = torch.stack([torch.cat([latents[p], cond[p]], dim=1) for p in range(len(cond))])  # [pass, F, 32, H, W]
        _ulp_close(x, d[f"hs_{i}"], f"model input of call {i}") if d[f"hs_{i}"].dtype == torch.bfloat16 else \
            _ulp_close(x.float(), d[f"hs_{i}"], f"model input of call {i}")
        assert int(timestep) == int(d[f"t_{i}"][0])
        self._texts(i, text)
        _ulp_close(rope[0].float(), d["rope_cos"], "rope cos")
        _ulp_close(rope[1].float(), d["rope_sin"], "rope sin")
        self.calls += 1
        return d[f"noise_{i}"]


@pytest.mark.parametrize("name", GL.names("cog"))
def test_cog_pipeline_matches_reference(name):
    from alg_b200.schedulers import CogVideoXDDIMScheduler, CogVideoXDPMScheduler
    from oracle.stub_vae import ArithVAE
    from pipeline_cogvideox_image2video_lowpass import CogVideoXImageToVideoPipeline
    meta, d = GL.load("cog", name, DEV)
    dt = getattr(torch, meta["dtype"])
    c = meta["cfg"]
    replay = ReplayCog(d, {"p": d["pos"][0], "n": d["neg"][0]},
                       dict(patch_size=2, patch_size_t=None, sample_width=c["sample_width"], sample_height=c["sample_height"],
                            sample_frames=c["sample_frames"], in_channels=32, attention_head_dim=c["attention_head_dim"],
                            use_rotary_positional_embeddings=True, ofs_embed_dim=None, text_embed_dim=c["text_embed_dim"]))
    replay.dtype = dt
    sched = CogVideoXDPMScheduler() if meta["dpm"] else CogVideoXDDIMScheduler()
    pipe = CogVideoXImageToVideoPipeline(tokenizer=None, text_encoder=None, vae=ArithVAE("cog", dtype=dt), transformer=replay,
                                         scheduler=sched).to(DEV)
    pipe.set_progress_bar_config(disable=True)
    seen = []
    out = pipe(image=d["image"], prompt_embeds=d["pos"], negative_prompt_embeds=d["neg"], height=meta["height"], width=meta["width"],
               num_frames=meta["num_frames"], num_inference_steps=meta["steps"], guidance_scale=meta["guidance"],
               use_dynamic_cfg=meta["use_dynamic_cfg"], generator=torch.Generator().manual_seed(meta["generator_seed"]),
               output_type="latent", callback_on_step_end=lambda p, i, t, k: seen.append(k["latents"].clone()) or {}, **meta["alg"])
    assert replay.calls == meta["n_calls"] and len(seen) == meta["steps"]
    for i, lat in enumerate(seen):
        _check_latents(lat, d[f"lat_{i}"], f"latents after step {i}")
    assert torch.equal(out.frames, seen[-1])


# ----------------------------------------------------------------------------------------------------------------
class ReplayHunyuan(_Replay):
    def forward_pass(self, latents, first, emb, pooled, t_model, guidance, out=None):
        i, d = self.calls, self.d
        hs = d[f"hs_{i}"]
        p = self.p
        x = torch.cat([first.reshape(16, 1, *latents.shape[-2:]), latents[:, 1:]], dim=1)
        _ulp_close(x, hs[p], f"model input of call {i} pass {p}")
        assert abs(t_model - float(d[f"t_{i}"][p])) == 0.0 and abs(guidance * 1.0 - float(d[f"guidance_{i}"][0])) == 0.0, (t_model, guidance)
        wt, wp = GL.split_names(d[f"text_{i}"]), GL.split_names(d[f"pooled_{i}"])
        n_valid = int(d[f"mask_{i}"][p].sum())
        assert emb.shape[0] == n_valid and torch.equal(emb, self.named[wt[p]][:n_valid]) and torch.equal(pooled, self.named_p[wp[p]])
        out.copy_(d[f"noise_{i}"][p])
        self.p += 1
        if self.p == hs.shape[0]:
            self.p, self.calls = 0, i + 1
        return out


@pytest.mark.parametrize("name", GL.names("hunyuan"))
def test_hunyuan_pipeline_matches_reference(name):
    from alg_b200.schedulers import FlowMatchEulerDiscreteScheduler
    from oracle.stub_vae import ArithVAE
    from pipeline_hunyuan_video_image2video_lowpass import HunyuanVideoImageToVideoPipeline
    meta, d = GL.load("hunyuan", name, DEV)
    replay = ReplayHunyuan(d, {"p": d["pos"][0], "n": d["neg"][0]},
                           dict(in_channels=16, guidance_embeds=True, patch_size=2, patch_size_t=1, image_condition_type="token_replace",
                                text_embed_dim=64, pooled_projection_dim=32))
    replay.named_p, replay.p = {"p": d["pooled_pos"][0], "n": d["pooled_neg"][0]}, 0
    pipe = HunyuanVideoImageToVideoPipeline(text_encoder=None, tokenizer=None, transformer=replay, vae=ArithVAE("hunyuan"),
                                            scheduler=FlowMatchEulerDiscreteScheduler(shift=meta["shift"]), text_encoder_2=None,
                                            tokenizer_2=None, image_processor=None).to(DEV)
    pipe.set_progress_bar_config(disable=True)
    kw = dict(prompt_embeds=d["pos"], pooled_prompt_embeds=d["pooled_pos"], prompt_attention_mask=d["mask_pos"])
    if meta["true_cfg"] > 1:
        kw.update(negative_prompt_embeds=d["neg"], negative_pooled_prompt_embeds=d["pooled_neg"], negative_prompt_attention_mask=d["mask_neg"])
    seen = []
    out = pipe(image=d["image"], height=meta["height"], width=meta["width"], num_frames=meta["num_frames"],
               num_inference_steps=meta["steps"], guidance_scale=meta["guidance"], true_cfg_scale=meta["true_cfg"],
               generator=torch.Generator().manual_seed(meta["generator_seed"]), output_type="latent",
               callback_on_step_end=lambda p, i, t, k: seen.append(k["latents"].clone()) or {}, **kw, **meta["alg"], **meta.get("extra", {}))
    assert replay.calls == meta["n_calls"] and len(seen) == meta["steps"]
    for i, lat in enumerate(seen):
        _check_latents(lat, d[f"lat_{i}"], f"latents after step {i}")
    assert torch.equal(out.frames, seen[-1])
