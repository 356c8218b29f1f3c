"""PyTorch restatement of ``WanTransformer3DModel.forward`` and of the Wan ALG loop.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Runs on CPU (fp32 or bf16) and,
inside ``-m gpu`` tests only, on the box's GPU as the eager-PyTorch checker.

PARITY UNPINNED for the DiT: the class lives in diffusers@be2fb77
(requirements.txt:13; ``models/transformers/transformer_wan.py``), absent here.
Restated from its published forward, dtype casts
included; anchored on the reference's call site wan:910-917.

The loop (``denoise_loop``) restates first-party code and follows
pipeline_wan_image2video_lowpass.py:839-946 branch for branch.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch
import torch.nn.functional as F


@dataclass
class WanConfig:
    patch_size: tuple = (1, 2, 2)
    num_attention_heads: int = 40
    attention_head_dim: int = 128
    in_channels: int = 36
    out_channels: int = 16
    text_dim: int = 4096
    freq_dim: int = 256
    ffn_dim: int = 13824
    num_layers: int = 40
    eps: float = 1e-6
    image_dim: int = 1280
    rope_max_seq_len: int = 1024
    text_len: int = 512  # hard-coded in the attention processor

    @property
    def dim(self):
        return self.num_attention_heads * self.attention_head_dim


def tiny_config(**kw):
    base = dict(num_attention_heads=2, attention_head_dim=128, text_dim=64, freq_dim=256, ffn_dim=512,
                num_layers=2, image_dim=64, text_len=32)
    base.update(kw)
    return WanConfig(**base)


FP32_KEYS = ("time_embedder", "scale_shift_table", "norm1", "norm2", "norm3")  # diffusers _keep_in_fp32_modules


def make_weights(cfg: WanConfig, seed: int = 0, device="cpu", dtype=torch.bfloat16, std: float = 0.02):
    """Seeded synthetic state_dict with diffusers' parameter names (no checkpoints offline)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    d, sd = cfg.dim, {}

    def lin(name, out_f, in_f, bias=True):
        sd[name + ".weight"] = torch.randn(out_f, in_f, generator=g) * std
        if bias:
            sd[name + ".bias"] = torch.randn(out_f, generator=g) * std

    pt, ph, pw = cfg.patch_size
    sd["patch_embedding.weight"] = torch.randn(d, cfg.in_channels, pt, ph, pw, generator=g) * std * 4
    sd["patch_embedding.bias"] = torch.randn(d, generator=g) * std
    lin("condition_embedder.time_embedder.linear_1", d, cfg.freq_dim)
    lin("condition_embedder.time_embedder.linear_2", d, d)
    lin("condition_embedder.time_proj", 6 * d, d)
    lin("condition_embedder.text_embedder.linear_1", d, cfg.text_dim)
    lin("condition_embedder.text_embedder.linear_2", d, d)
    ie = "condition_embedder.image_embedder."
    sd[ie + "norm1.weight"] = 1 + 0.1 * torch.randn(cfg.image_dim, generator=g)
    sd[ie + "norm1.bias"] = 0.1 * torch.randn(cfg.image_dim, generator=g)
    lin(ie + "ff.net.0.proj", cfg.image_dim, cfg.image_dim)
    lin(ie + "ff.net.2", d, cfg.image_dim)
    sd[ie + "norm2.weight"] = 1 + 0.1 * torch.randn(d, generator=g)
    sd[ie + "norm2.bias"] = 0.1 * torch.randn(d, generator=g)
    for i in range(cfg.num_layers):
        p = f"blocks.{i}."
        sd[p + "scale_shift_table"] = torch.randn(1, 6, d, generator=g) / d ** 0.5
        for a in ("attn1", "attn2"):
            for n in ("to_q", "to_k", "to_v", "to_out.0"):
                lin(p + f"{a}.{n}", d, d)
            sd[p + f"{a}.norm_q.weight"] = 1 + 0.1 * torch.randn(d, generator=g)
            sd[p + f"{a}.norm_k.weight"] = 1 + 0.1 * torch.randn(d, generator=g)
        lin(p + "attn2.add_k_proj", d, d)
        lin(p + "attn2.add_v_proj", d, d)
        sd[p + "attn2.norm_added_k.weight"] = 1 + 0.1 * torch.randn(d, generator=g)
        sd[p + "norm2.weight"] = 1 + 0.1 * torch.randn(d, generator=g)
        sd[p + "norm2.bias"] = 0.1 * torch.randn(d, generator=g)
        lin(p + "ffn.net.0.proj", cfg.ffn_dim, d)
        lin(p + "ffn.net.2", d, cfg.ffn_dim)
    sd["scale_shift_table"] = torch.randn(1, 2, d, generator=g) / d ** 0.5
    lin("proj_out", cfg.out_channels * pt * ph * pw, d)
    out = {}
    for k, v in sd.items():
        keep32 = any(s in k for s in FP32_KEYS)
        out[k] = v.to(device=device, dtype=torch.float32 if keep32 else dtype)
    return out


# ----------------------------------------------------------------------------
# building blocks (diffusers semantics)
# ----------------------------------------------------------------------------
def fp32_layer_norm(x, weight, bias, eps):
    return F.layer_norm(x.float(), (x.shape[-1],), None if weight is None else weight.float(),
                        None if bias is None else bias.float(), eps).to(x.dtype)


def rms_norm(x, weight, eps):
    var = x.to(torch.float32).pow(2).mean(-1, keepdim=True)
    h = x * torch.rsqrt(var + eps)
    if weight.dtype in (torch.float16, torch.bfloat16):
        h = h.to(weight.dtype)
    return h * weight


def rope_freqs(cfg: WanConfig, ppf, pph, ppw, device):
    """complex128 [1, 1, N, head_dim/2] (WanRotaryPosEmbed)."""
    hd = cfg.attention_head_dim
    h_dim = w_dim = 2 * (hd // 6)
    t_dim = hd - h_dim - w_dim
    parts = []
    for dim in (t_dim, h_dim, w_dim):
        f = 1.0 / (10000.0 ** (torch.arange(0, dim, 2, dtype=torch.float64)[: dim // 2] / dim))
        ang = torch.outer(torch.arange(cfg.rope_max_seq_len, dtype=torch.float64), f)
        parts.append(torch.polar(torch.ones_like(ang), ang))
    ff = parts[0][:ppf].view(ppf, 1, 1, -1).expand(ppf, pph, ppw, -1)
    fh = parts[1][:pph].view(1, pph, 1, -1).expand(ppf, pph, ppw, -1)
    fw = parts[2][:ppw].view(1, 1, ppw, -1).expand(ppf, pph, ppw, -1)
    return torch.cat([ff, fh, fw], dim=-1).reshape(1, 1, ppf * pph * ppw, -1).to(device)


def apply_rope(x, freqs):
    xr = torch.view_as_complex(x.to(torch.float64).unflatten(3, (-1, 2)))
    return torch.view_as_real(xr * freqs).flatten(3, 4).type_as(x)


def timestep_embedding(t, dim):
    half = dim // 2
    exponent = -math.log(10000) * torch.arange(0, half, dtype=torch.float32, device=t.device) / half
    emb = t[:, None].float() * torch.exp(exponent)[None, :]
    return torch.cat([torch.cos(emb), torch.sin(emb)], dim=-1)  # flip_sin_to_cos


def sdpa(q, k, v):
    if q.device.type == "cpu" and q.dtype == torch.bfloat16:  # keep CPU oracle exact-ish and deterministic
        return F.scaled_dot_product_attention(q.float(), k.float(), v.float()).to(q.dtype)
    return F.scaled_dot_product_attention(q, k, v)


def attention(sd, p, cfg, x, ctx=None, rotary=None):
    H = cfg.num_attention_heads
    ctx_img = None
    if ctx is not None and (p + "add_k_proj.weight") in sd:
        n_img = ctx.shape[1] - cfg.text_len
        ctx_img, ctx = ctx[:, :n_img], ctx[:, n_img:]
    kv_in = x if ctx is None else ctx
    q = F.linear(x, sd[p + "to_q.weight"], sd[p + "to_q.bias"])
    k = F.linear(kv_in, sd[p + "to_k.weight"], sd[p + "to_k.bias"])
    v = F.linear(kv_in, sd[p + "to_v.weight"], sd[p + "to_v.bias"])
    q = rms_norm(q, sd[p + "norm_q.weight"], cfg.eps)
    k = rms_norm(k, sd[p + "norm_k.weight"], cfg.eps)
    q, k, v = (t.unflatten(2, (H, -1)).transpose(1, 2) for t in (q, k, v))
    if rotary is not None:
        q, k = apply_rope(q, rotary), apply_rope(k, rotary)
    out_img = None
    if ctx_img is not None:
        ki = F.linear(ctx_img, sd[p + "add_k_proj.weight"], sd[p + "add_k_proj.bias"])
        ki = rms_norm(ki, sd[p + "norm_added_k.weight"], cfg.eps)
        vi = F.linear(ctx_img, sd[p + "add_v_proj.weight"], sd[p + "add_v_proj.bias"])
        ki, vi = (t.unflatten(2, (H, -1)).transpose(1, 2) for t in (ki, vi))
        out_img = sdpa(q, ki, vi).transpose(1, 2).flatten(2, 3).type_as(q)
    out = sdpa(q, k, v).transpose(1, 2).flatten(2, 3).type_as(q)
    if out_img is not None:
        out = out + out_img
    return F.linear(out, sd[p + "to_out.0.weight"], sd[p + "to_out.0.bias"])


def block(sd, i, cfg, x, ctx, temb6, rotary):
    p = f"blocks.{i}."
    shift, scale, gate, c_shift, c_scale, c_gate = (sd[p + "scale_shift_table"] + temb6.float()).chunk(6, dim=1)
    h = (fp32_layer_norm(x.float(), None, None, cfg.eps) * (1 + scale) + shift).type_as(x)
    a = attention(sd, p + "attn1.", cfg, h, None, rotary)
    x = (x.float() + a * gate).type_as(x)
    h = fp32_layer_norm(x.float(), sd[p + "norm2.weight"], sd[p + "norm2.bias"], cfg.eps).type_as(x)
    a = attention(sd, p + "attn2.", cfg, h, ctx, None)
    x = x + a
    h = (fp32_layer_norm(x.float(), None, None, cfg.eps) * (1 + c_scale) + c_shift).type_as(x)
    f = F.linear(h, sd[p + "ffn.net.0.proj.weight"], sd[p + "ffn.net.0.proj.bias"])
    f = F.gelu(f, approximate="tanh")
    f = F.linear(f, sd[p + "ffn.net.2.weight"], sd[p + "ffn.net.2.bias"])
    x = (x.float() + f.float() * c_gate).type_as(x)
    return x


def condition_embedder(sd, cfg, timestep, text, image):
    p = "condition_embedder."
    t = timestep_embedding(timestep, cfg.freq_dim)
    w1 = sd[p + "time_embedder.linear_1.weight"]
    t = t.to(w1.dtype)
    temb = F.linear(F.silu(F.linear(t, w1, sd[p + "time_embedder.linear_1.bias"])),
                    sd[p + "time_embedder.linear_2.weight"], sd[p + "time_embedder.linear_2.bias"]).type_as(text)
    tproj = F.linear(F.silu(temb), sd[p + "time_proj.weight"], sd[p + "time_proj.bias"])
    text = F.linear(text, sd[p + "text_embedder.linear_1.weight"], sd[p + "text_embedder.linear_1.bias"])
    text = F.gelu(text, approximate="tanh")
    text = F.linear(text, sd[p + "text_embedder.linear_2.weight"], sd[p + "text_embedder.linear_2.bias"])
    if image is not None:
        q = p + "image_embedder."
        image = fp32_layer_norm(image, sd[q + "norm1.weight"], sd[q + "norm1.bias"], 1e-5)
        image = F.gelu(F.linear(image, sd[q + "ff.net.0.proj.weight"], sd[q + "ff.net.0.proj.bias"]))
        image = F.linear(image, sd[q + "ff.net.2.weight"], sd[q + "ff.net.2.bias"])
        image = fp32_layer_norm(image, sd[q + "norm2.weight"], sd[q + "norm2.bias"], 1e-5)
    return temb, tproj, text, image


def forward(sd, cfg: WanConfig, hidden, timestep, text, image, return_intermediates=False):
    """hidden [B, 36, T, H, W] (model dtype); timestep int64 [B]; text [B, 512, text_dim]; image [B, 257, image_dim]."""
    B, C, T, Hh, Ww = hidden.shape
    pt, ph, pw = cfg.patch_size
    ppf, pph, ppw = T // pt, Hh // ph, Ww // pw
    rotary = rope_freqs(cfg, ppf, pph, ppw, hidden.device)
    x = F.conv3d(hidden, sd["patch_embedding.weight"], sd["patch_embedding.bias"], stride=cfg.patch_size)
    x = x.flatten(2).transpose(1, 2)
    temb, tproj, text, image = condition_embedder(sd, cfg, timestep, text, image)
    tproj = tproj.unflatten(1, (6, -1))
    ctx = torch.cat([image, text], dim=1) if image is not None else text
    inter = {"patch": x, "temb": temb, "tproj": tproj, "ctx": ctx}
    for i in range(cfg.num_layers):
        x = block(sd, i, cfg, x, ctx, tproj, rotary)
        if return_intermediates:
            inter[f"block{i}"] = x
    shift, scale = (sd["scale_shift_table"] + temb.unsqueeze(1)).chunk(2, dim=1)
    x = (fp32_layer_norm(x.float(), None, None, cfg.eps) * (1 + scale) + shift).type_as(x)
    x = F.linear(x, sd["proj_out.weight"], sd["proj_out.bias"])
    x = x.reshape(B, ppf, pph, ppw, pt, ph, pw, -1).permute(0, 7, 1, 4, 2, 5, 3, 6)
    out = x.flatten(6, 7).flatten(4, 5).flatten(2, 3)
    return (out, inter) if return_intermediates else out


# ----------------------------------------------------------------------------
# the ALG denoise loop, wan:839-946
# ----------------------------------------------------------------------------
def denoise_loop(transformer, scheduler, latents, condition, prompt_embeds, negative_prompt_embeds, image_embeds,
                 num_inference_steps, guidance_scale, alg, lp_filter, get_lp_strength, dtype=torch.bfloat16,
                 on_step=None, teacher=None, prepare_lp=None):
    """``transformer(x, t, text, img) -> noise``; ``lp_filter(cond, type, sigma, k, f) -> tensor`` (in-latent ALG), or
    ``prepare_lp(type, sigma, k, f) -> lp condition`` for the full wan:451-559 (``oracle/prepare_lp_oracle.wan_prepare_lp``:
    pixel-space filtering + VAE encode + mask rebuild).

    ``alg`` holds the 14 ALG kwargs (wan:612-633).  ``on_step(i, t, latents, noise_pred)`` observes every step.
    ``teacher[i]`` (optional) replaces ``latents`` before step i (teacher-forced parity).
    """
    scheduler.set_timesteps(num_inference_steps)
    for i, t in enumerate(scheduler.timesteps):
        if teacher is not None:
            latents = teacher[i]
        if guidance_scale > 1 and alg.get("use_low_pass_guidance", False):
            s = get_lp_strength(i, num_inference_steps, alg["lp_strength_schedule_type"],
                                alg["schedule_interval_start_time"], alg["schedule_interval_end_time"],
                                alg["schedule_linear_start_weight"], alg["schedule_linear_end_weight"],
                                alg["schedule_linear_end_time"], alg["schedule_exp_decay_rate"])
            sigma = alg["lp_blur_sigma"] * s
            k = alg["lp_blur_kernel_size"] * s if alg["schedule_blur_kernel_size"] else alg["lp_blur_kernel_size"]
            f = 1.0 - (1.0 - alg["lp_resize_factor"]) * s
            if prepare_lp is not None:
                lp = prepare_lp(alg["lp_filter_type"], sigma, k, f)
            else:
                lp = lp_filter(condition, alg["lp_filter_type"], sigma, k, f).to(condition.dtype)
            if s == 0.0:
                x = torch.cat([torch.cat([latents] * 2), torch.cat([condition, condition])], dim=1).to(dtype)
                text = torch.cat([negative_prompt_embeds, prompt_embeds])
            else:
                x = torch.cat([torch.cat([latents] * 3), torch.cat([condition, lp, lp])], dim=1).to(dtype)
                text = torch.cat([negative_prompt_embeds, negative_prompt_embeds, prompt_embeds])
        elif guidance_scale > 1:
            x = torch.cat([torch.cat([latents] * 2), torch.cat([condition, condition])], dim=1).to(dtype)
            text = torch.cat([negative_prompt_embeds, prompt_embeds])
        else:
            raise NameError("latent_model_input")  # quirk q2: the reference has no guidance_scale <= 1 branch
        timestep = t.expand(x.shape[0])
        img = image_embeds.repeat(x.shape[0], 1, 1) if image_embeds.shape[0] != x.shape[0] else image_embeds
        noise_pred = transformer(x, timestep, text, img)
        if noise_pred.shape[0] == 3:
            u0, u, tx = noise_pred.chunk(3)
            noise = u0 + guidance_scale * (tx - u)
        else:
            u, tx = noise_pred.chunk(2)
            noise = u + guidance_scale * (tx - u)
        latents = scheduler.step(noise, latents)
        if on_step is not None:
            on_step(i, t, latents, noise_pred)
    return latents
