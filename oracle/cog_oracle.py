"""PyTorch restatement of ``CogVideoXTransformer3DModel.forward`` and of the CogVideoX ALG loop.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Runs on CPU (fp32 or bf16) and, inside ``-m gpu`` tests only, on the
box's GPU as the eager-PyTorch checker.

PARITY UNPINNED for the DiT: the class lives in diffusers@be2fb77 (requirements.txt:13;
``models/transformers/cogvideox_transformer_3d.py``, ``models/embeddings.py``, ``models/normalization.py``), absent
here.  Restated from its published forward op by op, so that in bf16 every eager rounding
(LayerNorm output, ``1 + scale``, products, sums) happens where eager PyTorch would place it; anchored on the
reference's call site cog:1082-1090 and rotary preparation cog:542-584.

The loop (``denoise_loop``) restates first-party code: pipeline_cogvideox_image2video_lowpass.py:1005-1140.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional

import torch
import torch.nn.functional as F


@dataclass
class CogConfig:
    num_attention_heads: int = 48
    attention_head_dim: int = 64
    in_channels: int = 32
    out_channels: int = 16
    time_embed_dim: int = 512
    text_embed_dim: int = 4096
    num_layers: int = 42
    patch_size: int = 2
    sample_width: int = 90
    sample_height: int = 60
    sample_frames: int = 49
    temporal_compression_ratio: int = 4
    max_text_seq_length: int = 226
    ff_mult: int = 4
    norm_eps: float = 1e-5
    spatial_interpolation_scale: float = 1.875
    temporal_interpolation_scale: float = 1.0

    @property
    def dim(self):
        return self.num_attention_heads * self.attention_head_dim


def tiny_config(**kw):
    base = dict(num_attention_heads=2, attention_head_dim=64, time_embed_dim=64, text_embed_dim=64, num_layers=2,
                sample_width=12, sample_height=8, sample_frames=9, max_text_seq_length=16)
    base.update(kw)
    return CogConfig(**base)


# ----------------------------------------------------------------------------
# embeddings.py restatements (host tables)
# ----------------------------------------------------------------------------
def _sincos_1d(embed_dim, pos):
    omega = torch.arange(embed_dim // 2, dtype=torch.float64) / (embed_dim / 2.0)
    omega = 1.0 / 10000 ** omega
    out = torch.outer(pos.reshape(-1).to(torch.float64), omega)
    return torch.cat([torch.sin(out), torch.cos(out)], dim=1)


def sincos_pos_embed_3d(embed_dim, width, height, frames, spatial_scale, temporal_scale):
    """get_3d_sincos_pos_embed(output_type="pt"): [frames, height*width, embed_dim], temporal quarter first."""
    d_sp, d_t = 3 * embed_dim // 4, embed_dim // 4
    gh = torch.arange(height, dtype=torch.float32) / spatial_scale
    gw = torch.arange(width, dtype=torch.float32) / spatial_scale
    gw2, gh2 = torch.meshgrid(gw, gh, indexing="xy")  # w varies fastest
    emb_h = _sincos_1d(d_sp // 2, gw2)  # diffusers feeds grid[0] (the w mesh) first
    emb_w = _sincos_1d(d_sp // 2, gh2)
    spatial = torch.cat([emb_h, emb_w], dim=1)  # [H*W, d_sp]
    temporal = _sincos_1d(d_t, torch.arange(frames, dtype=torch.float32) / temporal_scale)  # [T, d_t]
    spatial = spatial[None].repeat_interleave(frames, dim=0)
    temporal = temporal[:, None].repeat_interleave(height * width, dim=1)
    return torch.cat([temporal, spatial], dim=-1).float()


def joint_pos_embedding(cfg: CogConfig, height, width, latent_frames):
    """CogVideoXPatchEmbed._get_positional_embeddings: zeros over the text tokens, 3-D sincos over the patches."""
    ph, pw = height // cfg.patch_size, width // cfg.patch_size
    pos = sincos_pos_embed_3d(cfg.dim, pw, ph, latent_frames, cfg.spatial_interpolation_scale,
                              cfg.temporal_interpolation_scale).flatten(0, 1)
    joint = torch.zeros(1, cfg.max_text_seq_length + pos.shape[0], cfg.dim)
    joint[:, cfg.max_text_seq_length:] = pos
    return joint


def crop_region(src, tgt_width, tgt_height):
    h, w = src
    if h / w > tgt_height / tgt_width:
        rh, rw = tgt_height, int(round(tgt_height / h * w))
    else:
        rw, rh = tgt_width, int(round(tgt_width / w * h))
    top, left = int(round((tgt_height - rh) / 2.0)), int(round((tgt_width - rw) / 2.0))
    return (top, left), (top + rh, left + rw)


def _rope_1d(dim, pos, theta=10000.0):
    freqs = 1.0 / (theta ** (torch.arange(0, dim, 2, dtype=torch.float32)[: dim // 2] / dim))
    f = torch.outer(pos, freqs)
    return f.cos().repeat_interleave(2, dim=1).float(), f.sin().repeat_interleave(2, dim=1).float()


def rotary_tables(cfg: CogConfig, grid_h, grid_w, frames):
    """cog:542-584 (CogVideoX 1.0 branch) + get_3d_rotary_pos_embed(linspace grid): cos, sin fp32 [T*h*w, head_dim]."""
    start, stop = crop_region((grid_h, grid_w), cfg.sample_width // cfg.patch_size, cfg.sample_height // cfg.patch_size)
    gh = torch.linspace(start[0], stop[0] * (grid_h - 1) / grid_h, grid_h, dtype=torch.float32)
    gw = torch.linspace(start[1], stop[1] * (grid_w - 1) / grid_w, grid_w, dtype=torch.float32)
    gt = torch.arange(frames, dtype=torch.float32)
    hd = cfg.attention_head_dim
    d_t, d_h, d_w = hd // 4, hd // 8 * 3, hd // 8 * 3
    out = []
    for k in range(2):
        t = _rope_1d(d_t, gt)[k][:, None, None, :].expand(-1, grid_h, grid_w, -1)
        h = _rope_1d(d_h, gh)[k][None, :, None, :].expand(frames, -1, grid_w, -1)
        w = _rope_1d(d_w, gw)[k][None, None, :, :].expand(frames, grid_h, -1, -1)
        out.append(torch.cat([t, h, w], dim=-1).reshape(frames * grid_h * grid_w, -1).contiguous())
    return out[0], out[1]


def apply_rotary(x, cos, sin):
    """diffusers apply_rotary_emb(use_real=True, use_real_unbind_dim=-1) on [B, H, S, D]."""
    cos, sin = cos[None, None].to(x.device), sin[None, None].to(x.device)
    xr, xi = x.reshape(*x.shape[:-1], -1, 2).unbind(-1)
    rot = torch.stack([-xi, xr], dim=-1).flatten(3)
    return (x.float() * cos + rot.float() * sin).to(x.dtype)


# ----------------------------------------------------------------------------
# weights
# ----------------------------------------------------------------------------
def parameter_shapes(cfg: CogConfig, latent_frames=None):
    d, te = cfg.dim, cfg.time_embed_dim
    s = {}

    def lin(name, o, i):
        s[name + ".weight"] = (o, i)
        s[name + ".bias"] = (o,)

    def ln(name, n):
        s[name + ".weight"] = (n,)
        s[name + ".bias"] = (n,)

    s["patch_embed.proj.weight"] = (d, cfg.in_channels, cfg.patch_size, cfg.patch_size)
    s["patch_embed.proj.bias"] = (d,)
    lin("patch_embed.text_proj", d, cfg.text_embed_dim)
    lf = latent_frames or (cfg.sample_frames - 1) // cfg.temporal_compression_ratio + 1
    n_patch = lf * (cfg.sample_height // cfg.patch_size) * (cfg.sample_width // cfg.patch_size)
    s["patch_embed.pos_embedding"] = (1, cfg.max_text_seq_length + n_patch, d)
    lin("time_embedding.linear_1", te, d)
    lin("time_embedding.linear_2", te, te)
    for i in range(cfg.num_layers):
        p = f"transformer_blocks.{i}."
        for n in ("norm1", "norm2"):
            lin(p + n + ".linear", 6 * d, te)
            ln(p + n + ".norm", d)
        for n in ("to_q", "to_k", "to_v", "to_out.0"):
            lin(p + "attn1." + n, d, d)
        ln(p + "attn1.norm_q", cfg.attention_head_dim)
        ln(p + "attn1.norm_k", cfg.attention_head_dim)
        lin(p + "ff.net.0.proj", cfg.ff_mult * d, d)
        lin(p + "ff.net.2", d, cfg.ff_mult * d)
    ln("norm_final", d)
    lin("norm_out.linear", 2 * d, te)
    ln("norm_out.norm", d)
    lin("proj_out", cfg.patch_size * cfg.patch_size * cfg.out_channels, d)
    return s


def make_weights(cfg: CogConfig, seed=0, device="cpu", dtype=torch.bfloat16, std=0.02):
    """Seeded synthetic state_dict with diffusers' parameter names (no checkpoints offline)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    sd = {}
    for name, shape in parameter_shapes(cfg).items():
        if "norm" in name and name.endswith(".weight") and "linear" not in name:
            w = 1 + 0.1 * torch.randn(shape, generator=g)
        elif "norm" in name and name.endswith(".bias") and "linear" not in name:
            w = 0.1 * torch.randn(shape, generator=g)
        elif name == "patch_embed.pos_embedding":
            w = 0.1 * torch.randn(shape, generator=g)
        elif name == "patch_embed.proj.weight":
            w = torch.randn(shape, generator=g) * std * 4
        elif ".linear.weight" in name and ("norm1" in name or "norm2" in name or "norm_out" in name):
            w = torch.randn(shape, generator=g) * std * 4  # modulation projections: keep scale/shift/gate O(0.3)
        else:
            w = torch.randn(shape, generator=g) * std
        sd[name] = w.to(device=device, dtype=dtype)
    return sd


# ----------------------------------------------------------------------------
# forward
# ----------------------------------------------------------------------------
def timestep_embedding(t, dim):
    half = dim // 2
    exponent = -math.log(10000) * torch.arange(0, half, dtype=torch.float32, device=t.device) / half
    emb = t[:, None].float() * torch.exp(exponent)[None, :]
    return torch.cat([torch.cos(emb), torch.sin(emb)], dim=-1)  # flip_sin_to_cos


def sdpa(q, k, v):
    if q.device.type == "cpu" and q.dtype == torch.bfloat16:
        return F.scaled_dot_product_attention(q.float(), k.float(), v.float()).to(q.dtype)
    return F.scaled_dot_product_attention(q, k, v)


def _ln(x, w, b, eps):
    return F.layer_norm(x, (x.shape[-1],), w, b, eps)


def layer_norm_zero(sd, p, cfg, hidden, enc, temb):
    mod = F.linear(F.silu(temb), sd[p + "linear.weight"], sd[p + "linear.bias"])
    shift, scale, gate, e_shift, e_scale, e_gate = mod.chunk(6, dim=1)
    w, b = sd[p + "norm.weight"], sd[p + "norm.bias"]
    hidden = _ln(hidden, w, b, cfg.norm_eps) * (1 + scale)[:, None, :] + shift[:, None, :]
    enc = _ln(enc, w, b, cfg.norm_eps) * (1 + e_scale)[:, None, :] + e_shift[:, None, :]
    return hidden, enc, gate[:, None, :], e_gate[:, None, :]


def attention(sd, p, cfg, hidden, enc, rope):
    H = cfg.num_attention_heads
    n_txt = enc.shape[1]
    x = torch.cat([enc, hidden], dim=1)
    q = F.linear(x, sd[p + "to_q.weight"], sd[p + "to_q.bias"])
    k = F.linear(x, sd[p + "to_k.weight"], sd[p + "to_k.bias"])
    v = F.linear(x, sd[p + "to_v.weight"], sd[p + "to_v.bias"])
    q, k, v = (t.unflatten(2, (H, -1)).transpose(1, 2) for t in (q, k, v))
    q = _ln(q, sd[p + "norm_q.weight"], sd[p + "norm_q.bias"], 1e-6)
    k = _ln(k, sd[p + "norm_k.weight"], sd[p + "norm_k.bias"], 1e-6)
    if rope is not None:
        q = torch.cat([q[:, :, :n_txt], apply_rotary(q[:, :, n_txt:], *rope)], dim=2)
        k = torch.cat([k[:, :, :n_txt], apply_rotary(k[:, :, n_txt:], *rope)], dim=2)
    o = sdpa(q, k, v).transpose(1, 2).flatten(2, 3)
    o = F.linear(o, sd[p + "to_out.0.weight"], sd[p + "to_out.0.bias"])
    return o[:, n_txt:], o[:, :n_txt]


def block(sd, i, cfg, hidden, enc, temb, rope):
    p = f"transformer_blocks.{i}."
    n_txt = enc.shape[1]
    nh, ne, gate, e_gate = layer_norm_zero(sd, p + "norm1.", cfg, hidden, enc, temb)
    ah, ae = attention(sd, p + "attn1.", cfg, nh, ne, rope)
    hidden = hidden + gate * ah
    enc = enc + e_gate * ae
    nh, ne, gate, e_gate = layer_norm_zero(sd, p + "norm2.", cfg, hidden, enc, temb)
    f = torch.cat([ne, nh], dim=1)
    f = F.gelu(F.linear(f, sd[p + "ff.net.0.proj.weight"], sd[p + "ff.net.0.proj.bias"]), approximate="tanh")
    f = F.linear(f, sd[p + "ff.net.2.weight"], sd[p + "ff.net.2.bias"])
    hidden = hidden + gate * f[:, n_txt:]
    enc = enc + e_gate * f[:, :n_txt]
    return hidden, enc


def forward(sd, cfg: CogConfig, hidden, text, timestep, rope=None, return_intermediates=False):
    """hidden [B, F, 32, H, W] (model dtype); text [B, L, text_dim]; timestep [B]; rope = (cos, sin) fp32 [F*h*w, 64]."""
    B, Fr, C, Hh, Ww = hidden.shape
    p = cfg.patch_size
    dt = hidden.dtype
    t_emb = timestep_embedding(timestep, cfg.dim).to(dt)
    emb = F.linear(F.silu(F.linear(t_emb, sd["time_embedding.linear_1.weight"], sd["time_embedding.linear_1.bias"])),
                   sd["time_embedding.linear_2.weight"], sd["time_embedding.linear_2.bias"])
    txt = F.linear(text, sd["patch_embed.text_proj.weight"], sd["patch_embed.text_proj.bias"])
    img = F.conv2d(hidden.reshape(-1, C, Hh, Ww), sd["patch_embed.proj.weight"], sd["patch_embed.proj.bias"], stride=p)
    img = img.view(B, Fr, *img.shape[1:]).flatten(3).transpose(2, 3).flatten(1, 2)
    x = torch.cat([txt, img], dim=1).contiguous()
    pos = sd["patch_embed.pos_embedding"]
    if pos.shape[1] != x.shape[1]:  # another frame count / resolution: diffusers recomputes the sincos table
        pos = joint_pos_embedding(cfg, Hh, Ww, Fr).to(x.device)
    x = x + pos.to(dtype=dt)
    n_txt = text.shape[1]
    enc, x = x[:, :n_txt], x[:, n_txt:]
    inter = {"embed": torch.cat([enc, x], dim=1), "emb": emb}
    for i in range(cfg.num_layers):
        x, enc = block(sd, i, cfg, x, enc, emb, rope)
        if return_intermediates:
            inter[f"block{i}"] = torch.cat([enc, x], dim=1)
    x = _ln(x, sd["norm_final.weight"], sd["norm_final.bias"], cfg.norm_eps)
    mod = F.linear(F.silu(emb), sd["norm_out.linear.weight"], sd["norm_out.linear.bias"])
    shift, scale = mod.chunk(2, dim=1)
    x = _ln(x, sd["norm_out.norm.weight"], sd["norm_out.norm.bias"], cfg.norm_eps) * (1 + scale)[:, None, :] + shift[:, None, :]
    x = F.linear(x, sd["proj_out.weight"], sd["proj_out.bias"])
    out = x.reshape(B, Fr, Hh // p, Ww // p, -1, p, p).permute(0, 1, 4, 2, 5, 3, 6).flatten(5, 6).flatten(3, 4)
    return (out, inter) if return_intermediates else out


# ----------------------------------------------------------------------------
# the ALG denoise loop, cog:1005-1140 (DDIM scheduler branch)
# ----------------------------------------------------------------------------
def denoise_loop(transformer, scheduler, latents, image_latents, prompt_embeds, negative_prompt_embeds,
                 num_inference_steps, guidance_scale, alg, prepare_lp, get_lp_strength, on_step=None, teacher=None,
                 use_dynamic_cfg=False, dpm_randn=None):
    """``transformer(x [B, F, 32, H, W], text [B, L, D], timestep [B]) -> noise``; ``prepare_lp(type, sigma, k, f)``
    returns the low-passed image latents [1, F, 16, H, W] (cog:586-703, in latent or pixel space)."""
    do_cfg = guidance_scale > 1.0
    use_lp = alg.get("use_low_pass_guidance", False)
    if do_cfg and use_lp:
        pe3 = torch.cat([negative_prompt_embeds, negative_prompt_embeds, prompt_embeds], dim=0)
    pe2 = torch.cat([negative_prompt_embeds, prompt_embeds], dim=0) if do_cfg else prompt_embeds
    scheduler.set_timesteps(num_inference_steps)
    for i, t in enumerate(scheduler.timesteps):
        if teacher is not None:
            latents = teacher[i]
        two_pass = True
        if do_cfg and use_lp:
            s = get_lp_strength(i, num_inference_steps, alg["lp_strength_schedule_type"],
                                alg["schedule_interval_start_time"], alg["schedule_interval_end_time"],
                                alg["schedule_linear_start_weight"], alg["schedule_linear_end_weight"],
                                alg["schedule_linear_end_time"], alg["schedule_exp_decay_rate"])
            two_pass = s == 0
            if alg["lp_strength_schedule_type"] == "exponential" and s < 0.1:
                two_pass = True
            sigma = alg["lp_blur_sigma"] * s
            k = alg["lp_blur_kernel_size"] * s if alg["schedule_blur_kernel_size"] else alg["lp_blur_kernel_size"]
            f = 1.0 - (1.0 - alg["lp_resize_factor"]) * s
            lp = prepare_lp(alg["lp_filter_type"], sigma, k, f)
            if two_pass:
                x = torch.cat([torch.cat([latents] * 2), torch.cat([lp] * 2, dim=0)], dim=2)
            else:
                x = torch.cat([torch.cat([latents] * 3), torch.cat([image_latents, lp, lp], dim=0)], dim=2)
        elif do_cfg:
            x = torch.cat([torch.cat([latents] * 2), torch.cat([image_latents] * 2, dim=0)], dim=2)
        else:
            if use_lp:
                raise NameError("two_pass")  # quirk q3: ALG on without CFG leaves two_pass undefined (cog:1084)
            x = torch.cat([latents, image_latents], dim=2)
        timestep = t.expand(x.shape[0])
        text = pe2 if two_pass else pe3
        noise_pred = transformer(x, text, timestep).float()
        if do_cfg and x.shape[0] == 3:
            u0, u, tx = noise_pred.chunk(3)
            noise = u0 + guidance_scale * (tx - u)
        elif do_cfg:
            u, tx = noise_pred.chunk(2)
            w = guidance_scale
            if use_dynamic_cfg and not use_lp:  # cog:1103-1108 (only the non-ALG branch has it)
                w = 1 + guidance_scale * ((1 - math.cos(math.pi * ((num_inference_steps - t.item()) / num_inference_steps) ** 5.0)) / 2)
            noise = u + w * (tx - u)
        else:
            noise = noise_pred
        if dpm_randn is not None:  # CogVideoXDPMScheduler: carries pred_original_sample (cog:1113-1122)
            if i == 0:
                old_pred = None
            latents, old_pred = scheduler.step(noise, old_pred, int(t), int(scheduler.timesteps[i - 1]) if i > 0 else None,
                                               latents, dpm_randn)
            latents = latents.to(prompt_embeds.dtype)
        else:
            latents = scheduler.step(noise, int(t), latents).to(prompt_embeds.dtype)
        if on_step is not None:
            on_step(i, t, latents, noise_pred)
    return latents
