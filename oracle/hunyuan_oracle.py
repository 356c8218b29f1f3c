"""PyTorch restatement of ``HunyuanVideoTransformer3DModel.forward`` (token_replace I2V) and of the HunyuanVideo ALG loop.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Runs on CPU (fp32 or bf16) and, inside ``-m gpu`` tests only, on the
box's GPU as the eager-PyTorch checker.

PARITY UNPINNED for the DiT: the class lives in diffusers@be2fb77 (requirements.txt:13;
``models/transformers/transformer_hunyuan_video.py``), absent here.  Restated from its published forward (SURVEY
Appendix A.3): token refiner, condition embedding with the timestep-0 "token replace" embedding, dual-stream and
single-stream token-replace blocks, key-padding mask, RoPE on latent tokens only; anchored on the reference's call site
hy:1243-1252.  The oracle keeps the PADDED text tokens and the boolean masks exactly like diffusers does, so it also
checks the engine's shortcut of dropping the padded tokens.

The loop (``denoise_loop``) restates first-party code: pipeline_hunyuan_video_image2video_lowpass.py:1126-1270.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch
import torch.nn.functional as F


@dataclass
class HunyuanConfig:
    in_channels: int = 16
    out_channels: int = 16
    num_attention_heads: int = 24
    attention_head_dim: int = 128
    num_layers: int = 20
    num_single_layers: int = 40
    num_refiner_layers: int = 2
    mlp_ratio: float = 4.0
    patch_size: int = 2
    patch_size_t: int = 1
    text_embed_dim: int = 4096
    pooled_projection_dim: int = 768
    rope_theta: float = 256.0
    rope_axes_dim: tuple = (16, 56, 56)

    @property
    def dim(self):
        return self.num_attention_heads * self.attention_head_dim


def tiny_config(**kw):
    base = dict(num_attention_heads=2, attention_head_dim=128, num_layers=2, num_single_layers=2, num_refiner_layers=1,
                text_embed_dim=64, pooled_projection_dim=32)
    base.update(kw)
    return HunyuanConfig(**base)


def parameter_shapes(cfg: HunyuanConfig):
    d, hd, mlp = cfg.dim, cfg.attention_head_dim, int(cfg.dim * cfg.mlp_ratio)
    s = {}

    def lin(name, o, i):
        s[name + ".weight"] = (o, i)
        s[name + ".bias"] = (o,)

    def temb(name, in_dim):
        lin(name + ".linear_1", d, in_dim)
        lin(name + ".linear_2", d, d)

    s["x_embedder.proj.weight"] = (d, cfg.in_channels, cfg.patch_size_t, cfg.patch_size, cfg.patch_size)
    s["x_embedder.proj.bias"] = (d,)
    ce = "context_embedder."
    temb(ce + "time_text_embed.timestep_embedder", 256)
    temb(ce + "time_text_embed.text_embedder", cfg.text_embed_dim)
    lin(ce + "proj_in", d, cfg.text_embed_dim)
    for i in range(cfg.num_refiner_layers):
        p = ce + f"token_refiner.refiner_blocks.{i}."
        for n in ("norm1", "norm2"):
            s[p + n + ".weight"] = s[p + n + ".bias"] = (d,)
        for n in ("to_q", "to_k", "to_v", "to_out.0"):
            lin(p + "attn." + n, d, d)
        lin(p + "ff.net.0.proj", mlp, d)
        lin(p + "ff.net.2", d, mlp)
        lin(p + "norm_out.linear", 2 * d, d)
    temb("time_text_embed.timestep_embedder", 256)
    temb("time_text_embed.guidance_embedder", 256)
    temb("time_text_embed.text_embedder", cfg.pooled_projection_dim)
    for i in range(cfg.num_layers):
        p = f"transformer_blocks.{i}."
        lin(p + "norm1.linear", 6 * d, d)
        lin(p + "norm1_context.linear", 6 * d, d)
        for n in ("to_q", "to_k", "to_v", "add_q_proj", "add_k_proj", "add_v_proj", "to_out.0", "to_add_out"):
            lin(p + "attn." + n, d, d)
        for n in ("norm_q", "norm_k", "norm_added_q", "norm_added_k"):
            s[p + "attn." + n + ".weight"] = (hd,)
        for f in ("ff", "ff_context"):
            lin(p + f + ".net.0.proj", mlp, d)
            lin(p + f + ".net.2", d, mlp)
    for i in range(cfg.num_single_layers):
        p = f"single_transformer_blocks.{i}."
        for n in ("to_q", "to_k", "to_v"):
            lin(p + "attn." + n, d, d)
        s[p + "attn.norm_q.weight"] = s[p + "attn.norm_k.weight"] = (hd,)
        lin(p + "norm.linear", 3 * d, d)
        lin(p + "proj_mlp", mlp, d)
        lin(p + "proj_out", d, d + mlp)
    lin("norm_out.linear", 2 * d, d)
    lin("proj_out", cfg.patch_size_t * cfg.patch_size * cfg.patch_size * cfg.out_channels, d)
    return s


def make_weights(cfg: HunyuanConfig, seed=0, device="cpu", dtype=torch.bfloat16, std=0.02):
    g = torch.Generator(device="cpu").manual_seed(seed)
    sd = {}
    for name, shape in parameter_shapes(cfg).items():
        if name.endswith("norm1.weight") or name.endswith("norm2.weight") or ".norm_q." in name or ".norm_k." in name or \
                "norm_added" in name:
            w = 1 + 0.1 * torch.randn(shape, generator=g)
        elif name.endswith("norm1.bias") or name.endswith("norm2.bias"):
            w = 0.1 * torch.randn(shape, generator=g)
        elif name == "x_embedder.proj.weight":
            w = torch.randn(shape, generator=g) * std * 6
        else:
            w = torch.randn(shape, generator=g) * std
        sd[name] = w.to(device=device, dtype=dtype)
    return sd


# ----------------------------------------------------------------------------
def timestep_proj(t, dim=256):
    half = dim // 2
    exponent = -math.log(10000) * torch.arange(0, half, dtype=torch.float32, device=t.device) / half
    emb = t[:, None].float() * torch.exp(exponent)[None, :]
    return torch.cat([torch.cos(emb), torch.sin(emb)], dim=-1)


def _mlp(sd, p, x, act=F.silu):
    return F.linear(act(F.linear(x, sd[p + ".linear_1.weight"], sd[p + ".linear_1.bias"])), sd[p + ".linear_2.weight"],
                    sd[p + ".linear_2.bias"])


def _lin(sd, p, x):
    return F.linear(x, sd[p + ".weight"], sd[p + ".bias"])


def _ln(x, eps=1e-6, w=None, b=None):
    return F.layer_norm(x, (x.shape[-1],), w, b, eps)


def rms_norm(x, weight, eps=1e-6):
    var = x.to(torch.float32).pow(2).mean(-1, keepdim=True)
    h = x * torch.rsqrt(var + eps)
    if weight.dtype in (torch.float16, torch.bfloat16):
        h = h.to(weight.dtype)
    return h * weight


def rope_tables(cfg: HunyuanConfig, T, gh, gw, device="cpu"):
    grids = torch.meshgrid(*[torch.arange(0, n, dtype=torch.float32, device=device) for n in (T, gh, gw)], indexing="ij")
    cos, sin = [], []
    for i in range(3):
        dim = cfg.rope_axes_dim[i]
        freqs = 1.0 / (cfg.rope_theta ** (torch.arange(0, dim, 2, dtype=torch.float32, device=device)[: dim // 2] / dim))
        f = torch.outer(grids[i].reshape(-1), freqs)
        cos.append(f.cos().repeat_interleave(2, dim=1).float())
        sin.append(f.sin().repeat_interleave(2, dim=1).float())
    return torch.cat(cos, dim=1), torch.cat(sin, dim=1)


def apply_rotary(x, cos, sin):
    cos, sin = cos[None, None], sin[None, None]
    xr, xi = x.reshape(*x.shape[:-1], -1, 2).unbind(-1)
    rot = torch.stack([-xi, xr], dim=-1).flatten(3)
    return (x.float() * cos + rot.float() * sin).to(x.dtype)


def sdpa(q, k, v, mask=None):
    if q.device.type == "cpu" and q.dtype == torch.bfloat16:
        return F.scaled_dot_product_attention(q.float(), k.float(), v.float(), attn_mask=mask).to(q.dtype)
    return F.scaled_dot_product_attention(q, k, v, attn_mask=mask)


def _heads(x, H):
    return x.unflatten(2, (H, -1)).transpose(1, 2)


def token_refiner(sd, cfg, text, timestep, mask):
    ce = "context_embedder."
    mask_f = mask.float().unsqueeze(-1)
    pooled = ((text * mask_f).sum(dim=1) / mask_f.sum(dim=1)).to(text.dtype)
    temb = _mlp(sd, ce + "time_text_embed.timestep_embedder", timestep_proj(timestep).to(text.dtype)) + \
        _mlp(sd, ce + "time_text_embed.text_embedder", pooled)
    hs = _lin(sd, ce + "proj_in", text)
    B, L = mask.shape
    m = mask.bool().view(B, 1, 1, L).expand(-1, -1, L, -1)
    attn_mask = (m & m.transpose(2, 3)).clone()
    attn_mask[:, :, :, 0] = True
    H = cfg.num_attention_heads
    for i in range(cfg.num_refiner_layers):
        p = ce + f"token_refiner.refiner_blocks.{i}."
        n = _ln(hs, 1e-6, sd[p + "norm1.weight"], sd[p + "norm1.bias"])
        q, k, v = (_heads(_lin(sd, p + "attn." + nm, n), H) for nm in ("to_q", "to_k", "to_v"))
        a = sdpa(q, k, v, attn_mask).transpose(1, 2).flatten(2, 3)
        a = _lin(sd, p + "attn.to_out.0", a)
        gate_msa, gate_mlp = _lin(sd, p + "norm_out.linear", F.silu(temb)).chunk(2, dim=1)
        hs = hs + a * gate_msa.unsqueeze(1)
        f = _ln(hs, 1e-6, sd[p + "norm2.weight"], sd[p + "norm2.bias"])
        f = _lin(sd, p + "ff.net.2", F.silu(_lin(sd, p + "ff.net.0.proj", f)))
        hs = hs + f * gate_mlp.unsqueeze(1)
    return hs


def _split_mod(norm, n_first, tr, base):
    """rows < n_first use the token-replace (scale, shift), the rest the regular ones."""
    zero = norm[:, :n_first] * (1 + tr[0][:, None]) + tr[1][:, None]
    orig = norm[:, n_first:] * (1 + base[0][:, None]) + base[1][:, None]
    return torch.cat([zero, orig], dim=1)


def _split_gate(res, y, n_first, tr_gate, gate):
    zero = res[:, :n_first] + y[:, :n_first] * tr_gate.unsqueeze(1)
    orig = res[:, n_first:] + y[:, n_first:] * gate.unsqueeze(1)
    return torch.cat([zero, orig], dim=1)


def joint_attention(sd, p, cfg, lat, txt, mask, rope, dual):
    H = cfg.num_attention_heads
    n_txt = txt.shape[1]
    x = lat if dual else torch.cat([lat, txt], dim=1)
    q, k, v = (_heads(_lin(sd, p + nm, x), H) for nm in ("to_q", "to_k", "to_v"))
    q, k = rms_norm(q, sd[p + "norm_q.weight"]), rms_norm(k, sd[p + "norm_k.weight"])
    if dual:
        q, k = apply_rotary(q, *rope), apply_rotary(k, *rope)
        eq, ek, ev = (_heads(_lin(sd, p + nm, txt), H) for nm in ("add_q_proj", "add_k_proj", "add_v_proj"))
        eq, ek = rms_norm(eq, sd[p + "norm_added_q.weight"]), rms_norm(ek, sd[p + "norm_added_k.weight"])
        q, k, v = torch.cat([q, eq], dim=2), torch.cat([k, ek], dim=2), torch.cat([v, ev], dim=2)
    else:
        q = torch.cat([apply_rotary(q[:, :, :-n_txt], *rope), q[:, :, -n_txt:]], dim=2)
        k = torch.cat([apply_rotary(k[:, :, :-n_txt], *rope), k[:, :, -n_txt:]], dim=2)
    o = sdpa(q, k, v, mask).transpose(1, 2).flatten(2, 3).to(q.dtype)
    o_lat, o_txt = o[:, :-n_txt], o[:, -n_txt:]
    if dual:
        o_lat, o_txt = _lin(sd, p + "to_out.0", o_lat), _lin(sd, p + "to_add_out", o_txt)
    return o_lat, o_txt


def dual_block(sd, i, cfg, x, c, temb, tr_emb, mask, rope, n_first):
    p = f"transformer_blocks.{i}."
    m = _lin(sd, p + "norm1.linear", F.silu(temb)).chunk(6, dim=1)      # shift, scale, gate (msa) | shift, scale, gate (mlp)
    t = _lin(sd, p + "norm1.linear", F.silu(tr_emb)).chunk(6, dim=1)
    cm = _lin(sd, p + "norm1_context.linear", F.silu(temb)).chunk(6, dim=1)
    nx = _split_mod(_ln(x), n_first, (t[1], t[0]), (m[1], m[0]))
    nc = _ln(c) * (1 + cm[1][:, None]) + cm[0][:, None]
    a, ac = joint_attention(sd, p + "attn.", cfg, nx, nc, mask, rope, True)
    x = _split_gate(x, a, n_first, t[2], m[2])
    c = c + ac * cm[2].unsqueeze(1)
    nx = _split_mod(_ln(x), n_first, (t[4], t[3]), (m[4], m[3]))
    nc = _ln(c) * (1 + cm[4][:, None]) + cm[3][:, None]
    f = _lin(sd, p + "ff.net.2", F.gelu(_lin(sd, p + "ff.net.0.proj", nx), approximate="tanh"))
    fc = _lin(sd, p + "ff_context.net.2", F.gelu(_lin(sd, p + "ff_context.net.0.proj", nc), approximate="tanh"))
    x = _split_gate(x, f, n_first, t[5], m[5])
    c = c + cm[5].unsqueeze(1) * fc
    return x, c


def single_block(sd, i, cfg, x, c, temb, tr_emb, mask, rope, n_first):
    p = f"single_transformer_blocks.{i}."
    n_txt = c.shape[1]
    h = torch.cat([x, c], dim=1)
    res = h
    m = _lin(sd, p + "norm.linear", F.silu(temb)).chunk(3, dim=1)   # shift, scale, gate
    t = _lin(sd, p + "norm.linear", F.silu(tr_emb)).chunk(3, dim=1)
    nh = _split_mod(_ln(h), n_first, (t[1], t[0]), (m[1], m[0]))
    mlp = F.gelu(_lin(sd, p + "proj_mlp", nh), approximate="tanh")
    a, ac = joint_attention(sd, p + "attn.", cfg, nh[:, :-n_txt], nh[:, -n_txt:], mask, rope, False)
    out = _lin(sd, p + "proj_out", torch.cat([torch.cat([a, ac], dim=1), mlp], dim=2))
    zero = out[:, :n_first] * t[2].unsqueeze(1)
    orig = out[:, n_first:] * m[2].unsqueeze(1)
    h = torch.cat([zero, orig], dim=1) + res
    return h[:, :-n_txt], h[:, -n_txt:]


def forward(sd, cfg: HunyuanConfig, hidden, timestep, text, text_mask, pooled, guidance, return_intermediates=False):
    """hidden [B, 16, T, H, W]; timestep [B] (already in the transformer dtype, hy:1237); text [B, L, 4096];
    text_mask [B, L]; pooled [B, 768]; guidance [B] (= guidance_scale * 1000 in the transformer dtype)."""
    B, C, T, Hh, Ww = hidden.shape
    p, pt = cfg.patch_size, cfg.patch_size_t
    gt, gh, gw = T // pt, Hh // p, Ww // p
    n_first = gh * gw
    dt = hidden.dtype
    rope = rope_tables(cfg, gt, gh, gw, hidden.device)
    te = "time_text_embed."
    pooled_p = _mlp(sd, te + "text_embedder", pooled)
    temb = _mlp(sd, te + "timestep_embedder", timestep_proj(timestep).to(dt)) + pooled_p
    tr_emb = _mlp(sd, te + "timestep_embedder", timestep_proj(torch.zeros_like(timestep)).to(dt)) + pooled_p
    temb = temb + _mlp(sd, te + "guidance_embedder", timestep_proj(guidance).to(dt))
    x = F.conv3d(hidden, sd["x_embedder.proj.weight"], sd["x_embedder.proj.bias"], stride=(pt, p, p)).flatten(2).transpose(1, 2)
    c = token_refiner(sd, cfg, text, timestep, text_mask)
    N, L = x.shape[1], c.shape[1]
    eff = N + text_mask.sum(dim=1, dtype=torch.int)
    mask = (torch.arange(N + L, device=x.device)[None] < eff[:, None])[:, None, None, :]
    inter = {"embed": x, "context": c, "temb": temb, "tr_emb": tr_emb}
    for i in range(cfg.num_layers):
        x, c = dual_block(sd, i, cfg, x, c, temb, tr_emb, mask, rope, n_first)
        if return_intermediates:
            inter[f"dual{i}"] = (x, c)
    for i in range(cfg.num_single_layers):
        x, c = single_block(sd, i, cfg, x, c, temb, tr_emb, mask, rope, n_first)
        if return_intermediates:
            inter[f"single{i}"] = (x, c)
    scale, shift = _lin(sd, "norm_out.linear", F.silu(temb).to(dt)).chunk(2, dim=1)
    x = _ln(x) * (1 + scale)[:, None, :] + shift[:, None, :]
    x = _lin(sd, "proj_out", x)
    x = x.reshape(B, gt, gh, gw, -1, pt, p, p).permute(0, 4, 1, 5, 2, 6, 3, 7).flatten(6, 7).flatten(4, 5).flatten(2, 3)
    return (x, inter) if return_intermediates else x


# ----------------------------------------------------------------------------
# the ALG denoise loop, hy:1126-1270 (token_replace checkpoints)
# ----------------------------------------------------------------------------
def denoise_loop(transformer, scheduler, latents, image_latents, pos, neg, num_inference_steps, guidance_scale,
                 true_cfg_scale, alg, lp_filter, get_lp_strength, dtype=torch.bfloat16, lp_on_noisy_latent=False,
                 on_step=None, teacher=None):
    """``transformer(x [B,16,T,H,W], timestep [B], text, mask, pooled, guidance) -> noise``; ``pos`` / ``neg`` are
    (prompt_embeds [1,L,D], pooled [1,P], mask [1,L]) triples (``neg`` None => no true CFG); ``lp_filter(x, type, sigma,
    k, f)`` filters the first-frame latent (in-latent mode, the only one that works in the reference: quirk q9)."""
    do_true_cfg = true_cfg_scale > 1 and neg is not None
    use_lp = alg.get("use_low_pass_guidance", False)
    sigmas = torch.linspace(1.0, 0.0, num_inference_steps + 1, dtype=torch.float64)[:-1].numpy()
    scheduler.set_timesteps(num_inference_steps, sigmas=sigmas)
    # hy:1115-1119: one entry per LATENT (not per pass): a [1] tensor that the embedder broadcasts over the 2-3 passes
    guidance = (torch.tensor([guidance_scale] * latents.shape[0], dtype=dtype, device=latents.device) * 1000.0)

    def strength(i):
        s = get_lp_strength(i, num_inference_steps, alg["lp_strength_schedule_type"], alg["schedule_interval_start_time"],
                            alg["schedule_interval_end_time"], alg["schedule_linear_start_weight"],
                            alg["schedule_linear_end_weight"], alg["schedule_linear_end_time"], alg["schedule_exp_decay_rate"])
        sigma = alg["lp_blur_sigma"] * s
        k = alg["lp_blur_kernel_size"] * s if alg["schedule_blur_kernel_size"] else alg["lp_blur_kernel_size"]
        f = 1.0 - (1.0 - alg["lp_resize_factor"]) * s
        return s, lp_filter(image_latents, alg["lp_filter_type"], sigma, k, f).to(image_latents.dtype)

    def cat3(a, b, c=None):
        return torch.cat([a, b] if c is None else [a, b, c], dim=0)

    for i, t in enumerate(scheduler.timesteps):
        if teacher is not None:
            latents = teacher[i]
        rest = latents[:, :, 1:]
        if do_true_cfg and use_lp:
            s, lp = strength(i)
            if s == 0.0 or lp_on_noisy_latent:
                x = torch.cat([cat3(image_latents, image_latents).to(dtype), cat3(rest, rest)], dim=2).to(dtype)
                ctx = [cat3(neg[k], pos[k]) for k in range(3)]
            else:
                x = torch.cat([cat3(image_latents, lp, lp), cat3(rest, rest, rest)], dim=2).to(dtype)
                ctx = [cat3(neg[k], neg[k], pos[k]) for k in range(3)]
        elif do_true_cfg:
            x = torch.cat([cat3(image_latents, image_latents).to(dtype), cat3(rest, rest)], dim=2).to(dtype)
            ctx = [cat3(neg[k], pos[k]) for k in range(3)]
        elif not use_lp:
            x = torch.cat([image_latents, rest], dim=2).to(dtype)
            ctx = list(pos)
        else:
            s, lp = strength(i)
            x = torch.cat([lp, rest], dim=2).to(dtype)
            ctx = list(pos)
        timestep = t.expand(x.shape[0]).to(dtype).to(x.device)
        noise_pred = transformer(x, timestep, ctx[0], ctx[2], ctx[1], guidance)
        if noise_pred.shape[0] == 3:
            u0, u, tx = noise_pred.chunk(3)
            noise = u0 + true_cfg_scale * (tx - u)
        elif noise_pred.shape[0] == 2:
            u, tx = noise_pred.chunk(2)
            noise = u + true_cfg_scale * (tx - u)
        else:
            noise = noise_pred
        stepped = scheduler.step(noise[:, :, 1:], latents[:, :, 1:])
        latents = torch.cat([image_latents, stepped], dim=2)
        if on_step is not None:
            on_step(i, t, latents, noise_pred)
    return latents
