"""``HunyuanVideoTransformer3DModel`` (I2V, ``image_condition_type="token_replace"``) on the sm_100a kernels of
``libalg_b200.so``.

Mirrors the interface the reference pipeline uses on ``self.transformer`` (hy:1021-1031, 1115-1119, 1243-1252):
``.config``, ``.dtype``, ``__call__(hidden_states=, timestep=, encoder_hidden_states=, encoder_attention_mask=,
pooled_projections=, guidance=, attention_kwargs=, return_dict=False)``.  This module only SEQUENCES C-ABI calls
(``alg_b200.ops``) the way diffusers' ``transformer_hunyuan_video.py`` sequences its modules:

    first-frame replacement + cast + im2col   alg_patch_gather   (hy:1171-1195, 1232: ``cat([image|lp, latents[:, :, 1:]])``
                                                                  is a frame-0 pointer override, never a tensor)
    every nn.Linear                           alg_gemm_bf16      tcgen05; epilogues GELU-tanh / SiLU / x + gate * y with
                                                                  the first-frame (token-replace) gate split
    AdaLayerNormZero(/Single/Continuous)      alg_layer_norm     bf16 rounding chain; first-frame rows use the timestep-0
                                                                  modulation
    per-head RMSNorm + RoPE (latent tokens)   alg_head_norm_rope
    joint latent+text attention               alg_attention_bf16 tcgen05 flash attention, head_dim 128

One CFG pass at a time: at ~119 k latent tokens per pass there is nothing to gain from batching passes, and the passes
differ in their text length and pooled conditioning.  The key-padding mask of diffusers (valid text tokens are a prefix)
is applied by DROPPING the padded tokens: they never influence a valid token (masked as keys, per-token elsewhere), so
the latent output is unchanged -- checked against the mask-keeping oracle (``oracle/hunyuan_oracle.py``).  The single
-stream block's ``cat([attn, mlp])`` is a shared [R, d + 4d] buffer the attention and the MLP GEMM write side by side.
Restated from diffusers@be2fb77 (not available offline: parity unpinned).
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import Dict, Optional

import torch

from . import _lib, embeddings, ops

HUNYUAN_VIDEO_I2V = dict(in_channels=16, out_channels=16, num_attention_heads=24, attention_head_dim=128, num_layers=20,
                         num_single_layers=40, num_refiner_layers=2, mlp_ratio=4.0, patch_size=2, patch_size_t=1,
                         qk_norm="rms_norm", guidance_embeds=True, text_embed_dim=4096, pooled_projection_dim=768,
                         rope_theta=256.0, rope_axes_dim=(16, 56, 56), image_condition_type="token_replace")


def parameter_shapes(cfg: dict) -> Dict[str, tuple]:
    """name -> shape of every parameter (diffusers naming)."""
    hd = cfg["attention_head_dim"]
    d = cfg["num_attention_heads"] * hd
    mlp = int(d * cfg["mlp_ratio"])
    s: Dict[str, tuple] = {}

    def lin(name, o, i):
        s[name + ".weight"] = (o, i)
        s[name + ".bias"] = (o,)

    def temb(name, in_dim):
        lin(name + ".linear_1", d, in_dim)
        lin(name + ".linear_2", d, d)

    s["x_embedder.proj.weight"] = (d, cfg["in_channels"], cfg["patch_size_t"], cfg["patch_size"], cfg["patch_size"])
    s["x_embedder.proj.bias"] = (d,)
    ce = "context_embedder."
    temb(ce + "time_text_embed.timestep_embedder", 256)
    temb(ce + "time_text_embed.text_embedder", cfg["text_embed_dim"])
    lin(ce + "proj_in", d, cfg["text_embed_dim"])
    for i in range(cfg["num_refiner_layers"]):
        p = ce + f"token_refiner.refiner_blocks.{i}."
        for n in ("norm1", "norm2"):
            s[p + n + ".weight"] = s[p + n + ".bias"] = (d,)
        for n in ("to_q", "to_k", "to_v", "to_out.0"):
            lin(p + "attn." + n, d, d)
        lin(p + "ff.net.0.proj", mlp, d)
        lin(p + "ff.net.2", d, mlp)
        lin(p + "norm_out.linear", 2 * d, d)
    temb("time_text_embed.timestep_embedder", 256)
    if cfg["guidance_embeds"]:
        temb("time_text_embed.guidance_embedder", 256)
    temb("time_text_embed.text_embedder", cfg["pooled_projection_dim"])
    for i in range(cfg["num_layers"]):
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
    for i in range(cfg["num_single_layers"]):
        p = f"single_transformer_blocks.{i}."
        for n in ("to_q", "to_k", "to_v"):
            lin(p + "attn." + n, d, d)
        s[p + "attn.norm_q.weight"] = s[p + "attn.norm_k.weight"] = (hd,)
        lin(p + "norm.linear", 3 * d, d)
        lin(p + "proj_mlp", mlp, d)
        lin(p + "proj_out", d, d + mlp)
    lin("norm_out.linear", 2 * d, d)
    lin("proj_out", cfg["patch_size_t"] * cfg["patch_size"] ** 2 * cfg["out_channels"], d)
    return s


def synthetic_state_dict(cfg: dict, seed: int = 0, device="cuda", std: float = 0.02) -> Dict[str, torch.Tensor]:
    """Seeded random-init weights at the true shapes (no checkpoints offline); deterministic per (seed, name)."""
    sd = {}
    for idx, (name, shape) in enumerate(parameter_shapes(cfg).items()):
        g = torch.Generator(device=device).manual_seed(seed * 1_000_003 + idx)
        is_norm_w = name.endswith(("norm1.weight", "norm2.weight", "norm_q.weight", "norm_k.weight", "norm_added_q.weight",
                                   "norm_added_k.weight"))
        if is_norm_w:
            w = 1 + 0.1 * torch.randn(shape, generator=g, device=device)
        elif name.endswith(("norm1.bias", "norm2.bias")):
            w = 0.1 * torch.randn(shape, generator=g, device=device)
        elif name == "x_embedder.proj.weight":
            w = torch.randn(shape, generator=g, device=device) * (std * 6)
        else:
            w = torch.randn(shape, generator=g, device=device) * std
        sd[name] = w.to(torch.bfloat16)
    return sd


class HunyuanVideoTransformer3DModel:
    """Native-kernel stand-in for diffusers' ``HunyuanVideoTransformer3DModel`` (inference only, token_replace I2V)."""

    def __init__(self, **config):
        cfg = dict(HUNYUAN_VIDEO_I2V)
        cfg.update(config)
        if cfg["image_condition_type"] != "token_replace":
            # the reference's loop conditions by frame replacement only (hy:1171-1232), so latent_concat checkpoints
            # cannot work with it either (quirk q9)
            raise NotImplementedError("only image_condition_type='token_replace' (HunyuanVideo-I2V) is built")
        if cfg["patch_size"] != 2 or cfg["patch_size_t"] != 1 or cfg["qk_norm"] != "rms_norm":
            raise NotImplementedError("only patch (1, 2, 2) with per-head RMSNorm q/k is built")
        if cfg["attention_head_dim"] not in (64, 128) or sum(cfg["rope_axes_dim"]) != cfg["attention_head_dim"]:
            raise NotImplementedError("attention_head_dim must be 64 or 128 and equal sum(rope_axes_dim)")
        self.config = SimpleNamespace(**cfg)
        self._cfg = cfg
        self._w: Dict[str, torch.Tensor] = {}
        self._ws: Dict[tuple, dict] = {}
        self._rope: Dict[tuple, tuple] = {}
        self.dtype = torch.bfloat16
        self.device = torch.device("cpu")

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path, subfolder: str = "transformer", torch_dtype=None,
                        cache_dir=None, device="cuda"):
        """run.py:72-77 surface: real weights from a LOCAL diffusers snapshot (directory, or hub id resolved in cache_dir)."""
        import os

        from . import checkpoint

        snap = checkpoint.resolve_snapshot(str(pretrained_model_name_or_path), cache_dir)
        if snap is None:
            raise FileNotFoundError(f"no local diffusers snapshot for {pretrained_model_name_or_path!r} (there is no network: "
                                    "pass a directory, or a hub id present under cache_dir)")
        folder = os.path.join(snap, subfolder)
        cfg = checkpoint.read_config(folder)
        for k in ("patch_size", "rope_axes_dim"):
            if isinstance(cfg.get(k), list):
                cfg[k] = tuple(cfg[k])
        return cls(**cfg).load_state_dict(checkpoint.load_safetensors_dir(folder, device))

    @classmethod
    def from_synthetic(cls, seed: int = 0, device="cuda", **config):
        m = cls(**config)
        return m.load_state_dict(synthetic_state_dict(m._cfg, seed=seed, device=device))

    def load_state_dict(self, sd: Dict[str, torch.Tensor]):
        shapes = parameter_shapes(self._cfg)
        missing = [k for k in shapes if k not in sd]
        if missing:
            raise KeyError(f"missing parameters: {missing[:4]}{'...' if len(missing) > 4 else ''}")
        self.device = next(iter(sd.values())).device
        for name, shape in shapes.items():
            t = sd[name]
            if tuple(t.shape) != tuple(shape):
                raise ValueError(f"{name}: expected shape {tuple(shape)}, got {tuple(t.shape)}")
            self._w[name] = t.to(device=self.device, dtype=torch.bfloat16).contiguous()
        self._ws.clear()
        self._rope.clear()
        return self

    def state_dict(self):
        return dict(self._w)

    def to(self, device=None, dtype=None):
        if device is not None and self._w:
            dev = torch.device(device)
            if dev.type == "cuda" and dev.index is None:
                dev = torch.device("cuda", torch.cuda.current_device())
            if dev != self.device:
                self.load_state_dict({k: v.to(dev) for k, v in self._w.items()})
        return self

    # ---- helpers ------------------------------------------------------------------------------------
    def _workspace(self, N: int, Lt: int) -> dict:
        key = (N, Lt)
        ws = self._ws.get(key)
        if ws is None:
            c = self._cfg
            d = c["num_attention_heads"] * c["attention_head_dim"]
            mlp = int(d * c["mlp_ratio"])
            R = N + Lt
            Rpad = (R + 7) // 8 * 8
            e = lambda *shape: torch.empty(*shape, device=self.device, dtype=torch.bfloat16)  # noqa: E731
            ws = dict(x=e(R, d), h=e(R, d), q=e(R, d), k=e(R, d), ao=e(R, d), vt=e(d, Rpad), cm=e(R, d + mlp),
                      proj=e(N, 4 * c["out_channels"]), A=e(N, 4 * c["in_channels"]))
            self._ws = {key: ws}
        return ws

    def _linear(self, name, a, **kw):
        return ops.gemm(a, self._w[name + ".weight"], self._w[name + ".bias"], **kw)

    def _embed_mlp(self, name, a):
        """TimestepEmbedding / PixArtAlphaTextProjection: linear_2(silu(linear_1(a)))."""
        return self._linear(name + ".linear_2", self._linear(name + ".linear_1", a, epilogue=_lib.EPI_SILU))

    def _token_refiner(self, text: torch.Tensor, tproj: torch.Tensor, out: torch.Tensor):
        """HunyuanVideoTokenRefiner on the VALID text tokens [Lt, text_dim] -> out [Lt, d] (all-true mask)."""
        c, w = self._cfg, self._w
        heads, hd = c["num_attention_heads"], c["attention_head_dim"]
        d = heads * hd
        Lt = text.shape[0]
        ce = "context_embedder."
        pooled = ops.mean_rows(text).view(1, -1)
        temb = ops.add(self._embed_mlp(ce + "time_text_embed.timestep_embedder", tproj),
                       self._embed_mlp(ce + "time_text_embed.text_embedder", pooled))
        silu_t = ops.silu(temb)
        hs = self._linear(ce + "proj_in", text)
        Lpad = (Lt + 7) // 8 * 8
        vt = torch.empty(d, Lpad, device=text.device, dtype=torch.bfloat16)
        for i in range(c["num_refiner_layers"]):
            p = ce + f"token_refiner.refiner_blocks.{i}."
            n = ops.layer_norm(hs, eps=1e-6, weight=w[p + "norm1.weight"], bias=w[p + "norm1.bias"])
            q, k = self._linear(p + "attn.to_q", n), self._linear(p + "attn.to_k", n)
            ops.gemm(w[p + "attn.to_v.weight"], n, w[p + "attn.to_v.bias"], bias_per_row=True, out=vt[:, :Lt])
            ao = ops.attention(q.view(1, Lt, heads, hd), k.view(1, Lt, heads, hd), vt.view(1, heads, hd, Lpad), n_kv=Lt)
            gates = self._linear(p + "norm_out.linear", silu_t).view(2, d)
            last = i == c["num_refiner_layers"] - 1
            hs = self._linear(p + "attn.to_out.0", ao.view(Lt, d), epilogue=_lib.EPI_GATE_RESIDUAL, residual=hs, gate=gates[0],
                              gate_round=True)
            n = ops.layer_norm(hs, eps=1e-6, weight=w[p + "norm2.weight"], bias=w[p + "norm2.bias"])
            f = self._linear(p + "ff.net.0.proj", n, epilogue=_lib.EPI_SILU)
            hs = self._linear(p + "ff.net.2", f, epilogue=_lib.EPI_GATE_RESIDUAL, residual=hs, gate=gates[1], gate_round=True,
                              out=out if last else None)
        if c["num_refiner_layers"] == 0:
            ops.copy_rows(hs, out)
        return out

    # ---- forward ------------------------------------------------------------------------------------
    def forward_pass(self, latents: torch.Tensor, first_frame: Optional[torch.Tensor], text: torch.Tensor,
                     pooled: torch.Tensor, timestep: float, guidance: Optional[float], out: Optional[torch.Tensor] = None):
        """One CFG pass (hy:1243-1252) without materialising the model input.

        latents [16, T, H, W] (fp32 or bf16): frames 1.. are read from it; frame 0 comes from ``first_frame`` [16, 1, H, W]
        (the image latent or its low-passed copy; ``None`` reads frame 0 of ``latents``).  text [Lt, text_dim] bf16: the
        VALID prompt tokens (mask prefix); pooled [pooled_dim] bf16; timestep / guidance: the values the reference hands
        over AFTER its cast to the transformer dtype (hy:1237, 1117).  -> noise [16, T, H, W] bf16.
        """
        c, w = self._cfg, self._w
        heads, hd = c["num_attention_heads"], c["attention_head_dim"]
        d = heads * hd
        Cl, T, H, W = latents.shape
        gh, gw = H // 2, W // 2
        N, F1, Lt = T * gh * gw, gh * gw, text.shape[0]
        if N % 8 != 0:
            raise ValueError(f"the latent token count {N} must be a multiple of 8 (16-byte aligned V^T column blocks)")
        if Lt < 1:
            raise ValueError("encoder_attention_mask selects no text token")
        R = N + Lt
        ws = self._workspace(N, Lt)
        x, h, q, k, ao, vt, cm, proj, A = (ws[n] for n in ("x", "h", "q", "k", "ao", "vt", "cm", "proj", "A"))
        key = (T, gh, gw)
        if key not in self._rope:
            self._rope = {key: embeddings.hunyuan_rotary_pos_embed(T, gh, gw, c["rope_axes_dim"], c["rope_theta"], self.device)}
        cos, sin = self._rope[key]

        # 1. condition embedding: temb (real timestep [+ guidance]) and the token-replace embedding (timestep 0)
        tp = torch.empty(2, 256, device=self.device, dtype=torch.bfloat16)
        ops.timestep_embedding(float(timestep), 256, torch.bfloat16, self.device, out=tp[0])
        ops.timestep_embedding(0.0, 256, torch.bfloat16, self.device, out=tp[1])
        te = self._embed_mlp("time_text_embed.timestep_embedder", tp)          # [2, d]
        pp = self._embed_mlp("time_text_embed.text_embedder", pooled.view(1, -1))  # [1, d]
        emb = torch.empty(2, d, device=self.device, dtype=torch.bfloat16)
        ops.add(te[0], pp[0], out=emb[0])
        ops.add(te[1], pp[0], out=emb[1])
        if c["guidance_embeds"]:
            if guidance is None:
                raise ValueError("guidance is required for guidance-distilled checkpoints (guidance_embeds=True)")
            gp = ops.timestep_embedding(float(guidance), 256, torch.bfloat16, self.device).view(1, 256)
            ge = self._embed_mlp("time_text_embed.guidance_embedder", gp)
            ops.add(emb[0], ge[0], out=emb[0])
        silu_emb = ops.silu(emb)  # row 0: temb, row 1: token_replace_emb; every consumer applies SiLU first

        # 2. patch embedding of [first_frame | latents[:, 1:]] -> latent rows; refined text -> text rows of the joint buffer
        ops.patch_gather([[latents if first_frame is None else (latents, first_frame)]], A)
        ops.gemm(A, w["x_embedder.proj.weight"].view(d, -1), w["x_embedder.proj.bias"], out=x[:N])
        self._token_refiner(text, tp[0:1], x[N:])

        def mod6(name, rows):
            return self._linear(name, rows).view(rows.shape[0], -1, d)

        q4, k4 = q.view(1, R, heads, hd), k.view(1, R, heads, hd)
        vt4 = vt.view(1, heads, hd, vt.shape[-1])
        lat, txt = slice(0, N), slice(N, R)

        def qk_norm(t, rows, name, rope):
            ops.head_norm_rope(t[rows], heads, hd, norm_kind=_lib.NORM_RMS, weight=w[name + ".weight"], eps=1e-6,
                               cos=cos if rope else None, sin=sin if rope else None, rope_row0=0, rope_rows=N)

        def gated(a, name, rows, gate, gate_alt=None):
            self._linear(name, a, epilogue=_lib.EPI_GATE_RESIDUAL, residual=x[rows], gate=gate, gate_alt=gate_alt,
                         gate_split_row=F1 if gate_alt is not None else 0, gate_round=True, out=x[rows])

        def ln_mod(rows, scale, shift, scale_alt=None, shift_alt=None):
            ops.layer_norm(x[rows], eps=1e-6, scale=scale, shift=shift, scale_alt=scale_alt, shift_alt=shift_alt,
                           split_row=F1 if scale_alt is not None else 0, chain_bf16=True, out=h[rows])

        # 3. dual-stream blocks: separate weights per stream, joint attention over [latent | text]
        for i in range(c["num_layers"]):
            p = f"transformer_blocks.{i}."
            m = mod6(p + "norm1.linear", silu_emb)          # [2, 6, d]: shift/scale/gate (msa), shift/scale/gate (mlp)
            cmod = mod6(p + "norm1_context.linear", silu_emb[0:1])[0]
            ln_mod(lat, m[0, 1], m[0, 0], m[1, 1], m[1, 0])
            ln_mod(txt, cmod[1], cmod[0])
            a = p + "attn."
            self._linear(a + "to_q", h[lat], out=q[lat])
            self._linear(a + "to_k", h[lat], out=k[lat])
            ops.gemm(w[a + "to_v.weight"], h[lat], w[a + "to_v.bias"], bias_per_row=True, out=vt[:, :N])
            self._linear(a + "add_q_proj", h[txt], out=q[txt])
            self._linear(a + "add_k_proj", h[txt], out=k[txt])
            ops.gemm(w[a + "add_v_proj.weight"], h[txt], w[a + "add_v_proj.bias"], bias_per_row=True, out=vt[:, N:R])
            qk_norm(q, lat, a + "norm_q", True)
            qk_norm(k, lat, a + "norm_k", True)
            qk_norm(q, txt, a + "norm_added_q", False)
            qk_norm(k, txt, a + "norm_added_k", False)
            ops.attention(q4, k4, vt4, n_kv=R, out=ao.view(1, R, heads, hd))
            gated(ao[lat], a + "to_out.0", lat, m[0, 2], m[1, 2])
            gated(ao[txt], a + "to_add_out", txt, cmod[2])
            ln_mod(lat, m[0, 4], m[0, 3], m[1, 4], m[1, 3])
            ln_mod(txt, cmod[4], cmod[3])
            ff = cm.view(-1)[: N * (cm.shape[1] - d)].view(N, -1)  # [N, 4d] scratch inside the single-stream buffer
            self._linear(p + "ff.net.0.proj", h[lat], epilogue=_lib.EPI_GELU_TANH, out=ff)
            gated(ff, p + "ff.net.2", lat, m[0, 5], m[1, 5])
            ffc = self._linear(p + "ff_context.net.0.proj", h[txt], epilogue=_lib.EPI_GELU_TANH)
            gated(ffc, p + "ff_context.net.2", txt, cmod[5])

        # 4. single-stream blocks on the joint sequence; attention output and MLP activations share the concat buffer
        allr = slice(0, R)
        cm_attn = cm[:, :d].unflatten(1, (heads, hd)).unsqueeze(0)
        for i in range(c["num_single_layers"]):
            p = f"single_transformer_blocks.{i}."
            m = mod6(p + "norm.linear", silu_emb)           # [2, 3, d]: shift, scale, gate
            ln_mod(allr, m[0, 1], m[0, 0], m[1, 1], m[1, 0])
            self._linear(p + "proj_mlp", h, epilogue=_lib.EPI_GELU_TANH, out=cm[:, d:])
            a = p + "attn."
            self._linear(a + "to_q", h, out=q)
            self._linear(a + "to_k", h, out=k)
            ops.gemm(w[a + "to_v.weight"], h, w[a + "to_v.bias"], bias_per_row=True, out=vt[:, :R])
            qk_norm(q, allr, a + "norm_q", True)  # rope_rows = N: the text rows are normalised but not rotated
            qk_norm(k, allr, a + "norm_k", True)
            ops.attention(q4, k4, vt4, n_kv=R, out=cm_attn)
            gated(cm, p + "proj_out", allr, m[0, 2], m[1, 2])

        # 5. AdaLayerNormContinuous -> proj_out -> unpatchify (latent rows only)
        m = self._linear("norm_out.linear", silu_emb[0:1]).view(2, d)  # scale, shift
        ln_mod(lat, m[0], m[1])
        ops.gemm(h[lat], w["proj_out.weight"], w["proj_out.bias"], out=proj)
        if out is None:
            out = torch.empty(c["out_channels"], T, H, W, device=self.device, dtype=torch.bfloat16)
        ops.unpatchify(proj, out.unsqueeze(0), channel_major=True)
        return out

    def __call__(self, hidden_states, timestep, encoder_hidden_states, encoder_attention_mask, pooled_projections,
                 guidance=None, attention_kwargs=None, return_dict: bool = True):
        """diffusers-compatible call on the pre-batched [B, 16, T, H, W] input (hy:1243-1252); passes run one by one."""
        B = hidden_states.shape[0]
        outs = []
        for b in range(B):
            n_valid = int(encoder_attention_mask[b].sum().item())
            g = None if guidance is None else float(guidance.flatten()[b if guidance.numel() > 1 else 0])
            outs.append(self.forward_pass(hidden_states[b], None, encoder_hidden_states[b, :n_valid].to(torch.bfloat16).contiguous(),
                                          pooled_projections[b].to(torch.bfloat16).contiguous(),
                                          float(timestep.flatten()[b if timestep.numel() > 1 else 0]), g))
        out = torch.stack(outs)
        if not return_dict:
            return (out,)
        return SimpleNamespace(sample=out)
