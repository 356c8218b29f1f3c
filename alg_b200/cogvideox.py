"""``CogVideoXTransformer3DModel`` on the sm_100a kernels of ``libalg_b200.so``.

Mirrors the interface the reference pipeline uses on ``self.transformer`` (cog:899-903, 1082-1090): ``.config``,
``.dtype``, ``__call__(hidden_states=, encoder_hidden_states=, timestep=, ofs=, image_rotary_emb=, attention_kwargs=,
return_dict=False)``.  This module only SEQUENCES C-ABI calls (``alg_b200.ops``) the way diffusers'
``cogvideox_transformer_3d.py`` sequences its modules -- every tensor operation is a CUDA kernel of the library:

    model-input assembly + im2col   alg_patch_gather      (cog:1059-1075: the [latents]*3 / cat / cast never exist)
    every nn.Linear                 alg_gemm_bf16         tcgen05, epilogues: +pos_embedding, GELU-tanh, SiLU,
                                                          x + gate * y with the text / video gate split
    LayerNormZero / AdaLayerNorm    alg_layer_norm        bf16 rounding chain, text rows use the enc_* modulation
    per-head LayerNorm(64) + RoPE   alg_head_norm_rope    video tokens only
    joint text+video attention      alg_attention_bf16    tcgen05 flash attention, head_dim 64

The text and video tokens of a pass live in ONE [226 + N, d] buffer (text first, like the attention processor's
concat), and the 2-3 CFG passes are stacked along the rows.  Restated from diffusers@be2fb77 (not available offline:
parity unpinned); checked against ``oracle/cog_oracle.py``.
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import Dict, List, Optional, Sequence

import torch

from . import _lib, embeddings, ops

COGVIDEOX_5B_I2V = dict(num_attention_heads=48, attention_head_dim=64, in_channels=32, out_channels=16,
                        flip_sin_to_cos=True, freq_shift=0, time_embed_dim=512, ofs_embed_dim=None, text_embed_dim=4096,
                        num_layers=42, dropout=0.0, attention_bias=True, sample_width=90, sample_height=60,
                        sample_frames=49, patch_size=2, patch_size_t=None, temporal_compression_ratio=4,
                        max_text_seq_length=226, activation_fn="gelu-approximate", timestep_activation_fn="silu",
                        norm_elementwise_affine=True, norm_eps=1e-5, spatial_interpolation_scale=1.875,
                        temporal_interpolation_scale=1.0, use_rotary_positional_embeddings=True,
                        use_learned_positional_embeddings=True, ff_mult=4)


def parameter_shapes(cfg: dict) -> Dict[str, tuple]:
    """name -> shape of every parameter / persistent buffer (diffusers naming)."""
    d, te, hd = cfg["num_attention_heads"] * cfg["attention_head_dim"], cfg["time_embed_dim"], cfg["attention_head_dim"]
    p = cfg["patch_size"]
    s: Dict[str, tuple] = {}

    def lin(name, o, i):
        s[name + ".weight"] = (o, i)
        s[name + ".bias"] = (o,)

    def ln(name, n):
        s[name + ".weight"] = s[name + ".bias"] = (n,)

    s["patch_embed.proj.weight"] = (d, cfg["in_channels"], p, p)
    s["patch_embed.proj.bias"] = (d,)
    lin("patch_embed.text_proj", d, cfg["text_embed_dim"])
    lat_frames = (cfg["sample_frames"] - 1) // cfg["temporal_compression_ratio"] + 1
    n_patch = lat_frames * (cfg["sample_height"] // p) * (cfg["sample_width"] // p)
    s["patch_embed.pos_embedding"] = (1, cfg["max_text_seq_length"] + n_patch, d)
    lin("time_embedding.linear_1", te, d)
    lin("time_embedding.linear_2", te, te)
    for i in range(cfg["num_layers"]):
        b = f"transformer_blocks.{i}."
        for n in ("norm1", "norm2"):
            lin(b + n + ".linear", 6 * d, te)
            ln(b + n + ".norm", d)
        for n in ("to_q", "to_k", "to_v", "to_out.0"):
            lin(b + "attn1." + n, d, d)
        ln(b + "attn1.norm_q", hd)
        ln(b + "attn1.norm_k", hd)
        lin(b + "ff.net.0.proj", cfg["ff_mult"] * d, d)
        lin(b + "ff.net.2", d, cfg["ff_mult"] * d)
    ln("norm_final", d)
    lin("norm_out.linear", 2 * d, te)
    ln("norm_out.norm", d)
    lin("proj_out", p * p * cfg["out_channels"], d)
    return s


def synthetic_state_dict(cfg: dict, seed: int = 0, device="cuda", std: float = 0.02) -> Dict[str, torch.Tensor]:
    """Seeded random-init weights at the true shapes (no checkpoints offline); deterministic per (seed, name)."""
    sd = {}
    for idx, (name, shape) in enumerate(parameter_shapes(cfg).items()):
        g = torch.Generator(device=device).manual_seed(seed * 1_000_003 + idx)
        plain_norm = "norm" in name and ".linear." not in name
        if plain_norm and name.endswith(".weight"):
            w = 1 + 0.1 * torch.randn(shape, generator=g, device=device)
        elif plain_norm and name.endswith(".bias"):
            w = 0.1 * torch.randn(shape, generator=g, device=device)
        elif name == "patch_embed.pos_embedding":
            w = 0.1 * torch.randn(shape, generator=g, device=device)
        elif name == "patch_embed.proj.weight" or (".linear.weight" in name and "norm" in name):
            w = torch.randn(shape, generator=g, device=device) * (std * 4)
        else:
            w = torch.randn(shape, generator=g, device=device) * std
        sd[name] = w.to(torch.bfloat16)
    return sd


class CogVideoXTransformer3DModel:
    """Native-kernel stand-in for diffusers' ``CogVideoXTransformer3DModel`` (inference only, patch_size_t=None)."""

    def __init__(self, **config):
        cfg = dict(COGVIDEOX_5B_I2V)
        cfg.update(config)
        if cfg["patch_size"] != 2 or cfg["patch_size_t"] is not None or cfg["ofs_embed_dim"] is not None:
            raise NotImplementedError("only the CogVideoX 1.0 layout (patch_size 2, no temporal patching, no ofs) is built")
        if cfg["attention_head_dim"] not in (64, 128):
            raise NotImplementedError("attention_head_dim must be 64 or 128")
        self.config = SimpleNamespace(**cfg)
        self._cfg = cfg
        self._w: Dict[str, torch.Tensor] = {}
        self._ws: Dict[tuple, dict] = {}
        self._pos_cache: Dict[tuple, torch.Tensor] = {}
        self.dtype = torch.bfloat16
        self.device = torch.device("cpu")

    # ---- construction -------------------------------------------------------------------------------
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
        dev = next(iter(sd.values())).device
        self.device = dev
        for name, shape in shapes.items():
            t = sd[name]
            if tuple(t.shape) != tuple(shape):
                raise ValueError(f"{name}: expected shape {tuple(shape)}, got {tuple(t.shape)}")
            self._w[name] = t.to(device=dev, dtype=torch.bfloat16).contiguous()
        self._ws.clear()
        self._pos_cache.clear()
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
    def _workspace(self, P: int, S: int) -> dict:
        key = (P, S)
        ws = self._ws.get(key)
        if ws is None:
            c = self._cfg
            d = c["num_attention_heads"] * c["attention_head_dim"]
            M, Spad = P * S, (S + 7) // 8 * 8
            e = lambda *shape: torch.empty(*shape, device=self.device, dtype=torch.bfloat16)  # noqa: E731
            ws = dict(x=e(M, d), h=e(M, d), q=e(M, d), k=e(M, d), ao=e(M, d), vt=e(P, d, Spad), ff=e(M, c["ff_mult"] * d),
                      proj=e(M, 4 * c["out_channels"]))
            self._ws = {key: ws}  # one live shape at a time (the loop alternates between at most two pass counts)
        return ws

    def _pos_embedding(self, Fr: int, H: int, W: int) -> torch.Tensor:
        c = self._cfg
        pos = self._w["patch_embed.pos_embedding"][0]
        n = c["max_text_seq_length"] + Fr * (H // 2) * (W // 2)
        if pos.shape[0] == n and H == c["sample_height"] and W == c["sample_width"]:
            return pos
        if c["use_learned_positional_embeddings"] and (H != c["sample_height"] or W != c["sample_width"]):
            raise ValueError("It is currently not possible to generate videos at a different resolution that the defaults. "
                             "This should only be the case with 'THUDM/CogVideoX-5b-I2V'.")
        key = (Fr, H, W)
        if key not in self._pos_cache:  # another frame count: diffusers recomputes the sincos table (host, once)
            d = c["num_attention_heads"] * c["attention_head_dim"]
            tab = embeddings.cogvideox_joint_pos_embedding(d, c["max_text_seq_length"], H // 2, W // 2, Fr,
                                                           c["spatial_interpolation_scale"], c["temporal_interpolation_scale"])
            self._pos_cache[key] = tab[0].to(self.device, torch.bfloat16).contiguous()
        return self._pos_cache[key]

    # ---- forward ------------------------------------------------------------------------------------
    def forward_passes(self, latents: Sequence[torch.Tensor], image_latents: Sequence[torch.Tensor],
                       text: Sequence[torch.Tensor], timestep, rope=None, out: Optional[torch.Tensor] = None):
        """The 1-3 CFG passes of one denoise step (cog:1059-1090) without materialising the batched model input.

        latents[p], image_latents[p] [F, 16, H, W] bf16; text[p] [L, text_dim] bf16; rope = (cos, sin) fp32 [F*h*w, 64]
        -> noise [n_pass, F, 16, H, W] bf16.
        """
        c, w = self._cfg, self._w
        P = len(latents)
        heads, hd = c["num_attention_heads"], c["attention_head_dim"]
        d = heads * hd
        Fr, Cl, H, W = latents[0].shape
        L = text[0].shape[0]
        N = Fr * (H // 2) * (W // 2)
        S = L + N
        if L != c["max_text_seq_length"]:
            raise ValueError(f"encoder_hidden_states must hold max_text_seq_length={c['max_text_seq_length']} tokens, got {L}")
        ws = self._workspace(P, S)
        x, h, q, k, ao, vt, ff, proj = (ws[n] for n in ("x", "h", "q", "k", "ao", "vt", "ff", "proj"))
        pos = self._pos_embedding(Fr, H, W)

        # 1. time embedding (every consumer applies SiLU first, so keep silu(emb))
        t_emb = ops.timestep_embedding(float(timestep), d, torch.bfloat16, self.device).view(1, d)
        e1 = ops.gemm(t_emb, w["time_embedding.linear_1.weight"], w["time_embedding.linear_1.bias"], epilogue=_lib.EPI_SILU)
        silu_emb = ops.gemm(e1, w["time_embedding.linear_2.weight"], w["time_embedding.linear_2.bias"], epilogue=_lib.EPI_SILU)

        # 2. patch + text embedding straight into the joint [text | video] rows, + positional table in the epilogue
        kdim = c["in_channels"] * 4
        A = torch.empty(P * N, kdim, device=self.device, dtype=torch.bfloat16)
        ops.patch_gather([[latents[p].transpose(0, 1), image_latents[p].transpose(0, 1)] for p in range(P)], A)
        wp = w["patch_embed.proj.weight"].view(d, kdim)
        for p in range(P):
            rows = x[p * S:(p + 1) * S]
            same = next((r for r in range(p) if text[r].data_ptr() == text[p].data_ptr()), None)
            if same is not None:  # [neg, neg, pos]: project each distinct prompt once
                ops.copy_rows(x[same * S:same * S + L], rows[:L])
            else:
                ops.gemm(text[p], w["patch_embed.text_proj.weight"], w["patch_embed.text_proj.bias"],
                         epilogue=_lib.EPI_RESIDUAL, residual=pos[:L], out=rows[:L])
            ops.gemm(A[p * N:(p + 1) * N], wp, w["patch_embed.proj.bias"], epilogue=_lib.EPI_RESIDUAL, residual=pos[L:],
                     out=rows[L:])

        cos, sin = rope if rope is not None else (None, None)
        q4, k4, ao4 = (t.view(P, S, heads, hd) for t in (q, k, ao))
        vt4 = vt.view(P, heads, hd, vt.shape[-1])

        def gated(a, wn, bn, gate, gate_alt):
            ops.gemm(a, w[wn], w[bn], epilogue=_lib.EPI_GATE_RESIDUAL, residual=x, gate=gate, gate_alt=gate_alt,
                     gate_split_row=L, gate_round=True, rows_per_batch=S, out=x)

        # 3. blocks
        for i in range(c["num_layers"]):
            b = f"transformer_blocks.{i}."
            for part in ("norm1", "norm2"):
                mod = ops.gemm(silu_emb, w[b + part + ".linear.weight"], w[b + part + ".linear.bias"]).view(6, d)
                shift, scale, gate, e_shift, e_scale, e_gate = mod.unbind(0)
                ops.layer_norm(x, eps=c["norm_eps"], weight=w[b + part + ".norm.weight"], bias=w[b + part + ".norm.bias"],
                               scale=scale, shift=shift, scale_alt=e_scale, shift_alt=e_shift, rows_per_batch=S,
                               split_row=L, chain_bf16=True, out=h)
                if part == "norm1":
                    a = b + "attn1."
                    ops.gemm(h, w[a + "to_q.weight"], w[a + "to_q.bias"], out=q)
                    ops.gemm(h, w[a + "to_k.weight"], w[a + "to_k.bias"], out=k)
                    for p in range(P):  # V^T = W_v h^T + b_v (swapped operands, bias per row): both MMAs K-major
                        ops.gemm(w[a + "to_v.weight"], h[p * S:(p + 1) * S], w[a + "to_v.bias"], bias_per_row=True,
                                 out=vt[p, :, :S])
                    for t, n in ((q, "norm_q"), (k, "norm_k")):
                        ops.head_norm_rope(t, heads, hd, norm_kind=_lib.NORM_LAYER, weight=w[a + n + ".weight"],
                                           bias=w[a + n + ".bias"], eps=1e-6, cos=cos, sin=sin, rows_per_batch=S,
                                           rope_row0=L, rope_rows=N)
                    ops.attention(q4, k4, vt4, n_kv=S, out=ao4)
                    gated(ao, a + "to_out.0.weight", a + "to_out.0.bias", gate, e_gate)
                else:
                    ops.gemm(h, w[b + "ff.net.0.proj.weight"], w[b + "ff.net.0.proj.bias"], epilogue=_lib.EPI_GELU_TANH, out=ff)
                    gated(ff, b + "ff.net.2.weight", b + "ff.net.2.bias", gate, e_gate)

        # 4. norm_final -> AdaLayerNorm(norm_out) -> proj_out -> unpatchify
        ops.layer_norm(x, eps=c["norm_eps"], weight=w["norm_final.weight"], bias=w["norm_final.bias"], out=h)
        mod = ops.gemm(silu_emb, w["norm_out.linear.weight"], w["norm_out.linear.bias"]).view(2, d)
        ops.layer_norm(h, eps=c["norm_eps"], weight=w["norm_out.norm.weight"], bias=w["norm_out.norm.bias"], scale=mod[1],
                       shift=mod[0], chain_bf16=True, out=q)
        ops.gemm(q, w["proj_out.weight"], w["proj_out.bias"], out=proj)
        if out is None:
            out = torch.empty(P, Fr, c["out_channels"], H, W, device=self.device, dtype=torch.bfloat16)
        for p in range(P):
            ops.unpatchify(proj[p * S + L:(p + 1) * S], out[p:p + 1].transpose(1, 2), channel_major=True)
        return out

    def __call__(self, hidden_states, encoder_hidden_states, timestep, timestep_cond=None, ofs=None,
                 image_rotary_emb=None, attention_kwargs=None, return_dict: bool = True):
        """diffusers-compatible call on the pre-batched [B, F, 32, H, W] input (cog:1082-1090)."""
        B = hidden_states.shape[0]
        if B > 3:
            raise NotImplementedError("the engine batches at most the 3 CFG passes of one sample")
        ts = timestep.flatten() if torch.is_tensor(timestep) else None
        if ts is not None and ts.numel() > 1 and not bool((ts == ts[0]).all()):
            raise NotImplementedError("all passes of a step share one timestep (cog:1081)")
        t0 = float(ts[0]) if ts is not None else float(timestep)
        oc = self._cfg["out_channels"]
        hs = hidden_states.to(torch.bfloat16)
        lat = [hs[b, :, :oc] for b in range(B)]
        img = [hs[b, :, oc:] for b in range(B)]
        text = [encoder_hidden_states[b].to(torch.bfloat16).contiguous() for b in range(B)]
        out = self.forward_passes(lat, img, text, t0, image_rotary_emb)
        if not return_dict:
            return (out,)
        return SimpleNamespace(sample=out)
