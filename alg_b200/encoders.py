"""Conditioning encoders on the sm_100a kernels of ``libalg_b200.so`` (run once per video).

``UMT5EncoderModel`` / ``T5EncoderModel`` and ``CLIPVisionModel`` keep the ``transformers`` surface the reference pipelines
drive -- ``text_encoder(input_ids, attention_mask).last_hidden_state`` (wan:212, cog:258 ``[0]``) and
``image_encoder(pixel_values=..., output_hidden_states=True).hidden_states[-2]`` (wan:233-234) -- with parameters named like
the HF state_dict, so a diffusers snapshot's ``text_encoder/`` and ``image_encoder/`` folders load without key mapping.
This module only SEQUENCES C-ABI calls; every tensor operation is a kernel of the library:

    nn.Embedding                      alg_gather_rows_bf16
    T5LayerNorm                       alg_t5_rms_norm_bf16           (bf16 rounding chain of the eager module)
    every nn.Linear                   alg_gemm_bf16                  tcgen05; q/k/v fused into one N = 3 * inner GEMM;
                                                                     gelu_new and "+ residual" in the epilogue
    T5 attention (relative bias, no   alg_small_attention            scores -> bf16, + bias -> bf16, fp32 softmax -> bf16
      scaling, key-padding mask)
    gated-GELU product                alg_mul_bf16
    CLIP-ViT-H in float32 (run.py:48) alg_patchify_f32, alg_clip_embed_f32, alg_layer_norm_f32, alg_bias_act_f32,
                                      alg_small_attention (fp32, head_dim 80), and alg_gemm_bf16 on bf16x3-split operands
                                      (alg_split3_bf16: hi*hi + hi*lo + lo*hi with fp32 accumulation ~ fp32 nn.Linear to 1e-5)

Arithmetic restated from transformers' ``modeling_umt5.py`` / ``modeling_t5.py`` / ``modeling_clip.py``; unlike diffusers,
transformers IS installed here, so ``tests/test_gpu_encoders.py`` pins these against the real modules (seeded weights).
"""
from __future__ import annotations

import ctypes as C
import math
from types import SimpleNamespace
from typing import Dict, List, Optional

import torch

from . import _lib, ops

UMT5_XXL = dict(vocab_size=256384, d_model=4096, d_kv=64, d_ff=10240, num_layers=24, num_heads=64,
                relative_attention_num_buckets=32, relative_attention_max_distance=128, layer_norm_epsilon=1e-6,
                per_layer_relative_bias=True)   # google/umt5-xxl (Wan text_encoder)
T5_V1_1_XXL = dict(UMT5_XXL, vocab_size=32128, per_layer_relative_bias=False)  # CogVideoX text_encoder: bias table in block 0 only
CLIP_L_TEXT = dict(vocab_size=49408, hidden_size=768, intermediate_size=3072, num_hidden_layers=12, num_attention_heads=12,
                   max_position_embeddings=77, hidden_act="quick_gelu", layer_norm_eps=1e-5, eos_token_id=2)  # HunyuanVideo text_encoder_2
CLIP_VIT_H_14 = dict(hidden_size=1280, intermediate_size=5120, num_hidden_layers=32, num_attention_heads=16, image_size=224,
                     patch_size=14, num_channels=3, hidden_act="gelu", layer_norm_eps=1e-5)  # Wan image_encoder


def _launch(fn, device, *args):
    with torch.cuda.device(device):
        _lib.check(fn(*args, _lib.stream_ptr(device)))


def split_weight(wt: torch.Tensor, k_pad: Optional[int] = None) -> torch.Tensor:
    """fp32 (or fp16) [N, K] -> bf16 [N, 3 K'] = [hi | lo | hi] (K zero-padded to K' so that rows stay 16-byte aligned)."""
    wt = wt.float().reshape(wt.shape[0], -1).contiguous()
    N, K = wt.shape
    k_pad = k_pad or K
    if k_pad != K:
        wt = torch.cat([wt, wt.new_zeros(N, k_pad - K)], dim=1).contiguous()
    out = torch.empty(N, 3 * k_pad, device=wt.device, dtype=torch.bfloat16)
    _launch(_lib.lib().alg_split3_bf16, wt.device, wt.data_ptr(), out.data_ptr(), N, k_pad, 1)
    return out


def split_act(x: torch.Tensor) -> torch.Tensor:
    """fp32 [rows, K] contiguous -> bf16 [rows, 3 K] = [hi | hi | lo] (the activation side of the split product)."""
    rows, K = x.shape
    assert x.is_contiguous() and x.dtype == torch.float32
    a3 = torch.empty(rows, 3 * K, device=x.device, dtype=torch.bfloat16)
    _launch(_lib.lib().alg_split3_bf16, x.device, x.data_ptr(), a3.data_ptr(), rows, K, 0)
    return a3


def linear_f32(x: torch.Tensor, w3: torch.Tensor, bias, act: int = 0, residual=None) -> torch.Tensor:
    """fp32 nn.Linear on the bf16 tensor cores: split the activations [hi | hi | lo], one K-tripled GEMM with fp32
    accumulation and fp32 output, then bias / activation / residual in fp32."""
    rows, K = x.shape
    if not act and (w3.shape[0] % 8 == 0) and (residual is None or (residual.is_contiguous() and residual.shape[1] == w3.shape[0])):
        # no activation: bias and residual ride in the GEMM epilogue (one pass less over the output)
        return ops.gemm(split_act(x), w3, None, out_dtype=torch.float32, bias_f32=bias, residual_f32=residual)
    y = ops.gemm(split_act(x), w3, None, out_dtype=torch.float32)
    if bias is not None or act or residual is not None:
        _launch(_lib.lib().alg_bias_act_f32, x.device, y.data_ptr(), None if bias is None else bias.data_ptr(),
                None if residual is None else residual.data_ptr(), rows, y.shape[1], act)
    return y


def small_attention(q, k, v, out, *, batch, heads, head_dim, n_q, n_kv, q_bs, q_rs, k_bs, k_rs, v_bs, v_rs, o_bs, o_rs, scale,
                    rel_bias=None, kv_valid=None, causal=False, kv_group=1, key_mask=None):
    _lib.require_cuda(q, k, v, out, rel_bias, kv_valid, key_mask)
    a = _lib.SmallAttention()
    a.q, a.k, a.v, a.out = q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr()
    a.dtype, a.batch, a.heads, a.head_dim, a.n_q, a.n_kv = _lib.dtype_code(q.dtype), batch, heads, head_dim, n_q, n_kv
    a.q_bs, a.q_rs, a.k_bs, a.k_rs, a.v_bs, a.v_rs, a.o_bs, a.o_rs = q_bs, q_rs, k_bs, k_rs, v_bs, v_rs, o_bs, o_rs
    a.scale, a.causal = float(scale), int(causal)
    a.rel_bias = None if rel_bias is None else rel_bias.data_ptr()
    a.kv_valid = None if kv_valid is None else kv_valid.data_ptr()
    a.kv_group = int(kv_group)
    a.key_mask = None if key_mask is None else key_mask.data_ptr()
    _launch(_lib.lib().alg_small_attention, q.device, C.byref(a))
    return out


# ======================================================================================================
# UMT5 / T5 encoder
# ======================================================================================================
def relative_position_buckets(n_q: int, n_kv: int, num_buckets: int, max_distance: int) -> torch.Tensor:
    """T5 bidirectional bucket of (key j - query i), as ``_relative_position_bucket``: int64 [2 * n - 1] indexed by j - i + n - 1."""
    n = max(n_q, n_kv)
    rel = torch.arange(-(n - 1), n, dtype=torch.long)
    nb = num_buckets // 2
    buckets = (rel > 0).to(torch.long) * nb
    rel = rel.abs()
    max_exact = nb // 2
    large = max_exact + (torch.log(rel.float() / max_exact) / math.log(max_distance / max_exact) * (nb - max_exact)).to(torch.long)
    large = torch.min(large, torch.full_like(large, nb - 1))
    return buckets + torch.where(rel < max_exact, rel, large)


class UMT5EncoderModel:
    """Native ``UMT5EncoderModel`` / ``T5EncoderModel`` (encoder stack only, inference, bf16)."""

    DEFAULTS = UMT5_XXL

    def __init__(self, **config):
        cfg = dict(self.DEFAULTS)
        cfg.update({k: v for k, v in config.items() if k in cfg or k in ("feed_forward_proj", "dense_act_fn")})
        if config.get("feed_forward_proj", "gated-gelu") != "gated-gelu":
            raise NotImplementedError("only the gated-GELU feed-forward of (U)MT5 / T5 v1.1 is built")
        self._cfg = cfg
        self.config = SimpleNamespace(**cfg)
        self.dtype = torch.bfloat16
        self.device = torch.device("cpu")
        self._w: Dict[str, torch.Tensor] = {}
        self._bias_cache: Dict[tuple, torch.Tensor] = {}

    # ---- construction -------------------------------------------------------------------------------
    def parameter_shapes(self) -> Dict[str, tuple]:
        c = self._cfg
        inner = c["num_heads"] * c["d_kv"]
        s = {"shared.weight": (c["vocab_size"], c["d_model"])}
        for i in range(c["num_layers"]):
            p = f"encoder.block.{i}.layer."
            for n in ("q", "k", "v"):
                s[p + f"0.SelfAttention.{n}.weight"] = (inner, c["d_model"])
            s[p + "0.SelfAttention.o.weight"] = (c["d_model"], inner)
            if c["per_layer_relative_bias"] or i == 0:
                s[p + "0.SelfAttention.relative_attention_bias.weight"] = (c["relative_attention_num_buckets"], c["num_heads"])
            s[p + "0.layer_norm.weight"] = (c["d_model"],)
            s[p + "1.DenseReluDense.wi_0.weight"] = (c["d_ff"], c["d_model"])
            s[p + "1.DenseReluDense.wi_1.weight"] = (c["d_ff"], c["d_model"])
            s[p + "1.DenseReluDense.wo.weight"] = (c["d_model"], c["d_ff"])
            s[p + "1.layer_norm.weight"] = (c["d_model"],)
        s["encoder.final_layer_norm.weight"] = (c["d_model"],)
        return s

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path, subfolder: str = "text_encoder", torch_dtype=None, cache_dir=None,
                        device="cuda"):
        import os

        from . import checkpoint
        snap = checkpoint.resolve_snapshot(str(pretrained_model_name_or_path), cache_dir, require=subfolder or "transformer")
        if snap is None:
            raise FileNotFoundError(f"no local snapshot for {pretrained_model_name_or_path!r} (there is no network)")
        folder = os.path.join(snap, subfolder) if subfolder else snap
        cfg = checkpoint.read_config(folder)
        cfg["per_layer_relative_bias"] = cfg.get("model_type", "umt5") == "umt5"
        return cls(**cfg).load_state_dict(checkpoint.load_safetensors_dir(folder, device))

    @classmethod
    def from_synthetic(cls, seed: int = 0, device="cuda", **config):
        m = cls(**config)
        sd = {}
        for idx, (name, shape) in enumerate(m.parameter_shapes().items()):
            g = torch.Generator(device=device).manual_seed(seed * 1_000_003 + idx)
            if name.endswith("layer_norm.weight"):
                w = 1 + 0.1 * torch.randn(shape, generator=g, device=device)
            elif "relative_attention_bias" in name:
                w = torch.randn(shape, generator=g, device=device)
            elif name == "shared.weight":
                w = torch.randn(shape, generator=g, device=device)
            else:
                w = torch.randn(shape, generator=g, device=device) * (shape[-1] ** -0.5)
            sd[name] = w.to(torch.bfloat16)
        return m.load_state_dict(sd)

    def load_state_dict(self, sd: Dict[str, torch.Tensor]):
        sd = dict(sd)
        if "shared.weight" not in sd and "encoder.embed_tokens.weight" in sd:
            sd["shared.weight"] = sd["encoder.embed_tokens.weight"]
        missing = [k for k in self.parameter_shapes() if k not in sd]
        if missing:
            raise KeyError(f"missing parameters: {missing[:4]}{'...' if len(missing) > 4 else ''}")
        dev = sd["shared.weight"].device
        if dev.type != "cuda":
            raise RuntimeError("encoder weights must live on a CUDA device (no CPU fallback)")
        self.device = dev
        w = {}
        for name, shape in self.parameter_shapes().items():
            t = sd[name]
            if tuple(t.shape) != tuple(shape):
                raise ValueError(f"{name}: expected shape {tuple(shape)}, got {tuple(t.shape)}")
            w[name] = t.to(device=dev, dtype=torch.bfloat16).contiguous()
        for i in range(self._cfg["num_layers"]):  # one N = 3 * inner projection instead of three launches reading h
            p = f"encoder.block.{i}.layer.0.SelfAttention."
            w[p + "qkv.weight"] = torch.cat([w.pop(p + "q.weight"), w.pop(p + "k.weight"), w.pop(p + "v.weight")], dim=0).contiguous()
        self._w = w
        self._bias_cache = {}
        return self

    def state_dict(self):
        out = {}
        inner = self._cfg["num_heads"] * self._cfg["d_kv"]
        for k, v in self._w.items():
            if k.endswith("qkv.weight"):
                for j, n in enumerate(("q", "k", "v")):
                    out[k.replace("qkv", n)] = v[j * inner:(j + 1) * inner]
            else:
                out[k] = v
        return out

    def to(self, device=None, dtype=None):
        if device is not None and self._w and torch.device(device).type == "cuda" and torch.device(device) != self.device:
            dev = torch.device(device)
            dev = torch.device("cuda", torch.cuda.current_device()) if dev.index is None else dev
            if dev != self.device:
                self._w = {k: v.to(dev) for k, v in self._w.items()}
                self.device, self._bias_cache = dev, {}
        return self

    # ---- forward ------------------------------------------------------------------------------------
    def _rel_bias(self, layer: int, n: int) -> torch.Tensor:
        """fp32 [heads, 2n - 1] table of this layer's bias by (key - query) offset (``compute_bias`` without the [n, n] blow-up)."""
        c = self._cfg
        src = layer if c["per_layer_relative_bias"] else 0
        key = (src, n)
        if key not in self._bias_cache:
            buckets = relative_position_buckets(n, n, c["relative_attention_num_buckets"], c["relative_attention_max_distance"])
            table = self._w[f"encoder.block.{src}.layer.0.SelfAttention.relative_attention_bias.weight"]  # [buckets, heads] bf16
            self._bias_cache[key] = table.float().t().contiguous()[:, buckets.to(table.device)].contiguous()
        return self._bias_cache[key]

    def __call__(self, input_ids, attention_mask=None, **kw):
        c, w, dev = self._cfg, self._w, self.device
        if not w:
            raise RuntimeError("UMT5EncoderModel has no weights loaded")
        ids = input_ids.to(dev).to(torch.int64).contiguous()
        B, L = ids.shape
        H, dk, d = c["num_heads"], c["d_kv"], c["d_model"]
        inner, eps = H * dk, c["layer_norm_epsilon"]
        valid = None
        if attention_mask is not None:
            # right-padded prompts (what the tokenizers produce): keys beyond the prompt length are masked
            valid = attention_mask.to(dev).gt(0).sum(dim=1).to(torch.int32).contiguous()
        lib = _lib.lib()
        rows = B * L
        x = torch.empty(rows, d, device=dev, dtype=torch.bfloat16)
        _launch(lib.alg_gather_rows_bf16, dev, w["shared.weight"].data_ptr(), c["vocab_size"], ids.data_ptr(), x.data_ptr(), rows, d)
        h = torch.empty_like(x)
        qkv = torch.empty(rows, 3 * inner, device=dev, dtype=torch.bfloat16)
        att = torch.empty(rows, inner, device=dev, dtype=torch.bfloat16)
        g = torch.empty(rows, c["d_ff"], device=dev, dtype=torch.bfloat16)
        lin = torch.empty_like(g)
        for i in range(c["num_layers"]):
            p = f"encoder.block.{i}.layer."
            _launch(lib.alg_t5_rms_norm_bf16, dev, x.data_ptr(), d, h.data_ptr(), d, rows, d, eps, w[p + "0.layer_norm.weight"].data_ptr())
            ops.gemm(h, w[p + "0.SelfAttention.qkv.weight"], None, out=qkv)
            small_attention(qkv, qkv[:, inner:], qkv[:, 2 * inner:], att, batch=B, heads=H, head_dim=dk, n_q=L, n_kv=L,
                            q_bs=L * 3 * inner, q_rs=3 * inner, k_bs=L * 3 * inner, k_rs=3 * inner, v_bs=L * 3 * inner,
                            v_rs=3 * inner, o_bs=L * inner, o_rs=inner, scale=1.0, rel_bias=self._rel_bias(i, L), kv_valid=valid)
            ops.gemm(att, w[p + "0.SelfAttention.o.weight"], None, epilogue=_lib.EPI_RESIDUAL, residual=x, out=x)
            _launch(lib.alg_t5_rms_norm_bf16, dev, x.data_ptr(), d, h.data_ptr(), d, rows, d, eps, w[p + "1.layer_norm.weight"].data_ptr())
            ops.gemm(h, w[p + "1.DenseReluDense.wi_0.weight"], None, epilogue=_lib.EPI_GELU_TANH, out=g)
            ops.gemm(h, w[p + "1.DenseReluDense.wi_1.weight"], None, out=lin)
            _launch(lib.alg_mul_bf16, dev, g.data_ptr(), lin.data_ptr(), g.data_ptr(), g.numel())
            ops.gemm(g, w[p + "1.DenseReluDense.wo.weight"], None, epilogue=_lib.EPI_RESIDUAL, residual=x, out=x)
        out = torch.empty_like(x)
        _launch(lib.alg_t5_rms_norm_bf16, dev, x.data_ptr(), d, out.data_ptr(), d, rows, d, eps, w["encoder.final_layer_norm.weight"].data_ptr())
        return _EncoderOutput(last_hidden_state=out.view(B, L, d))


class T5EncoderModel(UMT5EncoderModel):
    """T5 v1.1 encoder (CogVideoX ``text_encoder``): one relative-bias table, in block 0, shared by every block."""

    DEFAULTS = T5_V1_1_XXL


class _EncoderOutput(SimpleNamespace):
    def __getitem__(self, i):
        return (self.last_hidden_state,)[i]


# ======================================================================================================
# CLIP vision tower (float32, like run.py:48)
# ======================================================================================================
class CLIPVisionModel:
    """Native ``CLIPVisionModel``: embeddings + pre-LN + encoder layers, float32 arithmetic; returns every hidden state."""

    # transformers' CLIPVisionConfig defaults (ViT-B/32, quick_gelu); CLIP_VIT_H_14 is the Wan image_encoder's config.json
    HF_DEFAULTS = dict(hidden_size=768, intermediate_size=3072, num_hidden_layers=12, num_attention_heads=12, image_size=224,
                       patch_size=32, num_channels=3, hidden_act="quick_gelu", layer_norm_eps=1e-5)

    def __init__(self, **config):
        cfg = dict(self.HF_DEFAULTS)
        cfg.update({k: v for k, v in config.items() if k in cfg})
        if cfg["hidden_act"] not in ("gelu", "quick_gelu"):
            raise NotImplementedError(f"hidden_act {cfg['hidden_act']!r}")
        self._cfg = cfg
        self.config = SimpleNamespace(**cfg)
        self.dtype = torch.float32
        self.device = torch.device("cpu")
        self._w: Dict[str, torch.Tensor] = {}

    def parameter_shapes(self) -> Dict[str, tuple]:
        c = self._cfg
        d, f = c["hidden_size"], c["intermediate_size"]
        n_pos = (c["image_size"] // c["patch_size"]) ** 2 + 1
        s = {"vision_model.embeddings.class_embedding": (d,),
             "vision_model.embeddings.patch_embedding.weight": (d, c["num_channels"], c["patch_size"], c["patch_size"]),
             "vision_model.embeddings.position_embedding.weight": (n_pos, d),
             "vision_model.pre_layrnorm.weight": (d,), "vision_model.pre_layrnorm.bias": (d,)}
        for i in range(c["num_hidden_layers"]):
            p = f"vision_model.encoder.layers.{i}."
            for n in ("q_proj", "k_proj", "v_proj", "out_proj"):
                s[p + f"self_attn.{n}.weight"], s[p + f"self_attn.{n}.bias"] = (d, d), (d,)
            for n in ("layer_norm1", "layer_norm2"):
                s[p + n + ".weight"], s[p + n + ".bias"] = (d,), (d,)
            s[p + "mlp.fc1.weight"], s[p + "mlp.fc1.bias"] = (f, d), (f,)
            s[p + "mlp.fc2.weight"], s[p + "mlp.fc2.bias"] = (d, f), (d,)
        return s

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path, subfolder: str = "image_encoder", torch_dtype=None, cache_dir=None,
                        device="cuda"):
        import os

        from . import checkpoint
        snap = checkpoint.resolve_snapshot(str(pretrained_model_name_or_path), cache_dir, require=subfolder or "transformer")
        if snap is None:
            raise FileNotFoundError(f"no local snapshot for {pretrained_model_name_or_path!r} (there is no network)")
        folder = os.path.join(snap, subfolder) if subfolder else snap
        cfg = checkpoint.read_config(folder)
        cfg = cfg.get("vision_config", cfg)
        return cls(**cfg).load_state_dict(checkpoint.load_safetensors_dir(folder, device))

    @classmethod
    def from_synthetic(cls, seed: int = 0, device="cuda", **config):
        """Seeded weights at the CLIP-ViT-H/14 shape of the Wan checkpoints unless ``config`` says otherwise."""
        m = cls(**dict(CLIP_VIT_H_14, **config))
        sd = {}
        for idx, (name, shape) in enumerate(m.parameter_shapes().items()):
            g = torch.Generator(device=device).manual_seed(seed * 1_000_003 + idx)
            if "norm" in name and name.endswith(".weight"):
                w = 1 + 0.1 * torch.randn(shape, generator=g, device=device)
            elif name.endswith(".bias"):
                w = 0.02 * torch.randn(shape, generator=g, device=device)
            elif name.endswith("embedding") or "position_embedding" in name:
                w = 0.5 * torch.randn(shape, generator=g, device=device)
            else:
                fan_in = 1
                for s_ in shape[1:]:
                    fan_in *= s_
                w = torch.randn(shape, generator=g, device=device) * (fan_in ** -0.5)
            sd[name] = w.float()
        return m.load_state_dict(sd)

    def _split_weight(self, wt: torch.Tensor, k_pad: Optional[int] = None) -> torch.Tensor:
        return split_weight(wt, k_pad)

    def load_state_dict(self, sd: Dict[str, torch.Tensor]):
        shapes = self.parameter_shapes()
        missing = [k for k in shapes if k not in sd]
        if missing:
            raise KeyError(f"missing parameters: {missing[:4]}{'...' if len(missing) > 4 else ''}")
        dev = sd["vision_model.pre_layrnorm.weight"].device
        if dev.type != "cuda":
            raise RuntimeError("encoder weights must live on a CUDA device (no CPU fallback)")
        self.device = dev
        c = self._cfg
        self._sd = {k: sd[k].to(device=dev, dtype=torch.float32).contiguous() for k in shapes}
        w = dict(self._sd)
        self._k_patch = (c["num_channels"] * c["patch_size"] ** 2 + 7) // 8 * 8
        w["patch.w3"] = self._split_weight(w["vision_model.embeddings.patch_embedding.weight"], self._k_patch)
        self._w = w
        self._prepare_layers(w, "vision_model.encoder.layers.")
        return self

    def state_dict(self):
        return dict(self._sd)

    def to(self, device=None, dtype=None):
        if device is not None and self._w and torch.device(device).type == "cuda":
            dev = torch.device(device)
            dev = torch.device("cuda", torch.cuda.current_device()) if dev.index is None else dev
            if dev != self.device:
                self.load_state_dict({k: v.to(dev) for k, v in self._sd.items()})
        return self

    def _linear(self, x: torch.Tensor, w3: torch.Tensor, bias, act: int = 0, residual=None) -> torch.Tensor:
        return linear_f32(x, w3, bias, act, residual)

    def _layers(self, x: torch.Tensor, B: int, L: int, prefix: str, causal: bool, kv_valid=None) -> List[torch.Tensor]:
        """The pre-LN CLIP encoder stack on x [B*L, d] fp32; returns [input, after layer 0, after layer 1, ...] as [B, L, d] views."""
        c, w, dev, lib = self._cfg, self._w, self.device, _lib.lib()
        d, heads, eps, rows = c["hidden_size"], c["num_attention_heads"], c["layer_norm_eps"], x.shape[0]
        hidden: List[torch.Tensor] = [x.view(B, L, d)]
        act = 1 if c["hidden_act"] == "gelu" else 2
        dh = d // heads
        for i in range(c["num_hidden_layers"]):
            p = f"{prefix}{i}."
            h = torch.empty_like(x)
            _launch(lib.alg_layer_norm_f32, dev, x.data_ptr(), h.data_ptr(), rows, d, eps, w[p + "layer_norm1.weight"].data_ptr(),
                    w[p + "layer_norm1.bias"].data_ptr())
            qkv = self._linear(h, w[p + "qkv.w3"], w[p + "qkv.bias"])
            att = torch.empty(rows, d, device=dev, dtype=torch.float32)
            small_attention(qkv, qkv[:, d:], qkv[:, 2 * d:], att, batch=B, heads=heads, head_dim=dh, n_q=L, n_kv=L,
                            q_bs=L * 3 * d, q_rs=3 * d, k_bs=L * 3 * d, k_rs=3 * d, v_bs=L * 3 * d, v_rs=3 * d, o_bs=L * d, o_rs=d,
                            scale=dh ** -0.5, causal=causal, kv_valid=kv_valid)
            x = self._linear(att, w[p + "self_attn.out_proj.w3"], w[p + "self_attn.out_proj.bias"], residual=x)
            h = torch.empty_like(x)
            _launch(lib.alg_layer_norm_f32, dev, x.data_ptr(), h.data_ptr(), rows, d, eps, w[p + "layer_norm2.weight"].data_ptr(),
                    w[p + "layer_norm2.bias"].data_ptr())
            f = self._linear(h, w[p + "mlp.fc1.w3"], w[p + "mlp.fc1.bias"], act=act)
            x = self._linear(f, w[p + "mlp.fc2.w3"], w[p + "mlp.fc2.bias"], residual=x)
            hidden.append(x.view(B, L, d))
        return hidden

    def _prepare_layers(self, w: Dict[str, torch.Tensor], prefix: str) -> None:
        for i in range(self._cfg["num_hidden_layers"]):
            p = f"{prefix}{i}."
            qkv = torch.cat([w[p + f"self_attn.{n}.weight"] for n in ("q_proj", "k_proj", "v_proj")], dim=0)
            w[p + "qkv.w3"] = self._split_weight(qkv)
            w[p + "qkv.bias"] = torch.cat([w[p + f"self_attn.{n}.bias"] for n in ("q_proj", "k_proj", "v_proj")]).contiguous()
            for n in ("self_attn.out_proj", "mlp.fc1", "mlp.fc2"):
                w[p + n + ".w3"] = self._split_weight(w[p + n + ".weight"])

    def __call__(self, pixel_values=None, output_hidden_states: bool = True, **kw):
        c, w, dev = self._cfg, self._w, self.device
        if not w:
            raise RuntimeError("CLIPVisionModel has no weights loaded")
        lib = _lib.lib()
        px = pixel_values.to(device=dev, dtype=torch.float32).contiguous()
        B, Cc, Hh, Ww = px.shape
        P, d, heads = c["patch_size"], c["hidden_size"], c["num_attention_heads"]
        if Hh != c["image_size"] or Ww != c["image_size"] or Cc != c["num_channels"]:
            raise ValueError(f"pixel_values {tuple(px.shape)}: expected [B, {c['num_channels']}, {c['image_size']}, {c['image_size']}]")
        n_patch = (Hh // P) * (Ww // P)
        patches = torch.empty(B * n_patch, self._k_patch, device=dev, dtype=torch.float32)
        _launch(lib.alg_patchify_f32, dev, px.data_ptr(), patches.data_ptr(), B, Cc, Hh, Ww, P, self._k_patch)
        emb = self._linear(patches, w["patch.w3"], None)
        L = n_patch + 1
        rows = B * L
        x = torch.empty(rows, d, device=dev, dtype=torch.float32)
        _launch(lib.alg_clip_embed_f32, dev, emb.data_ptr(), w["vision_model.embeddings.class_embedding"].data_ptr(),
                w["vision_model.embeddings.position_embedding.weight"].data_ptr(), x.data_ptr(), B, n_patch, d)
        eps = c["layer_norm_eps"]
        h = torch.empty_like(x)
        _launch(lib.alg_layer_norm_f32, dev, x.data_ptr(), h.data_ptr(), rows, d, eps, w["vision_model.pre_layrnorm.weight"].data_ptr(),
                w["vision_model.pre_layrnorm.bias"].data_ptr())
        hidden = self._layers(h, B, L, "vision_model.encoder.layers.", causal=False)
        return SimpleNamespace(hidden_states=tuple(hidden), last_hidden_state=hidden[-1])


# ======================================================================================================
# CLIP text tower (HunyuanVideo ``text_encoder_2``: the pooled prompt embedding, hy:421-452)
# ======================================================================================================
class CLIPTextModel(CLIPVisionModel):
    """Native ``CLIPTextModel``: token + position embeddings, causal pre-LN encoder, final LayerNorm, EOS pooling.  Arithmetic
    is float32 (the bf16x3 tensor-core linears of the vision tower), a superset of the fp16 the reference loads it in."""

    HF_DEFAULTS = dict(vocab_size=49408, hidden_size=512, intermediate_size=2048, num_hidden_layers=12, num_attention_heads=8,
                       max_position_embeddings=77, hidden_act="quick_gelu", layer_norm_eps=1e-5, eos_token_id=2)

    def parameter_shapes(self) -> Dict[str, tuple]:
        c = self._cfg
        d, f = c["hidden_size"], c["intermediate_size"]
        s = {"text_model.embeddings.token_embedding.weight": (c["vocab_size"], d),
             "text_model.embeddings.position_embedding.weight": (c["max_position_embeddings"], d),
             "text_model.final_layer_norm.weight": (d,), "text_model.final_layer_norm.bias": (d,)}
        for i in range(c["num_hidden_layers"]):
            p = f"text_model.encoder.layers.{i}."
            for n in ("q_proj", "k_proj", "v_proj", "out_proj"):
                s[p + f"self_attn.{n}.weight"], s[p + f"self_attn.{n}.bias"] = (d, d), (d,)
            for n in ("layer_norm1", "layer_norm2"):
                s[p + n + ".weight"], s[p + n + ".bias"] = (d,), (d,)
            s[p + "mlp.fc1.weight"], s[p + "mlp.fc1.bias"] = (f, d), (f,)
            s[p + "mlp.fc2.weight"], s[p + "mlp.fc2.bias"] = (d, f), (d,)
        return s

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path, subfolder: str = "text_encoder_2", torch_dtype=None, cache_dir=None,
                        device="cuda"):
        import os

        from . import checkpoint
        snap = checkpoint.resolve_snapshot(str(pretrained_model_name_or_path), cache_dir, require=subfolder or "transformer")
        if snap is None:
            raise FileNotFoundError(f"no local snapshot for {pretrained_model_name_or_path!r} (there is no network)")
        folder = os.path.join(snap, subfolder) if subfolder else snap
        cfg = checkpoint.read_config(folder)
        cfg = cfg.get("text_config", cfg)
        return cls(**cfg).load_state_dict(checkpoint.load_safetensors_dir(folder, device))

    @classmethod
    def from_synthetic(cls, seed: int = 0, device="cuda", **config):
        m = cls(**config)
        sd = {}
        for idx, (name, shape) in enumerate(m.parameter_shapes().items()):
            g = torch.Generator(device=device).manual_seed(seed * 1_000_003 + 5_001 + idx)
            if "norm" in name and name.endswith(".weight"):
                w = 1 + 0.1 * torch.randn(shape, generator=g, device=device)
            elif name.endswith(".bias"):
                w = 0.02 * torch.randn(shape, generator=g, device=device)
            elif "embedding" in name:
                w = 0.5 * torch.randn(shape, generator=g, device=device)
            else:
                w = torch.randn(shape, generator=g, device=device) * (shape[1] ** -0.5)
            sd[name] = w.float()
        return m.load_state_dict(sd)

    def load_state_dict(self, sd: Dict[str, torch.Tensor]):
        shapes = self.parameter_shapes()
        missing = [k for k in shapes if k not in sd]
        if missing:
            raise KeyError(f"missing parameters: {missing[:4]}{'...' if len(missing) > 4 else ''}")
        dev = sd["text_model.final_layer_norm.weight"].device
        if dev.type != "cuda":
            raise RuntimeError("encoder weights must live on a CUDA device (no CPU fallback)")
        self.device = dev
        self._sd = {k: sd[k].to(device=dev, dtype=torch.float32).contiguous() for k in shapes}
        self._w = dict(self._sd)
        self._prepare_layers(self._w, "text_model.encoder.layers.")
        return self

    def __call__(self, input_ids=None, attention_mask=None, output_hidden_states: bool = False, **kw):
        c, w, dev, lib = self._cfg, self._w, self.device, _lib.lib()
        if not w:
            raise RuntimeError("CLIPTextModel has no weights loaded")
        ids = input_ids.to(dev).to(torch.int64).contiguous()
        B, L = ids.shape
        d = c["hidden_size"]
        if L > c["max_position_embeddings"]:
            raise ValueError(f"sequence length {L} exceeds max_position_embeddings {c['max_position_embeddings']}")
        x = torch.empty(B * L, d, device=dev, dtype=torch.float32)
        tok = w["text_model.embeddings.token_embedding.weight"]
        # fp32 rows moved as 2 d 16-bit lanes (the gather is a byte copy)
        _launch(lib.alg_gather_rows_bf16, dev, tok.data_ptr(), c["vocab_size"], ids.data_ptr(), x.data_ptr(), B * L, 2 * d)
        pos = w["text_model.embeddings.position_embedding.weight"]
        for b in range(B):
            _launch(lib.alg_bias_act_f32, dev, x[b * L:].data_ptr(), None, pos.data_ptr(), L, d, 0)
        valid = None if attention_mask is None else attention_mask.to(dev).gt(0).sum(dim=1).to(torch.int32).contiguous()
        hidden = self._layers(x, B, L, "text_model.encoder.layers.", causal=True, kv_valid=valid)
        last = torch.empty(B * L, d, device=dev, dtype=torch.float32)
        _launch(lib.alg_layer_norm_f32, dev, hidden[-1].data_ptr(), last.data_ptr(), B * L, d, c["layer_norm_eps"],
                w["text_model.final_layer_norm.weight"].data_ptr(), w["text_model.final_layer_norm.bias"].data_ptr())
        # pooled = the hidden state at the end-of-text token: argmax of the ids for the legacy eos id 2, else the first eos
        host = ids.cpu()
        if c["eos_token_id"] == 2:
            eos = host.argmax(dim=-1)
        else:
            eos = (host == c["eos_token_id"]).int().argmax(dim=-1)
        rows = (torch.arange(B) * L + eos).to(dev)
        pooled = torch.empty(B, d, device=dev, dtype=torch.float32)
        _launch(lib.alg_gather_rows_bf16, dev, last.data_ptr(), B * L, rows.data_ptr(), pooled.data_ptr(), B, 2 * d)
        return SimpleNamespace(last_hidden_state=last.view(B, L, d), pooler_output=pooled,
                               hidden_states=tuple(hidden) if output_hidden_states else None)
