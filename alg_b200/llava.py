"""LLaVA-Llama-3 prompt encoder of HunyuanVideo-I2V on the sm_100a kernels of ``libalg_b200.so``.

The reference calls ``self.text_encoder(**expanded_inputs, pixel_values=..., output_hidden_states=True).hidden_states[-3]``
(hy:333-337) on a ``transformers.LlavaForConditionalGeneration``: CLIP-ViT-L/14-336 vision tower -> two-layer GELU projector ->
the projected patch features replace the ``<image>`` token embeddings -> Llama-3-8B decoder stack (RMSNorm, rotate-half RoPE
from ``position_ids``, grouped-query causal attention under the prompt's padding mask, SwiGLU).  This module keeps that call
surface (``config.image_token_index`` / ``config.pad_token_id``, ``hidden_states`` tuple with the final-norm state last) and
the HF parameter names of both the 4.48 layout the HunyuanVideo snapshots were saved with (``language_model.model.layers...``)
and the current one (``model.language_model.layers...``).  It only SEQUENCES C-ABI calls:

    nn.Embedding                    alg_gather_rows_bf16 (fp32 rows moved as 16-bit lanes)
    LlamaRMSNorm                    alg_rms_norm_f32
    every nn.Linear                 alg_split3_bf16 + alg_gemm_bf16 (+ alg_bias_act_f32): fp32 product on the bf16 tensor
                                    cores (hi*hi + hi*lo + lo*hi, fp32 accumulation); q|k|v and gate|up fused into one launch
    apply_rotary_pos_emb            alg_rope_half_f32 (cos / sin table from position_ids, fp32 like LlamaRotaryEmbedding)
    attention                       alg_small_attention, streamed-key variant: fp32, causal, kv_valid / key_mask, kv_group
    silu(gate) * up                 alg_swiglu_f32
    vision tower                    alg_b200.encoders.CLIPVisionModel (fp32)

Arithmetic is float32 throughout -- a superset of the float16 the reference loads this encoder in (run.py:76-80): fp16
weights split exactly into bf16 hi + lo, and the result is closer to an fp32 evaluation than transformers' own fp16 forward
(``tests/test_gpu_llava.py`` pins it against the installed ``LlavaForConditionalGeneration`` with seeded weights).
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import Dict, List, Optional

import torch

from . import _lib
from .encoders import CLIPVisionModel, _launch, linear_f32, small_attention, split_weight

# xtuner/llava-llama-3-8b-v1_1-transformers, the text_encoder of hunyuanvideo-community/HunyuanVideo-I2V
LLAMA3_8B = dict(vocab_size=128320, hidden_size=4096, intermediate_size=14336, num_hidden_layers=32, num_attention_heads=32,
                 num_key_value_heads=8, rms_norm_eps=1e-5, rope_theta=500000.0, max_position_embeddings=8192, hidden_act="silu")
CLIP_VIT_L_14_336 = dict(hidden_size=1024, intermediate_size=4096, num_hidden_layers=24, num_attention_heads=16, image_size=336,
                         patch_size=14, num_channels=3, hidden_act="quick_gelu", layer_norm_eps=1e-5)
LLAVA_LLAMA3_8B = dict(text_config=LLAMA3_8B, vision_config=CLIP_VIT_L_14_336, image_token_index=128257, pad_token_id=128258,
                       vision_feature_layer=-2, vision_feature_select_strategy="default", projector_hidden_act="gelu",
                       multimodal_projector_bias=True)

_TEXT_KEYS = ("vocab_size", "hidden_size", "intermediate_size", "num_hidden_layers", "num_attention_heads", "num_key_value_heads",
              "rms_norm_eps", "rope_theta", "max_position_embeddings", "hidden_act", "head_dim", "attention_bias", "mlp_bias")
_VISION_KEYS = tuple(CLIP_VIT_L_14_336)


def _canonical(name: str) -> str:
    """transformers 4.48 parameter names -> the current (5.x) layout."""
    if name.startswith("language_model.model."):
        return "model.language_model." + name[len("language_model.model."):]
    if name.startswith("language_model.lm_head."):
        return name[len("language_model."):]
    if name.startswith(("vision_tower.", "multi_modal_projector.")):
        return "model." + name
    return name


class LlavaForConditionalGeneration:
    """Native LLaVA prompt encoder; see the module docstring."""

    def __init__(self, text_config: Optional[dict] = None, vision_config: Optional[dict] = None, image_token_index: int = 128257,
                 pad_token_id: int = 128258, vision_feature_layer: int = -2, vision_feature_select_strategy: str = "default",
                 projector_hidden_act: str = "gelu", multimodal_projector_bias: bool = True, image_token_id: Optional[int] = None,
                 torch_dtype=None, **unused):
        t = dict(LLAMA3_8B)
        t.update({k: v for k, v in (text_config or {}).items() if k in _TEXT_KEYS and v is not None})
        rp = (text_config or {}).get("rope_parameters") or (text_config or {}).get("rope_scaling")
        if isinstance(rp, dict):
            if rp.get("rope_type", rp.get("type", "default")) not in ("default", None):
                raise NotImplementedError(f"rope type {rp.get('rope_type', rp.get('type'))!r}: the LLaVA-Llama-3 encoder uses the default RoPE")
            if rp.get("rope_theta") is not None:
                t["rope_theta"] = rp["rope_theta"]
        if t.get("hidden_act", "silu") != "silu":
            raise NotImplementedError("Llama MLP activation other than silu")
        if isinstance(vision_feature_layer, (list, tuple)):
            raise NotImplementedError("a list of vision feature layers (LLaVA-Llama-3 uses the single layer -2)")
        if projector_hidden_act not in ("gelu", "quick_gelu"):
            raise NotImplementedError(f"projector activation {projector_hidden_act!r}")
        v = dict(CLIP_VIT_L_14_336)
        v.update({k: val for k, val in (vision_config or {}).items() if k in _VISION_KEYS and val is not None})
        t.setdefault("head_dim", None)
        t["head_dim"] = t["head_dim"] or t["hidden_size"] // t["num_attention_heads"]
        self._t, self._v = t, v
        self._proj_act = 1 if projector_hidden_act == "gelu" else 2
        self._proj_bias = bool(multimodal_projector_bias)
        self._feature_layer, self._select = int(vision_feature_layer), vision_feature_select_strategy
        itok = image_token_id if image_token_id is not None else image_token_index
        self.config = SimpleNamespace(image_token_index=itok, image_token_id=itok, pad_token_id=pad_token_id,
                                      text_config=SimpleNamespace(**t), vision_config=SimpleNamespace(**v),
                                      vision_feature_layer=vision_feature_layer,
                                      vision_feature_select_strategy=vision_feature_select_strategy)
        self.dtype = torch_dtype or torch.float32  # what callers cast the result to (run.py loads the encoder as float16)
        self.device = torch.device("cpu")
        self.vision_tower = CLIPVisionModel(**v)
        self._w: Dict[str, torch.Tensor] = {}

    # ---------------------------------------------------------------------------------------------------------------
    def parameter_shapes(self) -> Dict[str, tuple]:
        t = self._t
        d, f, dh = t["hidden_size"], t["intermediate_size"], t["head_dim"]
        H, Hk = t["num_attention_heads"], t["num_key_value_heads"]
        s = {"model.vision_tower." + k: v for k, v in self.vision_tower.parameter_shapes().items()}
        s["model.multi_modal_projector.linear_1.weight"] = (d, self._v["hidden_size"])
        s["model.multi_modal_projector.linear_2.weight"] = (d, d)
        if self._proj_bias:
            s["model.multi_modal_projector.linear_1.bias"] = (d,)
            s["model.multi_modal_projector.linear_2.bias"] = (d,)
        s["model.language_model.embed_tokens.weight"] = (t["vocab_size"], d)
        for i in range(t["num_hidden_layers"]):
            p = f"model.language_model.layers.{i}."
            s[p + "self_attn.q_proj.weight"] = (H * dh, d)
            s[p + "self_attn.k_proj.weight"] = (Hk * dh, d)
            s[p + "self_attn.v_proj.weight"] = (Hk * dh, d)
            s[p + "self_attn.o_proj.weight"] = (d, H * dh)
            s[p + "mlp.gate_proj.weight"], s[p + "mlp.up_proj.weight"], s[p + "mlp.down_proj.weight"] = (f, d), (f, d), (d, f)
            s[p + "input_layernorm.weight"], s[p + "post_attention_layernorm.weight"] = (d,), (d,)
        s["model.language_model.norm.weight"] = (d,)
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
        return cls(torch_dtype=torch_dtype, **checkpoint.read_config(folder)).load_state_dict(checkpoint.load_safetensors_dir(folder, device))

    @classmethod
    def from_synthetic(cls, seed: int = 0, device="cuda", **config):
        """Seeded weights at the LLaVA-Llama-3-8B shape unless ``config`` says otherwise."""
        m = cls(**dict(LLAVA_LLAMA3_8B, **config))
        sd = {}
        for idx, (name, shape) in enumerate(m.parameter_shapes().items()):
            g = torch.Generator(device=device).manual_seed(seed * 1_000_003 + 9_001 + idx)
            if "norm" in name and name.endswith(".weight"):
                w = 1 + 0.1 * torch.randn(shape, generator=g, device=device)
            elif name.endswith(".bias"):
                w = 0.02 * torch.randn(shape, generator=g, device=device)
            elif "embed" in name:
                w = 0.5 * torch.randn(shape, generator=g, device=device)
            else:
                fan_in = 1
                for s_ in shape[1:]:
                    fan_in *= s_
                w = torch.randn(shape, generator=g, device=device) * (fan_in ** -0.5)
            sd[name] = w
        return m.load_state_dict(sd)

    def load_state_dict(self, sd: Dict[str, torch.Tensor]):
        sd = {_canonical(k): v for k, v in sd.items()}
        shapes = self.parameter_shapes()
        missing = [k for k in shapes if k not in sd]
        if missing:
            raise KeyError(f"missing parameters: {missing[:4]}{'...' if len(missing) > 4 else ''}")
        dev = sd["model.language_model.norm.weight"].device
        if dev.type != "cuda":
            raise RuntimeError("encoder weights must live on a CUDA device (no CPU fallback)")
        self.device = dev
        vt = "model.vision_tower."
        self.vision_tower.load_state_dict({k[len(vt):]: v for k, v in sd.items() if k.startswith(vt)})
        t, w = self._t, {}
        f32 = lambda k: sd[k].to(device=dev, dtype=torch.float32).contiguous()
        for n in ("linear_1", "linear_2"):
            p = f"model.multi_modal_projector.{n}."
            w[p + "w3"] = split_weight(sd[p + "weight"].to(dev))
            w[p + "bias"] = f32(p + "bias") if self._proj_bias else None
        w["embed"] = f32("model.language_model.embed_tokens.weight")
        w["norm"] = f32("model.language_model.norm.weight")
        for i in range(t["num_hidden_layers"]):
            p = f"model.language_model.layers.{i}."
            # the big linears are kept ONLY as their bf16 [hi | lo | hi] split (6 bytes per parameter; state_dict() rebuilds fp32)
            w[p + "qkv.w3"] = split_weight(torch.cat([sd[p + f"self_attn.{n}_proj.weight"].to(dev).float() for n in "qkv"], dim=0))
            w[p + "o.w3"] = split_weight(sd[p + "self_attn.o_proj.weight"].to(dev))
            w[p + "gate_up.w3"] = split_weight(torch.cat([sd[p + f"mlp.{n}_proj.weight"].to(dev).float() for n in ("gate", "up")], dim=0))
            w[p + "down.w3"] = split_weight(sd[p + "mlp.down_proj.weight"].to(dev))
            w[p + "ln1"], w[p + "ln2"] = f32(p + "input_layernorm.weight"), f32(p + "post_attention_layernorm.weight")
        inv = 1.0 / (t["rope_theta"] ** (torch.arange(0, t["head_dim"], 2, dtype=torch.int64).to(dtype=torch.float) / t["head_dim"]))
        w["inv_freq"] = inv.to(dev)
        self._w = w
        return self

    def state_dict(self) -> Dict[str, torch.Tensor]:
        """fp32 parameters under the current transformers names (big linears rebuilt as hi + lo of their stored split)."""
        t, w = self._t, self._w

        def unsplit(w3):
            K = w3.shape[1] // 3
            return w3[:, :K].float() + w3[:, K:2 * K].float()

        d, dh, H, Hk, f = t["hidden_size"], t["head_dim"], t["num_attention_heads"], t["num_key_value_heads"], t["intermediate_size"]
        sd = {"model.vision_tower." + k: v for k, v in self.vision_tower.state_dict().items()}
        for n in ("linear_1", "linear_2"):
            p = f"model.multi_modal_projector.{n}."
            sd[p + "weight"] = unsplit(w[p + "w3"])
            if self._proj_bias:
                sd[p + "bias"] = w[p + "bias"]
        sd["model.language_model.embed_tokens.weight"] = w["embed"]
        for i in range(t["num_hidden_layers"]):
            p = f"model.language_model.layers.{i}."
            qkv = unsplit(w[p + "qkv.w3"])
            sd[p + "self_attn.q_proj.weight"], sd[p + "self_attn.k_proj.weight"], sd[p + "self_attn.v_proj.weight"] = \
                qkv[:H * dh], qkv[H * dh:(H + Hk) * dh], qkv[(H + Hk) * dh:]
            sd[p + "self_attn.o_proj.weight"] = unsplit(w[p + "o.w3"])
            gu = unsplit(w[p + "gate_up.w3"])
            sd[p + "mlp.gate_proj.weight"], sd[p + "mlp.up_proj.weight"] = gu[:f], gu[f:]
            sd[p + "mlp.down_proj.weight"] = unsplit(w[p + "down.w3"])
            sd[p + "input_layernorm.weight"], sd[p + "post_attention_layernorm.weight"] = w[p + "ln1"], w[p + "ln2"]
        sd["model.language_model.norm.weight"] = w["norm"]
        return sd

    def to(self, device=None, dtype=None):
        if device is not None and self._w and torch.device(device).type == "cuda":
            dev = torch.device(device)
            dev = torch.device("cuda", torch.cuda.current_device()) if dev.index is None else dev
            if dev != self.device:
                self.load_state_dict({k: v.to(dev) for k, v in self.state_dict().items()})
        return self

    # ---------------------------------------------------------------------------------------------------------------
    def get_image_features(self, pixel_values: torch.Tensor) -> torch.Tensor:
        """vision tower hidden state ``vision_feature_layer`` (CLS dropped under "default") through the projector:
        [n_images * patches, hidden_size] fp32 (modeling_llava.py ``get_image_features``)."""
        w = self._w
        hs = self.vision_tower(pixel_values=pixel_values, output_hidden_states=True).hidden_states[self._feature_layer]
        if self._select == "default":
            hs = hs[:, 1:]
        x = hs.reshape(-1, hs.shape[-1]).contiguous()
        p = "model.multi_modal_projector."
        x = linear_f32(x, w[p + "linear_1.w3"], w[p + "linear_1.bias"], act=self._proj_act)
        return linear_f32(x, w[p + "linear_2.w3"], w[p + "linear_2.bias"])

    def __call__(self, input_ids=None, attention_mask=None, position_ids=None, pixel_values=None,
                 output_hidden_states: bool = True, **kw):
        t, w, dev, lib = self._t, self._w, self.device, _lib.lib()
        if not w:
            raise RuntimeError("LlavaForConditionalGeneration has no weights loaded")
        ids = input_ids.to(dev).to(torch.int64).contiguous()
        B, L = ids.shape
        d, dh, H, Hk, f = t["hidden_size"], t["head_dim"], t["num_attention_heads"], t["num_key_value_heads"], t["intermediate_size"]
        rows = B * L
        x = torch.empty(rows, d, device=dev, dtype=torch.float32)
        _launch(lib.alg_gather_rows_bf16, dev, w["embed"].data_ptr(), t["vocab_size"], ids.data_ptr(), x.data_ptr(), rows, 2 * d)
        if pixel_values is not None:
            feats = self.get_image_features(pixel_values.to(dev).float())
            where = (ids.view(-1) == self.config.image_token_index).nonzero().view(-1)
            if where.numel() != feats.shape[0]:
                raise ValueError(f"Image features and image tokens do not match, tokens: {where.numel()}, features: {feats.shape[0]}")
            x.index_copy_(0, where, feats)  # masked_scatter: the k-th <image> token takes the k-th feature row
        # masks: right-padded prompts (the tokenizer pads on the right, hy:318-327) are a per-sample key count; anything else
        # goes to the kernel as a general key mask
        kv_valid = key_mask = None
        if attention_mask is not None:
            m = attention_mask.to(dev).gt(0)
            n = m.sum(dim=1)
            if bool((m == (torch.arange(L, device=dev)[None] < n[:, None])).all()):
                kv_valid = n.to(torch.int32).contiguous()
            else:
                key_mask = m.to(torch.uint8).contiguous()
        if position_ids is None:
            position_ids = torch.arange(L, device=dev)[None].expand(B, L)
        # LlamaRotaryEmbedding: freqs = inv_freq x position (fp32), cos / sin of [freqs | freqs]
        freqs = position_ids.to(dev).reshape(rows, 1).float() * w["inv_freq"][None, :]
        emb = torch.cat((freqs, freqs), dim=-1)
        cos, sin = emb.cos().contiguous(), emb.sin().contiguous()
        ld = (H + 2 * Hk) * dh
        hidden: List[torch.Tensor] = [x.view(B, L, d)]
        for i in range(t["num_hidden_layers"]):
            p = f"model.language_model.layers.{i}."
            h = torch.empty_like(x)
            _launch(lib.alg_rms_norm_f32, dev, x.data_ptr(), h.data_ptr(), rows, d, t["rms_norm_eps"], w[p + "ln1"].data_ptr())
            qkv = linear_f32(h, w[p + "qkv.w3"], None)
            assert qkv.stride(0) == ld
            _launch(lib.alg_rope_half_f32, dev, qkv.data_ptr(), ld, cos.data_ptr(), sin.data_ptr(), rows, H + Hk, dh)  # q heads, then k heads
            att = torch.empty(rows, H * dh, device=dev, dtype=torch.float32)
            small_attention(qkv, qkv[:, H * dh:], qkv[:, (H + Hk) * dh:], att, batch=B, heads=H, head_dim=dh, n_q=L, n_kv=L,
                            q_bs=L * ld, q_rs=ld, k_bs=L * ld, k_rs=ld, v_bs=L * ld, v_rs=ld, o_bs=L * H * dh, o_rs=H * dh,
                            scale=dh ** -0.5, causal=True, kv_valid=kv_valid, key_mask=key_mask, kv_group=H // Hk)
            x = linear_f32(att, w[p + "o.w3"], None, residual=x)
            h = torch.empty_like(x)
            _launch(lib.alg_rms_norm_f32, dev, x.data_ptr(), h.data_ptr(), rows, d, t["rms_norm_eps"], w[p + "ln2"].data_ptr())
            gu = linear_f32(h, w[p + "gate_up.w3"], None)
            assert gu.stride(0) == 2 * f
            a = torch.empty(rows, f, device=dev, dtype=torch.float32)
            _launch(lib.alg_swiglu_f32, dev, gu.data_ptr(), a.data_ptr(), rows, f)
            x = linear_f32(a, w[p + "down.w3"], None, residual=x)
            hidden.append(x.view(B, L, d))
        last = torch.empty_like(x)
        _launch(lib.alg_rms_norm_f32, dev, x.data_ptr(), last.data_ptr(), rows, d, t["rms_norm_eps"], w["norm"].data_ptr())
        hidden[-1] = last.view(B, L, d)  # transformers reports the final-norm state as the last hidden state
        return SimpleNamespace(last_hidden_state=hidden[-1], hidden_states=tuple(hidden) if output_hidden_states else None)
