"""``WanTransformer3DModel`` host wrapper over the native DiT engine (``alg_wan_*`` in libalg_b200.so).

Mirrors the interface the reference pipeline uses on ``self.transformer`` (wan:804-806, 910-917): ``.config``,
``.dtype``, ``__call__(hidden_states=, timestep=, encoder_hidden_states=, encoder_hidden_states_image=,
attention_kwargs=, return_dict=False)``.  Parameters are ordinary torch tensors named like the diffusers
state_dict; the engine only borrows their pointers.  The arithmetic is restated from diffusers@be2fb77
``transformer_wan.py`` (not available offline -- parity unpinned).
"""
from __future__ import annotations

import ctypes as C
from types import SimpleNamespace
from typing import Dict, List, Optional, Sequence

import torch

from . import _lib

FP32_KEYS = ("time_embedder", "scale_shift_table", "norm1", "norm2", "norm3")  # diffusers _keep_in_fp32_modules

WAN_I2V_14B = dict(patch_size=(1, 2, 2), num_attention_heads=40, attention_head_dim=128, in_channels=36,
                   out_channels=16, text_dim=4096, freq_dim=256, ffn_dim=13824, num_layers=40, cross_attn_norm=True,
                   qk_norm="rms_norm_across_heads", eps=1e-6, image_dim=1280, added_kv_proj_dim=5120,
                   rope_max_seq_len=1024, text_len=512)


def parameter_shapes(cfg: dict) -> Dict[str, tuple]:
    """name -> shape of every parameter of the configured model (diffusers naming)."""
    d = cfg["num_attention_heads"] * cfg["attention_head_dim"]
    pt, ph, pw = cfg["patch_size"]
    s: Dict[str, tuple] = {}

    def lin(name, o, i):
        s[name + ".weight"] = (o, i)
        s[name + ".bias"] = (o,)

    s["patch_embedding.weight"] = (d, cfg["in_channels"], pt, ph, pw)
    s["patch_embedding.bias"] = (d,)
    lin("condition_embedder.time_embedder.linear_1", d, cfg["freq_dim"])
    lin("condition_embedder.time_embedder.linear_2", d, d)
    lin("condition_embedder.time_proj", 6 * d, d)
    lin("condition_embedder.text_embedder.linear_1", d, cfg["text_dim"])
    lin("condition_embedder.text_embedder.linear_2", d, d)
    if cfg.get("image_dim"):
        ie, idim = "condition_embedder.image_embedder.", cfg["image_dim"]
        s[ie + "norm1.weight"] = s[ie + "norm1.bias"] = (idim,)
        lin(ie + "ff.net.0.proj", idim, idim)
        lin(ie + "ff.net.2", d, idim)
        s[ie + "norm2.weight"] = s[ie + "norm2.bias"] = (d,)
    for i in range(cfg["num_layers"]):
        p = f"blocks.{i}."
        s[p + "scale_shift_table"] = (1, 6, d)
        for a in ("attn1", "attn2"):
            for n in ("to_q", "to_k", "to_v", "to_out.0"):
                lin(p + f"{a}.{n}", d, d)
            s[p + f"{a}.norm_q.weight"] = s[p + f"{a}.norm_k.weight"] = (d,)
        if cfg.get("image_dim"):
            lin(p + "attn2.add_k_proj", d, d)
            lin(p + "attn2.add_v_proj", d, d)
            s[p + "attn2.norm_added_k.weight"] = (d,)
        s[p + "norm2.weight"] = s[p + "norm2.bias"] = (d,)
        lin(p + "ffn.net.0.proj", cfg["ffn_dim"], d)
        lin(p + "ffn.net.2", d, cfg["ffn_dim"])
    s["scale_shift_table"] = (1, 2, d)
    lin("proj_out", cfg["out_channels"] * pt * ph * pw, d)
    return s


def synthetic_state_dict(cfg: dict, seed: int = 0, device="cuda", std: float = 0.02) -> Dict[str, torch.Tensor]:
    """Seeded random-init weights at the true shapes (no checkpoints / network offline), generated on ``device``.

    Deterministic per (seed, parameter name): every rank regenerates identical tensors, or rank 0 broadcasts them.
    """
    sd = {}
    for idx, (name, shape) in enumerate(parameter_shapes(cfg).items()):
        g = torch.Generator(device=device).manual_seed(seed * 1_000_003 + idx)
        keep32 = any(k in name for k in FP32_KEYS)
        if name.endswith("scale_shift_table"):
            w = torch.randn(shape, generator=g, device=device) / shape[-1] ** 0.5
        elif ("norm" in name) and name.endswith(".weight"):
            w = 1 + 0.1 * torch.randn(shape, generator=g, device=device)
        elif ("norm" in name) and name.endswith(".bias"):
            w = 0.1 * torch.randn(shape, generator=g, device=device)
        elif name == "patch_embedding.weight":
            w = torch.randn(shape, generator=g, device=device) * (std * 4)
        else:
            w = torch.randn(shape, generator=g, device=device) * std
        sd[name] = w.to(torch.float32 if keep32 else torch.bfloat16)
    return sd


class WanTransformer3DModel:
    """Native-engine stand-in for diffusers' ``WanTransformer3DModel`` (inference only)."""

    def __init__(self, **config):
        cfg = dict(WAN_I2V_14B)
        cfg.update(config)
        self.config = SimpleNamespace(**cfg)
        self._cfg = cfg
        self._handle = C.c_void_p()
        self._state: Dict[str, torch.Tensor] = {}
        self._workspace: Optional[torch.Tensor] = None
        self._debug: Optional[torch.Tensor] = None
        self.dtype = torch.bfloat16
        self.device = torch.device("cpu")
        c = _lib.WanConfig()
        c.num_heads, c.head_dim = cfg["num_attention_heads"], cfg["attention_head_dim"]
        c.in_channels, c.out_channels = cfg["in_channels"], cfg["out_channels"]
        c.text_dim, c.freq_dim, c.ffn_dim = cfg["text_dim"], cfg["freq_dim"], cfg["ffn_dim"]
        c.num_layers, c.image_dim, c.text_len = cfg["num_layers"], cfg.get("image_dim") or 0, cfg["text_len"]
        c.patch_t, c.patch_h, c.patch_w = cfg["patch_size"]
        c.rope_max_seq_len, c.eps = cfg["rope_max_seq_len"], cfg["eps"]
        self._c = c

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
        m.load_state_dict(synthetic_state_dict(m._cfg, seed=seed, device=device))
        return m

    def load_state_dict(self, sd: Dict[str, torch.Tensor]):
        shapes = parameter_shapes(self._cfg)
        missing = [k for k in shapes if k not in sd]
        if missing:
            raise KeyError(f"missing parameters: {missing[:4]}{'...' if len(missing) > 4 else ''}")
        dev = next(iter(sd.values())).device
        if dev.type != "cuda":
            raise RuntimeError("WanTransformer3DModel weights must live on a CUDA device (no CPU fallback)")
        self.device = dev
        with torch.cuda.device(dev):
            if self._handle:
                _lib.lib().alg_wan_destroy(self._handle)
                self._handle = C.c_void_p()
            _lib.check(_lib.lib().alg_wan_create(C.byref(self._c), C.byref(self._handle)))
            for name, shape in shapes.items():
                t = sd[name]
                if tuple(t.shape) != tuple(shape):
                    raise ValueError(f"{name}: expected shape {tuple(shape)}, got {tuple(t.shape)}")
                want = torch.float32 if any(k in name for k in FP32_KEYS) else torch.bfloat16
                t = t.to(device=dev, dtype=want).contiguous()
                self._state[name] = t  # keeps the memory alive: the engine only stores the pointer
                _lib.check(_lib.lib().alg_wan_set_weight(self._handle, name.encode(), t.data_ptr(), t.numel(),
                                                        _lib.dtype_code(t.dtype)))
            _lib.check(_lib.lib().alg_wan_weights_complete(self._handle))
        return self

    def state_dict(self):
        return dict(self._state)

    def to(self, device=None, dtype=None):
        if device is not None and self._state:
            dev = torch.device(device)
            if dev.type == "cuda" and dev.index is None:
                dev = torch.device("cuda", torch.cuda.current_device())
            if dev != self.device:
                self.load_state_dict({k: v.to(dev) for k, v in self._state.items()})
        return self

    def __del__(self):
        try:
            if self._handle:
                _lib.lib().alg_wan_destroy(self._handle)
        except Exception:
            pass

    # ---- forward ------------------------------------------------------------------------------------
    def enable_debug(self, nbytes: int):
        self._debug = torch.zeros(nbytes, dtype=torch.uint8, device=self.device)
        _lib.check(_lib.lib().alg_wan_set_debug_buffer(self._handle, self._debug.data_ptr(), nbytes))
        return self._debug

    PROFILE_CLASSES = ("self_attention", "cross_attention", "gemm", "elementwise")

    def context_cache(self, enable: bool = True):
        """Memoise the step-invariant cross-attention context (text / image embedders + K, V^T of all layers) across forwards that
        receive the SAME prompt / image tensors (alg_wan_context_cache: keyed on their addresses).  Calling this -- with either
        value -- drops what was memoised: call it again whenever the contents of those tensors change."""
        _lib.check(_lib.lib().alg_wan_context_cache(self._handle, int(enable)))
        self._ctx_cache_on = bool(enable)

    def profile(self, enable: bool = True):
        """Per-kernel-class device timing of the forwards that follow (CUDA events on the launching stream)."""
        _lib.check(_lib.lib().alg_wan_profile(self._handle, int(enable)))

    def profile_read(self) -> Dict[str, dict]:
        ms = (C.c_float * 4)()
        n = (C.c_int32 * 4)()
        _lib.check(_lib.lib().alg_wan_profile_read(self._handle, ms, n, 4))
        return {k: {"ms": float(ms[i]), "launches": int(n[i])} for i, k in enumerate(self.PROFILE_CLASSES)}

    def forward_passes(self, latents: Sequence[torch.Tensor], cond: Sequence[torch.Tensor],
                       text: Sequence[torch.Tensor], image: Optional[torch.Tensor], timestep: int,
                       out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """The 1-3 CFG passes of one denoise step (wan:882-917) without materialising the batched model input.

        latents[p] [16, T, H, W] fp32, cond[p] [20, T, H, W] fp32, text[p] [512, text_dim] bf16,
        image [n_img, image_dim] bf16 -> noise [n_pass, 16, T, H, W] bf16.
        """
        n_pass = len(cond)
        _lib.require_cuda(*latents, *cond, *text, image)
        _, T, H, W = cond[0].shape[-4:]
        # the engine reads fixed extents through raw pointers: refuse anything else instead of reading past a buffer
        cfg = self._cfg
        if not 1 <= n_pass <= 3 or len(latents) != n_pass or len(text) != n_pass:
            raise ValueError(f"forward_passes: 1-3 passes with one latent / condition / text each, got {len(latents)}/{n_pass}/{len(text)}")
        for p in range(n_pass):
            lat_c, cond_c = latents[p].numel() // (T * H * W), cond[p].numel() // (T * H * W)
            if lat_c != cfg["out_channels"] or latents[p].numel() != lat_c * T * H * W:
                raise ValueError(f"latents[{p}]: expected {cfg['out_channels']} x {T} x {H} x {W}, got {tuple(latents[p].shape)}")
            if cond_c != cfg["in_channels"] - cfg["out_channels"] or cond[p].numel() != cond_c * T * H * W:
                raise ValueError(f"cond[{p}]: expected {cfg['in_channels'] - cfg['out_channels']} x {T} x {H} x {W}, got {tuple(cond[p].shape)}")
            if tuple(text[p].reshape(-1, text[p].shape[-1]).shape) != (cfg["text_len"], cfg["text_dim"]):
                raise ValueError(f"text[{p}]: the attention processor splits the context at text_len = {cfg['text_len']} rows of "
                                 f"{cfg['text_dim']} (wan: max_sequence_length), got {tuple(text[p].shape)}")
        if (image is None) != (not cfg.get("image_dim")) or (image is not None and image.shape[-1] != cfg["image_dim"]):
            raise ValueError(f"image context: expected {'[n, %d]' % cfg['image_dim'] if cfg.get('image_dim') else 'None'}, got "
                             f"{None if image is None else tuple(image.shape)}")
        if H % cfg["patch_size"][1] or W % cfg["patch_size"][2] or T % cfg["patch_size"][0]:
            raise ValueError(f"latent grid {T} x {H} x {W} is not a multiple of the patch size {cfg['patch_size']}")
        keep: List[torch.Tensor] = []

        def prep(t, dt):
            t = t.to(dtype=dt).contiguous()
            keep.append(t)
            return t.data_ptr()

        lat_p = (C.c_void_p * 3)(*[prep(t.reshape(-1, T, H, W), torch.float32) for t in latents])
        cond_p = (C.c_void_p * 3)(*[prep(t.reshape(-1, T, H, W), torch.float32) for t in cond])
        text_p = (C.c_void_p * 3)(*[prep(t.reshape(-1, t.shape[-1]), torch.bfloat16) for t in text])
        n_img = 0 if image is None else image.reshape(-1, image.shape[-1]).shape[0]
        img_ptr = None if image is None else prep(image.reshape(-1, image.shape[-1]), torch.bfloat16)
        if getattr(self, "_ctx_cache_on", False):
            # the context cache is keyed on addresses: a conditioning tensor that had to be cast / compacted lives in a temporary whose
            # address the allocator will reuse for other contents, so such a call must not hit (re-arming drops the memo)
            src = [t.reshape(-1, t.shape[-1]) for t in text] + ([] if image is None else [image.reshape(-1, image.shape[-1])])
            ptrs = list(text_p[:n_pass]) + ([] if image is None else [img_ptr])
            if any(t.data_ptr() != p_ for t, p_ in zip(src, ptrs)):
                self.context_cache(True)
        if out is None:
            out = torch.empty(n_pass, self._cfg["out_channels"], T, H, W, device=self.device, dtype=torch.bfloat16)
        L = _lib.lib()
        with torch.cuda.device(self.device):
            need = C.c_size_t()
            _lib.check(L.alg_wan_workspace_bytes(self._handle, n_pass, T, H, W, n_img, C.byref(need)))
            if self._workspace is None or self._workspace.numel() < need.value:
                self._workspace = None
                self._workspace = torch.empty(need.value, dtype=torch.uint8, device=self.device)
            _lib.check(L.alg_wan_forward(self._handle, lat_p, cond_p, text_p, img_ptr, n_img, n_pass, T, H, W,
                                         int(timestep), out.data_ptr(), self._workspace.data_ptr(),
                                         self._workspace.numel(), _lib.stream_ptr(self.device)))
        return out

    def __call__(self, hidden_states, timestep, encoder_hidden_states, encoder_hidden_states_image=None,
                 attention_kwargs=None, return_dict: bool = True):
        """diffusers-compatible call on the pre-batched [B, 36, T, H, W] input (wan:910-917)."""
        B = hidden_states.shape[0]
        if B > 3:
            raise NotImplementedError("the engine batches at most the 3 CFG passes of one sample")
        t0 = int(timestep.flatten()[0]) if torch.is_tensor(timestep) else int(timestep)
        if torch.is_tensor(timestep) and timestep.numel() > 1 and not bool((timestep == timestep.flatten()[0]).all()):
            raise NotImplementedError("all passes of a step share one timestep (wan:903)")
        oc = self._cfg["out_channels"]
        hs = hidden_states.float()
        lat = [hs[b, :oc] for b in range(B)]
        cond = [hs[b, oc:] for b in range(B)]
        text = [encoder_hidden_states[b] for b in range(B)]
        img = None
        if encoder_hidden_states_image is not None:
            img = encoder_hidden_states_image[0]
        out = self.forward_passes(lat, cond, text, img, t0)
        if not return_dict:
            return (out,)
        return SimpleNamespace(sample=out)
