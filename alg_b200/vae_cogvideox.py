"""``AutoencoderKLCogVideoX`` on the sm_100a kernels of ``libalg_b200.so``: ``encode`` of single-frame inputs (the per-step call of
pixel-space ALG) and the full video ``decode`` (cog:428-433, once per video: 13 latent frames -> 49 frames).

Why it is here: with ``lp_filter_in_latent=False`` (BASELINE configs[2], ``configs/cogvideox_alg.yaml``) the reference
low-pass-filters the IMAGE and runs ``self.vae.encode(image_lp.unsqueeze(2)).latent_dist.sample(generator)`` on EVERY
denoise step (cog:645 / cog:257 here); the conditioning image goes through the same call once (cog:166).  Both inputs
are one frame ``[1, 3, 1, H, W]``, so this module builds exactly that: the CogVideoX encoder on T = 1.

Mirrors the interface the pipeline uses on ``self.vae``: ``.config`` (``block_out_channels``,
``temporal_compression_ratio``, ``scaling_factor``, ``invert_scale_latents``), ``.dtype``, ``.encode(x)`` returning an
object with ``.latent_dist.sample(generator)`` / ``.mode()``, and ``.decode(z).sample`` -- native when the decoder weights are
loaded (``CogVideoXDecoder3D``: latent frames in batches of two with every causal convolution's last two input frames carried
to the next batch, ``CogVideoXSpatialNorm3D`` conditioning on the latent, nearest-neighbour ``CogVideoXUpsample3D``), otherwise
delegated to ``decoder=``.

This module only SEQUENCES C-ABI calls (``alg_b200.ops``); activations are channels-last ``[H*W, C]`` bf16:

    CogVideoXCausalConv3d (3x3x3)        alg_im2col_bf16 + ONE alg_gemm_bf16 with K = 27*C, bias in the epilogue.  Frames
                                         before t=0 replicate frame 0, so on one frame the three temporal taps read the
                                         same [3, 3, C] patch: it is gathered once and the GEMM's A operand repeats along
                                         K (``a_k_period``) against the three temporal weight slices
    ResnetBlock3D ``hidden + inputs``    the residual epilogue of conv2's GEMM
    1x1x1 conv_shortcut                  alg_gemm_bf16
    CogVideoXDownsample3D                compress_time is the identity on one frame; pad (0,1,0,1) + Conv2d stride 2 =
                                         alg_im2col_bf16 (stride 2, zero bottom/right) + alg_gemm_bf16
    GroupNorm(32) + SiLU                 alg_group_norm_bf16

Restated from diffusers@be2fb77 ``autoencoder_kl_cogvideox.py`` / ``downsampling.py`` (not available offline: parity
unpinned); checked against ``oracle/vae_oracle.py`` (torch ``F.conv3d`` / ``F.group_norm``).
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import Dict, Optional

import torch

from . import _lib, ops

COGVIDEOX_5B_VAE = dict(in_channels=3, out_channels=3, latent_channels=16, block_out_channels=(128, 256, 256, 512),
                        layers_per_block=3, act_fn="silu", norm_eps=1e-6, norm_num_groups=32,
                        temporal_compression_ratio=4, sample_height=480, sample_width=720, scaling_factor=0.7,
                        shift_factor=None, latents_mean=None, latents_std=None, force_upcast=True, use_quant_conv=False,
                        use_post_quant_conv=False, invert_scale_latents=False)


def encoder_parameter_shapes(cfg: dict) -> Dict[str, tuple]:
    """name -> shape of every ENCODER parameter (diffusers naming)."""
    s: Dict[str, tuple] = {}
    boc = list(cfg["block_out_channels"])

    def conv3(name, o, i, k=3):
        s[name + ".conv.weight"] = (o, i, k, k, k)
        s[name + ".conv.bias"] = (o,)

    def resnet(name, i, o):
        s[name + ".norm1.weight"] = s[name + ".norm1.bias"] = (i,)
        conv3(name + ".conv1", o, i)
        s[name + ".norm2.weight"] = s[name + ".norm2.bias"] = (o,)
        conv3(name + ".conv2", o, o)
        if i != o:
            s[name + ".conv_shortcut.weight"] = (o, i, 1, 1, 1)
            s[name + ".conv_shortcut.bias"] = (o,)

    conv3("encoder.conv_in", boc[0], cfg["in_channels"])
    ch = boc[0]
    for b, out_ch in enumerate(boc):
        for j in range(cfg["layers_per_block"]):
            resnet(f"encoder.down_blocks.{b}.resnets.{j}", ch if j == 0 else out_ch, out_ch)
        if b != len(boc) - 1:
            s[f"encoder.down_blocks.{b}.downsamplers.0.conv.weight"] = (out_ch, out_ch, 3, 3)
            s[f"encoder.down_blocks.{b}.downsamplers.0.conv.bias"] = (out_ch,)
        ch = out_ch
    for j in range(2):
        resnet(f"encoder.mid_block.resnets.{j}", ch, ch)
    s["encoder.norm_out.weight"] = s["encoder.norm_out.bias"] = (ch,)
    conv3("encoder.conv_out", 2 * cfg["latent_channels"], ch)
    return s


def decoder_parameter_shapes(cfg: dict) -> Dict[str, tuple]:
    """name -> shape of every DECODER parameter (diffusers naming: CogVideoXDecoder3D)."""
    s: Dict[str, tuple] = {}
    boc = list(reversed(cfg["block_out_channels"]))
    z = cfg["latent_channels"]

    def conv3(name, o, i, k=3):
        s[name + ".conv.weight"] = (o, i, k, k, k)
        s[name + ".conv.bias"] = (o,)

    def snorm(name, f):
        s[name + ".norm_layer.weight"] = s[name + ".norm_layer.bias"] = (f,)
        conv3(name + ".conv_y", f, z, 1)
        conv3(name + ".conv_b", f, z, 1)

    def resnet(name, i, o):
        snorm(name + ".norm1", i)
        conv3(name + ".conv1", o, i)
        snorm(name + ".norm2", o)
        conv3(name + ".conv2", o, o)
        if i != o:
            s[name + ".conv_shortcut.weight"] = (o, i, 1, 1, 1)
            s[name + ".conv_shortcut.bias"] = (o,)

    conv3("decoder.conv_in", boc[0], z)
    for j in range(2):
        resnet(f"decoder.mid_block.resnets.{j}", boc[0], boc[0])
    ch = boc[0]
    for b, out_ch in enumerate(boc):
        for j in range(cfg["layers_per_block"] + 1):
            resnet(f"decoder.up_blocks.{b}.resnets.{j}", ch if j == 0 else out_ch, out_ch)
        if b != len(boc) - 1:
            s[f"decoder.up_blocks.{b}.upsamplers.0.conv.weight"] = (out_ch, out_ch, 3, 3)
            s[f"decoder.up_blocks.{b}.upsamplers.0.conv.bias"] = (out_ch,)
        ch = out_ch
    snorm("decoder.norm_out", ch)
    conv3("decoder.conv_out", cfg["out_channels"], ch)
    return s


def synthetic_decoder_state_dict(cfg: dict, seed: int = 0, device="cuda") -> Dict[str, torch.Tensor]:
    """Seeded decoder weights at the true shapes (fan-in scaled; the multiplicative conv_y branch is centred on 1)."""
    sd = {}
    for idx, (name, shape) in enumerate(decoder_parameter_shapes(cfg).items()):
        g = torch.Generator(device=device).manual_seed(seed * 1_000_003 + 9_001 + idx)
        if "norm_layer" in name and name.endswith(".weight"):
            w = 1 + 0.1 * torch.randn(shape, generator=g, device=device)
        elif name.endswith(".conv_y.conv.bias"):
            w = 1 + 0.05 * torch.randn(shape, generator=g, device=device)
        elif name.endswith(".bias"):
            w = 0.05 * torch.randn(shape, generator=g, device=device)
        elif ".conv_y." in name:
            w = 0.1 * torch.randn(shape, generator=g, device=device)
        else:
            fan_in = 1
            for d in shape[1:]:
                fan_in *= d
            w = torch.randn(shape, generator=g, device=device) * (1.0 / fan_in) ** 0.5
        sd[name] = w.to(torch.bfloat16)
    return sd


def synthetic_state_dict(cfg: dict, seed: int = 0, device="cuda") -> Dict[str, torch.Tensor]:
    """Seeded random-init encoder weights at the true shapes (fan-in scaled so activations stay O(1) through 28 convs)."""
    sd = {}
    for idx, (name, shape) in enumerate(encoder_parameter_shapes(cfg).items()):
        g = torch.Generator(device=device).manual_seed(seed * 1_000_003 + 7_001 + idx)
        if ".norm" in name and name.endswith(".weight"):
            w = 1 + 0.1 * torch.randn(shape, generator=g, device=device)
        elif name.endswith(".bias"):
            w = 0.05 * torch.randn(shape, generator=g, device=device)
        else:
            fan_in = 1
            for d in shape[1:]:
                fan_in *= d
            w = torch.randn(shape, generator=g, device=device) * (1.0 / fan_in) ** 0.5
        sd[name] = w.to(torch.bfloat16)
    return sd


class DiagonalGaussianDistribution:
    """diffusers ``DiagonalGaussianDistribution`` on the moments [B, 2z, T, H, W] (mean | logvar)."""

    def __init__(self, parameters: torch.Tensor):
        self.parameters = parameters
        self.mean, self.logvar = torch.chunk(parameters, 2, dim=1)
        self.logvar = torch.clamp(self.logvar, -30.0, 20.0)
        self.std = torch.exp(0.5 * self.logvar)
        self.var = torch.exp(self.logvar)

    def sample(self, generator: Optional[torch.Generator] = None) -> torch.Tensor:
        from .pipeline_utils import randn_tensor
        noise = randn_tensor(self.mean.shape, generator=generator, device=self.parameters.device, dtype=self.parameters.dtype)
        return self.mean + self.std * noise

    def mode(self) -> torch.Tensor:
        return self.mean


class AutoencoderKLCogVideoX:
    """Native-kernel ``AutoencoderKLCogVideoX`` encoder (one frame per call); ``decode`` is delegated."""

    def __init__(self, decoder=None, **config):
        cfg = dict(COGVIDEOX_5B_VAE)
        cfg.update(config)
        cfg["block_out_channels"] = tuple(cfg["block_out_channels"])
        if cfg["use_quant_conv"]:
            raise NotImplementedError("use_quant_conv=True is not built (CogVideoX-5b ships without quant_conv)")
        if any(c % cfg["norm_num_groups"] for c in cfg["block_out_channels"]):
            raise ValueError("block_out_channels must be divisible by norm_num_groups")
        self.config = SimpleNamespace(**cfg)
        self._cfg = cfg
        self._w: Dict[str, torch.Tensor] = {}
        self._cols: Optional[torch.Tensor] = None
        self._cols_budget = 3 << 30  # bytes of patch matrix per im2col + GEMM call (whole frames)
        self.num_latent_frames_batch_size = 2
        import os
        self.implicit = os.environ.get("ALG_VAE_IMPLICIT", "1") == "1"  # 3x3x3 decoder convolutions without a patch matrix
        self._padded: Dict[tuple, torch.Tensor] = {}
        self.decoder = decoder
        self.has_decoder = False
        self.dtype = torch.bfloat16
        self.device = torch.device("cpu")

    # ---- construction -------------------------------------------------------------------------------
    @classmethod
    def from_synthetic(cls, seed: int = 0, device="cuda", decoder=None, with_decoder: bool = False, **config):
        m = cls(decoder=decoder, **config)
        sd = synthetic_state_dict(m._cfg, seed=seed, device=device)
        if with_decoder:
            sd.update(synthetic_decoder_state_dict(m._cfg, seed=seed, device=device))
        return m.load_state_dict(sd)

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path, subfolder: str = "vae", torch_dtype=None, cache_dir=None,
                        device="cuda", decoder=None):
        """Weights from a LOCAL diffusers snapshot (``vae/config.json`` + safetensors): encoder and decoder; with ``decoder=``
        given only the encoder keys are read and ``decode`` is delegated."""
        from . import checkpoint
        root = checkpoint.resolve_snapshot(str(pretrained_model_name_or_path), cache_dir, require=subfolder or "transformer")
        cfg, sd = checkpoint.load_component(root, subfolder, device=device, key_prefix=None if decoder is None else "encoder.")
        known = {k: v for k, v in cfg.items() if k in COGVIDEOX_5B_VAE}
        return cls(decoder=decoder, **known).load_state_dict(sd)

    def load_state_dict(self, sd: Dict[str, torch.Tensor], strict: bool = True):
        """diffusers names.  Conv weights are re-laid once as GEMM B operands [Co, (kt, kh, kw, Ci)] (Ci padded to 8)."""
        shapes = encoder_parameter_shapes(self._cfg)
        missing = [k for k in shapes if k not in sd]
        if strict and missing:
            raise KeyError(f"missing encoder weights: {missing[:4]}{'...' if len(missing) > 4 else ''}")
        dshapes = decoder_parameter_shapes(self._cfg)
        self.has_decoder = any(k.startswith("decoder.") for k in sd)
        if self.has_decoder:
            dmissing = [k for k in dshapes if k not in sd]
            if strict and dmissing:
                raise KeyError(f"missing decoder weights: {dmissing[:4]}{'...' if len(dmissing) > 4 else ''}")
            shapes = dict(shapes, **dshapes)
        w: Dict[str, torch.Tensor] = {}
        for name, shape in shapes.items():
            t = sd[name]
            if tuple(t.shape) != tuple(shape):
                raise ValueError(f"{name}: expected {tuple(shape)}, got {tuple(t.shape)}")
            t = t.to(torch.bfloat16)
            if t.dim() >= 4:  # conv weight [Co, Ci, (kt,) kh, kw] -> [Co, (kt,) kh, kw, Ci8] -> [Co, K]
                co, ci = t.shape[:2]
                t = t.movedim(1, -1)
                if ci % 8:
                    t = torch.nn.functional.pad(t, (0, 8 - ci % 8))
                t = t.reshape(co, -1)
            w[name] = t.contiguous()
            self.device = t.device
        if self.has_decoder:  # conv_y | conv_b of every spatial norm as ONE [2 f, z] projection of the latent
            for name in [k[:-len(".conv_y.conv.weight")] for k in dshapes if k.endswith(".conv_y.conv.weight")]:
                w[name + ".yb.weight"] = torch.cat([w.pop(name + ".conv_y.conv.weight"), w.pop(name + ".conv_b.conv.weight")]).contiguous()
                w[name + ".yb.bias"] = torch.cat([w.pop(name + ".conv_y.conv.bias"), w.pop(name + ".conv_b.conv.bias")]).contiguous()
        self._w = w
        return self

    def to(self, device=None, dtype=None):
        if device is not None and self._w:
            self._w = {k: v.to(device) for k, v in self._w.items()}
            self.device = torch.device(device)
            self._cols = None
        if self.decoder is not None and hasattr(self.decoder, "to"):
            self.decoder = self.decoder.to(device) if device is not None else self.decoder
        return self

    def eval(self):
        return self

    # ---- building blocks ----------------------------------------------------------------------------
    def _workspace(self, numel: int, device) -> torch.Tensor:
        if self._cols is None or self._cols.numel() < numel or self._cols.device != device:
            self._cols = torch.empty(numel, device=device, dtype=torch.bfloat16)
        return self._cols

    def _conv3(self, x, H, W, name, residual=None):
        """CogVideoXCausalConv3d(k=3) on one frame: [H*W, Ci] -> [H*W, Co]."""
        wt, b = self._w[name + ".conv.weight"], self._w[name + ".conv.bias"]
        Cc = x.shape[1]
        if self.implicit and Cc % 64 == 0 and wt.shape[0] % 8 == 0:
            # no patch matrix: the frame is laid out once with a zero border; the GEMM's tap mode reads it at the 9 spatial
            # offsets, three times over (on one frame the three temporal taps of the replicate-padded clip are the same frame)
            dev, plane, row = x.device, (H + 2) * (W + 2), W + 2
            xp = self._padded.get((1, H, W, Cc))
            if xp is None or xp.device != dev:
                xp = self._padded[(1, H, W, Cc)] = torch.zeros(plane, Cc, device=dev, dtype=torch.bfloat16)
            ops.pad_frames(x, xp, 1, H, W, to_padded=True)
            offs = [(ih - 1) * row + (iw - 1) for _ in range(3) for ih in range(3) for iw in range(3)]
            d = ops.gemm(xp, wt, b, a_tap_kblocks=Cc // 64, a_tap_offsets=offs)
            out = torch.empty(H * W, wt.shape[0], device=dev, dtype=torch.bfloat16)
            return ops.pad_frames(d, out, 1, H, W, to_padded=False, residual=residual)
        if (9 * Cc) % 64 == 0:
            # one frame: the three temporal taps read the same [3, 3, C] patch, so it is gathered once and the GEMM walks
            # it three times along K (a_k_period) against the three temporal weight slices
            cols = ops.im2col(x, 1, H, W, kernel=(1, 3, 3), pad_top=1, pad_left=1, out=self._workspace(H * W * 9 * Cc, x.device))
            kw = dict(a_k_period=9 * Cc)
        else:  # conv_in (3 -> 8 padded channels): K = 216, materialised in full
            cols = ops.im2col(x, 1, H, W, kernel=(3, 3, 3), pad_t=2, pad_top=1, pad_left=1,
                              out=self._workspace(H * W * wt.shape[1], x.device))
            kw = {}
        if residual is not None:
            return ops.gemm(cols, wt, b, epilogue=_lib.EPI_RESIDUAL, residual=residual, **kw)
        return ops.gemm(cols, wt, b, **kw)

    def _resnet(self, x, H, W, name):
        g, eps = self._cfg["norm_num_groups"], self._cfg["norm_eps"]
        h = ops.group_norm(x, g, self._w[name + ".norm1.weight"], self._w[name + ".norm1.bias"], eps=eps, silu=True)
        h = self._conv3(h, H, W, name + ".conv1")
        h = ops.group_norm(h, g, self._w[name + ".norm2.weight"], self._w[name + ".norm2.bias"], eps=eps, silu=True, out=h)
        if name + ".conv_shortcut.weight" in self._w:
            x = ops.gemm(x, self._w[name + ".conv_shortcut.weight"], self._w[name + ".conv_shortcut.bias"])
        return self._conv3(h, H, W, name + ".conv2", residual=x)

    def _downsample(self, x, H, W, name):
        """CogVideoXDownsample3D on one frame: F.pad (0,1,0,1) + Conv2d(3, stride 2, padding 0)."""
        wt, b = self._w[name + ".conv.weight"], self._w[name + ".conv.bias"]
        Ho, Wo = (H + 1 - 3) // 2 + 1, (W + 1 - 3) // 2 + 1
        cols = ops.im2col(x, 1, H, W, kernel=(1, 3, 3), stride=(1, 2, 2), out_hw=(Ho, Wo),
                          out=self._workspace(Ho * Wo * wt.shape[1], x.device))
        return ops.gemm(cols, wt, b), Ho, Wo

    def _encode_frame(self, img: torch.Tensor) -> torch.Tensor:
        """img [3, H, W] -> moments [2z, H/8, W/8] bf16."""
        cfg = self._cfg
        Cin, H, W = img.shape
        x = torch.zeros(H * W, (Cin + 7) // 8 * 8, device=img.device, dtype=torch.bfloat16)
        x[:, :Cin] = img.to(torch.bfloat16).permute(1, 2, 0).reshape(H * W, Cin)  # channels-last, zero-padded to 8
        x = self._conv3(x, H, W, "encoder.conv_in")
        n_blocks = len(cfg["block_out_channels"])
        for bi in range(n_blocks):
            for j in range(cfg["layers_per_block"]):
                x = self._resnet(x, H, W, f"encoder.down_blocks.{bi}.resnets.{j}")
            if bi != n_blocks - 1:
                x, H, W = self._downsample(x, H, W, f"encoder.down_blocks.{bi}.downsamplers.0")
        for j in range(2):
            x = self._resnet(x, H, W, f"encoder.mid_block.resnets.{j}")
        x = ops.group_norm(x, cfg["norm_num_groups"], self._w["encoder.norm_out.weight"], self._w["encoder.norm_out.bias"],
                           eps=cfg["norm_eps"], silu=True, out=x)
        x = self._conv3(x, H, W, "encoder.conv_out")
        return x.view(H, W, -1).permute(2, 0, 1)

    # ---- diffusers surface --------------------------------------------------------------------------
    def encode(self, x: torch.Tensor, return_dict: bool = True):
        """x [B, 3, 1, H, W] -> ``.latent_dist`` over [B, 2z -> z, 1, H/8, W/8] (cog:166, cog:257)."""
        if not self._w:
            raise RuntimeError("AutoencoderKLCogVideoX: no weights loaded")
        if x.dim() != 5 or x.shape[1] != self._cfg["in_channels"]:
            raise ValueError(f"encode expects [B, {self._cfg['in_channels']}, T, H, W], got {tuple(x.shape)}")
        if x.shape[2] != 1:
            raise NotImplementedError("the native CogVideoX encoder handles one frame per call (the per-step path); "
                                      "multi-frame video encoding runs once per video and is out of scope")
        n_down = len(self._cfg["block_out_channels"]) - 1
        if x.shape[3] % (1 << n_down) or x.shape[4] % (1 << n_down):
            raise ValueError("height and width must be multiples of the spatial compression ratio")
        moments = torch.stack([self._encode_frame(x[b, :, 0]) for b in range(x.shape[0])]).unsqueeze(2).to(x.dtype)
        posterior = DiagonalGaussianDistribution(moments)
        if not return_dict:
            return (posterior,)
        return SimpleNamespace(latent_dist=posterior)

    # ---- decoder -------------------------------------------------------------------------------------
    def _conv3_cached(self, x, T, H, W, name, cache, residual=None):
        """CogVideoXCausalConv3d(k=3) on a chunk of T frames [T*H*W, Ci] -> [T*H*W, Co].  The two frames in front come from
        the previous chunk (``conv_cache``) or replicate frame 0; the last two input frames are kept for the next chunk."""
        wt, b = self._w[name + ".conv.weight"], self._w[name + ".conv.bias"]
        Ci, HW = x.shape[1], H * W
        if self.implicit and Ci % 64 == 0:
            return self._conv3_implicit(x, T, H, W, name, cache, residual)
        xin = torch.empty((T + 2) * HW, Ci, device=x.device, dtype=torch.bfloat16)
        prev = cache.get(name)
        if prev is None:
            ops.copy_rows(x[:HW], xin[:HW])
            ops.copy_rows(x[:HW], xin[HW:2 * HW])
        else:
            ops.copy_rows(prev, xin[:2 * HW])
        ops.copy_rows(x, xin[2 * HW:])
        cache[name] = xin[T * HW:]  # the last two frames of the padded input (a view: xin stays alive through it)
        K = wt.shape[1]
        out = torch.empty(T * HW, wt.shape[0], device=x.device, dtype=torch.bfloat16) if wt.shape[0] % 8 == 0 else None
        per_call = max(1, min(T, self._cols_budget // max(1, HW * K * 2)))
        outs = []
        for t0 in range(0, T, per_call):
            n = min(per_call, T - t0)
            cols = ops.im2col(xin[t0 * HW:(t0 + n + 2) * HW], n + 2, H, W, kernel=(3, 3, 3), pad_t=0, pad_top=1, pad_left=1,
                              out=self._workspace(n * HW * K, x.device))
            res = None if residual is None else residual[t0 * HW:(t0 + n) * HW]
            kw = dict(epilogue=_lib.EPI_RESIDUAL, residual=res) if res is not None else {}
            if out is not None:
                ops.gemm(cols, wt, b, out=out[t0 * HW:(t0 + n) * HW], **kw)
            else:
                outs.append(ops.gemm(cols, wt, b, **kw))
        return out if out is not None else torch.cat(outs)

    def _conv3_implicit(self, x, T, H, W, name, cache, residual=None):
        """The same convolution without a patch matrix: the chunk (with its two cached / replicated frames in front) is laid out once
        as spatially zero-padded frames [(T + 2), H + 2, W + 2, Ci]; the GEMM's tap mode reads it 27 times at row offsets
        (it * plane + (ih - 1) * row + (iw - 1)); the bf16 result comes back from the padded raster with the residual added in the
        same pass (the rounding chain of the GEMM's residual epilogue: bf16(R + bf16(acc + bias)))."""
        dev = x.device
        wt, b = self._w[name + ".conv.weight"], self._w[name + ".conv.bias"]
        Ci, HW, plane, row = x.shape[1], H * W, (H + 2) * (W + 2), W + 2
        key = (T, H, W, Ci)
        xp = self._padded.get(key)
        if xp is None or xp.device != dev:  # borders are zeroed once; every use rewrites all interiors
            xp = self._padded[key] = torch.zeros((T + 2) * plane, Ci, device=dev, dtype=torch.bfloat16)
        prev = cache.get(name)

        def pad_in(src, frame0, n):
            ops.pad_frames(src, xp[frame0 * plane:(frame0 + n) * plane], n, H, W, to_padded=True)
        if prev is None:
            pad_in(x[:HW], 0, 1)
            pad_in(x[:HW], 1, 1)
        else:
            pad_in(prev, 0, 2)
        pad_in(x, 2, T)
        keep = torch.empty(2 * HW, Ci, device=dev, dtype=torch.bfloat16)  # the last two frames of (front frames + chunk)
        if T >= 2:
            ops.copy_rows(x[(T - 2) * HW:], keep)
        else:
            ops.copy_rows(prev[HW:] if prev is not None else x[:HW], keep[:HW])
            ops.copy_rows(x, keep[HW:])
        cache[name] = keep
        offs = [it * plane + (ih - 1) * row + (iw - 1) for it in range(3) for ih in range(3) for iw in range(3)]
        d = ops.gemm(xp, wt, b, a_tap_kblocks=Ci // 64, a_tap_offsets=offs, m_rows=T * plane)  # padded raster [T * plane, Co]
        Co = wt.shape[0]
        out = torch.empty(T * HW, Co, device=dev, dtype=torch.bfloat16)
        return ops.pad_frames(d, out, T, H, W, to_padded=False, residual=residual)

    def _spatial_norm(self, f, T, H, W, z, zt, zh, zw, name, silu=True):
        """CogVideoXSpatialNorm3D: GroupNorm(f) * conv_y(zq') + conv_b(zq') (+ SiLU); f [T*H*W, C], z [zt*zh*zw, zc]."""
        lib = _lib.lib()
        fn = ops.group_norm(f, self._cfg["norm_num_groups"], self._w[name + ".norm_layer.weight"], self._w[name + ".norm_layer.bias"],
                            eps=1e-6, silu=False)
        yb = ops.gemm(z, self._w[name + ".yb.weight"], self._w[name + ".yb.bias"])  # [zrows, 2C] at latent resolution
        out = torch.empty_like(f)
        Cc, HW, zhw = f.shape[1], H * W, zh * zw

        def apply(t0, t1, z0, z1):
            with torch.cuda.device(f.device):
                _lib.check(lib.alg_spatial_norm_apply_bf16(fn[t0 * HW:].data_ptr(), yb[z0 * zhw:].data_ptr(), out[t0 * HW:].data_ptr(), Cc,
                                                           t1 - t0, H, W, z1 - z0, zh, zw, int(silu), _lib.stream_ptr(f.device)))
        if T > 1 and T % 2 == 1:  # first frame <- first latent frame, the rest resized together
            apply(0, 1, 0, 1)
            apply(1, T, 1, zt)
        else:
            apply(0, T, 0, zt)
        return out

    def _dec_resnet(self, x, T, H, W, z, zdims, name, cache):
        h = self._spatial_norm(x, T, H, W, z, *zdims, name + ".norm1")
        h = self._conv3_cached(h, T, H, W, name + ".conv1", cache)
        h = self._spatial_norm(h, T, H, W, z, *zdims, name + ".norm2")
        if name + ".conv_shortcut.weight" in self._w:
            x = ops.gemm(x, self._w[name + ".conv_shortcut.weight"], self._w[name + ".conv_shortcut.bias"])
        return self._conv3_cached(h, T, H, W, name + ".conv2", cache, residual=x)

    def _upsample(self, x, T, H, W, name, compress_time):
        """CogVideoXUpsample3D: nearest x2 (space; time too when ``compress_time``, the first frame of an odd chunk kept single),
        then a per-frame Conv2d(3, padding 1)."""
        lib = _lib.lib()
        Cc, HW = x.shape[1], H * W
        if compress_time and T > 1 and T % 2 == 1:
            To, parts = 1 + 2 * (T - 1), [(0, 1, 0, 1), (1, T, 1, 1 + 2 * (T - 1))]
        elif compress_time and T > 1:
            To, parts = 2 * T, [(0, T, 0, 2 * T)]
        else:
            To, parts = T, [(0, T, 0, T)]
        up = torch.empty(To * 4 * HW, Cc, device=x.device, dtype=torch.bfloat16)
        for (i0, i1, o0, o1) in parts:
            with torch.cuda.device(x.device):
                _lib.check(lib.alg_upsample_nearest_bf16(x[i0 * HW:].data_ptr(), up[o0 * 4 * HW:].data_ptr(), Cc, i1 - i0, H, W, o1 - o0,
                                                         2 * H, 2 * W, _lib.stream_ptr(x.device)))
        H, W, HW = 2 * H, 2 * W, 4 * HW
        wt, b = self._w[name + ".conv.weight"], self._w[name + ".conv.bias"]
        out = torch.empty(To * HW, wt.shape[0], device=x.device, dtype=torch.bfloat16)
        per_call = max(1, min(To, self._cols_budget // max(1, HW * wt.shape[1] * 2)))
        for t0 in range(0, To, per_call):
            n = min(per_call, To - t0)
            cols = ops.im2col(up[t0 * HW:(t0 + n) * HW], n, H, W, kernel=(1, 3, 3), pad_top=1, pad_left=1,
                              out=self._workspace(n * HW * wt.shape[1], x.device))
            ops.gemm(cols, wt, b, out=out[t0 * HW:(t0 + n) * HW])
        return out, To, H, W

    def _decode_chunk(self, zc: torch.Tensor, cache: dict) -> torch.Tensor:
        """zc [zc, t, h, w] (one latent frame batch) -> [3, T, 8h, 8w] bf16."""
        cfg = self._cfg
        Cz, T, H, W = zc.shape
        z = zc.to(torch.bfloat16).permute(1, 2, 3, 0).reshape(T * H * W, Cz).contiguous()  # channels-last
        zdims = (T, H, W)
        boc = list(reversed(cfg["block_out_channels"]))
        n_compress = {1: 0, 2: 1, 4: 2, 8: 3}[cfg["temporal_compression_ratio"]]
        x = self._conv3_cached(z, T, H, W, "decoder.conv_in", cache)
        for j in range(2):
            x = self._dec_resnet(x, T, H, W, z, zdims, f"decoder.mid_block.resnets.{j}", cache)
        for bi in range(len(boc)):
            for j in range(cfg["layers_per_block"] + 1):
                x = self._dec_resnet(x, T, H, W, z, zdims, f"decoder.up_blocks.{bi}.resnets.{j}", cache)
            if bi != len(boc) - 1:
                x, T, H, W = self._upsample(x, T, H, W, f"decoder.up_blocks.{bi}.upsamplers.0", bi < n_compress)
        x = self._spatial_norm(x, T, H, W, z, *zdims, "decoder.norm_out", silu=True)
        x = self._conv3_cached(x, T, H, W, "decoder.conv_out", cache)
        return x.reshape(T, H, W, -1).permute(3, 0, 1, 2)

    def decode(self, z: torch.Tensor, return_dict: bool = True):
        """z [B, zc, T, h, w] -> ``.sample`` [B, 3, 1 + 4 (T - 1), 8h, 8w] (cog:428-433; ``_decode``: latent frame batches of
        ``num_latent_frames_batch_size`` = 2, the first batch also takes the remainder)."""
        if not self.has_decoder:
            if self.decoder is None:
                raise NotImplementedError("AutoencoderKLCogVideoX.decode: no decoder weights loaded and no decoder= object given")
            return self.decoder.decode(z, return_dict=return_dict)
        if z.dim() != 5 or z.shape[1] != self._cfg["latent_channels"]:
            raise ValueError(f"decode expects [B, {self._cfg['latent_channels']}, T, h, w], got {tuple(z.shape)}")
        _lib.require_cuda(z)
        fb, T = self.num_latent_frames_batch_size, z.shape[2]
        n_batches, rem = max(T // fb, 1), T % fb
        videos = []
        for b in range(z.shape[0]):
            cache: dict = {}
            parts = []
            for i in range(n_batches):
                start = fb * i + (0 if i == 0 else rem)
                parts.append(self._decode_chunk(z[b, :, start:fb * (i + 1) + rem], cache))
            videos.append(torch.cat(parts, dim=1))
        video = torch.stack(videos).to(z.dtype)
        self._cols = None  # the patch-matrix workspace is large at 480 x 720: give it back once the video is decoded
        self._padded.clear()
        if not return_dict:
            return (video,)
        return SimpleNamespace(sample=video)
