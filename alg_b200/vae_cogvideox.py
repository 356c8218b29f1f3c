"""``AutoencoderKLCogVideoX.encode`` on the sm_100a kernels of ``libalg_b200.so`` (single-frame inputs).

Why it is here: with ``lp_filter_in_latent=False`` (BASELINE configs[2], ``configs/cogvideox_alg.yaml``) the reference
low-pass-filters the IMAGE and runs ``self.vae.encode(image_lp.unsqueeze(2)).latent_dist.sample(generator)`` on EVERY
denoise step (cog:645 / cog:257 here); the conditioning image goes through the same call once (cog:166).  Both inputs
are one frame ``[1, 3, 1, H, W]``, so this module builds exactly that: the CogVideoX encoder on T = 1.

Mirrors the interface the pipeline uses on ``self.vae``: ``.config`` (``block_out_channels``,
``temporal_compression_ratio``, ``scaling_factor``, ``invert_scale_latents``), ``.dtype``, ``.encode(x)`` returning an
object with ``.latent_dist.sample(generator)`` / ``.mode()``, and ``.decode`` (delegated: decoding 13 latent frames to 49
video frames happens once per video and is out of scope, SURVEY 8(f).1 -- pass ``decoder=`` or the call raises).

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
        self.decoder = decoder
        self.dtype = torch.bfloat16
        self.device = torch.device("cpu")

    # ---- construction -------------------------------------------------------------------------------
    @classmethod
    def from_synthetic(cls, seed: int = 0, device="cuda", decoder=None, **config):
        m = cls(decoder=decoder, **config)
        return m.load_state_dict(synthetic_state_dict(m._cfg, seed=seed, device=device))

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path, subfolder: str = "vae", torch_dtype=None, cache_dir=None,
                        device="cuda", decoder=None):
        """Encoder weights from a LOCAL diffusers snapshot (``vae/config.json`` + safetensors); decoder keys are ignored."""
        from . import checkpoint
        root = checkpoint.resolve_snapshot(pretrained_model_name_or_path, cache_dir)
        cfg, sd = checkpoint.load_component(root, subfolder, device=device, key_prefix="encoder.")
        known = {k: v for k, v in cfg.items() if k in COGVIDEOX_5B_VAE}
        return cls(decoder=decoder, **known).load_state_dict(sd)

    def load_state_dict(self, sd: Dict[str, torch.Tensor], strict: bool = True):
        """diffusers names.  Conv weights are re-laid once as GEMM B operands [Co, (kt, kh, kw, Ci)] (Ci padded to 8)."""
        shapes = encoder_parameter_shapes(self._cfg)
        missing = [k for k in shapes if k not in sd]
        if strict and missing:
            raise KeyError(f"missing encoder weights: {missing[:4]}{'...' if len(missing) > 4 else ''}")
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

    def decode(self, z: torch.Tensor, return_dict: bool = True):
        if self.decoder is None:
            raise NotImplementedError("AutoencoderKLCogVideoX.decode is not built (once per video, SURVEY 8(f).1): "
                                      "pass decoder=<object with .decode(z)>")
        return self.decoder.decode(z, return_dict=return_dict)
