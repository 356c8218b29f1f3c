"""CPU/GPU restatement of ``AutoencoderKLCogVideoX.encode`` -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows diffusers@be2fb77 (``requirements.txt:13`` of the reference pins it; the package is NOT available offline, so
this restatement is written from its published source and is **parity unpinned**):

  * ``models/autoencoders/autoencoder_kl_cogvideox.py``: ``CogVideoXCausalConv3d`` (constant-pad mode: the first frame is
    repeated ``kt - 1`` times in front when there is no conv cache, spatial zero padding inside the conv),
    ``CogVideoXResnetBlock3D`` (GroupNorm -> SiLU -> conv1 -> GroupNorm -> SiLU -> dropout -> conv2, 1x1x1
    ``conv_shortcut`` when the channel count changes, ``hidden + inputs``), ``CogVideoXDownBlock3D``,
    ``CogVideoXMidBlock3D`` (two resnets), ``CogVideoXEncoder3D`` (conv_in, 4 down blocks, mid block, GroupNorm, SiLU,
    conv_out), ``AutoencoderKLCogVideoX._encode`` / ``encode`` (frame batches of 8: one batch for T <= 8);
  * ``models/downsampling.py``: ``CogVideoXDownsample3D`` (``compress_time`` average-pools frame pairs after the first
    frame; ``F.pad(x, (0, 1, 0, 1))``; per-frame ``Conv2d(3, stride 2, padding 0)``);
  * ``models/autoencoders/vae.py``: ``DiagonalGaussianDistribution``.

Call sites in the reference: cog:166 (conditioning image) and cog:645 (the low-pass-filtered image, EVERY step in pixel
mode), both with one frame.  Works on any T that fits one frame batch; ``dtype=torch.float32`` evaluates the same bf16
weights in fp32 (the ground truth of the parity protocol), ``torch.bfloat16`` is the eager chain the reference runs.
"""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn.functional as F


def _causal_conv3d(x, w, b):
    kt = w.shape[2]
    if kt > 1:
        x = torch.cat([x[:, :, :1]] * (kt - 1) + [x], dim=2)
    return F.conv3d(x, w, b, padding=(0, w.shape[3] // 2, w.shape[4] // 2))


def _resnet(x, sd: Dict[str, torch.Tensor], name: str, groups: int, eps: float, dt):
    g = lambda k: sd[name + k].to(dt)
    h = F.silu(F.group_norm(x, groups, g(".norm1.weight"), g(".norm1.bias"), eps))
    h = _causal_conv3d(h, g(".conv1.conv.weight"), g(".conv1.conv.bias"))
    h = F.silu(F.group_norm(h, groups, g(".norm2.weight"), g(".norm2.bias"), eps))
    h = _causal_conv3d(h, g(".conv2.conv.weight"), g(".conv2.conv.bias"))
    if name + ".conv_shortcut.weight" in sd:
        x = F.conv3d(x, g(".conv_shortcut.weight"), g(".conv_shortcut.bias"))
    return h + x


def _downsample(x, w, b, compress_time: bool):
    B, Cc, T, H, W = x.shape
    if compress_time:
        y = x.permute(0, 3, 4, 1, 2).reshape(B * H * W, Cc, T)
        if T % 2 == 1:
            first, rest = y[..., 0], y[..., 1:]
            if rest.shape[-1] > 0:
                rest = F.avg_pool1d(rest, kernel_size=2, stride=2)
            y = torch.cat([first[..., None], rest], dim=-1)
        else:
            y = F.avg_pool1d(y, kernel_size=2, stride=2)
        T = y.shape[-1]
        x = y.reshape(B, H, W, Cc, T).permute(0, 3, 4, 1, 2)
    x = F.pad(x, (0, 1, 0, 1), mode="constant", value=0)
    B, Cc, T, H, W = x.shape
    y = F.conv2d(x.permute(0, 2, 1, 3, 4).reshape(B * T, Cc, H, W), w, b, stride=2, padding=0)
    return y.reshape(B, T, y.shape[1], y.shape[2], y.shape[3]).permute(0, 2, 1, 3, 4)


def encode_moments(x: torch.Tensor, sd: Dict[str, torch.Tensor], cfg: dict, dtype=torch.float32) -> torch.Tensor:
    """x [B, 3, T, H, W] -> moments [B, 2z, T', H/8, W/8] in ``dtype`` (weights are taken from ``sd`` and cast)."""
    if x.shape[2] > 8:
        raise NotImplementedError("oracle covers one frame batch (T <= 8)")
    groups, eps = cfg["norm_num_groups"], cfg["norm_eps"]
    boc = list(cfg["block_out_channels"])
    n_compress = {1: 0, 2: 1, 4: 2, 8: 3}[cfg["temporal_compression_ratio"]]
    h = _causal_conv3d(x.to(dtype), sd["encoder.conv_in.conv.weight"].to(dtype), sd["encoder.conv_in.conv.bias"].to(dtype))
    for bi in range(len(boc)):
        for j in range(cfg["layers_per_block"]):
            h = _resnet(h, sd, f"encoder.down_blocks.{bi}.resnets.{j}", groups, eps, dtype)
        if bi != len(boc) - 1:
            n = f"encoder.down_blocks.{bi}.downsamplers.0.conv."
            h = _downsample(h, sd[n + "weight"].to(dtype), sd[n + "bias"].to(dtype), compress_time=bi < n_compress)
    for j in range(2):
        h = _resnet(h, sd, f"encoder.mid_block.resnets.{j}", groups, eps, dtype)
    h = F.silu(F.group_norm(h, groups, sd["encoder.norm_out.weight"].to(dtype), sd["encoder.norm_out.bias"].to(dtype), eps))
    return _causal_conv3d(h, sd["encoder.conv_out.conv.weight"].to(dtype), sd["encoder.conv_out.conv.bias"].to(dtype))


def sample(moments: torch.Tensor, noise: torch.Tensor) -> torch.Tensor:
    """DiagonalGaussianDistribution.sample with the N(0,1) draw given."""
    mean, logvar = torch.chunk(moments, 2, dim=1)
    return mean + torch.exp(0.5 * torch.clamp(logvar, -30.0, 20.0)) * noise
