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


# ----------------------------------------------------------------------------------------------------------------
# decode (cog:428-433 ``decode_latents`` -> ``self.vae.decode(latents).sample``): CogVideoXDecoder3D, parity unpinned like
# the encoder.  Restated from the same diffusers file:
#   * ``CogVideoXSpatialNorm3D``: GroupNorm(f) * conv_y(zq') + conv_b(zq') with zq' = the latent chunk resized (nearest) to
#     f's (T, H, W) -- first frame and the rest separately when f has an odd number (> 1) of frames;
#   * ``CogVideoXResnetBlock3D`` with ``spatial_norm_dim``; ``CogVideoXMidBlock3D``; ``CogVideoXUpBlock3D`` (layers_per_block
#     + 1 resnets) + ``CogVideoXUpsample3D`` (nearest x2 in space, and in time for the first log2(temporal ratio) blocks,
#     first frame kept single; then per-frame Conv2d 3x3);
#   * ``AutoencoderKLCogVideoX._decode``: latent frames in batches of 2 (the first batch takes the remainder), every causal
#     convolution carrying its last kt - 1 input frames to the next batch (``conv_cache``).
# ----------------------------------------------------------------------------------------------------------------
def decoder_parameter_shapes(cfg: dict) -> Dict[str, tuple]:
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


def make_decoder_weights(cfg: dict, seed: int = 0, device="cpu", dtype=torch.bfloat16) -> Dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for name, shape in decoder_parameter_shapes(cfg).items():
        if "norm_layer" in name and name.endswith(".weight"):
            w = 1 + 0.1 * torch.randn(shape, generator=g)
        elif name.endswith(".bias"):
            w = 0.05 * torch.randn(shape, generator=g)
        elif ".conv_y." in name:  # multiplicative modulation around 1
            w = torch.randn(shape, generator=g) * 0.1
        else:
            fan_in = 1
            for d in shape[1:]:
                fan_in *= d
            w = torch.randn(shape, generator=g) * (1.0 / fan_in) ** 0.5
        if ".conv_y.conv.bias" in name:
            w = 1 + w
        sd[name] = w.to(device=device, dtype=dtype)
    return sd


def _cached_conv3d(x, w, b, cache: dict, key: str):
    kt = w.shape[2]
    if kt > 1:
        prev = cache.get(key)
        front = [x[:, :, :1]] * (kt - 1) if prev is None else [prev]
        xin = torch.cat(front + [x], dim=2)
        cache[key] = xin[:, :, -(kt - 1):].clone()
    else:
        xin = x
    return F.conv3d(xin, w, b, padding=(0, w.shape[3] // 2, w.shape[4] // 2))


def _spatial_norm(f, zq, sd, name, groups, dt, cache):
    g = lambda k: sd[name + k].to(dt)
    if f.shape[2] > 1 and f.shape[2] % 2 == 1:
        z = torch.cat([F.interpolate(zq[:, :, :1], size=(1,) + tuple(f.shape[-2:])),
                       F.interpolate(zq[:, :, 1:], size=(f.shape[2] - 1,) + tuple(f.shape[-2:]))], dim=2)
    else:
        z = F.interpolate(zq, size=tuple(f.shape[-3:]))
    y = _cached_conv3d(z, g(".conv_y.conv.weight"), g(".conv_y.conv.bias"), cache, name + ".conv_y")
    b = _cached_conv3d(z, g(".conv_b.conv.weight"), g(".conv_b.conv.bias"), cache, name + ".conv_b")
    return F.group_norm(f, groups, g(".norm_layer.weight"), g(".norm_layer.bias"), 1e-6) * y + b


def _dec_resnet(x, zq, sd, name, groups, dt, cache):
    g = lambda k: sd[name + k].to(dt)
    h = F.silu(_spatial_norm(x, zq, sd, name + ".norm1", groups, dt, cache))
    h = _cached_conv3d(h, g(".conv1.conv.weight"), g(".conv1.conv.bias"), cache, name + ".conv1")
    h = F.silu(_spatial_norm(h, zq, sd, name + ".norm2", groups, dt, cache))
    h = _cached_conv3d(h, g(".conv2.conv.weight"), g(".conv2.conv.bias"), cache, name + ".conv2")
    if name + ".conv_shortcut.weight" in sd:
        x = F.conv3d(x, g(".conv_shortcut.weight"), g(".conv_shortcut.bias"))
    return h + x


def _upsample(x, w, b, compress_time: bool):
    if compress_time:
        if x.shape[2] > 1 and x.shape[2] % 2 == 1:
            first = F.interpolate(x[:, :, 0], scale_factor=2.0)[:, :, None]
            x = torch.cat([first, F.interpolate(x[:, :, 1:], scale_factor=2.0)], dim=2)
        elif x.shape[2] > 1:
            x = F.interpolate(x, scale_factor=2.0)
        else:
            x = F.interpolate(x.squeeze(2), scale_factor=2.0)[:, :, None]
    else:
        B, Cc, T, H, W = x.shape
        y = F.interpolate(x.permute(0, 2, 1, 3, 4).reshape(B * T, Cc, H, W), scale_factor=2.0)
        x = y.reshape(B, T, Cc, 2 * H, 2 * W).permute(0, 2, 1, 3, 4)
    B, Cc, T, H, W = x.shape
    y = F.conv2d(x.permute(0, 2, 1, 3, 4).reshape(B * T, Cc, H, W), w, b, padding=1)
    return y.reshape(B, T, y.shape[1], H, W).permute(0, 2, 1, 3, 4)


def _decoder(z, sd, cfg, dt, cache):
    groups = cfg["norm_num_groups"]
    boc = list(reversed(cfg["block_out_channels"]))
    n_compress = {1: 0, 2: 1, 4: 2, 8: 3}[cfg["temporal_compression_ratio"]]
    g = lambda k: sd[k].to(dt)
    h = _cached_conv3d(z, g("decoder.conv_in.conv.weight"), g("decoder.conv_in.conv.bias"), cache, "conv_in")
    for j in range(2):
        h = _dec_resnet(h, z, sd, f"decoder.mid_block.resnets.{j}", groups, dt, cache)
    for bi in range(len(boc)):
        for j in range(cfg["layers_per_block"] + 1):
            h = _dec_resnet(h, z, sd, f"decoder.up_blocks.{bi}.resnets.{j}", groups, dt, cache)
        if bi != len(boc) - 1:
            n = f"decoder.up_blocks.{bi}.upsamplers.0.conv."
            h = _upsample(h, g(n + "weight"), g(n + "bias"), compress_time=bi < n_compress)
    h = F.silu(_spatial_norm(h, z, sd, "decoder.norm_out", groups, dt, cache))
    return _cached_conv3d(h, g("decoder.conv_out.conv.weight"), g("decoder.conv_out.conv.bias"), cache, "conv_out")


def decode(z: torch.Tensor, sd: Dict[str, torch.Tensor], cfg: dict, dtype=torch.float32, frame_batch: int = 2) -> torch.Tensor:
    """z [B, zc, T, h, w] (already divided by the scaling factor, cog:430) -> video [B, 3, 1 + 4 (T - 1), 8h, 8w]."""
    z = z.to(dtype)
    T = z.shape[2]
    n_batches = max(T // frame_batch, 1)
    rem = T % frame_batch
    cache: dict = {}
    out = []
    for i in range(n_batches):
        start = frame_batch * i + (0 if i == 0 else rem)
        end = frame_batch * (i + 1) + rem
        out.append(_decoder(z[:, :, start:end], sd, cfg, dtype, cache))
    return torch.cat(out, dim=2)
