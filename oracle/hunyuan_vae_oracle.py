"""CPU/GPU restatement of ``AutoencoderKLHunyuanVideo.encode`` / ``decode`` -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows diffusers@be2fb77 ``models/autoencoders/autoencoder_kl_hunyuan_video.py`` (absent offline: **parity unpinned**).  Call
sites in the reference: hy:576-581 (``retrieve_latents(self.vae.encode(image[i].unsqueeze(0)), ..., "argmax")`` on ONE frame)
and hy:1292 (``decode``); run.py:76-80 loads the VAE in float16.

  HunyuanVideoCausalConv3d      F.pad(replicate): (k//2, k//2, k//2, k//2, k - 1, 0) over (W, H, T-front), then Conv3d (stride s)
  HunyuanVideoResnetBlockCausal3D  GroupNorm(32) -> SiLU -> conv1 -> GroupNorm -> SiLU -> conv2, + (1x1x1 conv_shortcut | identity)
  HunyuanVideoDownsampleCausal3D   the causal conv with stride (1|2, 2, 2)
  HunyuanVideoUpsampleCausal3D     nearest: frame 0 spatially x2, later frames x(2|1, 2, 2); then the causal conv
  HunyuanVideoMidBlock3D           resnet, attention (ONE head of C channels over ALL T*H*W tokens, GroupNorm first, frame-causal
                                   mask: a token sees the frames up to its own, + residual), resnet
"""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn.functional as F

HUNYUAN_VAE = dict(in_channels=3, out_channels=3, latent_channels=16, block_out_channels=(128, 256, 512, 512), layers_per_block=2,
                   act_fn="silu", norm_num_groups=32, scaling_factor=0.476986, spatial_compression_ratio=8,
                   temporal_compression_ratio=4, mid_block_add_attention=True)


def stage_plan(cfg: dict):
    """[(c_in, c_out, spatial_resample, temporal_resample)] per encoder down block; the decoder mirrors it."""
    boc = list(cfg["block_out_channels"])
    n = len(boc)
    ns = {1: 0, 2: 1, 4: 2, 8: 3}[cfg["spatial_compression_ratio"]]
    nt = {1: 0, 2: 1, 4: 2, 8: 3}[cfg["temporal_compression_ratio"]]
    if cfg["temporal_compression_ratio"] != 4:
        raise NotImplementedError("temporal_compression_ratio other than 4")
    enc, dec = [], []
    cin = boc[0]
    for i, cout in enumerate(boc):
        final = i == n - 1
        enc.append((cin, cout, i < ns, i >= (n - 1 - nt) and not final))
        cin = cout
    rev = boc[::-1]
    cin = rev[0]
    for i, cout in enumerate(rev):
        final = i == n - 1
        dec.append((cin, cout, i < ns, i >= (n - 1 - nt) and not final))
        cin = cout
    return enc, dec


def parameter_shapes(cfg: dict) -> Dict[str, tuple]:
    enc, dec = stage_plan(cfg)
    boc, z, L = list(cfg["block_out_channels"]), cfg["latent_channels"], cfg["layers_per_block"]
    s: Dict[str, tuple] = {}

    def conv(name, co, ci, k=3):
        s[name + ".conv.weight"], s[name + ".conv.bias"] = (co, ci, k, k, k), (co,)

    def res(name, ci, co):
        s[name + ".norm1.weight"] = s[name + ".norm1.bias"] = (ci,)
        s[name + ".norm2.weight"] = s[name + ".norm2.bias"] = (co,)
        conv(name + ".conv1", co, ci)
        conv(name + ".conv2", co, co)
        if ci != co:
            conv(name + ".conv_shortcut", co, ci, 1)

    def mid(name, c):
        res(name + ".resnets.0", c, c)
        a = name + ".attentions.0"
        s[a + ".group_norm.weight"] = s[a + ".group_norm.bias"] = (c,)
        for n in ("to_q", "to_k", "to_v", "to_out.0"):
            s[f"{a}.{n}.weight"], s[f"{a}.{n}.bias"] = (c, c), (c,)
        res(name + ".resnets.1", c, c)

    conv("encoder.conv_in", boc[0], cfg["in_channels"])
    for i, (ci, co, sp, tp) in enumerate(enc):
        for j in range(L):
            res(f"encoder.down_blocks.{i}.resnets.{j}", ci if j == 0 else co, co)
        if sp or tp:
            conv(f"encoder.down_blocks.{i}.downsamplers.0.conv", co, co)
    mid("encoder.mid_block", boc[-1])
    s["encoder.conv_norm_out.weight"] = s["encoder.conv_norm_out.bias"] = (boc[-1],)
    conv("encoder.conv_out", 2 * z, boc[-1])
    s["quant_conv.weight"], s["quant_conv.bias"] = (2 * z, 2 * z, 1, 1, 1), (2 * z,)
    s["post_quant_conv.weight"], s["post_quant_conv.bias"] = (z, z, 1, 1, 1), (z,)
    conv("decoder.conv_in", boc[-1], z)
    mid("decoder.mid_block", boc[-1])
    for i, (ci, co, sp, tp) in enumerate(dec):
        for j in range(L + 1):
            res(f"decoder.up_blocks.{i}.resnets.{j}", ci if j == 0 else co, co)
        if sp or tp:
            conv(f"decoder.up_blocks.{i}.upsamplers.0.conv", co, co)
    s["decoder.conv_norm_out.weight"] = s["decoder.conv_norm_out.bias"] = (boc[0],)
    conv("decoder.conv_out", cfg["out_channels"], boc[0])
    return s


def make_weights(cfg: dict, seed: int = 0, device="cpu", dtype=torch.float32) -> Dict[str, torch.Tensor]:
    sd = {}
    for idx, (name, shape) in enumerate(parameter_shapes(cfg).items()):
        g = torch.Generator().manual_seed(seed * 1_000_003 + 31 + idx)
        if "norm" in name and name.endswith(".weight"):
            w = 1 + 0.1 * torch.randn(shape, generator=g)
        elif name.endswith(".bias"):
            w = 0.05 * torch.randn(shape, generator=g)
        else:
            fan_in = 1
            for d in shape[1:]:
                fan_in *= d
            w = torch.randn(shape, generator=g) * (1.2 * fan_in ** -0.5)
        sd[name] = w.to(device=device, dtype=dtype)
    return sd


class _Net:
    def __init__(self, sd, cfg, dtype):
        self.sd, self.cfg, self.dt = sd, cfg, dtype
        self.groups = cfg["norm_num_groups"]

    def p(self, n):
        return self.sd[n].to(self.dt)

    def conv(self, x, name, stride=(1, 1, 1)):
        w, b = self.p(name + ".conv.weight"), self.p(name + ".conv.bias")
        k = w.shape[2]
        if k > 1:
            x = F.pad(x, (k // 2, k // 2, k // 2, k // 2, k - 1, 0), mode="replicate")
        return F.conv3d(x, w, b, stride=stride)

    def res(self, x, name):
        h = F.silu(F.group_norm(x, self.groups, self.p(name + ".norm1.weight"), self.p(name + ".norm1.bias"), 1e-6))
        h = self.conv(h, name + ".conv1")
        h = F.silu(F.group_norm(h, self.groups, self.p(name + ".norm2.weight"), self.p(name + ".norm2.bias"), 1e-6))
        h = self.conv(h, name + ".conv2")
        if name + ".conv_shortcut.conv.weight" in self.sd:
            x = self.conv(x, name + ".conv_shortcut")
        return h + x

    def mid(self, x, name):
        x = self.res(x, name + ".resnets.0")
        if self.cfg.get("mid_block_add_attention", True):
            a = name + ".attentions.0"
            B, C, T, H, W = x.shape
            t = x.permute(0, 2, 3, 4, 1).flatten(1, 3)  # [B, N, C]
            n = F.group_norm(t.transpose(1, 2), self.groups, self.p(a + ".group_norm.weight"), self.p(a + ".group_norm.bias"), 1e-6).transpose(1, 2)
            q = F.linear(n, self.p(a + ".to_q.weight"), self.p(a + ".to_q.bias"))
            k = F.linear(n, self.p(a + ".to_k.weight"), self.p(a + ".to_k.bias"))
            v = F.linear(n, self.p(a + ".to_v.weight"), self.p(a + ".to_v.bias"))
            frame = torch.arange(T * H * W, device=x.device) // (H * W)
            mask = (frame[None, :] <= frame[:, None])  # token i sees the frames up to its own
            o = F.scaled_dot_product_attention(q[:, None], k[:, None], v[:, None], attn_mask=mask[None, None])[:, 0]
            o = F.linear(o, self.p(a + ".to_out.0.weight"), self.p(a + ".to_out.0.bias")) + t
            x = o.unflatten(1, (T, H, W)).permute(0, 4, 1, 2, 3)
        return self.res(x, name + ".resnets.1")


def encode_moments(x, sd, cfg, dtype=torch.float32):
    """x [B, 3, T, H, W] -> [B, 2 z, 1 + (T - 1) / 4, H/8, W/8] (mean | logvar)."""
    net = _Net(sd, cfg, dtype)
    enc, _ = stage_plan(cfg)
    h = net.conv(x.to(dtype), "encoder.conv_in")
    for i, (ci, co, sp, tp) in enumerate(enc):
        for j in range(cfg["layers_per_block"]):
            h = net.res(h, f"encoder.down_blocks.{i}.resnets.{j}")
        if sp or tp:
            h = net.conv(h, f"encoder.down_blocks.{i}.downsamplers.0.conv", stride=(2 if tp else 1, 2 if sp else 1, 2 if sp else 1))
    h = net.mid(h, "encoder.mid_block")
    h = F.silu(F.group_norm(h, net.groups, net.p("encoder.conv_norm_out.weight"), net.p("encoder.conv_norm_out.bias"), 1e-6))
    h = net.conv(h, "encoder.conv_out")
    return F.conv3d(h, net.p("quant_conv.weight"), net.p("quant_conv.bias"))


def decode(z, sd, cfg, dtype=torch.float32, tile_min_frames=16, tile_stride_frames=12, framewise=True):
    """``AutoencoderKLHunyuanVideo._decode``: z [B, z, T, h, w] -> [B, 3, 4 (T - 1) + 1, 8 h, 8 w].  With diffusers' default
    ``use_framewise_decoding`` a clip of more than 4 latent frames is decoded in overlapping temporal tiles of 5 latent frames
    (stride 3) and cross-faded over 4 frames (``_temporal_tiled_decode`` + ``blend_t``)."""
    r = cfg["temporal_compression_ratio"]
    t_min, t_stride = tile_min_frames // r, tile_stride_frames // r
    T = z.shape[2]
    if not (framewise and T > t_min):
        return decode_tile(z, sd, cfg, dtype)
    blend = tile_min_frames - tile_stride_frames
    row = []
    for i in range(0, T, t_stride):
        d = decode_tile(z[:, :, i:i + t_min + 1], sd, cfg, dtype)
        row.append(d if i == 0 else d[:, :, 1:])
    out = []
    for i, tile in enumerate(row):
        if i > 0:
            a = row[i - 1]
            n = min(a.shape[2], tile.shape[2], blend)
            tile = tile.clone()
            for x in range(n):
                tile[:, :, x] = a[:, :, -n + x] * (1 - x / n) + tile[:, :, x] * (x / n)
            row[i] = tile
            out.append(tile[:, :, :tile_stride_frames])
        else:
            out.append(tile[:, :, :tile_stride_frames + 1])
    return torch.cat(out, dim=2)[:, :, :(T - 1) * r + 1]


def decode_tile(z, sd, cfg, dtype=torch.float32):
    """One temporal tile: post_quant_conv + decoder."""
    net = _Net(sd, cfg, dtype)
    _, dec = stage_plan(cfg)
    h = F.conv3d(z.to(dtype), net.p("post_quant_conv.weight"), net.p("post_quant_conv.bias"))
    h = net.conv(h, "decoder.conv_in")
    h = net.mid(h, "decoder.mid_block")
    for i, (ci, co, sp, tp) in enumerate(dec):
        for j in range(cfg["layers_per_block"] + 1):
            h = net.res(h, f"decoder.up_blocks.{i}.resnets.{j}")
        if sp or tp:
            T = h.shape[2]
            f = (2.0 if sp else 1.0, 2.0 if sp else 1.0)
            first = F.interpolate(h[:, :, 0], scale_factor=f, mode="nearest").unsqueeze(2)
            if T > 1:
                other = F.interpolate(h[:, :, 1:].contiguous(), scale_factor=(2.0 if tp else 1.0, *f), mode="nearest")
                h = torch.cat((first, other), dim=2)
            else:
                h = first
            h = net.conv(h, f"decoder.up_blocks.{i}.upsamplers.0.conv")
    h = F.silu(F.group_norm(h, net.groups, net.p("decoder.conv_norm_out.weight"), net.p("decoder.conv_norm_out.bias"), 1e-6))
    return net.conv(h, "decoder.conv_out")
