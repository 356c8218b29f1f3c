"""A small ARITHMETIC video VAE with the diffusers surface the three pipelines touch.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Not a restatement of any reference network: the loop fixtures
(``oracle/gen_golden_loops.py`` -> ``tests/golden/loop_*.npz``) need *some* deterministic ``vae.encode`` behind
``prepare_latents`` / ``prepare_lp`` (wan:402-449, 493-540; cog:428-433, 628-680; hy:574-580) so that the REAL reference
code and the drop-in can be driven with the same object.  ``encode`` = causal 4x temporal / 8x spatial average pooling, a
fixed 3 -> z_dim channel mix, and a diagonal Gaussian whose ``sample(generator)`` draws with diffusers' ``randn_tensor``
rule (a CPU generator draws on the CPU), so RNG stream positions are comparable between the two sides.
"""
from __future__ import annotations

import math
from types import SimpleNamespace

import torch
import torch.nn.functional as F

WAN_MEAN = [-0.7571, -0.7089, -0.9113, 0.1075, -0.1745, 0.9653, -0.1517, 1.5508, 0.4134, -0.0715, 0.5517, -0.3632,
            -0.1922, -0.9497, 0.2503, -0.2921]
WAN_STD = [2.8184, 1.4541, 2.3275, 2.6558, 1.2196, 1.7708, 2.6052, 2.0743, 3.2687, 2.1526, 2.8652, 1.5579, 1.6382,
           1.1253, 2.8251, 1.9160]


def _randn(shape, generator, device, dtype):
    if isinstance(generator, list):
        return torch.cat([_randn((1,) + tuple(shape[1:]), g, device, dtype) for g in generator], dim=0)
    if generator is not None and generator.device.type != torch.device(device).type:
        return torch.randn(shape, generator=generator, device="cpu", dtype=dtype).to(device)
    return torch.randn(shape, generator=generator, device=device, dtype=dtype)


class _Gaussian:
    def __init__(self, mean, std):
        self.mean, self.std = mean, std

    def sample(self, generator=None):
        return self.mean + self.std * _randn(self.mean.shape, generator, self.mean.device, self.mean.dtype)

    def mode(self):
        return self.mean


class ArithVAE:
    """``kind`` in {"wan", "cog", "hunyuan"} only selects the config fields each pipeline reads."""

    def __init__(self, kind: str, z_dim: int = 16, dtype=torch.float32, invert_scale_latents: bool = False):
        self.kind, self.z_dim, self.dtype = kind, z_dim, dtype
        self.temperal_downsample = [False, True, True]  # (sic) wan:180-181
        self.temporal_compression_ratio, self.spatial_compression_ratio = 4, 8  # hy:278-279
        cfg = dict(z_dim=z_dim, latent_channels=z_dim)
        if kind == "wan":
            cfg.update(latents_mean=WAN_MEAN[:z_dim], latents_std=WAN_STD[:z_dim])
        elif kind == "cog":
            cfg.update(scaling_factor=0.7, invert_scale_latents=invert_scale_latents, block_out_channels=[8, 16, 16, 32],
                       temporal_compression_ratio=4)
        else:
            cfg.update(scaling_factor=0.476986)
        self.config = SimpleNamespace(**cfg)
        i = torch.arange(z_dim, dtype=torch.float64)[:, None]
        j = torch.arange(3, dtype=torch.float64)[None, :]
        self.mix = (torch.cos(1.7 * i + 2.3 * j + 0.4) * 0.9).to(torch.float32)  # [z, 3]
        self.bias = (torch.sin(0.9 * i[:, 0] + 0.2) * 0.3).to(torch.float32)

    def to(self, *a, **k):
        return self

    def encode(self, x):
        """x [B, 3, T, H, W] -> latent_dist over [B, z, 1 + (T-1)//4, H/8, W/8]."""
        B, C, T, H, W = x.shape
        xf = x.to(torch.float32)
        xs = F.avg_pool3d(xf, kernel_size=(1, 8, 8))
        parts = [xs[:, :, :1]]
        if T > 1:
            rest = xs[:, :, 1:]
            pad = (-rest.shape[2]) % 4
            if pad:
                rest = torch.cat([rest, rest[:, :, -1:].expand(-1, -1, pad, -1, -1)], dim=2)
            parts.append(F.avg_pool3d(rest, kernel_size=(4, 1, 1)))
        pooled = torch.cat(parts, dim=2)
        mean = torch.einsum("zc,bcthw->bzthw", self.mix.to(x.device), pooled) + self.bias.to(x.device).view(1, -1, 1, 1, 1)
        std = 0.05 + 0.1 * torch.sigmoid(mean)
        return SimpleNamespace(latent_dist=_Gaussian(mean.to(x.dtype), std.to(x.dtype)))

    def decode(self, z, return_dict=True):
        B, Cz, T, H, W = z.shape
        rgb = torch.einsum("zc,bzthw->bcthw", self.mix.to(z.device), z.to(torch.float32)) / math.sqrt(self.z_dim)
        rgb = F.interpolate(rgb, scale_factor=(1, 8, 8), mode="nearest")
        rgb = torch.cat([rgb[:, :, :1], rgb[:, :, 1:].repeat_interleave(4, dim=2)], dim=2).to(z.dtype)
        if not return_dict:
            return (rgb,)
        return SimpleNamespace(sample=rgb)
