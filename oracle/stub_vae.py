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


class _Batch(dict):
    def to(self, *a, **k):
        return self


class StubImageProcessor:
    """``CLIPImageProcessor`` call surface (wan:232): ``processor(images=..., return_tensors="pt").to(device)``."""

    def __call__(self, images=None, return_tensors="pt"):
        n = len(images) if isinstance(images, (list, tuple)) else (images.shape[0] if torch.is_tensor(images) and images.ndim == 4 else 1)
        return _Batch(pixel_values=torch.zeros(n, 3, 2, 2))


class StubImageEncoder:
    """``CLIPVisionModel`` call surface (wan:233-234): ``hidden_states[-2]`` of image k is ``table[k]`` ([n, tokens, dim])."""

    def __init__(self, table: torch.Tensor):
        self.table = table
        self.dtype = table.dtype

    def to(self, *a, **k):
        return self

    def __call__(self, pixel_values=None, output_hidden_states=True, **kw):
        h = self.table[: pixel_values.shape[0]]
        return SimpleNamespace(hidden_states=[h * 0, h, h * 0])


class StubTokenizer:
    """HF tokenizer call surface (wan:200-209, cog:244-253): whitespace words -> ids in [2, vocab), one EOS (id 1), padded
    with 0 to ``max_length``; returns ``input_ids`` / ``attention_mask``."""

    def __init__(self, vocab=997):
        self.vocab = vocab

    def __call__(self, prompt, padding="max_length", max_length=32, truncation=True, add_special_tokens=True,
                 return_attention_mask=True, return_tensors="pt"):
        prompt = [prompt] if isinstance(prompt, str) else prompt
        ids = torch.zeros(len(prompt), max_length, dtype=torch.int64)
        mask = torch.zeros(len(prompt), max_length, dtype=torch.int64)
        for b, text in enumerate(prompt):
            toks = [2 + sum(ord(ch) * (i + 1) for i, ch in enumerate(w)) % (self.vocab - 2) for w in text.split()][: max_length - 1] + [1]
            ids[b, : len(toks)] = torch.tensor(toks)
            mask[b, : len(toks)] = 1
        return SimpleNamespace(input_ids=ids, attention_mask=mask)


class _EncoderOutput(SimpleNamespace):
    def __getitem__(self, i):
        return (self.last_hidden_state,)[i]


class StubTextEncoder:
    """HF encoder call surface (wan:212, cog:258): ``encoder(ids, mask).last_hidden_state`` (``encoder(ids)[0]`` for T5)."""

    def __init__(self, dim=64, vocab=997, dtype=torch.bfloat16, seed=99):
        g = torch.Generator().manual_seed(seed)
        self.table = torch.randn(vocab, dim, generator=g).to(dtype)
        self.dtype = dtype

    def to(self, *a, **k):
        return self

    def __call__(self, input_ids, attention_mask=None, **kw):
        h = self.table.to(input_ids.device)[input_ids]
        return _EncoderOutput(last_hidden_state=h)
