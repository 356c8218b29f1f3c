"""CPU/GPU restatement of ``AutoencoderKLWan.encode`` / ``decode`` -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows diffusers@be2fb77 ``models/autoencoders/autoencoder_kl_wan.py`` (``requirements.txt:13`` of the reference pins it; the
package is NOT available offline, so this is written from its published source and is **parity unpinned**).  Call sites in the
reference: wan:429-434 (``retrieve_latents(self.vae.encode(video_condition), "argmax")`` on the image + zero frames), wan:526
(pixel-space low-pass: ``encode(...).latent_dist.sample(generator)`` EVERY step), wan:959 (``decode`` -- the frames the metric
counts); ``run.py:51-55`` loads the VAE in float32.

The restatement keeps diffusers' CHUNKED evaluation with per-convolution feature caches (encoder: 1 frame, then 4 at a time;
decoder: one latent frame at a time), because two of its quirks only exist there:
  * ``WanResample`` "downsample3d" / "upsample3d" skip their temporal convolution on the first chunk (frame 0 passes through);
  * "upsample3d" marks its cache "Rep" on the first chunk and then front-pads with ZEROS: frame 0 never enters a temporal window.
The product (alg_b200/vae_wan.py) evaluates the same network in its closed, whole-clip form; tests compare the two.

  WanCausalConv3d   zero padding: (kt - 1) frames in front (minus the cached frames), kh // 2, kw // 2 around
  WanRMS_norm       F.normalize(x, dim=channel) * sqrt(C) * gamma
  WanResidualBlock  norm1 -> SiLU -> conv1 -> norm2 -> SiLU -> conv2, + (1x1x1 conv_shortcut | identity)
  WanAttentionBlock per frame, ONE head of C channels over the H*W pixels: norm -> to_qkv (1x1) -> SDPA -> proj (1x1), + identity
  WanResample       upsample: nearest-exact x2 + Conv2d(C, C/2, 3, pad 1); downsample: ZeroPad2d((0,1,0,1)) + Conv2d(C, C, 3, stride 2)
                    3-D variants add time_conv (3,1,1): C -> 2C then channel halves -> two frames (up); stride 2 in time (down)
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

CACHE_T = 2

WAN21_VAE = dict(base_dim=96, z_dim=16, dim_mult=[1, 2, 4, 4], num_res_blocks=2, attn_scales=[], temperal_downsample=[False, True, True],
                 latents_mean=[-0.7571, -0.7089, -0.9113, 0.1075, -0.1745, 0.9653, -0.1517, 1.5508, 0.4134, -0.0715, 0.5517, -0.3632,
                               -0.1922, -0.9497, 0.2503, -0.2921],
                 latents_std=[2.8184, 1.4541, 2.3275, 2.6558, 1.2196, 1.7708, 2.6052, 2.0743, 3.2687, 2.1526, 2.8652, 1.5579, 1.6382,
                              1.1253, 2.8251, 1.9160])


def layer_plan(cfg: dict):
    """(encoder down_blocks, decoder up_blocks) as lists of ("res", cin, cout) / ("attn", c) / ("down2d" | "down3d" | "up2d" |
    "up3d", c) in module order, with the diffusers module name of each."""
    dim, mult, nres = cfg["base_dim"], list(cfg["dim_mult"]), cfg["num_res_blocks"]
    tdown, attn_scales = list(cfg["temperal_downsample"]), list(cfg.get("attn_scales", []))
    dims = [dim * u for u in [1] + mult]
    enc, scale, k = [], 1.0, 0
    for i, (cin, cout) in enumerate(zip(dims[:-1], dims[1:])):
        for _ in range(nres):
            enc.append((f"encoder.down_blocks.{k}", "res", cin, cout))
            k += 1
            if scale in attn_scales:
                enc.append((f"encoder.down_blocks.{k}", "attn", cout, cout))
                k += 1
            cin = cout
        if i != len(mult) - 1:
            enc.append((f"encoder.down_blocks.{k}", "down3d" if tdown[i] else "down2d", cout, cout))
            k += 1
            scale /= 2.0
    ddims = [dim * u for u in [mult[-1]] + mult[::-1]]
    tup = tdown[::-1]
    dec = []
    for i, (cin, cout) in enumerate(zip(ddims[:-1], ddims[1:])):
        if i > 0:
            cin = cin // 2
        for j in range(nres + 1):
            dec.append((f"decoder.up_blocks.{i}.resnets.{j}", "res", cin, cout))
            cin = cout
        if i != len(mult) - 1:
            dec.append((f"decoder.up_blocks.{i}.upsamplers.0", "up3d" if tup[i] else "up2d", cout, cout // 2))
    return enc, dec, dims[-1], ddims[0], ddims[-1]


def parameter_shapes(cfg: dict) -> Dict[str, tuple]:
    enc, dec, c_enc, c_dec_in, c_dec_out = layer_plan(cfg)
    z = cfg["z_dim"]
    s: Dict[str, tuple] = {}

    def conv3(name, cout, cin, k):
        s[name + ".weight"], s[name + ".bias"] = (cout, cin, *k), (cout,)

    def res(name, cin, cout):
        s[name + ".norm1.gamma"], s[name + ".norm2.gamma"] = (cin, 1, 1, 1), (cout, 1, 1, 1)
        conv3(name + ".conv1", cout, cin, (3, 3, 3))
        conv3(name + ".conv2", cout, cout, (3, 3, 3))
        if cin != cout:
            conv3(name + ".conv_shortcut", cout, cin, (1, 1, 1))

    def attn(name, c):
        s[name + ".norm.gamma"] = (c, 1, 1)
        s[name + ".to_qkv.weight"], s[name + ".to_qkv.bias"] = (3 * c, c, 1, 1), (3 * c,)
        s[name + ".proj.weight"], s[name + ".proj.bias"] = (c, c, 1, 1), (c,)

    def mid(name, c):
        res(name + ".resnets.0", c, c)
        attn(name + ".attentions.0", c)
        res(name + ".resnets.1", c, c)

    conv3("encoder.conv_in", cfg["base_dim"], 3, (3, 3, 3))
    for name, kind, cin, cout in enc:
        if kind == "res":
            res(name, cin, cout)
        elif kind == "attn":
            attn(name, cin)
        else:
            s[name + ".resample.1.weight"], s[name + ".resample.1.bias"] = (cout, cin, 3, 3), (cout,)
            if kind == "down3d":
                conv3(name + ".time_conv", cout, cin, (3, 1, 1))
    mid("encoder.mid_block", c_enc)
    s["encoder.norm_out.gamma"] = (c_enc, 1, 1, 1)
    conv3("encoder.conv_out", 2 * z, c_enc, (3, 3, 3))
    conv3("quant_conv", 2 * z, 2 * z, (1, 1, 1))
    conv3("post_quant_conv", z, z, (1, 1, 1))
    conv3("decoder.conv_in", c_dec_in, z, (3, 3, 3))
    mid("decoder.mid_block", c_dec_in)
    for name, kind, cin, cout in dec:
        if kind == "res":
            res(name, cin, cout)
        else:
            s[name + ".resample.1.weight"], s[name + ".resample.1.bias"] = (cout, cin, 3, 3), (cout,)
            if kind == "up3d":
                conv3(name + ".time_conv", 2 * cin, cin, (3, 1, 1))
    s["decoder.norm_out.gamma"] = (c_dec_out, 1, 1, 1)
    conv3("decoder.conv_out", 3, c_dec_out, (3, 3, 3))
    return s


def make_weights(cfg: dict, seed: int = 0, device="cpu", dtype=torch.float32) -> Dict[str, torch.Tensor]:
    sd = {}
    for idx, (name, shape) in enumerate(parameter_shapes(cfg).items()):
        g = torch.Generator().manual_seed(seed * 1_000_003 + idx)
        if name.endswith(".gamma"):
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


# ---- modules ---------------------------------------------------------------------------------------------------------
def causal_conv3d(x, w, b, cache_x=None, stride=(1, 1, 1)):
    kt, kh, kw = w.shape[2:]
    pad_t = kt - 1
    if cache_x is not None and pad_t > 0:
        x = torch.cat([cache_x.to(x.device), x], dim=2)
        pad_t -= cache_x.shape[2]
    x = F.pad(x, (kw // 2, kw // 2, kh // 2, kh // 2, pad_t, 0))
    return F.conv3d(x, w, b, stride=stride)


def rms_norm(x, gamma, channel_dim=1):
    return F.normalize(x, dim=channel_dim) * (x.shape[channel_dim] ** 0.5) * gamma


class _Net:
    def __init__(self, sd: Dict[str, torch.Tensor], cfg: dict, dtype):
        self.sd, self.cfg, self.dt = sd, cfg, dtype

    def p(self, name):
        return self.sd[name].to(self.dt)

    def cached_conv(self, x, name, cache: Optional[list], idx: List[int]):
        """The cache protocol WanResidualBlock / encoder / decoder wrap around every 3x3x3 convolution."""
        if cache is None:
            return causal_conv3d(x, self.p(name + ".weight"), self.p(name + ".bias"))
        i = idx[0]
        cache_x = x[:, :, -CACHE_T:].clone()
        if cache_x.shape[2] < 2 and cache[i] is not None:
            cache_x = torch.cat([cache[i][:, :, -1:].to(cache_x.device), cache_x], dim=2)
        y = causal_conv3d(x, self.p(name + ".weight"), self.p(name + ".bias"), cache[i])
        cache[i] = cache_x
        idx[0] += 1
        return y

    def res(self, x, name, cache, idx):
        h = x
        if name + ".conv_shortcut.weight" in self.sd:
            h = causal_conv3d(x, self.p(name + ".conv_shortcut.weight"), self.p(name + ".conv_shortcut.bias"))
        x = F.silu(rms_norm(x, self.p(name + ".norm1.gamma")))
        x = self.cached_conv(x, name + ".conv1", cache, idx)
        x = F.silu(rms_norm(x, self.p(name + ".norm2.gamma")))
        x = self.cached_conv(x, name + ".conv2", cache, idx)
        return x + h

    def attn(self, x, name):
        B, C, T, H, W = x.shape
        y = x.permute(0, 2, 1, 3, 4).reshape(B * T, C, H, W)
        y = rms_norm(y, self.p(name + ".norm.gamma"))
        qkv = F.conv2d(y, self.p(name + ".to_qkv.weight"), self.p(name + ".to_qkv.bias"))
        qkv = qkv.reshape(B * T, 1, C * 3, -1).permute(0, 1, 3, 2).contiguous()
        q, k, v = qkv.chunk(3, dim=-1)
        y = F.scaled_dot_product_attention(q, k, v)
        y = y.squeeze(1).permute(0, 2, 1).reshape(B * T, C, H, W)
        y = F.conv2d(y, self.p(name + ".proj.weight"), self.p(name + ".proj.bias"))
        return y.view(B, T, C, H, W).permute(0, 2, 1, 3, 4) + x

    def mid(self, x, name, cache, idx):
        x = self.res(x, name + ".resnets.0", cache, idx)
        x = self.attn(x, name + ".attentions.0")
        return self.res(x, name + ".resnets.1", cache, idx)

    def resample(self, x, name, kind, cache, idx):
        B, C, T, H, W = x.shape
        if kind == "up3d" and cache is not None:
            i = idx[0]
            if cache[i] is None:
                cache[i] = "Rep"
                idx[0] += 1
            else:
                cache_x = x[:, :, -CACHE_T:].clone()
                if cache_x.shape[2] < 2 and not isinstance(cache[i], str):
                    cache_x = torch.cat([cache[i][:, :, -1:].to(cache_x.device), cache_x], dim=2)
                if cache_x.shape[2] < 2 and isinstance(cache[i], str):
                    cache_x = torch.cat([torch.zeros_like(cache_x), cache_x], dim=2)
                w, b = self.p(name + ".time_conv.weight"), self.p(name + ".time_conv.bias")
                x = causal_conv3d(x, w, b) if isinstance(cache[i], str) else causal_conv3d(x, w, b, cache[i])
                cache[i] = cache_x
                idx[0] += 1
                x = x.reshape(B, 2, C, T, H, W)
                x = torch.stack((x[:, 0], x[:, 1]), 3).reshape(B, C, T * 2, H, W)
        T = x.shape[2]
        y = x.permute(0, 2, 1, 3, 4).reshape(B * T, C, H, W)
        w, b = self.p(name + ".resample.1.weight"), self.p(name + ".resample.1.bias")
        if kind in ("up2d", "up3d"):
            y = F.interpolate(y.float(), scale_factor=(2.0, 2.0), mode="nearest-exact").type_as(y)
            y = F.conv2d(y, w, b, padding=1)
        else:
            y = F.conv2d(F.pad(y, (0, 1, 0, 1)), w, b, stride=2)
        x = y.view(B, T, y.shape[1], y.shape[2], y.shape[3]).permute(0, 2, 1, 3, 4)
        if kind == "down3d" and cache is not None:
            i = idx[0]
            if cache[i] is None:
                cache[i] = x.clone()
                idx[0] += 1
            else:
                cache_x = x[:, :, -1:].clone()
                x = F.conv3d(torch.cat([cache[i][:, :, -1:], x], 2), self.p(name + ".time_conv.weight"),
                             self.p(name + ".time_conv.bias"), stride=(2, 1, 1))
                cache[i] = cache_x
                idx[0] += 1
        return x

    # ---- encoder / decoder --------------------------------------------------------------------------------------------
    def encoder(self, x, cache, idx):
        enc, _, _, _, _ = layer_plan(self.cfg)
        x = self.cached_conv(x, "encoder.conv_in", cache, idx)
        for name, kind, cin, cout in enc:
            if kind == "res":
                x = self.res(x, name, cache, idx)
            elif kind == "attn":
                x = self.attn(x, name)
            else:
                x = self.resample(x, name, kind, cache, idx)
        x = self.mid(x, "encoder.mid_block", cache, idx)
        x = F.silu(rms_norm(x, self.p("encoder.norm_out.gamma")))
        return self.cached_conv(x, "encoder.conv_out", cache, idx)

    def decoder(self, x, cache, idx):
        _, dec, _, _, _ = layer_plan(self.cfg)
        x = self.cached_conv(x, "decoder.conv_in", cache, idx)
        x = self.mid(x, "decoder.mid_block", cache, idx)
        for name, kind, cin, cout in dec:
            x = self.res(x, name, cache, idx) if kind == "res" else self.resample(x, name, kind, cache, idx)
        x = F.silu(rms_norm(x, self.p("decoder.norm_out.gamma")))
        return self.cached_conv(x, "decoder.conv_out", cache, idx)


def _count_cached_convs(cfg: dict, decoder: bool) -> int:
    enc, dec, _, _, _ = layer_plan(cfg)
    n = 2 + 4  # conv_in, conv_out, two mid resnets
    for name, kind, cin, cout in (dec if decoder else enc):
        n += 2 if kind == "res" else (1 if kind in ("up3d", "down3d") else 0)
    return n


def encode_moments(x: torch.Tensor, sd: Dict[str, torch.Tensor], cfg: dict, dtype=torch.float32) -> torch.Tensor:
    """``AutoencoderKLWan._encode``: x [B, 3, 1 + 4n, H, W] -> [B, 2 z, 1 + n, H/8, W/8] (mean | logvar)."""
    net = _Net(sd, cfg, dtype)
    x = x.to(dtype)
    T = x.shape[2]
    cache = [None] * _count_cached_convs(cfg, decoder=False)
    out = None
    for i in range(1 + (T - 1) // 4):
        idx = [0]
        chunk = x[:, :, :1] if i == 0 else x[:, :, 1 + 4 * (i - 1):1 + 4 * i]
        o = net.encoder(chunk, cache, idx)
        out = o if out is None else torch.cat([out, o], 2)
    return causal_conv3d(out, net.p("quant_conv.weight"), net.p("quant_conv.bias"))


def decode(z: torch.Tensor, sd: Dict[str, torch.Tensor], cfg: dict, dtype=torch.float32) -> torch.Tensor:
    """``AutoencoderKLWan._decode``: z [B, z, T, h, w] -> [B, 3, 4 T - 3, 8 h, 8 w], clamped to [-1, 1]."""
    net = _Net(sd, cfg, dtype)
    x = causal_conv3d(z.to(dtype), net.p("post_quant_conv.weight"), net.p("post_quant_conv.bias"))
    cache = [None] * _count_cached_convs(cfg, decoder=True)
    out = None
    for i in range(x.shape[2]):
        idx = [0]
        o = net.decoder(x[:, :, i:i + 1], cache, idx)
        out = o if out is None else torch.cat([out, o], 2)
    return torch.clamp(out, min=-1.0, max=1.0)


# ---- the same network in closed (whole-clip) form: what alg_b200/vae_wan.py implements ------------------------------------
def decode_closed_form(z, sd, cfg, dtype=torch.float32):
    """Whole-clip evaluation with no caches: every 3x3x3 convolution is a zero-padded causal convolution over all frames,
    "upsample3d" = frame 0 passes, frames t >= 1 see the window (t-2, t-1, t) with frame 0 (and anything before) read as zero.
    Must equal ``decode`` (tests/test_oracle_wan_vae.py)."""
    net = _Net(sd, cfg, dtype)
    _, dec, _, _, _ = layer_plan(cfg)
    x = causal_conv3d(z.to(dtype), net.p("post_quant_conv.weight"), net.p("post_quant_conv.bias"))
    x = net.cached_conv(x, "decoder.conv_in", None, None)
    x = net.mid(x, "decoder.mid_block", None, None)
    for name, kind, cin, cout in dec:
        if kind == "res":
            x = net.res(x, name, None, None)
            continue
        if kind == "up3d" and x.shape[2] > 1:
            B, C, T, H, W = x.shape
            xz = x.clone()
            xz[:, :, 0] = 0
            y = causal_conv3d(xz, net.p(name + ".time_conv.weight"), net.p(name + ".time_conv.bias"))[:, :, 1:]
            y = y.reshape(B, 2, C, T - 1, H, W)
            y = torch.stack((y[:, 0], y[:, 1]), 3).reshape(B, C, 2 * (T - 1), H, W)
            x = torch.cat([x[:, :, :1], y], dim=2)
        x = net.resample(x, name, "up2d", None, None)
    x = F.silu(rms_norm(x, net.p("decoder.norm_out.gamma")))
    return torch.clamp(net.cached_conv(x, "decoder.conv_out", None, None), -1.0, 1.0)


def encode_closed_form(x, sd, cfg, dtype=torch.float32):
    """Whole-clip encoder: "downsample3d" = frame 0 passes, output k >= 1 = time_conv over frames (2k-2, 2k-1, 2k)."""
    net = _Net(sd, cfg, dtype)
    enc, _, _, _, _ = layer_plan(cfg)
    x = net.cached_conv(x.to(dtype), "encoder.conv_in", None, None)
    for name, kind, cin, cout in enc:
        if kind == "res":
            x = net.res(x, name, None, None)
        elif kind == "attn":
            x = net.attn(x, name)
        else:
            x = net.resample(x, name, "down2d", None, None)
            if kind == "down3d" and x.shape[2] > 1:
                y = F.conv3d(x, net.p(name + ".time_conv.weight"), net.p(name + ".time_conv.bias"), stride=(2, 1, 1))
                x = torch.cat([x[:, :, :1], y], dim=2)
    x = net.mid(x, "encoder.mid_block", None, None)
    x = F.silu(rms_norm(x, net.p("encoder.norm_out.gamma")))
    x = net.cached_conv(x, "encoder.conv_out", None, None)
    return causal_conv3d(x, net.p("quant_conv.weight"), net.p("quant_conv.bias"))
