"""``AutoencoderKLHunyuanVideo`` on the sm_100a kernels of ``libalg_b200.so``: ``encode`` and ``decode`` through the float32
split-GEMM path of ``alg_b200/vae_wan.py``.

Reference call sites: hy:576-581 (``retrieve_latents(self.vae.encode(image[i].unsqueeze(0)), ..., "argmax")``: ONE frame -> the
first-frame latent every step's model input carries) and hy:1292 (``decode``).  run.py:76-80 loads the VAE in float16; the
arithmetic here is float32 (a superset).  The network is diffusers@be2fb77 ``autoencoder_kl_hunyuan_video.py`` (absent offline:
parity unpinned; ``oracle/hunyuan_vae_oracle.py`` restates it, ``tests/test_gpu_vae_hunyuan.py`` compares).

    HunyuanVideoCausalConv3d          multi-frame stride-1: implicit -- alg_norm_split_pad_f32 (split only) + alg_replicate_border_bf16
                                      (F.pad(mode="replicate") = the padding of the operand raster filled from the nearest interior
                                      pixel) + tap-mode alg_gemm_bf16; otherwise alg_im2col_split3_f32 with replicate = 1 (index
                                      clamping inside the gather) + alg_gemm_bf16
    HunyuanVideoDownsampleCausal3D    the same gather with stride (1|2, 2, 2)
    HunyuanVideoUpsampleCausal3D      the same gather with up = 2 / tdup = 2: frame 0 once, later frames twice, pixels x2 -- the
                                      nearest-neighbour resize is addressing, never materialised
    GroupNorm(32) (+ SiLU)            alg_group_norm_f32
    mid-block attention               one head of C channels over ALL T*H*W tokens, frame-causal: scores and P V through the split
                                      GEMM, alg_softmax_rows_f32 with causal_block = H*W.  The score matrix is materialised, as
                                      the reference materialises the N x N mask; diffusers' default temporal tiling (below) keeps
                                      N at 5 latent frames: 72 000 tokens at 720 x 1280.
    _temporal_tiled_decode / blend_t  clips of more than 4 latent frames: overlapping tiles of 5 latent frames (stride 3), each
                                      decoded on its own, cross-faded over 4 frames with alg_axpby_f32 (diffusers' default
                                      use_framewise_decoding; 129 x 720 x 1280 = 11 tiles, 17.0 s)
"""
from __future__ import annotations

from typing import Dict

import torch

from . import _lib
from .encoders import _launch, linear_f32
from .vae_wan import SplitConvVAE, _Act

HUNYUAN_VAE = dict(in_channels=3, out_channels=3, latent_channels=16, block_out_channels=(128, 256, 512, 512), layers_per_block=2,
                   act_fn="silu", norm_num_groups=32, scaling_factor=0.476986, spatial_compression_ratio=8,
                   temporal_compression_ratio=4, mid_block_add_attention=True)


def _stages(cfg: dict):
    """Per encoder / decoder block: (c_in, c_out, spatial resample, temporal resample)."""
    boc = list(cfg["block_out_channels"])
    n = len(boc)
    log2 = {1: 0, 2: 1, 4: 2, 8: 3}
    ns, nt = log2[cfg["spatial_compression_ratio"]], log2[cfg["temporal_compression_ratio"]]
    if cfg["temporal_compression_ratio"] != 4:
        raise NotImplementedError("temporal_compression_ratio other than 4 (HunyuanVideo ships 4)")

    def plan(chs):
        out, cin = [], chs[0]
        for i, cout in enumerate(chs):
            out.append((cin, cout, i < ns, i >= (n - 1 - nt) and i != n - 1))
            cin = cout
        return out
    return plan(boc), plan(boc[::-1])


def parameter_shapes(cfg: dict) -> Dict[str, tuple]:
    """name -> shape of every parameter (diffusers naming)."""
    enc, dec = _stages(cfg)
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
        if cfg.get("mid_block_add_attention", True):
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


class AutoencoderKLHunyuanVideo(SplitConvVAE):
    """Native-kernel ``AutoencoderKLHunyuanVideo``: ``encode(x).latent_dist`` / ``decode(z).sample``."""

    Z_KEY = "latent_channels"
    TRUNCATE_FRAMES = False
    BUILD_TAPS = True

    def __init__(self, **config):
        cfg = dict(HUNYUAN_VAE)
        cfg.update({k: v for k, v in config.items() if k in cfg})
        cfg["block_out_channels"] = tuple(cfg["block_out_channels"])
        if any(c % cfg["norm_num_groups"] for c in cfg["block_out_channels"]):
            raise ValueError("block_out_channels must be divisible by norm_num_groups")
        self._init_common(cfg)
        self.temporal_compression_ratio = cfg["temporal_compression_ratio"]  # hy:277-278 read these off the module
        self.spatial_compression_ratio = cfg["spatial_compression_ratio"]
        # diffusers decodes long clips in overlapping TEMPORAL tiles by default (use_framewise_decoding = True: 16-frame tiles, stride
        # 12, linear cross-fade over the 4 shared frames) -- that is what hy:1292 runs for 129 frames, and it is what bounds the
        # mid-block attention to 5 latent frames.  Spatial tiling is opt-in there (enable_tiling) and run.py never opts in.
        self.use_framewise_decoding = True
        self.tile_sample_min_num_frames, self.tile_sample_stride_num_frames = 16, 12
        self._attn_budget = 64 << 30

    def enable_tiling(self, *a, **k):
        raise NotImplementedError("spatial VAE tiling is not built (run.py does not enable it); temporal tiling is always on, like diffusers' default")

    def _shapes(self) -> Dict[str, tuple]:
        return parameter_shapes(self._cfg)

    # ---- layers ---------------------------------------------------------------------------------------------------------
    def _gn(self, x: torch.Tensor, name: str, silu: bool) -> torch.Tensor:
        g = self._cfg["norm_num_groups"]
        if self._stats is None or self._stats.device != self.device or self._stats.numel() < 2 * g:
            self._stats = torch.empty(2 * g, device=self.device, dtype=torch.float64)
        out = torch.empty_like(x)
        _launch(_lib.lib().alg_group_norm_f32, self.device, x.data_ptr(), out.data_ptr(), x.shape[0], x.shape[1], g, 1e-6,
                self._w[name + ".weight"].data_ptr(), self._w[name + ".bias"].data_ptr(), int(silu), self._stats.data_ptr())
        return out

    def _cconv(self, x: _Act, name: str, **kw) -> _Act:
        """HunyuanVideoCausalConv3d.  Stride-1 convolutions at the input resolution take the implicit path: the input split into the
        padded raster (alg_norm_split_pad_f32 without a norm), its padding filled by replication (alg_replicate_border_bf16), the
        27 taps as row-shifted reads in ONE GEMM with bias / residual in the epilogue, the result copied back out of the padded
        raster.  Strided and upsampling convolutions gather patches (the resize and the padding are addressing there)."""
        plain = not kw.get("stride") and not kw.get("up") and kw.get("tdup", 1) == 1 and not kw.get("frames") and not kw.get("out_hw")
        # (a single frame -- the pipeline's encode -- is faster through the patch gather: 84 vs 103 ms at 720 x 1280)
        if self.implicit and plain and x.T > 1 and name + ".conv.weight_taps" in self._w and self._fits_implicit_hy(x, self._w[name + ".conv.bias"].numel()):
            x = self._to_compact(x)
            s3p = self._split_pad(x, None, False)
            _launch(_lib.lib().alg_replicate_border_bf16, self.device, s3p.data_ptr(), x.T, x.H, x.W, 2, s3p.shape[1])
            res = kw.get("residual")
            if res is not None:  # the residual lives in the compact layout: add it after un-padding (one fused pass)
                y = self._to_compact(self._conv_taps(s3p, (x.T, x.H, x.W), name + ".conv"))
                _launch(_lib.lib().alg_axpby_f32, self.device, y.t.data_ptr(), res.data_ptr(), y.t.data_ptr(), y.t.numel(), 1.0, 1.0)
                return y
            return self._to_compact(self._conv_taps(s3p, (x.T, x.H, x.W), name + ".conv"))
        return self._conv(x, name + ".conv", (3, 3, 3), replicate=True, **kw)

    def _fits_implicit_hy(self, x: _Act, c_out: int) -> bool:
        rows = (x.T + 2) * (x.H + 2) * (x.W + 2)
        cs = (3 * x.C + 63) // 64 * 64
        need = rows * (2 * cs + 4 * c_out) + x.T * x.H * x.W * 4 * c_out
        free, _ = torch.cuda.mem_get_info(self.device)
        free += torch.cuda.memory_reserved(self.device) - torch.cuda.memory_allocated(self.device)
        return need < 0.8 * free

    def _res(self, x: _Act, name: str) -> _Act:
        h = x.t
        if name + ".conv_shortcut.conv.weight" in self._w:
            h = self._pointwise(x.t, name + ".conv_shortcut.conv")
        y = _Act(self._gn(x.t, name + ".norm1", True), x.T, x.H, x.W, x.C)
        y = self._cconv(y, name + ".conv1")
        y.t = self._gn(y.t, name + ".norm2", True)
        return self._cconv(y, name + ".conv2", residual=h)

    def _mid(self, x: _Act, name: str) -> _Act:
        x = self._res(x, name + ".resnets.0")
        if self._cfg.get("mid_block_add_attention", True):
            a, w = name + ".attentions.0", self._w
            y = self._gn(x.t, a + ".group_norm", False)
            o = self._attention_core(y, w[a + ".to_q.weight"], w[a + ".to_q.bias"], w[a + ".to_k.weight"], w[a + ".to_k.bias"],
                                     w[a + ".to_v.weight"], w[a + ".to_v.bias"], causal_block=x.H * x.W if x.T > 1 else 0)
            x = _Act(linear_f32(o, w[a + ".to_out.0.weight"], w[a + ".to_out.0.bias"], residual=x.t), x.T, x.H, x.W, x.C)
        return self._res(x, name + ".resnets.1")

    def _down(self, x: _Act, name: str, sp: bool, tp: bool) -> _Act:
        st, ss = (2 if tp else 1), (2 if sp else 1)
        To, Ho, Wo = (x.T - 1) // st + 1, (x.H - 1) // ss + 1, (x.W - 1) // ss + 1
        return self._cconv(x, name, stride=(st, ss, ss), frames=(0, To), out_hw=(Ho, Wo))

    def _up(self, x: _Act, name: str, sp: bool, tp: bool) -> _Act:
        up, tdup = (2 if sp else 1), (2 if tp and x.T > 1 else 1)
        return self._cconv(x, name, up=up, tdup=tdup)

    # ---- one sample -------------------------------------------------------------------------------------------------------
    def _encode_one(self, x: torch.Tensor) -> torch.Tensor:
        enc, _ = _stages(self._cfg)
        a = self._cconv(self._to_cl(x), "encoder.conv_in")
        for i, (ci, co, sp, tp) in enumerate(enc):
            for j in range(self._cfg["layers_per_block"]):
                a = self._res(a, f"encoder.down_blocks.{i}.resnets.{j}")
            if sp or tp:
                a = self._down(a, f"encoder.down_blocks.{i}.downsamplers.0.conv", sp, tp)
        a = self._mid(a, "encoder.mid_block")
        a.t = self._gn(a.t, "encoder.conv_norm_out", True)
        a = self._cconv(a, "encoder.conv_out")
        a.t = self._pointwise(a.t, "quant_conv")
        out = self._from_cl(a, 2 * self._cfg["latent_channels"], clamp=False)
        self._release_operands()
        return out

    def _decode_one(self, z: torch.Tensor) -> torch.Tensor:
        """``AutoencoderKLHunyuanVideo._decode``: clips longer than one temporal tile go through ``_temporal_tiled_decode``."""
        r = self._cfg["temporal_compression_ratio"]
        t_min, t_stride = self.tile_sample_min_num_frames // r, self.tile_sample_stride_num_frames // r
        T = z.shape[1]
        if not (self.use_framewise_decoding and T > t_min):
            return self._decode_tile(z)
        lib, dev = _lib.lib(), self.device
        blend = self.tile_sample_min_num_frames - self.tile_sample_stride_num_frames
        stride_f = self.tile_sample_stride_num_frames
        tiles = []
        for i in range(0, T, t_stride):  # tiles of t_min + 1 latent frames; every tile but the first drops its first decoded frame
            d = self._decode_tile(z[:, i:i + t_min + 1].contiguous())
            tiles.append(d if i == 0 else d[:, 1:].contiguous())
        out = []
        for i, tile in enumerate(tiles):
            if i > 0:  # blend_t: the first frames of this tile cross-fade from the last frames of the previous one
                prev = tiles[i - 1]
                n = min(prev.shape[1], tile.shape[1], blend)
                for x in range(n):
                    a_, b_ = prev[:, prev.shape[1] - n + x].contiguous(), tile[:, x].contiguous()
                    mixed = torch.empty_like(b_)
                    _launch(lib.alg_axpby_f32, dev, a_.data_ptr(), b_.data_ptr(), mixed.data_ptr(), b_.numel(), 1.0 - x / n, x / n)
                    tile[:, x] = mixed
                out.append(tile[:, :stride_f])
            else:
                out.append(tile[:, :stride_f + 1])
        return torch.cat(out, dim=1)[:, :(T - 1) * r + 1].contiguous()

    def _decode_tile(self, z: torch.Tensor) -> torch.Tensor:
        _, dec = _stages(self._cfg)
        a = self._to_cl(z)
        a.t = self._pointwise(a.t, "post_quant_conv")[:, :self._cfg["latent_channels"]].contiguous()
        a = self._cconv(a, "decoder.conv_in")
        a = self._mid(a, "decoder.mid_block")
        for i, (ci, co, sp, tp) in enumerate(dec):
            for j in range(self._cfg["layers_per_block"] + 1):
                a = self._res(a, f"decoder.up_blocks.{i}.resnets.{j}")
            if sp or tp:
                a = self._up(a, f"decoder.up_blocks.{i}.upsamplers.0.conv", sp, tp)
        a.t = self._gn(a.t, "decoder.conv_norm_out", True)
        a = self._cconv(a, "decoder.conv_out")
        out = self._from_cl(a, self._cfg["out_channels"], clamp=False)
        self._release_operands()
        return out
