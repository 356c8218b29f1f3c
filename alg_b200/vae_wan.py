"""``AutoencoderKLWan`` (Wan2.1 VAE) on the sm_100a kernels of ``libalg_b200.so``: ``encode`` and ``decode`` in float32.

Reference call sites: wan:429-434 (``retrieve_latents(self.vae.encode(video_condition), "argmax")`` on image + zero frames),
wan:526 (pixel-space ALG: ``encode(...).latent_dist.sample(generator)`` on EVERY step), wan:959 (``decode``: the frames the
metric counts); ``run.py:51-55`` loads this VAE in float32, and wan:180-181 reads ``vae.temperal_downsample`` off the object.
The network is diffusers@be2fb77 ``autoencoder_kl_wan.py`` (absent offline: parity unpinned; ``oracle/wan_vae_oracle.py``
restates it WITH diffusers' chunked feature-cache evaluation, and ``tests/test_gpu_vae_wan.py`` compares).

diffusers walks the clip in chunks (encoder: 1 frame then 4 at a time, decoder: one latent frame at a time) and carries the
last frames of every 3x3x3 convolution's input in a cache.  Worked through, that is a WHOLE-CLIP network with
  * every ``WanCausalConv3d`` = a convolution with two ZERO frames in front (kh // 2, kw // 2 zeros around),
  * ``WanResample`` "downsample3d": frame 0 passes; output frame k >= 1 = time_conv over frames (2k-2, 2k-1, 2k),
  * ``WanResample`` "upsample3d": frame 0 passes; frames t >= 1 each give two frames from time_conv over (t-2, t-1, t) where
    frame 0 and everything before it read as ZERO (the "Rep" cache marker of the first chunk),
and with 180 GB of HBM the whole clip of every level stays resident (81 x 480 x 832 x 96 fp32 = 12.4 GB), so there are no caches
here: a convolution is a frame-chunked patch gather + ONE GEMM per chunk.  This module only SEQUENCES C-ABI calls:

    WanCausalConv3d / Conv2d(3x3)     alg_im2col_split3_f32 (gather + bf16 [hi | hi | lo] split in one pass; stride 2 and the
                                      nearest-exact x2 upsample are addressing) + alg_gemm_bf16 vs the weight's [hi | lo | hi]
                                      split, fp32 accumulate / output: the fp32 product to 2^-16 + alg_bias_act_f32 (bias, residual)
    1x1x1 convolutions, to_qkv, proj  alg_split3_bf16 + alg_gemm_bf16 (+ alg_bias_act_f32)
    WanRMS_norm (+ SiLU)              alg_rms_norm_cl_f32
    WanAttentionBlock                 per frame, one head of C channels: scores and P V through the split GEMM (V produced
                                      transposed by swapping the GEMM operands), alg_softmax_rows_f32
    [B, C, T, H, W] <-> channels-last alg_nchw_to_cl_f32 / alg_cl_to_nchw_f32 (clamp to [-1, 1] on the way out)
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import Dict, List, Optional, Tuple

import torch

from . import _lib, ops
from .encoders import _launch, linear_f32, split_act, split_weight
from .vae_cogvideox import DiagonalGaussianDistribution

import ctypes as C

WAN21_VAE = dict(base_dim=96, z_dim=16, dim_mult=[1, 2, 4, 4], num_res_blocks=2, attn_scales=[], temperal_downsample=[False, True, True],
                 dropout=0.0,
                 latents_mean=[-0.7571, -0.7089, -0.9113, 0.1075, -0.1745, 0.9653, -0.1517, 1.5508, 0.4134, -0.0715, 0.5517, -0.3632,
                               -0.1922, -0.9497, 0.2503, -0.2921],
                 latents_std=[2.8184, 1.4541, 2.3275, 2.6558, 1.2196, 1.7708, 2.6052, 2.0743, 3.2687, 2.1526, 2.8652, 1.5579, 1.6382,
                              1.1253, 2.8251, 1.9160])


def _plan(cfg: dict):
    """Module order of ``WanEncoder3d.down_blocks`` / ``WanDecoder3d.up_blocks``: (diffusers name, kind, c_in, c_out)."""
    dim, mult, nres = cfg["base_dim"], list(cfg["dim_mult"]), cfg["num_res_blocks"]
    tdown, attn_scales = list(cfg["temperal_downsample"]), list(cfg.get("attn_scales") or [])
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
    tup, dec = tdown[::-1], []
    for i, (cin, cout) in enumerate(zip(ddims[:-1], ddims[1:])):
        cin = cin // 2 if i > 0 else cin
        for j in range(nres + 1):
            dec.append((f"decoder.up_blocks.{i}.resnets.{j}", "res", cin, cout))
            cin = cout
        if i != len(mult) - 1:
            dec.append((f"decoder.up_blocks.{i}.upsamplers.0", "up3d" if tup[i] else "up2d", cout, cout // 2))
    return enc, dec, dims[-1], ddims[0], ddims[-1]


def parameter_shapes(cfg: dict) -> Dict[str, tuple]:
    """name -> shape of every parameter (diffusers naming)."""
    enc, dec, c_enc, c_dec_in, c_dec_out = _plan(cfg)
    z, s = cfg["z_dim"], {}

    def conv(name, co, ci, k):
        s[name + ".weight"], s[name + ".bias"] = (co, ci, *k), (co,)

    def res(name, ci, co):
        s[name + ".norm1.gamma"], s[name + ".norm2.gamma"] = (ci, 1, 1, 1), (co, 1, 1, 1)
        conv(name + ".conv1", co, ci, (3, 3, 3))
        conv(name + ".conv2", co, co, (3, 3, 3))
        if ci != co:
            conv(name + ".conv_shortcut", co, ci, (1, 1, 1))

    def attn(name, c):
        s[name + ".norm.gamma"] = (c, 1, 1)
        conv(name + ".to_qkv", 3 * c, c, (1, 1))
        conv(name + ".proj", c, c, (1, 1))

    def mid(name, c):
        res(name + ".resnets.0", c, c)
        attn(name + ".attentions.0", c)
        res(name + ".resnets.1", c, c)

    conv("encoder.conv_in", cfg["base_dim"], 3, (3, 3, 3))
    for name, kind, ci, co in enc:
        if kind == "res":
            res(name, ci, co)
        elif kind == "attn":
            attn(name, ci)
        else:
            conv(name + ".resample.1", co, ci, (3, 3))
            if kind == "down3d":
                conv(name + ".time_conv", co, ci, (3, 1, 1))
    mid("encoder.mid_block", c_enc)
    s["encoder.norm_out.gamma"] = (c_enc, 1, 1, 1)
    conv("encoder.conv_out", 2 * z, c_enc, (3, 3, 3))
    conv("quant_conv", 2 * z, 2 * z, (1, 1, 1))
    conv("post_quant_conv", z, z, (1, 1, 1))
    conv("decoder.conv_in", c_dec_in, z, (3, 3, 3))
    mid("decoder.mid_block", c_dec_in)
    for name, kind, ci, co in dec:
        if kind == "res":
            res(name, ci, co)
        else:
            conv(name + ".resample.1", co, ci, (3, 3))
            if kind == "up3d":
                conv(name + ".time_conv", 2 * ci, ci, (3, 1, 1))
    s["decoder.norm_out.gamma"] = (c_dec_out, 1, 1, 1)
    conv("decoder.conv_out", 3, c_dec_out, (3, 3, 3))
    return s


FRONT_PAD = 2  # zero frames in front of a clip in the padded raster (kt - 1 of the causal 3x3x3 convolutions)


class _Act:
    """One sample's activation, channels-last fp32: ``t`` is [T*H*W, C] (compact), or, with ``padded``, the zero-padded raster
    [(T + FRONT_PAD) * (H + 2) * (W + 2), C] in which the implicit convolutions produce their output (pixel (t, y, x) sits in row
    ((t + FRONT_PAD) * (H + 2) + y + 1) * (W + 2) + x + 1; the padding rows hold don't-care values)."""
    __slots__ = ("t", "T", "H", "W", "C", "padded")

    def __init__(self, t, T, H, W, C_, padded=False):
        self.t, self.T, self.H, self.W, self.C, self.padded = t, T, H, W, C_, padded

    def frame(self, i, n=1):
        hw = self.H * self.W
        return self.t[i * hw:(i + n) * hw]


class SplitConvVAE:
    """What the float32 video VAEs share: weights as split GEMM operands, the frame-chunked (gather + split, GEMM, bias) convolution,
    pointwise linears, layout changes, the diffusers-facing ``encode`` / ``decode`` wrappers.  Subclasses provide
    ``_shapes()``, ``_encode_one`` and ``_decode_one``."""

    Z_KEY = "z_dim"
    BUILD_TAPS = False      # also lay the 3x3x3 weights out tap-major for the implicit (zero-padded) convolution
    TRUNCATE_FRAMES = True  # AutoencoderKLWan's chunk loop (1 frame, then 4 at a time) drops a trailing partial chunk

    def _init_common(self, cfg: dict):
        self._cfg = cfg
        self.config = SimpleNamespace(**cfg)
        self.dtype = torch.float32
        self.device = torch.device("cpu")
        self._w: Dict[str, torch.Tensor] = {}
        self._sd: Dict[str, torch.Tensor] = {}
        self._cols: Optional[torch.Tensor] = None
        self._cols_budget = 6 << 30  # bytes of patch matrix per gather + GEMM call (whole output frames)
        self._stats: Optional[torch.Tensor] = None
        self._attn_budget = 48 << 30  # bytes the materialised attention scores may take
        self._s3p: Dict[tuple, torch.Tensor] = {}  # zero-padded split operands of the implicit convolutions, by geometry
        import os
        self.implicit = os.environ.get("ALG_VAE_IMPLICIT", "1") == "1"

    def _shapes(self) -> Dict[str, tuple]:
        raise NotImplementedError

    # ---- construction -------------------------------------------------------------------------------------------------
    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path, subfolder: str = "vae", torch_dtype=None, cache_dir=None, device="cuda"):
        from . import checkpoint
        root = checkpoint.resolve_snapshot(str(pretrained_model_name_or_path), cache_dir, require=subfolder or "transformer")
        if root is None:
            raise FileNotFoundError(f"no local snapshot for {pretrained_model_name_or_path!r} (there is no network)")
        cfg, sd = checkpoint.load_component(root, subfolder, device=device)
        return cls(**cfg).load_state_dict(sd)

    @classmethod
    def from_synthetic(cls, seed: int = 0, device="cuda", **config):
        """Seeded weights at the model's true shape unless ``config`` says otherwise (fan-in scaled: activations stay O(1))."""
        m = cls(**config)
        sd = {}
        for idx, (name, shape) in enumerate(m._shapes().items()):
            g = torch.Generator(device=device).manual_seed(seed * 1_000_003 + 11_001 + idx)
            if name.endswith(".gamma") or ("norm" in name and name.endswith(".weight")):
                w = 1 + 0.1 * torch.randn(shape, generator=g, device=device)
            elif name.endswith(".bias"):
                w = 0.05 * torch.randn(shape, generator=g, device=device)
            else:
                fan_in = 1
                for d in shape[1:]:
                    fan_in *= d
                w = torch.randn(shape, generator=g, device=device) * (1.2 * fan_in ** -0.5)
            sd[name] = w
        return m.load_state_dict(sd)

    def load_state_dict(self, sd: Dict[str, torch.Tensor]):
        """diffusers names.  Convolution weights are re-laid once as split GEMM B operands [Co8, 3 * K8] with
        K = (kt, kh, kw, Ci) flattened (Ci innermost, like the patch rows) and Co / K zero-padded to multiples of 8."""
        shapes = self._shapes()
        missing = [k for k in shapes if k not in sd]
        if missing:
            raise KeyError(f"missing VAE weights: {missing[:4]}{'...' if len(missing) > 4 else ''}")
        dev = sd[next(iter(shapes))].device
        if dev.type != "cuda":
            raise RuntimeError("VAE weights must live on a CUDA device (no CPU fallback)")
        self.device = dev
        w: Dict[str, torch.Tensor] = {}
        self._sd = {}
        for name, shape in shapes.items():
            t = sd[name]
            if tuple(t.shape) != tuple(shape):
                raise ValueError(f"{name}: expected {tuple(shape)}, got {tuple(t.shape)}")
            t = t.to(device=dev, dtype=torch.float32)
            self._sd[name] = t
            if name.endswith(".gamma"):
                w[name] = t.reshape(-1).contiguous()
            elif t.dim() == 1:  # biases (padded like their weight's output channels) and GroupNorm affine parameters
                co8 = (t.numel() + 7) // 8 * 8
                w[name] = torch.nn.functional.pad(t, (0, co8 - t.numel())).contiguous()
            else:
                co = t.shape[0]
                m = t.movedim(1, -1).reshape(co, -1)  # [Co, (kt,) kh, kw, Ci] flattened
                co8, k8 = (co + 7) // 8 * 8, (m.shape[1] + 7) // 8 * 8
                if co8 != co:
                    m = torch.cat([m, m.new_zeros(co8 - co, m.shape[1])])
                w[name] = split_weight(m.contiguous(), k8)
                if self.BUILD_TAPS and t.dim() == 5 and tuple(t.shape[2:]) == (3, 3, 3) and t.shape[1] % 2 == 0 and t.shape[1] <= 512:
                    # implicit convolution: per tap [w_hi | w_lo | w_hi | 0] padded to a multiple of 64 columns, taps along K
                    ci = t.shape[1]
                    cs = (3 * ci + 63) // 64 * 64
                    taps = t.permute(0, 2, 3, 4, 1).reshape(co, 27, ci)
                    if co8 != co:
                        taps = torch.cat([taps, taps.new_zeros(co8 - co, 27, ci)])
                    w3 = split_weight(taps.reshape(co8 * 27, ci).contiguous()).view(co8, 27, 3 * ci)
                    w[name + "_taps"] = torch.nn.functional.pad(w3, (0, cs - 3 * ci)).reshape(co8, 27 * cs).contiguous()
        self._w = w
        self._cols = None
        return self

    def state_dict(self) -> Dict[str, torch.Tensor]:
        return dict(self._sd)

    def to(self, device=None, dtype=None):
        if device is not None and self._sd and torch.device(device).type == "cuda":
            dev = torch.device(device)
            dev = torch.device("cuda", torch.cuda.current_device()) if dev.index is None else dev
            if dev != self.device:
                self.load_state_dict({k: v.to(dev) for k, v in self._sd.items()})
        return self

    # ---- layers ---------------------------------------------------------------------------------------------------------
    def _patch_buffer(self, nbytes: int) -> torch.Tensor:
        if self._cols is None or self._cols.numel() * 2 < nbytes or self._cols.device != self.device:
            self._cols = None
            self._cols = torch.empty((nbytes + 1) // 2, device=self.device, dtype=torch.bfloat16)
        return self._cols

    def _conv(self, x: _Act, name: str, k: Tuple[int, int, int], *, stride=(1, 1, 1), pad=(None, None, None), up: int = 1,
              t_min: int = 0, frames: Optional[Tuple[int, int]] = None, out_hw: Optional[Tuple[int, int]] = None,
              residual: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None, w_rows: Optional[Tuple[int, int]] = None,
              out_frame0: int = 0, out_frame_step: int = 1, replicate: bool = False, tdup: int = 1) -> _Act:
        """Convolution of the whole clip as frame-chunked (gather + split, GEMM, bias [+ residual]).

        ``pad`` = (front frames, top, left), default (kt - 1, kh // 2, kw // 2); ``frames`` = (first output frame, count) of the
        temporal output range, default every frame; ``w_rows`` restricts the output channels to a row range of the weight and
        ``out`` / ``out_frame0`` / ``out_frame_step`` scatter output frame i to frame ``out_frame0 + i * out_frame_step`` of an
        existing clip (the two channel halves of "upsample3d" become alternating frames)."""
        lib, dev = _lib.lib(), self.device
        x = self._to_compact(x)
        kt, kh, kw = k
        st, sh, sw = stride
        pad_t = kt - 1 if pad[0] is None else pad[0]
        pad_top = kh // 2 if pad[1] is None else pad[1]
        pad_left = kw // 2 if pad[2] is None else pad[2]
        Ho, Wo = out_hw if out_hw is not None else (x.H * up, x.W * up)
        to0, To = frames if frames is not None else (0, x.T if tdup == 1 else 2 * x.T - 1)
        w3, bias = self._w[name + ".weight"], self._w[name + ".bias"]
        if w_rows is not None:
            w3, bias = w3[w_rows[0]:w_rows[1]], bias[w_rows[0]:w_rows[1]]
        N = w3.shape[0]
        ld = w3.shape[1] // 3
        assert ld >= kt * kh * kw * x.C and N % 8 == 0
        hw = Ho * Wo
        if out is None:
            out = torch.empty(To * hw, N, device=dev, dtype=torch.float32)
            out_frame0, out_frame_step = 0, 1
            res = _Act(out, To, Ho, Wo, N)
        else:
            res = None
        per_frame = hw * 3 * ld * 2
        chunk = max(1, min(To, self._cols_budget // per_frame)) if out_frame_step == 1 else 1
        cols = self._patch_buffer(min(To, chunk) * per_frame)
        p = _lib.Im2colF32()
        p.x, p.cols = x.t.data_ptr(), cols.data_ptr()
        p.T, p.H, p.W, p.C = x.T, x.H, x.W, x.C
        p.kt, p.kh, p.kw, p.st, p.sh, p.sw = kt, kh, kw, st, sh, sw
        p.pad_t, p.pad_top, p.pad_left = pad_t, pad_top, pad_left
        p.Ho, p.Wo, p.up, p.t_min, p.ld = Ho, Wo, up, t_min, ld
        p.replicate, p.tdup = int(replicate), tdup
        for f0 in range(0, To, chunk):
            n = min(chunk, To - f0)
            p.To, p.to0 = n, to0 + f0
            _launch(lib.alg_im2col_split3_f32, dev, C.byref(p))
            a = cols[:n * hw * 3 * ld].view(n * hw, 3 * ld)
            r0 = (out_frame0 + f0 * out_frame_step) * hw
            y = out[r0:r0 + n * hw]
            rs = None if residual is None else residual[r0:r0 + n * hw]
            ops.gemm(a, w3, None, out=y, out_dtype=torch.float32, bias_f32=bias, residual_f32=rs)
        return res

    def _pointwise(self, x: torch.Tensor, name: str, residual: Optional[torch.Tensor] = None) -> torch.Tensor:
        """1x1(x1) convolution = nn.Linear over the channels of [rows, C]."""
        w3 = self._w[name + ".weight"]
        if w3.shape[1] // 3 != x.shape[1]:  # K was padded to a multiple of 8
            x = torch.nn.functional.pad(x, (0, w3.shape[1] // 3 - x.shape[1]))
        return linear_f32(x.contiguous(), w3, self._w[name + ".bias"], residual=residual)

    def _norm(self, x: torch.Tensor, name: str, silu: bool) -> torch.Tensor:
        out = torch.empty_like(x)
        _launch(_lib.lib().alg_rms_norm_cl_f32, self.device, x.data_ptr(), out.data_ptr(), x.shape[0], x.shape[1],
                self._w[name + ".gamma"].data_ptr(), None, int(silu))
        return out

    # ---- implicit (patch-matrix-free) 3x3x3 stride-1 convolution -----------------------------------------------------------------
    def _to_padded(self, x: _Act) -> _Act:
        if x.padded:
            return x
        out = torch.zeros((x.T + FRONT_PAD) * (x.H + 2) * (x.W + 2), x.C, device=self.device, dtype=torch.float32)
        _launch(_lib.lib().alg_pad_copy_f32, self.device, x.t.data_ptr(), out.data_ptr(), x.T, x.H, x.W, x.C, FRONT_PAD, 1)
        return _Act(out, x.T, x.H, x.W, x.C, padded=True)

    def _to_compact(self, x: _Act) -> _Act:
        if not x.padded:
            return x
        out = torch.empty(x.T * x.H * x.W, x.C, device=self.device, dtype=torch.float32)
        _launch(_lib.lib().alg_pad_copy_f32, self.device, x.t.data_ptr(), out.data_ptr(), x.T, x.H, x.W, x.C, FRONT_PAD, 0)
        return _Act(out, x.T, x.H, x.W, x.C)

    def _split_pad(self, x: _Act, gamma: Optional[torch.Tensor], silu: bool) -> torch.Tensor:
        """(WanRMS_norm) (+ SiLU) + bf16 split of ``x`` into the zero-padded operand buffer of its geometry (reused by every
        convolution of that geometry: only interior pixels are rewritten, the padding stays zero)."""
        cs = (3 * x.C + 63) // 64 * 64
        key = (x.T, x.H, x.W, cs)
        buf = self._s3p.get(key)
        if buf is None or buf.device != self.device:
            buf = self._s3p[key] = torch.zeros((x.T + FRONT_PAD) * (x.H + 2) * (x.W + 2), cs, device=self.device, dtype=torch.bfloat16)
        _launch(_lib.lib().alg_norm_split_pad_f32, self.device, x.t.data_ptr(), buf.data_ptr(), x.T, x.H, x.W, x.C, FRONT_PAD,
                int(x.padded), cs, None if gamma is None else gamma.data_ptr(), int(silu))
        return buf

    def _conv_taps(self, s3p: torch.Tensor, geom: Tuple[int, int, int], name: str, residual: Optional[torch.Tensor] = None) -> _Act:
        """The 27 taps as row-shifted reads of ONE padded operand (alg_gemm_bf16 tap mode): out = padded raster [P_pad, Co8]."""
        T, H, W = geom
        w3, bias = self._w[name + ".weight_taps"], self._w[name + ".bias"]
        cs = s3p.shape[1]
        plane, row = (H + 2) * (W + 2), W + 2
        offs = [(it - FRONT_PAD) * plane + (ih - 1) * row + (iw - 1) for it in range(3) for ih in range(3) for iw in range(3)]
        out = torch.empty(s3p.shape[0], w3.shape[0], device=self.device, dtype=torch.float32)
        ops.gemm(s3p, w3, None, out=out, out_dtype=torch.float32, a_tap_kblocks=cs // 64, a_tap_offsets=offs, bias_f32=bias,
                 residual_f32=residual)  # bias and the block's residual ride in the GEMM epilogue
        return _Act(out, T, H, W, out.shape[1], padded=True)

    def _release_operands(self):
        self._s3p.clear()
        self._cols = None

    def _attention_core(self, y: torch.Tensor, wq3, bq, wk3, bk, wv3, bv, causal_block: int = 0) -> torch.Tensor:
        """softmax(q k^T / sqrt(C)) v for ONE head of C channels over the rows of y [N, C] (already normalised), fp32 through the
        split GEMM: scores [N, N] are materialised (like the N x N mask of the reference), V is produced transposed by swapping the
        GEMM operands, and the value bias is added after P V (softmax rows sum to one).  ``causal_block`` = tokens per frame for the
        frame-causal mask of the HunyuanVideo VAE."""
        lib, dev = _lib.lib(), self.device
        N, Cc = y.shape
        n8 = (N + 7) // 8 * 8
        if N * n8 * 10 > self._attn_budget:
            raise NotImplementedError(
                f"VAE mid-block attention over {N} tokens needs a {N} x {N} score matrix ({N * n8 * 10 / 2 ** 30:.0f} GiB with its "
                "split copy); the reference materialises the same N x N mask -- decode a shorter / smaller clip")
        q = linear_f32(y, wq3, bq)
        k = linear_f32(y, wk3, bk)
        # W_v as the swapped GEMM's A operand: its stored [hi | lo | hi] sections reordered to the activation order [hi | hi | lo]
        wv_a = wv3.reshape(Cc, 3, -1)[:, [0, 2, 1]].reshape(Cc, -1).contiguous()
        vt = torch.zeros(Cc, n8, device=dev, dtype=torch.float32)  # columns >= N stay zero (K padding of the P V product)
        ops.gemm(wv_a, split_weight(y), None, out=vt[:, :N], out_dtype=torch.float32)
        s = torch.zeros(N, n8, device=dev, dtype=torch.float32)
        ops.gemm(split_act(q), split_weight(k), None, out=s[:, :N], out_dtype=torch.float32)
        _launch(lib.alg_softmax_rows_f32, dev, s.data_ptr(), N, N, n8, float(Cc) ** -0.5, int(causal_block))
        return linear_f32(s, split_weight(vt), bv)

    # ---- public surface ---------------------------------------------------------------------------------------------------
    def _to_cl(self, x: torch.Tensor) -> _Act:
        Cc, T, H, W = x.shape
        out = torch.empty(T * H * W, Cc, device=self.device, dtype=torch.float32)
        _launch(_lib.lib().alg_nchw_to_cl_f32, self.device, x.data_ptr(), out.data_ptr(), Cc, T * H * W, Cc)
        return _Act(out, T, H, W, Cc)

    def _from_cl(self, a: _Act, Cc: int, clamp: bool) -> torch.Tensor:
        out = torch.empty(Cc, a.T, a.H, a.W, device=self.device, dtype=torch.float32)
        lo, hi = (-1.0, 1.0) if clamp else (0.0, 0.0)
        _launch(_lib.lib().alg_cl_to_nchw_f32, self.device, a.t.data_ptr(), out.data_ptr(), Cc, a.T * a.H * a.W, a.t.shape[1], lo, hi)
        return out

    def _check(self, x: torch.Tensor, channels: int, what: str) -> torch.Tensor:
        if not self._w:
            raise RuntimeError(f"{type(self).__name__} has no weights loaded")
        if x.dim() != 5 or x.shape[1] != channels:
            raise ValueError(f"{what}: expected [B, {channels}, T, H, W], got {tuple(x.shape)}")
        if not x.is_cuda:
            raise RuntimeError(f"{type(self).__name__} runs on CUDA tensors only (no CPU fallback)")
        return x.to(device=self.device, dtype=torch.float32).contiguous()

    @torch.no_grad()
    def encode(self, x: torch.Tensor, return_dict: bool = True):
        """x [B, 3, 1 + 4n, H, W] -> ``.latent_dist`` over [B, 2 z, 1 + n, H/8, W/8] (frames beyond 1 + 4n are dropped like
        diffusers' chunk loop drops them)."""
        x = self._check(x, 3, "encode")
        T = 1 + (x.shape[2] - 1) // 4 * 4 if self.TRUNCATE_FRAMES else x.shape[2]
        moments = torch.stack([self._encode_one(x[b, :, :T].contiguous()) for b in range(x.shape[0])])
        dist = DiagonalGaussianDistribution(moments)
        return SimpleNamespace(latent_dist=dist) if return_dict else (dist,)

    @torch.no_grad()
    def decode(self, z: torch.Tensor, return_dict: bool = True):
        """z [B, z, T, h, w] -> ``.sample`` [B, 3, 4 T - 3, 8 h, 8 w]."""
        z = self._check(z, self._cfg[self.Z_KEY], "decode")
        video = torch.stack([self._decode_one(z[b].contiguous()) for b in range(z.shape[0])])
        return SimpleNamespace(sample=video) if return_dict else (video,)


class AutoencoderKLWan(SplitConvVAE):
    """Native-kernel ``AutoencoderKLWan``: ``encode(x).latent_dist`` / ``decode(z).sample`` in float32."""

    BUILD_TAPS = True

    def __init__(self, **config):
        cfg = dict(WAN21_VAE)
        cfg.update({k: v for k, v in config.items() if k in cfg})
        if cfg.get("attn_scales"):
            raise NotImplementedError("attn_scales other than [] (Wan2.1's VAE has attention only in the mid blocks)")
        self._init_common(cfg)
        self.temperal_downsample = list(cfg["temperal_downsample"])  # (sic) wan:180-181 reads it off the module

    def _shapes(self) -> Dict[str, tuple]:
        return parameter_shapes(self._cfg)

    def _fits_implicit(self, x: _Act, c_out: int) -> bool:
        """The implicit path keeps the padded input, two padded outputs and the split operands of a whole level resident; when that
        does not fit next to what is already allocated (720p clips beside a resident DiT), fall back to the frame-chunked gather."""
        rows = (x.T + FRONT_PAD) * (x.H + 2) * (x.W + 2)
        cs_in, cs_out = (3 * x.C + 63) // 64 * 64, (3 * c_out + 63) // 64 * 64
        need = rows * (2 * (cs_in + cs_out) + 4 * (x.C + 3 * c_out))
        free, _ = torch.cuda.mem_get_info(self.device)
        free += torch.cuda.memory_reserved(self.device) - torch.cuda.memory_allocated(self.device)
        return need < 0.8 * free

    def _res(self, x: _Act, name: str) -> _Act:
        if self.implicit and self._fits_implicit(x, self._w[name + ".conv1.bias"].numel()):
            # norm + SiLU + split straight into the padded operand; both convolutions without a patch matrix
            x = self._to_padded(x)
            geom = (x.T, x.H, x.W)
            h = x.t
            if name + ".conv_shortcut.weight" in self._w:
                h = self._pointwise(x.t, name + ".conv_shortcut")
            y = self._conv_taps(self._split_pad(x, self._w[name + ".norm1.gamma"], True), geom, name + ".conv1")
            return self._conv_taps(self._split_pad(y, self._w[name + ".norm2.gamma"], True), geom, name + ".conv2", residual=h)
        h = x.t
        if name + ".conv_shortcut.weight" in self._w:
            h = self._pointwise(x.t, name + ".conv_shortcut")
        y = _Act(self._norm(x.t, name + ".norm1", True), x.T, x.H, x.W, x.C)
        y = self._conv(y, name + ".conv1", (3, 3, 3))
        y.t = self._norm(y.t, name + ".norm2", True)
        return self._conv(y, name + ".conv2", (3, 3, 3), residual=h)

    def _attn(self, x: _Act, name: str) -> _Act:
        """WanAttentionBlock: per frame, ONE head of C channels over the H*W pixels."""
        x = self._to_compact(x)
        Cc, hw = x.C, x.H * x.W
        wq = self._w[name + ".to_qkv.weight"]  # [3C, 3C] split, rows = (q | k | v) output channels
        bq = self._w[name + ".to_qkv.bias"]
        out = torch.empty_like(x.t)
        for f in range(x.T):
            xf = x.frame(f)
            y = self._norm(xf, name + ".norm", False)
            o = self._attention_core(y, wq[:Cc], bq[:Cc], wq[Cc:2 * Cc], bq[Cc:2 * Cc], wq[2 * Cc:3 * Cc], bq[2 * Cc:3 * Cc])
            out[f * hw:(f + 1) * hw] = self._pointwise(o, name + ".proj", residual=xf)
        return _Act(out, x.T, x.H, x.W, Cc)

    def _mid(self, x: _Act, name: str) -> _Act:
        x = self._res(x, name + ".resnets.0")
        x = self._attn(x, name + ".attentions.0")
        return self._res(x, name + ".resnets.1")

    def _down(self, x: _Act, name: str, kind: str) -> _Act:
        x = self._to_compact(x)
        Ho, Wo = (x.H + 1 - 3) // 2 + 1, (x.W + 1 - 3) // 2 + 1  # ZeroPad2d((0, 1, 0, 1)) + Conv2d(3, stride 2)
        y = self._conv(x, name + ".resample.1", (1, 3, 3), stride=(1, 2, 2), pad=(0, 0, 0), out_hw=(Ho, Wo))
        if kind == "down3d" and y.T > 1:
            n = (y.T - 1) // 2
            out = torch.empty((1 + n) * Ho * Wo, y.C, device=self.device, dtype=torch.float32)
            out[:Ho * Wo] = y.frame(0)  # frame 0 passes (first chunk of diffusers' feature cache)
            self._conv(y, name + ".time_conv", (3, 1, 1), stride=(2, 1, 1), pad=(2, 0, 0), frames=(1, n), out_hw=(Ho, Wo),
                       out=out, out_frame0=1)
            y = _Act(out, 1 + n, Ho, Wo, y.C)
        return y

    def _up(self, x: _Act, name: str, kind: str) -> _Act:
        x = self._to_compact(x)
        if kind == "up3d" and x.T > 1:
            hw, Cc = x.H * x.W, x.C
            out = torch.empty((2 * x.T - 1) * hw, Cc, device=self.device, dtype=torch.float32)
            out[:hw] = x.frame(0)
            for half in (0, 1):  # channel halves of time_conv's 2C outputs are the even / odd new frames
                self._conv(x, name + ".time_conv", (3, 1, 1), pad=(2, 0, 0), t_min=1, frames=(1, x.T - 1), out_hw=(x.H, x.W),
                           w_rows=(half * Cc, (half + 1) * Cc), out=out, out_frame0=1 + half, out_frame_step=2)
            x = _Act(out, 2 * x.T - 1, x.H, x.W, Cc)
        return self._conv(x, name + ".resample.1", (1, 3, 3), pad=(0, 1, 1), up=2)

    def _conv3(self, a: _Act, name: str, norm: Optional[str] = None) -> _Act:
        """A causal 3x3x3 convolution, optionally behind ``norm`` (WanRMS_norm) + SiLU: implicit when its tap-major weight exists."""
        gamma = None if norm is None else self._w[norm + ".gamma"]
        if self.implicit and name + ".weight_taps" in self._w and self._fits_implicit(a, self._w[name + ".bias"].numel()):
            return self._conv_taps(self._split_pad(a, gamma, norm is not None), (a.T, a.H, a.W), name)
        a = self._to_compact(a)
        if norm is not None:
            a = _Act(self._norm(a.t, norm, True), a.T, a.H, a.W, a.C)
        return self._conv(a, name, (3, 3, 3))

    def _encode_one(self, x: torch.Tensor) -> torch.Tensor:
        enc, _, _, _, _ = _plan(self._cfg)
        a = self._conv3(self._to_cl(x), "encoder.conv_in")
        for name, kind, ci, co in enc:
            a = self._res(a, name) if kind == "res" else self._down(a, name, kind)
        a = self._mid(a, "encoder.mid_block")
        a = self._to_compact(self._conv3(a, "encoder.conv_out", norm="encoder.norm_out"))
        a.t = self._pointwise(a.t, "quant_conv")
        out = self._from_cl(a, 2 * self._cfg["z_dim"], clamp=False)
        self._release_operands()
        return out

    def _decode_one(self, z: torch.Tensor) -> torch.Tensor:
        _, dec, _, _, _ = _plan(self._cfg)
        a = self._to_cl(z)
        a.t = self._pointwise(a.t, "post_quant_conv")[:, :self._cfg["z_dim"]].contiguous()
        a = self._conv3(a, "decoder.conv_in")
        a = self._mid(a, "decoder.mid_block")
        for name, kind, ci, co in dec:
            a = self._res(a, name) if kind == "res" else self._up(a, name, kind)
        a = self._to_compact(self._conv3(a, "decoder.conv_out", norm="decoder.norm_out"))
        out = self._from_cl(a, 3, clamp=True)
        self._release_operands()
        return out
