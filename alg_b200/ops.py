"""Tensor-level wrappers over the tcgen05 building blocks of ``libalg_b200.so``.

These take ``torch.Tensor`` only to borrow ``data_ptr()`` and the current CUDA
stream; all arithmetic happens in the library.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional

import torch

from . import _lib


# ------------------------------------------------------------------------------------------------------
# per-kernel-class device timing for the Python sequencers (CogVideoX / HunyuanVideo; the Wan engine times its own
# classes in C++): CUDA events on the launching stream around every GEMM / attention launch while enabled
# ------------------------------------------------------------------------------------------------------
_PROFILE = None


def profile_start():
    global _PROFILE
    _PROFILE = {}


def profile_stop():
    """-> {class: {"ms", "launches", "flops"}} (synchronises the device)."""
    global _PROFILE
    rec, _PROFILE = _PROFILE or {}, None
    torch.cuda.synchronize()
    return {k: {"ms": sum(e0.elapsed_time(e1) for e0, e1, _ in v), "launches": len(v), "flops": sum(f for _, _, f in v)}
            for k, v in rec.items()}


class _Span:
    def __init__(self, key, flops):
        self.key, self.flops = key, flops

    def __enter__(self):
        if _PROFILE is not None:
            self.e0, self.e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            self.e0.record()

    def __exit__(self, *a):
        if _PROFILE is not None:
            self.e1.record()
            _PROFILE.setdefault(self.key, []).append((self.e0, self.e1, self.flops))


def gemm(a: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None, *, epilogue: int = _lib.EPI_NONE,
         residual: Optional[torch.Tensor] = None, gate: Optional[torch.Tensor] = None, rows_per_batch: int = 0,
         bias_per_row: bool = False, out: Optional[torch.Tensor] = None, out_dtype=torch.bfloat16,
         gate_alt: Optional[torch.Tensor] = None, gate_split_row: int = 0, gate_round: bool = False,
         a_k_period: int = 0, a_tap_kblocks: int = 0, a_tap_offsets=None, m_rows: int = 0,
         bias_f32: Optional[torch.Tensor] = None, residual_f32: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``epilogue(a @ w.T + bias)``: a [M, K] bf16, w [N, K] bf16 (nn.Linear layout).  See alg_gemm_bf16.

    ``a_k_period``: a is [M, period] and repeats along K (``a.repeat(1, K // period)`` without materialising it).

    ``gate`` fp32 or bf16, [N] or [batches, N]; ``gate_alt`` replaces it for rows whose index inside their sample is
    below ``gate_split_row``; ``gate_round`` rounds ``gate * y`` to bf16 before the residual add (eager bf16 chain)."""
    _lib.require_cuda(a, w, bias, residual, gate, out, gate_alt, bias_f32, residual_f32)
    assert a.dtype == torch.bfloat16 and w.dtype == torch.bfloat16 and a.dim() == 2 and w.dim() == 2
    assert a.stride(1) == 1 and w.stride(1) == 1 and a.shape[1] == (a_k_period or a_tap_kblocks * 64 or w.shape[1])
    M, K = (int(m_rows) if (a_tap_kblocks and m_rows) else a.shape[0]), w.shape[1]
    N = w.shape[0]
    if out is None:
        ldd = (N + 7) // 8 * 8
        buf = torch.empty(M, ldd, device=a.device, dtype=out_dtype)
        out = buf[:, :N]
    assert out.stride(1) == 1 and out.dtype in (torch.bfloat16, torch.float32)
    g = _lib.Gemm()
    g.A, g.B, g.D = a.data_ptr(), w.data_ptr(), out.data_ptr()
    g.bias = bias.data_ptr() if bias is not None else None
    g.R = residual.data_ptr() if residual is not None else None
    g.gate = gate.data_ptr() if gate is not None else None
    g.M, g.N, g.K = M, N, K
    g.lda, g.ldb, g.ldd = a.stride(0), w.stride(0), out.stride(0)
    g.rows_per_batch = rows_per_batch or M
    g.gate_ld = gate.stride(0) if gate is not None and gate.dim() == 2 else 0
    g.epilogue = epilogue
    g.bias_per_row = int(bias_per_row)
    g.out_f32 = int(out.dtype == torch.float32)
    g.a_k_period = int(a_k_period)
    if bias_f32 is not None:  # fp32 output only: fp32 bias / residual added in the epilogue (no extra pass over D)
        assert bias_f32.dtype == torch.float32 and bias_f32.numel() >= N and out.dtype == torch.float32
        g.bias_f32 = bias_f32.data_ptr()
    if residual_f32 is not None:
        assert residual_f32.dtype == torch.float32 and residual_f32.stride(0) == out.stride(0) and out.dtype == torch.float32
        g.residual_f32 = residual_f32.data_ptr()
    if a_tap_kblocks:  # implicit convolution: K groups read row-shifted copies of the same A columns (see alg_gemm_t)
        offs = (C.c_int32 * len(a_tap_offsets))(*[int(o) for o in a_tap_offsets])
        g.a_tap_kblocks, g.a_n_taps, g.a_tap_offsets = int(a_tap_kblocks), len(a_tap_offsets), offs
        g.a_rows = a.shape[0]  # the offsets may reach past the last output row (m_rows < rows of a)
    if gate is not None:
        assert gate.dtype in (torch.float32, torch.bfloat16) and gate.stride(-1) == 1
        g.gate_dtype = _lib.dtype_code(gate.dtype)
        g.gate_round = int(gate_round)
        if gate_alt is not None:
            assert gate_alt.dtype == gate.dtype and gate_alt.stride(-1) == 1
            assert gate_alt.dim() == 1 or gate.dim() == 1 or gate_alt.stride(0) == gate.stride(0)
            g.gate_alt = gate_alt.data_ptr()
            g.gate_split_row = int(gate_split_row)
    if residual is not None:
        assert residual.dtype == torch.bfloat16 and residual.stride(0) == out.stride(0)
    with torch.cuda.device(a.device), _Span("gemm", 2 * g.M * g.N * g.K):
        _lib.check(_lib.lib().alg_gemm_bf16(C.byref(g), _lib.stream_ptr(a.device)))
    return out


def attention(q: torch.Tensor, k: torch.Tensor, vt: torch.Tensor, *, n_kv: Optional[int] = None,
              out: Optional[torch.Tensor] = None, accumulate: bool = False, scale: Optional[float] = None):
    """Non-causal attention.  q [B, Nq, H, D], k [B, Nkv, H, D], vt [B, H, D, Nkv_padded] (V transposed), bf16.

    Returns o [B, Nq, H, D].  See alg_attention_bf16 (replaces F.scaled_dot_product_attention).
    """
    _lib.require_cuda(q, k, vt, out)
    B, Nq, H, D = q.shape
    n_kv = n_kv or k.shape[1]
    assert q.dtype == k.dtype == vt.dtype == torch.bfloat16
    assert q.stride(3) == 1 and q.stride(2) == D and k.stride(3) == 1 and k.stride(2) == D
    assert vt.stride(3) == 1 and vt.stride(1) == D * vt.stride(2)
    if out is None:
        out = torch.empty(B, Nq, H, D, device=q.device, dtype=torch.bfloat16)
    a = _lib.Attention()
    a.Q, a.K, a.Vt, a.O = q.data_ptr(), k.data_ptr(), vt.data_ptr(), out.data_ptr()
    a.batch, a.heads, a.head_dim, a.n_q, a.n_kv = B, H, D, Nq, n_kv
    a.q_bs, a.q_rs = q.stride(0), q.stride(1)
    a.k_bs, a.k_rs = k.stride(0), k.stride(1)
    a.v_bs, a.v_rs = vt.stride(0), vt.stride(2)
    a.o_bs, a.o_rs = out.stride(0), out.stride(1)
    a.scale = scale if scale is not None else 1.0 / math.sqrt(D)
    a.accumulate = int(accumulate)
    with torch.cuda.device(q.device), _Span("self_attention" if n_kv > 1024 else "short_attention", 4 * B * H * Nq * n_kv * D):
        _lib.check(_lib.lib().alg_attention_bf16(C.byref(a), _lib.stream_ptr(q.device)))
    return out


# ------------------------------------------------------------------------------------------------------
# HBM-bound DiT building blocks (alg_layer_norm, alg_head_norm_rope, alg_patch_gather, ...)
# ------------------------------------------------------------------------------------------------------
def _ptr(t):
    return None if t is None else t.data_ptr()


def _launch(fn, device, *args):
    with torch.cuda.device(device):
        _lib.check(fn(*args, _lib.stream_ptr(device)))


def layer_norm(x: torch.Tensor, *, eps: float, weight=None, bias=None, scale=None, shift=None, scale_alt=None,
               shift_alt=None, rows_per_batch: int = 0, split_row: int = 0, chain_bf16: bool = False,
               out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """LayerNorm over the rows of x [rows, d] bf16 (+affine) (+AdaLN modulate).  See alg_layer_norm."""
    _lib.require_cuda(x, weight, bias, scale, shift, scale_alt, shift_alt, out)
    assert x.dtype == torch.bfloat16 and x.dim() == 2 and x.is_contiguous()
    rows, d = x.shape
    if out is None:
        out = torch.empty_like(x)
    assert out.is_contiguous() and out.shape == x.shape and out.dtype == torch.bfloat16
    p = _lib.LayerNorm()
    p.x, p.out, p.rows, p.d, p.eps = x.data_ptr(), out.data_ptr(), rows, d, eps
    if weight is not None:
        assert weight.dtype == bias.dtype and weight.numel() == d and weight.is_contiguous() and bias.is_contiguous()
        p.weight, p.bias, p.affine_dtype = weight.data_ptr(), bias.data_ptr(), _lib.dtype_code(weight.dtype)
    if scale is not None:
        assert scale.dtype == shift.dtype and scale.stride(-1) == 1 and shift.stride(-1) == 1
        p.scale, p.shift, p.mod_dtype = scale.data_ptr(), shift.data_ptr(), _lib.dtype_code(scale.dtype)
        p.rows_per_batch = rows_per_batch or max(rows, 1)
        if scale.dim() == 2 and scale.shape[0] > 1:
            assert shift.stride(0) == scale.stride(0)
            p.mod_batch_stride = scale.stride(0)
        if scale_alt is not None:
            assert scale_alt.dtype == scale.dtype and shift_alt.dtype == scale.dtype
            if scale.dim() == 2 and scale.shape[0] > 1:
                assert scale_alt.stride(0) == scale.stride(0) and shift_alt.stride(0) == scale.stride(0)
            p.scale_alt, p.shift_alt, p.split_row = scale_alt.data_ptr(), shift_alt.data_ptr(), int(split_row)
    p.chain_bf16 = int(chain_bf16)
    _launch(_lib.lib().alg_layer_norm, x.device, C.byref(p))
    return out


def head_norm_rope(x: torch.Tensor, heads: int, head_dim: int, *, norm_kind: int = _lib.NORM_NONE, weight=None,
                   bias=None, eps: float = 1e-6, cos=None, sin=None, rows_per_batch: int = 0, rope_row0: int = 0,
                   rope_rows: int = 0) -> torch.Tensor:
    """In-place per-head q/k norm + rotary embedding on x [rows, >= heads*head_dim] bf16.  See alg_head_norm_rope."""
    _lib.require_cuda(x, weight, bias, cos, sin)
    assert x.dtype == torch.bfloat16 and x.dim() == 2 and x.stride(1) == 1
    p = _lib.HeadNormRope()
    p.x, p.rows, p.ld, p.heads, p.head_dim = x.data_ptr(), x.shape[0], x.stride(0), heads, head_dim
    p.norm_kind, p.eps, p.weight, p.bias = norm_kind, eps, _ptr(weight), _ptr(bias)
    if weight is not None:
        assert weight.dtype == torch.bfloat16 and weight.numel() == head_dim
    if cos is not None:
        assert cos.dtype == sin.dtype == torch.float32 and cos.is_contiguous() and sin.is_contiguous()
        assert cos.shape[-1] == head_dim and cos.shape == sin.shape
        p.cos, p.sin = cos.data_ptr(), sin.data_ptr()
        p.rope_rows = rope_rows or cos.shape[0]
        assert p.rope_rows <= cos.shape[0]
    p.rows_per_batch, p.rope_row0 = rows_per_batch or x.shape[0], rope_row0
    _launch(_lib.lib().alg_head_norm_rope, x.device, C.byref(p))
    return x


def patch_gather(passes, out: torch.Tensor) -> torch.Tensor:
    """Model-input assembly + 2x2 im2col.  ``passes``: per CFG pass, a list of sources; a source is a [C, T, H, W] tensor
    view (stride(-1) == 1, any float dtype) or a tuple (view, frame0_view [C, 1, H, W]).  out [n_pass*N, >= 4*C_total] bf16."""
    n_pass, n_src = len(passes), len(passes[0])
    arr = (_lib.PatchSrc * (n_pass * n_src))()
    keep = []
    T = H = W = None
    for pi, srcs in enumerate(passes):
        assert len(srcs) == n_src
        for si, s in enumerate(srcs):
            t, t0 = s if isinstance(s, tuple) else (s, None)
            _lib.require_cuda(t, t0)
            assert t.dim() == 4 and t.stride(3) == 1
            Cc, T_, H_, W_ = t.shape
            assert (T, H, W) in ((None, None, None), (T_, H_, W_))
            T, H, W = T_, H_, W_
            d = arr[pi * n_src + si]
            d.ptr, d.dtype, d.channels = t.data_ptr(), _lib.dtype_code(t.dtype), Cc
            d.sc, d.st, d.sy = t.stride(0), t.stride(1), t.stride(2)
            if t0 is not None:
                assert t0.dtype == t.dtype and t0.stride(-1) == 1 and t0.stride(-2) == t.stride(2) and t0.shape[0] == Cc
                d.ptr_t0, d.sc_t0 = t0.data_ptr(), t0.stride(0)
            keep.append((t, t0))
    assert out.dtype == torch.bfloat16 and out.dim() == 2 and out.stride(1) == 1
    _launch(_lib.lib().alg_patch_gather, out.device, arr, n_pass, n_src, T, H, W, out.data_ptr(), out.stride(0))
    return out


def unpatchify(proj: torch.Tensor, out: torch.Tensor, channel_major: bool) -> torch.Tensor:
    """proj [n_pass*N, >= 4*C] bf16 -> out, a [n_pass, C, T, H, W] bf16 view (stride(-1) == 1).  See alg_unpatchify."""
    _lib.require_cuda(proj, out)
    assert proj.dtype == out.dtype == torch.bfloat16 and out.dim() == 5 and out.stride(4) == 1 and proj.stride(1) == 1
    n_pass, Cc, T, H, W = out.shape
    _launch(_lib.lib().alg_unpatchify, out.device, proj.data_ptr(), proj.stride(0), out.data_ptr(), n_pass, Cc, T, H, W,
            out.stride(0), out.stride(1), out.stride(2), out.stride(3), int(channel_major))
    return out


def timestep_embedding(t: float, dim: int, dtype, device, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """diffusers get_timestep_embedding(flip_sin_to_cos=True, downscale_freq_shift=0) of one scalar -> [dim]."""
    if out is None:
        out = torch.empty(dim, device=device, dtype=dtype)
    assert out.numel() == dim and out.is_contiguous() and out.dtype == dtype
    _lib.require_cuda(out)
    _launch(_lib.lib().alg_timestep_embedding, out.device, float(t), dim, out.data_ptr(), _lib.dtype_code(dtype))
    return out


def _ew(op, a, b=None, out=None):
    _lib.require_cuda(a, b, out)
    assert a.dtype == torch.bfloat16 and a.is_contiguous() and (b is None or (b.shape == a.shape and b.is_contiguous()))
    if out is None:
        out = torch.empty_like(a)
    _launch(_lib.lib().alg_elementwise_bf16, a.device, op, a.data_ptr(), _ptr(b), out.data_ptr(), a.numel())
    return out


def add(a, b, out=None):
    return _ew(_lib.EW_ADD, a, b, out)


def silu(a, out=None):
    return _ew(_lib.EW_SILU, a, None, out)


def mean_rows(x: torch.Tensor) -> torch.Tensor:
    """bf16 [rows, d] -> bf16 [d], fp32 accumulation."""
    _lib.require_cuda(x)
    assert x.dtype == torch.bfloat16 and x.dim() == 2 and x.stride(1) == 1
    out = torch.empty(x.shape[1], device=x.device, dtype=torch.bfloat16)
    _launch(_lib.lib().alg_mean_rows_bf16, x.device, x.data_ptr(), x.shape[0], x.shape[1], x.stride(0), out.data_ptr())
    return out


def copy_rows(src: torch.Tensor, dst: torch.Tensor) -> torch.Tensor:
    """dst[r, :d] = src[r, :d] (bf16 2-D views with unit inner stride)."""
    _lib.require_cuda(src, dst)
    assert src.dtype == dst.dtype == torch.bfloat16 and src.shape == dst.shape and src.stride(1) == dst.stride(1) == 1
    _launch(_lib.lib().alg_copy_rows_bf16, src.device, src.data_ptr(), src.stride(0), dst.data_ptr(), dst.stride(0),
            src.shape[0], src.shape[1])
    return dst


def pad_frames(src: torch.Tensor, dst: torch.Tensor, frames: int, H: int, W: int, *, to_padded: bool,
               residual: Optional[torch.Tensor] = None) -> torch.Tensor:
    """bf16 frames <-> their zero-bordered raster [frames, H + 2, W + 2, ld] (alg_pad_frames_bf16, operand / result of the implicit
    convolution).  ``to_padded``: src compact [frames*H*W, C] -> interior of dst (borders untouched); else src padded (row length
    ``src.stride(0)``) -> dst compact [frames*H*W, C] (+ bf16 ``residual``: dst = bf16(float(src) + float(residual)))."""
    _lib.require_cuda(src, dst, residual)
    assert src.dtype == dst.dtype == torch.bfloat16 and src.stride(1) == dst.stride(1) == 1
    Cc = dst.shape[1] if not to_padded else src.shape[1]
    ld = dst.stride(0) if to_padded else src.stride(0)
    _launch(_lib.lib().alg_pad_frames_bf16, src.device, src.data_ptr(), dst.data_ptr(), None if residual is None else residual.data_ptr(),
            frames, H, W, Cc, ld, int(to_padded))
    return dst


def im2col(x: torch.Tensor, T: int, H: int, W: int, *, kernel, stride=(1, 1, 1), pad_t: int = 0, pad_top: int = 0,
           pad_left: int = 0, out_hw=None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Patch gather for a (causal 3-D / strided 2-D) convolution on channels-last x [T*H*W, C] bf16.  See alg_im2col_bf16.

    Returns cols [To*Ho*Wo, kt*kh*kw*C] (column order (it, ih, iw, c)); ``out_hw`` = (Ho, Wo), default: same size.
    ``out`` may be a larger flat bf16 workspace: the result is a view of its head."""
    _lib.require_cuda(x, out)
    assert x.dtype == torch.bfloat16 and x.dim() == 2 and x.is_contiguous() and x.shape[0] == T * H * W
    Cc = x.shape[1]
    kt, kh, kw = kernel
    st, sh, sw = stride
    Ho, Wo = out_hw if out_hw is not None else (H, W)
    To = (T + pad_t - kt) // st + 1
    K = kt * kh * kw * Cc
    M = To * Ho * Wo
    if out is None:
        cols = torch.empty(M, K, device=x.device, dtype=torch.bfloat16)
    else:
        assert out.dtype == torch.bfloat16 and out.is_contiguous() and out.numel() >= M * K
        cols = out.view(-1)[: M * K].view(M, K)
    p = _lib.Im2col()
    p.x, p.cols = x.data_ptr(), cols.data_ptr()
    p.T, p.H, p.W, p.C = T, H, W, Cc
    p.kt, p.kh, p.kw, p.st, p.sh, p.sw = kt, kh, kw, st, sh, sw
    p.pad_t, p.pad_top, p.pad_left = pad_t, pad_top, pad_left
    p.To, p.Ho, p.Wo, p.ld = To, Ho, Wo, K
    _launch(_lib.lib().alg_im2col_bf16, x.device, C.byref(p))
    return cols


def group_norm(x: torch.Tensor, groups: int, weight=None, bias=None, *, eps: float = 1e-6, silu: bool = False,
               out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """nn.GroupNorm (+ SiLU) over channels-last x [rows, C] bf16 (one sample).  See alg_group_norm_bf16."""
    _lib.require_cuda(x, weight, bias, out)
    assert x.dtype == torch.bfloat16 and x.dim() == 2 and x.is_contiguous()
    if out is None:
        out = torch.empty_like(x)
    assert out.shape == x.shape and out.dtype == torch.bfloat16 and out.is_contiguous()
    stats = torch.empty(2 * groups, device=x.device, dtype=torch.float64)
    p = _lib.GroupNorm()
    p.x, p.y, p.stats = x.data_ptr(), out.data_ptr(), stats.data_ptr()
    if weight is not None:
        assert weight.dtype == torch.bfloat16 and weight.numel() == x.shape[1] and weight.is_contiguous()
        p.weight = weight.data_ptr()
    if bias is not None:
        assert bias.dtype == torch.bfloat16 and bias.numel() == x.shape[1] and bias.is_contiguous()
        p.bias = bias.data_ptr()
    p.rows, p.C, p.groups, p.eps, p.silu = x.shape[0], x.shape[1], groups, eps, int(silu)
    _launch(_lib.lib().alg_group_norm_bf16, x.device, C.byref(p))
    return out
