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


def gemm(a: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None, *, epilogue: int = _lib.EPI_NONE,
         residual: Optional[torch.Tensor] = None, gate: Optional[torch.Tensor] = None, rows_per_batch: int = 0,
         bias_per_row: bool = False, out: Optional[torch.Tensor] = None, out_dtype=torch.bfloat16) -> torch.Tensor:
    """``epilogue(a @ w.T + bias)``: a [M, K] bf16, w [N, K] bf16 (nn.Linear layout).  See alg_gemm_bf16."""
    _lib.require_cuda(a, w, bias, residual, gate, out)
    assert a.dtype == torch.bfloat16 and w.dtype == torch.bfloat16 and a.dim() == 2 and w.dim() == 2
    assert a.stride(1) == 1 and w.stride(1) == 1 and a.shape[1] == w.shape[1]
    M, K = a.shape
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
    if residual is not None:
        assert residual.dtype == torch.bfloat16 and residual.stride(0) == out.stride(0)
    with torch.cuda.device(a.device):
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
    with torch.cuda.device(q.device):
        _lib.check(_lib.lib().alg_attention_bf16(C.byref(a), _lib.stream_ptr(q.device)))
    return out
