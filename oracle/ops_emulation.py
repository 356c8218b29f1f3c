"""PyTorch emulation of the tensor-level C-ABI wrappers in ``alg_b200/ops.py`` -- TEST INFRASTRUCTURE ONLY.

Same function names and signatures as ``alg_b200.ops``; each function restates, in eager PyTorch on any device, what the
corresponding CUDA kernel computes, rounding to bf16 exactly where the kernel does.  Two uses, both in ``tests/``:

  * ``-m "not gpu"``: the host-side DiT sequencers (``alg_b200/cogvideox.py``, ``alg_b200/hunyuan.py``) are run on CPU
    with their ``ops`` module monkeypatched to this one and compared with the model oracles -- that checks weight
    naming, buffer layout, row splits, modulation chunk order and RoPE tables without a GPU;
  * ``-m gpu``: every CUDA op is compared with its emulation on the same inputs (per-op parity).

The product never imports this module; ``alg_b200.ops`` raises without the CUDA library.
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn.functional as F

EPI_NONE, EPI_GELU_TANH, EPI_GATE_RESIDUAL, EPI_RESIDUAL, EPI_GELU_ERF, EPI_SILU = range(6)
NORM_NONE, NORM_RMS, NORM_LAYER = 0, 1, 2


def _r(x):  # round to bf16, keep fp32 container
    return x.to(torch.bfloat16).float()


def gemm(a, w, bias=None, *, epilogue=EPI_NONE, residual=None, gate=None, rows_per_batch=0, bias_per_row=False, out=None,
         out_dtype=torch.bfloat16, gate_alt=None, gate_split_row=0, gate_round=False, a_k_period=0, a_tap_kblocks=0,
         a_tap_offsets=None, m_rows=0, bias_f32=None, residual_f32=None):
    if a_k_period:
        a = a.repeat(1, w.shape[1] // a_k_period)
    if a_tap_kblocks:  # implicit convolution: K group g = the same columns of a, rows shifted by offsets[g] (zero outside)
        rows, m = a.shape[0], int(m_rows or a.shape[0])
        cols = []
        for off in a_tap_offsets:
            idx = torch.arange(m, device=a.device) + int(off)
            ok = (idx >= 0) & (idx < rows)
            cols.append(a[idx.clamp(0, rows - 1)] * ok[:, None].to(a.dtype))
        a = torch.cat(cols, dim=1)
    M, N = a.shape[0], w.shape[0]
    acc = a.float() @ w.float().t()
    if bias is not None:
        acc = acc + (bias.float()[:, None] if bias_per_row else bias.float()[None, :])
    if bias_f32 is not None:
        acc = acc + bias_f32[None, :acc.shape[1]]
    if residual_f32 is not None:
        acc = acc + residual_f32
    if epilogue != EPI_NONE:
        y = _r(acc)
        if epilogue == EPI_GELU_TANH:
            y = F.gelu(y, approximate="tanh")
        elif epilogue == EPI_GELU_ERF:
            y = F.gelu(y)
        elif epilogue == EPI_SILU:
            y = F.silu(y)
        else:
            if epilogue == EPI_GATE_RESIDUAL:
                rpb = rows_per_batch or M
                rows = torch.arange(M, device=a.device)
                b_idx, r_in = rows // rpb, rows % rpb
                g = gate.float().reshape(-1, N)
                g_rows = g[b_idx if g.shape[0] > 1 else torch.zeros_like(b_idx)]
                if gate_alt is not None and gate_split_row > 0:
                    ga = gate_alt.float().reshape(-1, N)
                    ga_rows = ga[b_idx if ga.shape[0] > 1 else torch.zeros_like(b_idx)]
                    g_rows = torch.where((r_in < gate_split_row)[:, None], ga_rows, g_rows)
                y = y * g_rows
                if gate_round:
                    y = _r(y)
            y = residual.float() + y
        acc = y
    if out is None:
        out = torch.empty(M, N, device=a.device, dtype=out_dtype)
    out.copy_(acc.to(out.dtype))
    return out


def attention(q, k, vt, *, n_kv=None, out=None, accumulate=False, scale=None):
    B, Nq, H, D = q.shape
    n_kv = n_kv or k.shape[1]
    scale = scale if scale is not None else 1.0 / math.sqrt(D)
    qf, kf = q.float().transpose(1, 2), k[:, :n_kv].float().transpose(1, 2)
    vf = vt[..., :n_kv].float().transpose(2, 3)
    o = F.scaled_dot_product_attention(qf, kf, vf, scale=scale).transpose(1, 2)
    if out is None:
        out = torch.empty(B, Nq, H, D, device=q.device, dtype=torch.bfloat16)
    if accumulate:
        o = _r(o) + out.float()
    out.copy_(o.to(torch.bfloat16))
    return out


def layer_norm(x, *, eps, weight=None, bias=None, scale=None, shift=None, scale_alt=None, shift_alt=None,
               rows_per_batch=0, split_row=0, chain_bf16=False, out=None):
    rows, d = x.shape
    y = F.layer_norm(x.float(), (d,), None if weight is None else weight.float(), None if bias is None else bias.float(), eps)
    if scale is not None:
        rpb = rows_per_batch or max(rows, 1)
        r = torch.arange(rows, device=x.device)
        b_idx, r_in = r // rpb, r % rpb

        def per_row(v, v_alt):
            v2 = v.float().reshape(-1, d)
            sel = v2[b_idx if v2.shape[0] > 1 else torch.zeros_like(b_idx)]
            if v_alt is not None and split_row > 0:
                a2 = v_alt.float().reshape(-1, d)
                sel = torch.where((r_in < split_row)[:, None], a2[b_idx if a2.shape[0] > 1 else torch.zeros_like(b_idx)], sel)
            return sel

        sc, sh = per_row(scale, scale_alt), per_row(shift, shift_alt)
        if chain_bf16:
            y = _r(_r(y) * _r(1.0 + sc)) + sh
        else:
            y = y * (1.0 + sc) + sh
    if out is None:
        out = torch.empty_like(x)
    out.copy_(y.to(torch.bfloat16))
    return out


def head_norm_rope(x, heads, head_dim, *, norm_kind=NORM_NONE, weight=None, bias=None, eps=1e-6, cos=None, sin=None,
                   rows_per_batch=0, rope_row0=0, rope_rows=0):
    rows = x.shape[0]
    v = x[:, : heads * head_dim].float().reshape(rows, heads, head_dim)
    if norm_kind == NORM_RMS:
        var = v.pow(2).mean(-1, keepdim=True)
        v = _r(_r(v * torch.rsqrt(var + eps)) * weight.float())
    elif norm_kind == NORM_LAYER:
        v = _r(F.layer_norm(v, (head_dim,), weight.float(), None if bias is None else bias.float(), eps))
    if cos is not None:
        rpb = rows_per_batch or rows
        rope_rows = rope_rows or cos.shape[0]
        r_in = torch.arange(rows, device=x.device) % rpb
        sel = (r_in >= rope_row0) & (r_in < rope_row0 + rope_rows)
        idx = (r_in - rope_row0).clamp(0, cos.shape[0] - 1)
        c, s = cos.to(x.device)[idx][:, None, :], sin.to(x.device)[idx][:, None, :]
        re, im = v.reshape(rows, heads, -1, 2).unbind(-1)
        rot = torch.stack([-im, re], dim=-1).flatten(2)
        v = torch.where(sel[:, None, None], v * c + rot * s, v)
    x[:, : heads * head_dim] = v.reshape(rows, -1).to(torch.bfloat16)
    return x


def patch_gather(passes, out):
    rows = []
    for srcs in passes:
        chans = []
        for s in srcs:
            t, t0 = s if isinstance(s, tuple) else (s, None)
            t = t.float()
            if t0 is not None:
                t = torch.cat([t0.float(), t[:, 1:]], dim=1)
            chans.append(t)
        x = torch.cat(chans, dim=0)  # [C, T, H, W]
        C, T, H, W = x.shape
        x = x.reshape(C, T, H // 2, 2, W // 2, 2).permute(1, 2, 4, 0, 3, 5).reshape(T * (H // 2) * (W // 2), C * 4)
        rows.append(x)
    a = torch.cat(rows, dim=0)
    out[:, : a.shape[1]] = a.to(torch.bfloat16)
    return out


def unpatchify(proj, out, channel_major):
    n_pass, C, T, H, W = out.shape
    N = T * (H // 2) * (W // 2)
    p = proj[:, : 4 * C].reshape(n_pass, T, H // 2, W // 2, *((C, 2, 2) if channel_major else (2, 2, C)))
    p = p.permute(0, 4, 1, 2, 5, 3, 6) if channel_major else p.permute(0, 6, 1, 2, 4, 3, 5)
    out.copy_(p.reshape(n_pass, C, T, H, W))
    return out


def timestep_embedding(t, dim, dtype, device, out=None):
    half = dim // 2
    exponent = -math.log(10000) * torch.arange(0, half, dtype=torch.float32, device=device) / half
    emb = float(t) * torch.exp(exponent)
    r = torch.cat([torch.cos(emb), torch.sin(emb)]).to(dtype)
    return r if out is None else out.copy_(r)


def add(a, b, out=None):
    r = (a.float() + b.float()).to(torch.bfloat16)
    return r if out is None else out.copy_(r)


def silu(a, out=None):
    r = F.silu(a.float()).to(torch.bfloat16)
    return r if out is None else out.copy_(r)


def mean_rows(x):
    return x.float().mean(0).to(torch.bfloat16)


def pad_frames(src, dst, frames, H, W, *, to_padded, residual=None):
    """alg_pad_frames_bf16: compact [frames*H*W, C] <-> zero-bordered raster [frames, H + 2, W + 2, ld]."""
    if to_padded:
        Cc = src.shape[1]
        dst.view(frames, H + 2, W + 2, -1)[:, 1:H + 1, 1:W + 1, :Cc] = src.view(frames, H, W, Cc)
    else:
        Cc = dst.shape[1]
        v = src.reshape(frames, H + 2, W + 2, -1)[:, 1:H + 1, 1:W + 1, :Cc].reshape(frames * H * W, Cc).float()
        if residual is not None:
            v = v + residual.float()
        dst.copy_(v.to(dst.dtype))
    return dst


def copy_rows(src, dst):
    dst.copy_(src)
    return dst


def im2col(x, T, H, W, *, kernel, stride=(1, 1, 1), pad_t=0, pad_top=0, pad_left=0, out_hw=None, out=None):
    """alg_im2col_bf16: columns ordered (it, ih, iw, c); frames before 0 replicate frame 0; zero spatial padding."""
    Cc = x.shape[1]
    kt, kh, kw = kernel
    st, sh, sw = stride
    Ho, Wo = out_hw if out_hw is not None else (H, W)
    To = (T + pad_t - kt) // st + 1
    v = x.view(T, H, W, Cc)
    need_h, need_w = (Ho - 1) * sh + kh, (Wo - 1) * sw + kw
    vp = torch.zeros(T, need_h, need_w, Cc, dtype=x.dtype, device=x.device)
    hh, ww = min(H, need_h - pad_top), min(W, need_w - pad_left)
    vp[:, pad_top:pad_top + hh, pad_left:pad_left + ww] = v[:, :hh, :ww]
    taps = []
    for it in range(kt):
        tidx = torch.clamp(torch.arange(To) * st + it - pad_t, min=0)
        for ih in range(kh):
            for iw in range(kw):
                taps.append(vp[tidx][:, ih:ih + (Ho - 1) * sh + 1:sh, iw:iw + (Wo - 1) * sw + 1:sw])
    cols = torch.stack(taps, dim=3).reshape(To * Ho * Wo, kt * kh * kw * Cc)
    if out is not None:
        out.view(-1)[: cols.numel()].view_as(cols).copy_(cols)
        return out.view(-1)[: cols.numel()].view_as(cols)
    return cols.contiguous()


def group_norm(x, groups, weight=None, bias=None, *, eps=1e-6, silu=False, out=None):
    """alg_group_norm_bf16: fp32 statistics over (rows, C/groups); y = bf16(a*x + b); optional SiLU rounded again."""
    rows, Cc = x.shape
    xf = x.float().view(rows, groups, Cc // groups)
    mean = xf.double().mean(dim=(0, 2))
    var = (xf.double() ** 2).mean(dim=(0, 2)) - mean ** 2
    rstd = torch.rsqrt(var.clamp_min(0).float() + eps)
    ga = weight.float() if weight is not None else torch.ones(Cc, device=x.device)
    be = bias.float() if bias is not None else torch.zeros(Cc, device=x.device)
    a = rstd.repeat_interleave(Cc // groups) * ga
    b = be - a * mean.float().repeat_interleave(Cc // groups)
    y = _r(x.float() * a[None, :] + b[None, :])
    if silu:
        y = _r(y / (1.0 + torch.exp(-y)))
    y = y.to(torch.bfloat16)
    if out is not None:
        out.copy_(y)
        return out
    return y
