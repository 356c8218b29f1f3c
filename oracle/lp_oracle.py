"""numpy restatement of the reference's low-pass filter + strength schedule.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Pinned against the real
``/root/reference/lp_utils.py`` by ``tests/golden/lp_*.npz`` (made by
``oracle/gen_golden.py``; checked in ``tests/test_oracle_lp.py``).

Follows:
  * ``get_lp_strength``        -- lp_utils.py:63-111
  * ``modulate``               -- pipeline_wan_image2video_lowpass.py:863-867,
                                  pipeline_cogvideox_image2video_lowpass.py:1034-1040,
                                  pipeline_hunyuan_video_image2video_lowpass.py:1144-1151
  * ``apply_low_pass_filter``  -- lp_utils.py:8-60 (early exits :23-28, 5-D view :31-35)
  * ``down_up``                -- lp_utils.py:49-54, i.e. two ATen
                                  ``upsample_bilinear2d_aa`` calls (triangle filter,
                                  align_corners=False, antialias=True)
  * ``gaussian_blur``          -- lp_utils.py:40-47 -> torchvision
                                  ``_functional_tensor.gaussian_blur`` (weights in the
                                  tensor dtype, reflect pad, depthwise conv2d)
"""
from __future__ import annotations

import math

import numpy as np


# ----------------------------------------------------------------------------
# dtype helpers (bf16 / fp16 rounding emulated on fp32 containers)
# ----------------------------------------------------------------------------
def round_bf16(x: np.ndarray) -> np.ndarray:
    """Round fp32 values to the nearest bf16 (ties to even); result stays fp32."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    u = x.view(np.uint32).astype(np.uint64)
    lsb = (u >> 16) & 1
    u = (u + 0x7FFF + lsb) & 0xFFFF0000
    out = u.astype(np.uint32).view(np.float32).reshape(x.shape)
    return np.where(np.isnan(x), x, out)


def round_to(x: np.ndarray, dtype: str) -> np.ndarray:
    if dtype == "float32":
        return np.asarray(x, dtype=np.float32)
    if dtype == "bfloat16":
        return round_bf16(np.asarray(x, dtype=np.float32))
    if dtype == "float16":
        return np.asarray(x, dtype=np.float32).astype(np.float16).astype(np.float32)
    if dtype == "float64":
        return np.asarray(x, dtype=np.float64)
    raise ValueError(dtype)


# ----------------------------------------------------------------------------
# strength schedule  (lp_utils.py:63-111)
# ----------------------------------------------------------------------------
def get_lp_strength(
    step_index,
    total_steps,
    lp_strength_schedule_type,
    schedule_interval_start_time,
    schedule_interval_end_time,
    schedule_linear_start_weight,
    schedule_linear_end_weight,
    schedule_linear_end_time,
    schedule_exp_decay_rate,
):
    step_norm = step_index / max(total_steps - 1, 1)
    kind = lp_strength_schedule_type
    if kind == "linear":
        if schedule_linear_end_time <= 0:
            return schedule_linear_start_weight
        if step_norm >= schedule_linear_end_time:
            return schedule_linear_end_weight
        p = step_norm / schedule_linear_end_time
        return schedule_linear_start_weight * (1 - p) + schedule_linear_end_weight * p
    if kind == "interval":
        return 1.0 if schedule_interval_start_time <= step_norm <= schedule_interval_end_time else 0.0
    if kind == "exponential":
        return math.exp(-abs(schedule_exp_decay_rate) * step_norm)
    return 1.0  # "none" and unknown types


def modulate(lp_blur_sigma, lp_blur_kernel_size, lp_resize_factor, strength, schedule_blur_kernel_size):
    """(sigma', k', f') -- wan:863-867.  NB ``k * s`` is always a float (quirk q8)."""
    sigma = lp_blur_sigma * strength
    k = lp_blur_kernel_size * strength if schedule_blur_kernel_size else lp_blur_kernel_size
    f = 1.0 - (1.0 - lp_resize_factor) * strength
    return sigma, k, f


# ----------------------------------------------------------------------------
# down_up  (lp_utils.py:49-54)
# ----------------------------------------------------------------------------
def down_size(n: int, factor: float) -> int:
    """lp_utils.py:51-52 -- Python round() is half-to-even (quirk q14)."""
    return max(1, int(round(n * factor)))


def aa_weights(in_size: int, out_size: int, dtype=np.float64):
    """Dense [out, in] ATen anti-aliased triangle-filter resampling matrix.

    ATen ``_compute_indices_min_size_weights_aa`` (align_corners=False):
      scale = in/out; support = max(scale, 1); center = scale*(i+0.5)
      xmin = max(int(center - support + 0.5), 0)
      xmax = min(int(center + support + 0.5), in)
      w_j = max(0, 1 - |(j + 0.5 - center) / max(scale, 1)|), row-normalised.
    Returns (M, xmin[out], xsize[out]).
    """
    ft = np.dtype(dtype).type
    scale = ft(in_size) / ft(out_size)
    support = scale if scale >= 1.0 else ft(1.0)
    invscale = ft(1.0) / scale if scale >= 1.0 else ft(1.0)
    M = np.zeros((out_size, in_size), dtype=dtype)
    xmins = np.zeros(out_size, dtype=np.int64)
    xsizes = np.zeros(out_size, dtype=np.int64)
    for i in range(out_size):
        center = scale * ft(i + 0.5)
        xmin = max(int(center - support + ft(0.5)), 0)
        xmax = min(int(center + support + ft(0.5)), in_size)
        j = np.arange(xmin, xmax)
        w = ft(1.0) - np.abs((j.astype(dtype) - center + ft(0.5)) * invscale)
        w = np.maximum(w, ft(0.0)).astype(dtype)
        tot = w.sum(dtype=dtype)
        if tot != 0:
            w = w / tot
        M[i, xmin:xmax] = w
        xmins[i] = xmin
        xsizes[i] = xmax - xmin
    return M, xmins, xsizes


def resample_aa(x: np.ndarray, out_h: int, out_w: int, dtype: str) -> np.ndarray:
    """One ``F.interpolate(..., bilinear, antialias=True)`` on [..., H, W]."""
    H, W = x.shape[-2:]
    acc = np.float64 if dtype == "float64" else np.float32
    Mh, _, _ = aa_weights(H, out_h, acc)
    Mw, _, _ = aa_weights(W, out_w, acc)
    xa = np.asarray(x, dtype=acc)
    # horizontal then vertical, fp32 (fp64) accumulation, one rounding to `dtype`
    y = np.einsum("...hw,jw->...hj", xa, Mw)
    y = np.einsum("ih,...hj->...ij", Mh, y)
    return round_to(y, dtype)


def down_up(x: np.ndarray, resize_factor: float, dtype: str = "float32") -> np.ndarray:
    H, W = x.shape[-2:]
    h1, w1 = down_size(H, resize_factor), down_size(W, resize_factor)
    small = resample_aa(x, h1, w1, dtype)  # rounded to `dtype` between the two calls
    return resample_aa(small, H, W, dtype)


# ----------------------------------------------------------------------------
# gaussian_blur  (lp_utils.py:40-47 -> torchvision)
# ----------------------------------------------------------------------------
def gaussian_kernel_size(blur_kernel_size, H: int) -> int:
    """lp_utils.py:41-46: float => fraction of H, then forced odd."""
    if isinstance(blur_kernel_size, float):
        k = max(int(blur_kernel_size * H), 1)
    else:
        k = int(blur_kernel_size)
    if k % 2 == 0:
        k += 1
    return k


def gaussian_kernel1d(k: int, sigma: float, dtype: str) -> np.ndarray:
    """torchvision ``_get_gaussian_kernel1d`` with every op rounded to ``dtype``."""
    half = (k - 1) * 0.5
    if k == 1:
        x = np.array([-half], dtype=np.float32)
    else:
        # torch.linspace: symmetric two-sided formula, step computed in fp32
        step = np.float32((half - (-half)) / (k - 1))
        i = np.arange(k)
        lo = np.float32(-half) + step * i.astype(np.float32)
        hi = np.float32(half) - step * (k - 1 - i).astype(np.float32)
        x = np.where(i < k // 2, lo, hi).astype(np.float32)
    x = round_to(x, dtype)
    q = round_to(x / np.float32(sigma), dtype)
    q = round_to(q * q, dtype)
    q = round_to(np.float32(-0.5) * q, dtype)
    pdf = round_to(np.exp(q.astype(np.float32)), dtype)
    tot = round_to(np.array(pdf.sum(dtype=np.float32)), dtype)
    return round_to(pdf / tot, dtype)


def gaussian_blur(x: np.ndarray, k: int, sigma: float, dtype: str = "float32") -> np.ndarray:
    """Dense k x k depthwise conv with reflect padding; fp32 accumulate, one rounding."""
    w1 = gaussian_kernel1d(k, sigma, dtype)
    w2 = round_to(np.outer(w1, w1), dtype)  # torch.mm in the tensor dtype
    r = k // 2
    H, W = x.shape[-2:]
    xa = np.asarray(x, dtype=np.float32)
    pad = [(0, 0)] * (xa.ndim - 2) + [(r, r), (r, r)]
    xp = np.pad(xa, pad, mode="reflect")
    out = np.zeros_like(xa, dtype=np.float32)
    for dy in range(k):
        for dx in range(k):
            out += np.float32(w2[dy, dx]) * xp[..., dy : dy + H, dx : dx + W]
    return round_to(out, dtype)


# ----------------------------------------------------------------------------
# dispatcher  (lp_utils.py:8-60)
# ----------------------------------------------------------------------------
def apply_low_pass_filter(x, filter_type, blur_sigma, blur_kernel_size, resize_factor, dtype="float32"):
    if filter_type == "none":
        return x
    if filter_type == "down_up" and resize_factor == 1.0:
        return x
    if filter_type == "gaussian_blur" and blur_sigma == 0:
        return x
    shape = x.shape
    if x.ndim == 5:  # view(B*K, C, H, W) without a permute (quirk q6)
        B, C, K, H, W = shape
        x = x.reshape(B * K, C, H, W)
    H = x.shape[-2]
    if filter_type == "gaussian_blur":
        k = gaussian_kernel_size(blur_kernel_size, H)
        x = gaussian_blur(x, k, blur_sigma, dtype)
    elif filter_type == "down_up":
        x = down_up(x, resize_factor, dtype)
    return x.reshape(shape)
