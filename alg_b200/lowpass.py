"""Host-side mirror of the reference's ``lp_utils`` for the hot path.

Same names, argument meaning and error behaviour as ``/root/reference/lp_utils.py``;
the filtering itself runs in ``libalg_b200.so`` (sm_100a CUDA) -- there is no
PyTorch / CPU fallback.

  * ``apply_low_pass_filter``  <- lp_utils.py:8-60
  * ``get_lp_strength``        <- lp_utils.py:63-111
  * ``get_hunyuan_video_size`` <- lp_utils.py:113-189 (pure host integer math)
"""
from __future__ import annotations

import math

import numpy as np
import torch

from . import _lib


def apply_low_pass_filter(
    tensor: torch.Tensor,
    filter_type: str,
    # Gaussian Blur Params
    blur_sigma: float,
    blur_kernel_size: float,  # Can be float (relative) or int (absolute)
    # Down/Up Sampling Params
    resize_factor: float,
):
    """Low-pass filter a [B, C, H, W] or [B, C, F, H, W] CUDA tensor (lp_utils.py:8-60).

    Early exits return the *same tensor object* like the reference (lp_utils.py:23-28).  A 5-D input is treated as
    ``B*C*F`` independent H x W planes -- the reference's ``view(B*K, C, H, W)`` without a permute is legal only
    because both filters act per plane (lp_utils.py:31-35).
    """
    if filter_type == "none":
        return tensor
    if filter_type == "down_up" and resize_factor == 1.0:
        return tensor
    if filter_type == "gaussian_blur" and blur_sigma == 0:
        return tensor

    if tensor.ndim not in (4, 5):
        # the reference fails on the tuple unpack of tensor.shape
        raise ValueError(f"not enough values to unpack (expected 4 or 5 dims, got {tensor.ndim})")
    if filter_type not in ("gaussian_blur", "down_up"):
        return tensor  # the reference applies no filter for unknown types (lp_utils.py:40-58)

    _lib.require_cuda(tensor)
    H, W = int(tensor.shape[-2]), int(tensor.shape[-1])
    src = tensor if tensor.is_contiguous() else tensor.contiguous()
    planes = src.numel() // (H * W) if H * W else 0
    dt = _lib.dtype_code(src.dtype)
    out = torch.empty_like(src)
    if planes == 0:
        return out.view(tensor.shape)  # empty batch: nothing to launch
    L = _lib.lib()
    stream = _lib.stream_ptr(src.device)
    with torch.cuda.device(src.device):
        if filter_type == "gaussian_blur":
            if isinstance(blur_kernel_size, float):
                kernel_val = max(int(blur_kernel_size * H), 1)
            else:
                kernel_val = int(blur_kernel_size)
            if kernel_val % 2 == 0:
                kernel_val += 1
            if blur_sigma <= 0:
                raise ValueError(f"If sigma is a single number, it must be positive. Got {blur_sigma}")
            _lib.check(L.alg_lowpass_gaussian(src.data_ptr(), out.data_ptr(), planes, H, W, kernel_val,
                                              float(blur_sigma), dt, stream))
        else:
            h1 = max(1, int(round(H * resize_factor)))
            w1 = max(1, int(round(W * resize_factor)))
            _lib.check(L.alg_lowpass_down_up(src.data_ptr(), out.data_ptr(), planes, H, W, h1, w1, dt, stream))
    return out.view(tensor.shape)


def get_lp_strength(
    step_index: int,
    total_steps: int,
    lp_strength_schedule_type: str,
    # Interval params
    schedule_interval_start_time: float,
    schedule_interval_end_time: float,
    # Linear params
    schedule_linear_start_weight: float,
    schedule_linear_end_weight: float,
    schedule_linear_end_time: float,
    # Exponential params
    schedule_exp_decay_rate: float,
) -> float:
    """Low-pass strength multiplier of one denoising step (lp_utils.py:63-111); pure host math."""
    step_norm = step_index / max(total_steps - 1, 1)

    if lp_strength_schedule_type == "linear":
        if schedule_linear_end_time <= 0:
            return schedule_linear_start_weight
        if step_norm >= schedule_linear_end_time:
            return schedule_linear_end_weight
        progress = step_norm / schedule_linear_end_time
        return schedule_linear_start_weight * (1 - progress) + schedule_linear_end_weight * progress
    if lp_strength_schedule_type == "interval":
        return 1.0 if schedule_interval_start_time <= step_norm <= schedule_interval_end_time else 0.0
    if lp_strength_schedule_type == "exponential":
        decay_rate = schedule_exp_decay_rate
        if decay_rate < 0:
            print(f"Warning: Negative exponential_decay_rate ({decay_rate}) is unusual. Using abs value.")
            decay_rate = abs(decay_rate)
        return math.exp(-decay_rate * step_norm)
    if lp_strength_schedule_type == "none":
        return 1.0
    print(f"Warning: Unknown lp_strength_schedule_type '{lp_strength_schedule_type}'. Using constant strength 1.0.")
    return 1.0


def modulate_lp_params(lp_blur_sigma, lp_blur_kernel_size, lp_resize_factor, lp_strength, schedule_blur_kernel_size):
    """Strength -> filter parameters (wan:863-867, cog:1034-1040, hy:1144-1151).  ``k * s`` stays a float (quirk q8)."""
    sigma = lp_blur_sigma * lp_strength
    ksize = lp_blur_kernel_size * lp_strength if schedule_blur_kernel_size else lp_blur_kernel_size
    factor = 1.0 - (1.0 - lp_resize_factor) * lp_strength
    return sigma, ksize, factor


# ---- HunyuanVideo resolution bucketing (lp_utils.py:113-189) -------------------------------------------
def _generate_crop_size_list(base_size=256, patch_size=32, max_ratio=4.0):
    num_patches = round((base_size / patch_size) ** 2)
    assert max_ratio >= 1.0
    sizes, wp, hp = [], num_patches, 1
    while wp > 0:
        if max(wp, hp) / min(wp, hp) <= max_ratio:
            sizes.append((wp * patch_size, hp * patch_size))
        if (hp + 1) * wp <= num_patches:
            hp += 1
        else:
            wp -= 1
    return sizes


def _get_closest_ratio(height: float, width: float, ratios, buckets):
    aspect_ratio = float(height) / float(width)
    diff = ratios - aspect_ratio
    if aspect_ratio >= 1:
        cand = [(i, x) for i, x in enumerate(diff) if x <= 0]
    else:
        cand = [(i, x) for i, x in enumerate(diff) if x > 0]
    idx = min(cand, key=lambda pair: abs(pair[1]))[0]
    return buckets[idx], ratios[idx]


def get_hunyuan_video_size(i2v_resolution, input_image):
    """(height, width) bucket for HunyuanVideo-I2V from the resolution tag and the image aspect (lp_utils.py:165-189)."""
    if i2v_resolution == "720p":
        base = 960
    elif i2v_resolution == "540p":
        base = 720
    elif i2v_resolution == "360p":
        base = 480
    else:  # the reference leaves the variable unbound
        raise UnboundLocalError("bucket_hw_base_size")
    origin_size = input_image.size
    sizes = _generate_crop_size_list(base, 32)
    ratios = np.array([round(float(h) / float(w), 5) for h, w in sizes])
    closest, _ = _get_closest_ratio(origin_size[1], origin_size[0], ratios, sizes)
    return closest[0], closest[1]
