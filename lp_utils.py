"""Drop-in for the reference's ``lp_utils`` module name (imported by run.py:19 and the three pipelines).

The implementations live in ``alg_b200.lowpass`` (host mirrors of lp_utils.py:8-189); the filters run as sm_100a
CUDA kernels from ``libalg_b200.so``.
"""
from alg_b200.lowpass import (  # noqa: F401
    _generate_crop_size_list,
    _get_closest_ratio,
    apply_low_pass_filter,
    get_hunyuan_video_size,
    get_lp_strength,
    modulate_lp_params,
)
