"""alg_b200 -- Blackwell (sm_100a) implementation of the ALG image-to-video denoise loop.

Only what the hot path needs lives here: ``csrc/`` (CUDA kernels + the C ABI of ``include/alg_b200.h``),
the ctypes binding, and host-side mirrors of the reference's interfaces (``lowpass`` <- lp_utils.py,
``schedulers`` <- diffusers schedulers as the pipelines drive them, ``wan`` <- WanTransformer3DModel).
"""
__version__ = "0.1.0"

from . import _lib  # noqa: F401
