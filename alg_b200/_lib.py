"""ctypes binding of ``libalg_b200.so`` (the C ABI declared in ``include/alg_b200.h``).

There is no CPU fallback: if the shared library is missing the import of any
compute entry point raises, and every call checks the status code and raises
``RuntimeError`` with ``alg_last_error()``.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libalg_b200.so")
CSRC = os.path.join(_HERE, "csrc")

ALG_F32, ALG_BF16, ALG_F16 = 0, 1, 2
ABI_VERSION = 5
NORM_NONE, NORM_RMS, NORM_LAYER = 0, 1, 2
EW_ADD, EW_SILU, EW_COPY, EW_GELU_TANH = 0, 1, 2, 3
EPI_NONE, EPI_GELU_TANH, EPI_GATE_RESIDUAL, EPI_RESIDUAL, EPI_GELU_ERF, EPI_SILU = range(6)


class UniPCStep(C.Structure):
    _fields_ = [
        ("n_pass", C.c_int32), ("cfg_fp32", C.c_int32), ("guidance", C.c_float), ("sigma_t", C.c_float),
        ("use_corrector", C.c_int32), ("order_c", C.c_int32), ("c_ratio", C.c_float), ("c_a", C.c_float),
        ("c_b", C.c_float), ("c_rk_inv", C.c_float), ("c_rho0", C.c_float), ("c_rho_last", C.c_float),
        ("order_p", C.c_int32), ("p_ratio", C.c_float), ("p_a", C.c_float), ("p_b", C.c_float),
        ("p_rk_inv", C.c_float), ("p_rho0", C.c_float),
    ]


class DpmStep(C.Structure):
    _fields_ = [
        ("n_pass", C.c_int32), ("second_order", C.c_int32), ("guidance", C.c_float), ("sqrt_alpha_t", C.c_float),
        ("sqrt_beta_t", C.c_float), ("m0", C.c_float), ("m1", C.c_float), ("m2", C.c_float), ("m3", C.c_float),
        ("m_noise", C.c_float),
    ]


class Im2col(C.Structure):
    _fields_ = [
        ("x", C.c_void_p), ("cols", C.c_void_p), ("T", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("C", C.c_int32),
        ("kt", C.c_int32), ("kh", C.c_int32), ("kw", C.c_int32), ("st", C.c_int32), ("sh", C.c_int32), ("sw", C.c_int32),
        ("pad_t", C.c_int32), ("pad_top", C.c_int32), ("pad_left", C.c_int32), ("To", C.c_int32), ("Ho", C.c_int32),
        ("Wo", C.c_int32), ("ld", C.c_int64),
    ]


class GroupNorm(C.Structure):
    _fields_ = [
        ("x", C.c_void_p), ("y", C.c_void_p), ("weight", C.c_void_p), ("bias", C.c_void_p), ("stats", C.c_void_p),
        ("rows", C.c_int64), ("C", C.c_int32), ("groups", C.c_int32), ("eps", C.c_float), ("silu", C.c_int32),
    ]


class Gemm(C.Structure):
    _fields_ = [
        ("A", C.c_void_p), ("B", C.c_void_p), ("D", C.c_void_p), ("bias", C.c_void_p), ("R", C.c_void_p),
        ("gate", C.c_void_p), ("M", C.c_int64), ("N", C.c_int64), ("K", C.c_int64), ("lda", C.c_int64),
        ("ldb", C.c_int64), ("ldd", C.c_int64), ("rows_per_batch", C.c_int64), ("gate_ld", C.c_int64),
        ("epilogue", C.c_int32), ("bias_per_row", C.c_int32), ("out_f32", C.c_int32),
        ("gate_dtype", C.c_int32), ("gate_round", C.c_int32), ("gate_split_row", C.c_int64), ("gate_alt", C.c_void_p),
        ("a_k_period", C.c_int64), ("a_tap_kblocks", C.c_int32), ("a_n_taps", C.c_int32), ("a_tap_offsets", C.POINTER(C.c_int32)),
        ("a_rows", C.c_int64), ("bias_f32", C.c_void_p), ("residual_f32", C.c_void_p),
    ]


class LayerNorm(C.Structure):
    _fields_ = [
        ("x", C.c_void_p), ("out", C.c_void_p), ("rows", C.c_int64), ("d", C.c_int32), ("eps", C.c_float),
        ("weight", C.c_void_p), ("bias", C.c_void_p), ("affine_dtype", C.c_int32), ("mod_dtype", C.c_int32),
        ("scale", C.c_void_p), ("shift", C.c_void_p), ("scale_alt", C.c_void_p), ("shift_alt", C.c_void_p),
        ("rows_per_batch", C.c_int64), ("mod_batch_stride", C.c_int64), ("split_row", C.c_int64),
        ("chain_bf16", C.c_int32),
    ]


class HeadNormRope(C.Structure):
    _fields_ = [
        ("x", C.c_void_p), ("rows", C.c_int64), ("ld", C.c_int64), ("heads", C.c_int32), ("head_dim", C.c_int32),
        ("norm_kind", C.c_int32), ("eps", C.c_float), ("weight", C.c_void_p), ("bias", C.c_void_p),
        ("cos", C.c_void_p), ("sin", C.c_void_p), ("rows_per_batch", C.c_int64), ("rope_row0", C.c_int64),
        ("rope_rows", C.c_int64),
    ]


class PatchSrc(C.Structure):
    _fields_ = [
        ("ptr", C.c_void_p), ("ptr_t0", C.c_void_p), ("dtype", C.c_int32), ("channels", C.c_int32),
        ("sc", C.c_int64), ("st", C.c_int64), ("sy", C.c_int64), ("sc_t0", C.c_int64),
    ]


class Attention(C.Structure):
    _fields_ = [
        ("Q", C.c_void_p), ("K", C.c_void_p), ("Vt", C.c_void_p), ("O", C.c_void_p), ("batch", C.c_int32),
        ("heads", C.c_int32), ("head_dim", C.c_int32), ("n_q", C.c_int64), ("n_kv", C.c_int64),
        ("q_bs", C.c_int64), ("q_rs", C.c_int64), ("k_bs", C.c_int64), ("k_rs", C.c_int64), ("v_bs", C.c_int64),
        ("v_rs", C.c_int64), ("o_bs", C.c_int64), ("o_rs", C.c_int64), ("scale", C.c_float),
        ("accumulate", C.c_int32),
    ]


class SmallAttention(C.Structure):
    _fields_ = [
        ("q", C.c_void_p), ("k", C.c_void_p), ("v", C.c_void_p), ("out", C.c_void_p), ("dtype", C.c_int32),
        ("batch", C.c_int32), ("heads", C.c_int32), ("head_dim", C.c_int32), ("n_q", C.c_int64), ("n_kv", C.c_int64),
        ("q_bs", C.c_int64), ("q_rs", C.c_int64), ("k_bs", C.c_int64), ("k_rs", C.c_int64), ("v_bs", C.c_int64),
        ("v_rs", C.c_int64), ("o_bs", C.c_int64), ("o_rs", C.c_int64), ("scale", C.c_float), ("rel_bias", C.c_void_p),
        ("kv_valid", C.c_void_p), ("causal", C.c_int32), ("kv_group", C.c_int32), ("key_mask", C.c_void_p),
    ]


class Im2colF32(C.Structure):
    _fields_ = [
        ("x", C.c_void_p), ("cols", C.c_void_p), ("T", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("C", C.c_int32),
        ("kt", C.c_int32), ("kh", C.c_int32), ("kw", C.c_int32), ("st", C.c_int32), ("sh", C.c_int32), ("sw", C.c_int32),
        ("pad_t", C.c_int32), ("pad_top", C.c_int32), ("pad_left", C.c_int32), ("To", C.c_int32), ("Ho", C.c_int32),
        ("Wo", C.c_int32), ("to0", C.c_int32), ("up", C.c_int32), ("t_min", C.c_int32), ("replicate", C.c_int32),
        ("tdup", C.c_int32), ("ld", C.c_int64),
    ]


class WanConfig(C.Structure):
    _fields_ = [
        ("num_heads", C.c_int32), ("head_dim", C.c_int32), ("in_channels", C.c_int32), ("out_channels", C.c_int32),
        ("text_dim", C.c_int32), ("freq_dim", C.c_int32), ("ffn_dim", C.c_int32), ("num_layers", C.c_int32),
        ("image_dim", C.c_int32), ("text_len", C.c_int32), ("patch_t", C.c_int32), ("patch_h", C.c_int32),
        ("patch_w", C.c_int32), ("rope_max_seq_len", C.c_int32), ("eps", C.c_float),
    ]


# name -> (restype, argtypes); every symbol include/alg_b200.h declares
SIGNATURES = {
    "alg_abi_version": (C.c_int, []),
    "alg_last_error": (C.c_char_p, []),
    "alg_check_device": (C.c_int, []),
    "alg_launch_count": (C.c_int64, []),
    "alg_lowpass_down_up": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "alg_lowpass_gaussian": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_void_p]),
    "alg_gaussian_kernel1d": (C.c_int, [C.c_int, C.c_double, C.c_int, C.POINTER(C.c_float)]),
    "alg_cfg_unipc_step": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(UniPCStep), C.c_void_p]),
    "alg_cfg_ddim_step": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_void_p]),
    "alg_cfg_dpm_step": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(DpmStep), C.c_void_p]),
    "alg_cfg_euler_step": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int64, C.c_int, C.c_float, C.c_float, C.c_void_p]),
    "alg_gemm_bf16": (C.c_int, [C.POINTER(Gemm), C.c_void_p]),
    "alg_attention_bf16": (C.c_int, [C.POINTER(Attention), C.c_void_p]),
    "alg_layer_norm": (C.c_int, [C.POINTER(LayerNorm), C.c_void_p]),
    "alg_im2col_bf16": (C.c_int, [C.POINTER(Im2col), C.c_void_p]),
    "alg_wan_rms_norm_rope": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "alg_group_norm_bf16": (C.c_int, [C.POINTER(GroupNorm), C.c_void_p]),
    "alg_head_norm_rope": (C.c_int, [C.POINTER(HeadNormRope), C.c_void_p]),
    "alg_upsample_nearest_bf16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "alg_spatial_norm_apply_bf16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                               C.c_int, C.c_int, C.c_void_p]),
    "alg_patch_gather": (C.c_int, [C.POINTER(PatchSrc), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int64, C.c_void_p]),
    "alg_unpatchify": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int, C.c_void_p]),
    "alg_timestep_embedding": (C.c_int, [C.c_float, C.c_int, C.c_void_p, C.c_int, C.c_void_p]),
    "alg_elementwise_bf16": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "alg_mean_rows_bf16": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_int64, C.c_void_p, C.c_void_p]),
    "alg_copy_rows_bf16": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_void_p]),
    "alg_t5_rms_norm_bf16": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_float, C.c_void_p, C.c_void_p]),
    "alg_gather_rows_bf16": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p]),
    "alg_small_attention": (C.c_int, [C.POINTER(SmallAttention), C.c_void_p]),
    "alg_layer_norm_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]),
    "alg_bias_act_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p]),
    "alg_split3_bf16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p]),
    "alg_rms_norm_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_float, C.c_void_p, C.c_void_p]),
    "alg_rope_half_f32": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p]),
    "alg_swiglu_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p]),
    "alg_im2col_split3_f32": (C.c_int, [C.POINTER(Im2colF32), C.c_void_p]),
    "alg_rms_norm_cl_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "alg_softmax_rows_f32": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_int64, C.c_float, C.c_int, C.c_void_p]),
    "alg_group_norm_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_void_p, C.c_int,
                                     C.c_void_p, C.c_void_p]),
    "alg_nchw_to_cl_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_int, C.c_void_p]),
    "alg_cl_to_nchw_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_int, C.c_float, C.c_float, C.c_void_p]),
    "alg_norm_split_pad_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                         C.c_int, C.c_void_p]),
    "alg_pad_copy_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "alg_pad_frames_bf16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int, C.c_void_p]),
    "alg_axpby_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_float, C.c_float, C.c_void_p]),
    "alg_replicate_border_bf16": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "alg_mul_bf16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "alg_patchify_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "alg_clip_embed_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "alg_wan_create": (C.c_int, [C.POINTER(WanConfig), C.POINTER(C.c_void_p)]),
    "alg_wan_destroy": (None, [C.c_void_p]),
    "alg_wan_set_weight": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int64, C.c_int]),
    "alg_wan_weights_complete": (C.c_int, [C.c_void_p]),
    "alg_wan_set_debug_buffer": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "alg_wan_context_cache": (C.c_int, [C.c_void_p, C.c_int]),
    "alg_wan_profile": (C.c_int, [C.c_void_p, C.c_int]),
    "alg_wan_profile_read": (C.c_int, [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_int32), C.c_int]),
    "alg_wan_workspace_bytes": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_size_t)]),
    "alg_wan_forward": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
}

_lib = None


def build(verbose: bool = False) -> str:
    """Compile the library in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    cmd = ["make", "-C", CSRC, "-j", str(os.cpu_count() or 4)]
    r = subprocess.run(cmd, capture_output=not verbose, text=True)
    if r.returncode != 0:
        raise RuntimeError("building libalg_b200.so failed:\n" + (r.stdout or "") + (r.stderr or ""))
    return LIB_PATH


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(alg_b200 has no CPU or PyTorch fallback)"
            )
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        if l.alg_abi_version() != ABI_VERSION:
            raise RuntimeError("libalg_b200.so ABI version mismatch")
        _lib = l
    return _lib


def check(status: int) -> None:
    if status != 0:
        raise RuntimeError(lib().alg_last_error().decode() or f"alg_b200 call failed with status {status}")


def launch_count() -> int:
    return int(lib().alg_launch_count())


def dtype_code(dt) -> int:
    import torch

    try:
        return {torch.float32: ALG_F32, torch.bfloat16: ALG_BF16, torch.float16: ALG_F16}[dt]
    except KeyError:
        raise TypeError(f"alg_b200 supports float32 / bfloat16 / float16 tensors, got {dt}") from None


def stream_ptr(device=None) -> int:
    import torch

    return torch.cuda.current_stream(device).cuda_stream


def require_cuda(*tensors) -> None:
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("alg_b200 operators need CUDA tensors (there is no CPU fallback)")
