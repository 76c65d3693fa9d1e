"""ctypes binding of libvnqa_b200.so (C ABI in include/vnqa_b200.h).

There is no fallback: if the shared library has not been built (``python -c "import __graft_entry__ as g;
g.build()"`` or ``make -C videonavqa_b200/csrc``) importing a symbol raises, and every compute entry point
needs a CUDA sm_100 device.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_float, c_int32, c_int64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
# VNQA_LIB_PATH selects another build of the same library (the -DVNQA_DEBUG_MBAR bring-up build)
LIB_PATH = os.environ.get("VNQA_LIB_PATH") or os.path.join(_HERE, "libvnqa_b200.so")


class VnqaError(RuntimeError):
    pass


class ConvArgs(ctypes.Structure):
    """Mirror of vnqa_conv_args (include/vnqa_b200.h)."""
    _fields_ = [
        ("w_packed", c_void_p), ("x", c_void_p),
        ("taps", c_int32), ("cin", c_int32), ("cout", c_int32), ("rows_pad", c_int32), ("n_img", c_int32),
        ("flat", c_int32), ("x_ld", c_int32), ("splits", c_int32), ("epi", c_int32), ("no_round", c_int32),
        ("out", c_void_p), ("bias", c_void_p), ("scale", c_void_p), ("shift", c_void_p),
        ("img_frame", c_void_p), ("stat_sum", c_void_p), ("stat_sq", c_void_p),
        ("film", c_void_p), ("film_row", c_void_p), ("film_ld", c_int32), ("alpha_off", c_int32),
        ("res", c_void_p), ("z_out", c_void_p), ("add", c_void_p), ("dbias", c_void_p),
        ("act_dtype", c_int32), ("out_f32", c_int32), ("grad_unscale", c_float),
        ("img_h", c_int32), ("img_w", c_int32), ("pool", c_int32),
    ]


EPI_BIAS_RELU, EPI_BIAS_RELU_AFFINE, EPI_BIAS_RELU_STATS, EPI_FILM, EPI_PARTIAL, EPI_DGRAD_RES = 0, 1, 2, 3, 4, 5
EPI_PLAIN = 7
EPI_BIAS, EPI_BIAS_AFFINE_RELU = 8, 9      # frozen stem (models/obj_detector.py)
ACT_TF32, ACT_FP16 = 0, 1

P, I32, I64, F32 = c_void_p, c_int32, c_int64, c_float

# name -> argtypes (every function returns int unless listed in _RESTYPES)
SIGNATURES = {
    "vnqa_abi_version": [],
    "vnqa_last_error": [],
    "vnqa_device_check": [],
    "vnqa_sm_count": [],
    "vnqa_debug_install": [P],
    "vnqa_debug_build": [],
    "vnqa_frames_to_nhwc": [P, P, P, I32, I32, I32, I32, P],
    "vnqa_frames_blocked_to_nhwc": [P, P, P, I32, I32, I32, I64, I64, I32, P],
    "vnqa_frames_u8_to_video": [P, P, P, I32, I32, I32, I32, P],
    "vnqa_pack_conv_weight": [P, P, I32, I32, I32, I32, I32, I32, P],
    "vnqa_to_half": [P, P, I64, F32, P],
    "vnqa_permute_fc_weight": [P, P, I32, I32, I32, I32, P],
    "vnqa_conv_gemm": [POINTER(ConvArgs), P],
    "vnqa_nchw_to_nhwc_affine": [P, P, P, P, I32, I32, I32, I32, P],
    "vnqa_nhwc_to_nchw": [P, P, I32, I32, I32, I32, P],
    "vnqa_conv_wgrad_ws_bytes": [I32, I32, I32, POINTER(c_int64)],
    "vnqa_conv_wgrad": [P, P, P, P, I32, I32, I32, I32, I32, I32, F32, P],
    "vnqa_sgemm": [P, P, P, P, I32, I32, I32, I64, I64, I64, I64, I64, I32, I32, I32, P],
    "vnqa_colsum": [P, I32, I32, I64, P, I32, P],
    "vnqa_bn_finalize_train": [P, P, P, I32, I32, P, P, P, P, F32, F32, P, P, P, P, P],
    "vnqa_bn_eval_affine": [P, P, P, P, F32, I32, P, P, P],
    "vnqa_bn_apply": [P, P, P, P, P, I32, I32, I32, P],
    "vnqa_bn_backward": [P, P, P, P, I32, P, P, P, I32, I32, P, P, P, P, P, P, I32, F32, P],
    "vnqa_film_backward": [P, P, P, P, I32, I32, I32, I32, P, P, P, I32, I32, F32, P],
    "vnqa_splitk_reduce": [P, I32, I32, I32, P, P, P, P],
    "vnqa_gather_rows": [P, P, I32, I32, P, I32, P],
    "vnqa_gp_maxpool": [P, P, I32, I32, I32, P, P, P],
    "vnqa_gp_maxpool_backward": [P, P, P, I32, I32, P, P],
    "vnqa_embed_gates": [P, P, P, P, P, P, I32, I32, I32, I32, P, P],
    "vnqa_lstm_chain_forward": [P, P, P, P, P, I32, I32, I32, I32, P, P, P, P, P, P],
    "vnqa_lstm_chain_backward": [P, P, P, P, P, P, I32, I32, I32, I32, P, P, P, P],
    "vnqa_embed_backward": [P, P, P, I32, I32, I32, P, P],
    "vnqa_embed_gather": [P, P, P, I32, I32, I32, P, P],
    "vnqa_tmh_hops_forward": [P, P, P, P, I32, I32, I32, I32, I32, I32, P, P, P, P, P, P, P, P, P],
    "vnqa_tmh_hops_backward": [P, P, P, P, P, I32, I32, I32, I32, I32, I32, P, P, P, P, P, P, P, P, P, P, P, P, P],
    "vnqa_ln_slice_forward": [P, I32, I32, P, P, P, I32, I32, I32, P, P, P],
    "vnqa_ln_slice_backward": [P, I32, I32, P, P, I32, I32, I32, P, P, P, P, P, P],
    "vnqa_nccl_load": [c_char_p],
    "vnqa_nccl_version": [],
    "vnqa_nccl_unique_id": [P],
    "vnqa_nccl_init": [P, I32, I32],
    "vnqa_nccl_rank": [],
    "vnqa_nccl_nranks": [],
    "vnqa_nccl_stream": [],
    "vnqa_nccl_allreduce_sum_f32": [P, I64, P],
    "vnqa_nccl_allgather_f64": [P, P, I64, P],
    "vnqa_nccl_broadcast_f32": [P, I64, I32, P],
    "vnqa_nccl_join": [P],
    "vnqa_nccl_destroy": [],
    "vnqa_nccl_abort": [],
    "vnqa_opt_num_partials": [],
    "vnqa_sumsq_partials": [P, I64, P, P],
    "vnqa_clip_adam": [P, P, P, P, I64, P, I32, F32, F32, c_double, c_double, c_double, c_double, I64, P, P],
    "vnqa_clip_adam_dev": [P, P, P, P, I64, P, I32, F32, F32, c_double, c_double, c_double, c_double, P, P, P],
    "vnqa_attn_tail_forward": [P, P, I32, I32, I32, I32, P, P, P, P, P, P, P, P, P, P, P, P, P, P],
    "vnqa_attn_tail_backward": [P, P, I32, I32, I32, I32, P, P, P, P, P, P, P, P, P, P, P, P, P, P, P, P],
}
_RESTYPES = {"vnqa_last_error": c_char_p, "vnqa_nccl_stream": c_void_p, "vnqa_conv_wgrad_ws_bytes": c_int64}

_lib = None


def lib() -> ctypes.CDLL:
    """Load the shared library once; fail loudly when it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise VnqaError(
                f"{LIB_PATH} is missing: build it with `make -C videonavqa_b200/csrc` "
                "(or __graft_entry__.build()); videonavqa_b200 has no CPU or PyTorch fallback")
        handle = ctypes.CDLL(LIB_PATH)
        for name, argtypes in SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError if the library does not export a declared symbol
            fn.argtypes = argtypes
            fn.restype = _RESTYPES.get(name, c_int32)
        if handle.vnqa_abi_version() != 3:
            raise VnqaError("libvnqa_b200.so ABI version mismatch")
        _lib = handle
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().vnqa_last_error()
        raise VnqaError(f"{what} failed ({rc}): {msg.decode() if msg else '?'}")
