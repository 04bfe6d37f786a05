"""ctypes binding of libmauv_b200.so (the C-ABI in include/mauv_b200.h).

There is deliberately no fallback: if the shared library is missing or the device is not
sm_100a, the product path raises. (The CPU oracle under /oracle is test infrastructure only.)
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

_LIB_PATH = Path(__file__).resolve().parent / "lib" / "libmauv_b200.so"
_lib = None

vp, i32, i64, u32, u64, f32 = C.c_void_p, C.c_int, C.c_longlong, C.c_uint32, C.c_uint64, C.c_float

# name -> (restype, argtypes). Must list every symbol include/mauv_b200.h declares
# (tests/test_abi.py checks the two against each other).
SIGNATURES = {
    "mauv_version": (i32, []),
    "mauv_last_error": (C.c_char_p, []),
    "mauv_device_check": (i32, []),
    "mauv_num_sms_c": (i32, []),
    "mauv_set_sample_base": (i32, [vp]),
    "mauv_sample_weights_f16": (i32, [vp, vp, vp, u64, u32, u32, i32, i32, i32, i32, i32, i32, vp, vp]),
    "mauv_sample_vector_f32": (i32, [vp, vp, vp, u64, u32, u32, i32, i32, vp, vp]),
    "mauv_philox_normal_f32": (i32, [u64, u32, u32, i64, vp, vp]),
    "mauv_gemm_m_tiles": (i32, [i64]),
    "mauv_gemm_f16": (i32, [vp, i64, vp, vp, vp, vp, i32, i64, i32, i32, vp]),
    "mauv_gemm_bn_stats_tiles": (i32, [i64]),
    "mauv_gemm_bn_f16": (i32, [vp, vp, vp, vp, vp, vp, i32, i32, i32, i64, i32, i32, vp]),
    "mauv_conv2d_im2col_f16": (i32, [vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, i32, i32, i32, vp]),
    "mauv_stem_im2col_f16": (i32, [vp, i32, i32, i32, i32, i32, i32, i32, i32, i32, vp, vp]),
    "mauv_bn_finalize_ws_bytes": (i64, [i32, i32, i32]),
    "mauv_bn_finalize": (i32, [vp, i32, i32, i32, i64, vp, vp, f32, f32, vp, vp, vp, vp, vp, vp, vp]),
    "mauv_bn_act_blocks": (i32, [i32, i64, i32]),
    "mauv_bn_act_f16": (i32, [vp, vp, vp, vp, vp, i32, i32, i64, i32, vp, vp, vp]),
    "mauv_colsum_f16": (i32, [vp, i32, i64, i32, vp, vp]),
    "mauv_bn_stats_from_gram_ws_bytes": (i64, [i32, i32, i32]),
    "mauv_bn_stats_from_gram": (i32, [vp, i32, vp, i32, vp, i32, i32, i32, i64, vp, vp, f32, f32, vp, vp, vp, vp, vp, vp, vp]),
    "mauv_bn_relu_maxpool_f16": (i32, [vp, vp, i32, i32, i32, i32, i32, vp, vp]),
    "mauv_stem_conv_pool_f16": (i32, [vp, vp, vp, vp, vp, i32, i32, i32, i32, vp]),
    "mauv_avgpool_f16": (i32, [vp, i64, i32, i32, vp, vp]),
    "mauv_nchw_f32_to_nhwc_f16": (i32, [vp, i64, i32, i32, i32, vp, vp]),
    "mauv_nhwc_f16_to_nchw_f32": (i32, [vp, i64, i32, i32, vp, vp]),
    "mauv_sampled_linear_f32": (i32, [vp, i64, i32, vp, vp, vp, vp, vp, vp, u64, u32, u32, i32, i32, i32, i32,
                                      vp, i64, i32, vp]),
    "mauv_tanh_add_f32": (i32, [vp, vp, i64, vp, vp]),
    "mauv_softmax_gate_f32": (i32, [vp, vp, i64, i32, vp, i32, vp]),
    "mauv_mc_reduce": (i32, [vp, i32, i64, i32, i32, f32, vp, vp, vp, vp, vp, vp, vp, vp, vp]),
    "mauv_sample_weights_dgrad_f16": (i32, [vp, vp, vp, u64, u32, u32, i32, i32, i32, i32, i32, vp, vp]),
    "mauv_dilate_f16": (i32, [vp, i64, i32, i32, i32, i32, i32, i32, vp, vp]),
    "mauv_transpose_chunks_f16": (i32, [vp, i64, i32, i32, f32, vp, vp]),
    "mauv_im2col_t_f16": (i32, [vp, i64, i32, i32, i32, i32, i32, i32, i32, i32, i32, vp, vp]),
    "mauv_wgrad_finalize": (i32, [vp, i32, i32, i32, i32, i32, i32, f32, vp, vp, u64, u32, u32, vp, vp, vp]),
    "mauv_sampled_linear_bwd_f32": (i32, [vp, vp, vp, vp, vp, vp, vp, u64, u32, u32, i32, i32, i32,
                                          vp, vp, vp, vp, vp, vp]),
    "mauv_gemm_x3_f16": (i32, [vp, i64, vp, vp, vp, i32, i64, i32, i32, vp]),
    "mauv_conv2d_im2col_x3_f16": (i32, [vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, i32, i32, i32, vp]),
    "mauv_sample_weights_x3_f16": (i32, [vp, vp, vp, u64, u32, u32, i32, i32, i32, i32, i32, i32, i32, f32, vp, vp]),
    "mauv_stem_im2col_x3_f16": (i32, [vp, i32, i32, i32, i32, i32, i32, i32, i32, i32, vp, vp]),
    "mauv_bn_act_x3_f16": (i32, [vp, vp, vp, vp, vp, i32, i32, i64, i32, vp, vp]),
    "mauv_bn_relu_maxpool_x3_f16": (i32, [vp, vp, i32, i32, i32, i32, i32, vp, vp]),
    "mauv_avgpool_x3_f16": (i32, [vp, i64, i32, i32, vp, vp]),
    "mauv_bn_bwd_blocks": (i32, [i32, i64, i32]),
    "mauv_weights_to_dgrad_f16": (i32, [vp, i32, i32, i32, i32, i32, vp, vp]),
    "mauv_bn_bwd_reduce": (i32, [vp, vp, vp, vp, vp, vp, vp, i32, i64, i32, vp, vp, vp]),
    "mauv_bn_bwd_coeffs": (i32, [vp, i32, i64, i32, i32, vp, vp, f32, vp, vp, vp, vp, vp, vp, vp]),
    "mauv_bn_bwd_apply": (i32, [vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, f32, i32, i64, i32, vp, vp, vp, vp, vp, vp]),
    "mauv_maxpool_bwd_f16": (i32, [vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, vp, vp, vp]),
    "mauv_avgpool_bwd_f16": (i32, [vp, i64, i32, i32, f32, vp, vp, vp, vp]),
    "mauv_gemm_bn_cat_f16": (i32, [vp, i32, vp, i32, vp, vp, vp, i32, i32, i64, i32, vp]),
    "mauv_gram_bn_f16": (i32, [vp, vp, vp, vp, i32, i32, i64, i32, vp]),
    "mauv_gemm_bn_xf_f16": (i32, [vp, vp, vp, vp, vp, vp, i32, i32, i64, i32, i32, vp]),
    "mauv_gemm_bn_cat_xf_f16": (i32, [vp, vp, i32, vp, i32, vp, vp, vp, i32, i32, i64, i32, vp]),
    "mauv_sample_weights_scaled_f16": (i32, [vp, vp, vp, u64, u32, u32, i32, i32, i32, vp, i32, i32, vp, vp]),
    "mauv_bn_shift_sum": (i32, [vp, vp, i64, vp, vp]),
    "mauv_subsample_f16": (i32, [vp, i64, i32, i32, i32, i32, vp, vp]),
    "mauv_conv3x3_c64_tiles": (i32, [i32, i32, i32]),
    "mauv_conv3x3_c64_f16": (i32, [vp, vp, vp, vp, vp, i32, i32, i32, i32, vp]),
    "mauv_gemm_wmod_f16": (i32, [vp, vp, i32, vp, i32, i32, i64, i32, i32, vp]),
    "mauv_wgrad_f16": (i32, [vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, i32, i32, i32, i32, vp]),
    "mauv_wgrad_finalize_group": (i32, [vp, i32, i32, i32, i32, i32, i32, i32, i32, f32, vp, vp, vp, u64, u32, u32, i32, vp, vp, vp]),
    "mauv_sampled_linear_bwd_group_f32": (i32, [vp, i64, i32, vp, i64, i32, vp, vp, vp, vp, vp, u64, u32, u32, i32, i32,
                                                i32, i32, i32, vp, i64, i32, i32, vp, vp, vp, vp, vp]),
    "mauv_tanh_bwd_f32": (i32, [vp, vp, i64, vp, vp]),
    "mauv_softmax_gate_bwd_f32": (i32, [vp, vp, vp, i32, i64, i32, vp, vp, vp]),
    "mauv_ce_mean_fwd_bwd_f32": (i32, [vp, vp, i32, i32, i32, vp, vp, vp, vp]),
    "mauv_adam_state_bytes": (i32, []),
    "mauv_adam_step_f32": (i32, [vp, vp, vp, vp, i64, f32, f32, f32, f32, f32, vp, vp]),
    "mauv_membench_fill": (i32, [vp, i64, u32, i32, vp]),
    "mauv_kl_chunk_elems": (i32, []),
    "mauv_kl_ws_bytes": (i64, []),
    "mauv_kl_fwd_bwd": (i32, [vp, vp, i32, i64, f32, f32, f32, vp, vp, vp]),
}


class MauvError(RuntimeError):
    pass


def lib_path() -> Path:
    return _LIB_PATH


def load() -> C.CDLL:
    """Load the shared library and bind every declared symbol (no GPU needed for this)."""
    global _lib
    if _lib is not None:
        return _lib
    if not _LIB_PATH.exists():
        raise MauvError(
            f"{_LIB_PATH} is missing: build it with `python multimodal-auv_b200/build.py` "
            "(or __graft_entry__.build()). mauv_b200 has no CPU / PyTorch fallback.")
    lib = C.CDLL(str(_LIB_PATH))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != 0:
        msg = load().mauv_last_error().decode("utf-8", "replace")
        raise MauvError(f"mauv_b200 error {rc}: {msg}")


_device_ok = False


def require_device() -> C.CDLL:
    """Library + an sm_100a current device, or raise."""
    global _device_ok
    lib = load()
    if not _device_ok:
        import torch
        if not torch.cuda.is_available():
            raise MauvError("mauv_b200 needs a CUDA device (sm_100a); none is visible and there is no CPU path")
        torch.cuda.init()
        check(lib.mauv_device_check())
        _device_ok = True
    return lib
