"""ctypes binding of libistvt_b200.so (the C ABI declared in include/istvt_b200.h).

There is no fallback: if the shared library is missing, `lib()` raises and tells the caller how to
build it; if a kernel returns a non-zero code, `check()` raises with the library's own error text.
"""
from __future__ import annotations

import ctypes
import os
import threading
from ctypes import c_char_p, c_float, c_int, c_int64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libistvt_b200.so")
ABI_VERSION = 1

_P, _I, _L, _F = c_void_p, c_int, c_int64, c_float

# name -> argtypes, exactly mirroring include/istvt_b200.h (restype is int unless listed in _RESTYPES)
SIGNATURES = {
    "istvt_abi_version": [],
    "istvt_error_string": [_I],
    "istvt_launch_count": [],
    "istvt_layernorm_fwd": [_P, _I, _P, _P, _P, _I, _L, _I, _F, _P],
    "istvt_layernorm_diff_fwd": [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _F, _P],
    "istvt_layernorm_fwd_ld": [_P, _I, _L, _P, _P, _P, _I, _L, _L, _I, _F, _P],
    "istvt_layernorm_diff_fwd_ld": [_P, _L, _P, _P, _P, _P, _I, _L, _I, _I, _I, _I, _F, _P],
    "istvt_gemm_fwd": [_P, _L, _P, _L, _P, _L, _I, _L, _I, _I, _P, _P, _L, _I, _P],
    "istvt_gemm_f32_fwd": [_P, _L, _P, _L, _P, _L, _L, _I, _I, _P, _P, _L, _I, _P],
    "istvt_gemm_act_dual_fwd": [_P, _L, _P, _L, _P, _L, _P, _L, _L, _I, _I, _P, _I, _P],
    "istvt_gemm_dgelu_fwd": [_P, _L, _P, _L, _P, _L, _P, _L, _L, _I, _I, _P],
    "istvt_conv3x3_fwd": [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P],
    "istvt_conv3x3_pair_pack": [_P, _P, _I, _P],
    "istvt_conv3x3_pair_fwd": [_P, _P, _P, _P, _I, _I, _I, _I, _P],
    "istvt_conv_stem_fwd": [_P, _P, _P, _P, _I, _I, _I, _I, _I, _P],
    "istvt_conv_stem_u8_fwd": [_P, _P, _P, _P, _I, _I, _I, _I, _I, _P],
    "istvt_dwconv3x3_fwd": [_P, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "istvt_subsample2_fwd": [_P, _P, _I, _I, _I, _I, _I, _P],
    "istvt_sepconv_fused_fwd": [_P, _P, _P, _L, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P],
    "istvt_pool_add_fwd": [_P, _P, _P, _I, _I, _I, _I, _I, _P],
    "istvt_pool_add_tokens_fwd": [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "istvt_token_fill_fwd": [_P, _P, _P, _P, _I, _I, _I, _I, _P],
    "istvt_attn_temporal_fwd": [_P, _P, _P, _P, _I, _I, _I, _I, _I, _F, _P],
    "istvt_attn_spatial_fwd": [_P, _P, _P, _I, _I, _I, _I, _F, _P],
    "istvt_attn_joint_fwd": [_P, _P, _I, _I, _I, _I, _F, _P],
    "istvt_token_build_fwd": [_P, _I, _P, _P, _P, _I, _I, _I, _I, _P],
    "istvt_mean_rows_fwd": [_P, _P, _I, _I, _I, _P],
    "istvt_add_fwd": [_P, _P, _P, _I, _L, _P],
    "istvt_pool_linear_fwd": [_P, _I, _P, _P, _P, _I, _I, _I, _I, _I, _P],
    "istvt_head_fwd": [_P, _L, _P, _P, _P, _P, _P, _P, _P, _I, _I, _F, _P],
    # training step
    "istvt_attn_spatial_fwd_lse": [_P, _P, _P, _I, _I, _I, _F, _P],
    "istvt_attn_spatial_bwd": [_P, _P, _P, _P, _P, _P, _I, _I, _I, _F, _P],
    "istvt_attn_temporal_bwd": [_P, _P, _P, _P, _P, _I, _I, _I, _I, _F, _P],
    "istvt_layernorm_bwd": [_P, _P, _I, _I, _P, _I, _P, _P, _P, _P, _P, _P, _L, _I, _F, _P],
    "istvt_layernorm_bwd_ld": [_P, _P, _L, _I, _I, _P, _I, _L, _P, _P, _L, _P, _L, _P, _L, _P, _P, _P, _L, _I, _F, _P],
    "istvt_gelu_bwd_colsum": [_P, _P, _P, _P, _L, _I, _P],
    "istvt_cast_f32_bf16_rows": [_P, _L, _P, _L, _L, _I, _P],
    "istvt_colsum_ld": [_P, _L, _P, _L, _I, _P],
    "istvt_gelu_fwd": [_P, _P, _L, _P],
    "istvt_gelu_bwd": [_P, _P, _P, _L, _P],
    "istvt_cast_f32_bf16": [_P, _P, _L, _P],
    "istvt_transpose_colsum": [_P, _P, _P, _L, _I, _L, _P],
    "istvt_gemm_splitk_accum": [_P, _L, _P, _L, _P, _L, _L, _I, _L, _P],
    "istvt_head_bwd": [_P, _L, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _F, _P],
    "istvt_token_bwd": [_P, _P, _P, _P, _I, _I, _I, _I, _P],
    "istvt_adamw_step": [_P, _P, _P, _P, _L, _F, _F, _F, _F, _F, _I, _F, _P],
    "istvt_conv_stem_raw_fwd": [_P, _P, _P, _P, _I, _I, _I, _I, _I, _P],
    "istvt_bn_stats_fwd": [_P, _P, _P, _L, _I, _P],
    "istvt_bn_finalize_fwd": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _L, _I, _F, _F, _P],
    "istvt_bn_apply_fwd": [_P, _P, _P, _P, _L, _I, _I, _P],
    "istvt_bn_bwd": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _L, _I, _I, _P],
    "istvt_pool_add_idx_fwd": [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P],
    "istvt_pool_bwd": [_P, _P, _P, _I, _I, _I, _I, _P],
    "istvt_token_grad_gather": [_P, _P, _I, _I, _I, _I, _P],
    "istvt_dwconv3x3_wgrad": [_P, _P, _P, _I, _I, _I, _I, _I, _P],
    "istvt_block_input_grad": [_P, _P, _P, _P, _I, _I, _I, _I, _I, _P],
    "istvt_im2col_t": [_P, _P, _I, _I, _I, _I, _L, _P],
    "istvt_im2col_t_stem": [_P, _P, _I, _I, _I, _L, _P],
    # relevance pass
    "istvt_attn_spatial_bwd_cam": [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _F, _P],
    "istvt_attn_temporal_bwd_cam": [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _F, _P],
    "istvt_rollout_row": [_P, _P, _L, _I, _P],
    "istvt_gather_rows": [_P, _P, _L, _L, _L, _L, _L, _P],
    "istvt_gemm_wgrad_accum": [_P, _L, _P, _L, _P, _L, _L, _I, _I, _P],
    "istvt_colsum": [_P, _P, _L, _I, _P],
    "istvt_gemm_rowstats_fwd": [_P, _L, _P, _L, _P, _L, _L, _I, _I, _P, _P, _P],
    "istvt_ln_stats_finalize": [_P, _P, _L, _I, _F, _P],
    "istvt_gemm_lnfold_fwd": [_P, _L, _P, _L, _P, _L, _L, _I, _I, _P, _P, _P, _P],
}
_RESTYPES = {"istvt_error_string": c_char_p, "istvt_launch_count": c_int64}

_lock = threading.Lock()
_lib = None


def lib() -> ctypes.CDLL:
    """Load (once) and return the shared library; raise loudly if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: the CUDA extension is not built. Run "
                f"`python {os.path.join(_HERE, 'build.py')}` (or __graft_entry__.build()). "
                "There is no CPU / eager fallback for this path."
            )
        handle = ctypes.CDLL(LIB_PATH)
        for name, argtypes in SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError if the .so is stale
            fn.argtypes = argtypes
            fn.restype = _RESTYPES.get(name, c_int)
        got = handle.istvt_abi_version()
        if got != ABI_VERSION:
            raise RuntimeError(f"libistvt_b200.so ABI {got} != expected {ABI_VERSION}; rebuild it")
        _lib = handle
    return _lib


def check(code: int, what: str) -> None:
    if code != 0:
        msg = lib().istvt_error_string(code)
        raise RuntimeError(f"{what} failed with code {code}: {msg.decode() if msg else '?'}")


def launch_count() -> int:
    return int(lib().istvt_launch_count())
