"""ctypes binding of libmoma_b200.so (the C ABI declared in include/moma_b200.h).

There is no fallback: if the library is missing, or a call returns an error
code, a RuntimeError is raised with ``moma_last_error()``.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import (POINTER, c_char_p, c_float, c_int, c_int32, c_int64, c_size_t, c_void_p)

from . import _build

F32, BF16 = 0, 1
ABI_VERSION = 3

_vp = c_void_p
_i64 = c_int64

# name -> (restype, argtypes); the single source of truth checked by tests against the header
SIGNATURES = {
    "moma_abi_version": (c_int, []),
    "moma_last_error": (c_char_p, []),
    "moma_has_tcgen05": (c_int, []),
    "moma_ema_plan_size": (c_int, [c_int, POINTER(_i64), POINTER(_i64), POINTER(c_size_t)]),
    "moma_ema_plan_fill": (c_int, [c_int, POINTER(_vp), POINTER(_vp), POINTER(_i64), _vp, c_size_t]),
    "moma_ema_multi": (c_int, [_vp, _i64, c_float, c_float, _vp]),
    "moma_sgd_ema_plan_size": (c_int, [c_int, POINTER(_i64), POINTER(_i64), POINTER(c_size_t)]),
    "moma_sgd_ema_plan_fill": (c_int, [c_int, POINTER(_vp), POINTER(_vp), POINTER(_vp), POINTER(_vp), POINTER(_i64), _vp,
                                       c_size_t]),
    "moma_sgd_ema_multi": (c_int, [_vp, _i64, c_float, c_float, c_float, c_int, c_float, c_float, _vp]),
    "moma_l2norm_fwd": (c_int, [_vp, _vp, _i64, _i64, c_float, _vp]),
    "moma_l2norm_bwd": (c_int, [_vp, _vp, _vp, _i64, _i64, c_float, _vp]),
    "moma_enqueue": (c_int, [_vp, _i64, _i64, _vp, _vp, _i64, _i64, _vp, c_int, c_int, c_int, c_float, _vp]),
    "moma_enqueue_strided": (c_int, [_vp, _i64, _i64, _vp, _vp, _i64, _i64, _vp, c_int, c_int, _i64, _i64, _vp]),
    "moma_enqueue_ids": (c_int, [_i64, _i64, _vp, _i64, _vp, _vp]),
    "moma_pointer_advance": (c_int, [_vp, _i64, _i64, _vp]),
    "moma_cast_bf16": (c_int, [_vp, _vp, _i64, _vp]),
    "moma_scale_by_scalar": (c_int, [_vp, _vp, _vp, _i64, _vp]),
    "moma_cls_kd": (c_int, [_vp, _vp, _vp, _i64, _i64, c_float, _vp, _vp, _vp, _vp]),
    "moma_nce_num_splits": (c_int, [_i64, _i64, _i64, c_int]),
    "moma_nce_partial": (c_int, [_vp, _vp, _i64, _i64, _i64, c_float, c_int, c_int, _vp, _vp, _vp, _vp, _vp]),
    "moma_nce_combine": (c_int, [_vp, _vp, _vp, _vp, c_int, _vp, _vp, _i64, _i64, c_float, c_int, c_float,
                                 _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "moma_nce_merge": (c_int, [_vp, _vp, _vp, _vp, c_int, _i64, _i64, _vp, _vp, _vp, _vp, _vp]),
    "moma_nce_merge_packed": (c_int, [_vp, _vp, _vp, _vp, c_int, _i64, _i64, _vp, _vp]),
    "moma_nce_merge_push": (c_int, [_vp, _vp, _vp, _vp, c_int, _i64, _i64, _vp, _i64, _i64, _i64, c_int, c_int, c_int, _vp]),
    "moma_nce_combine_poll": (c_int, [_vp, _vp, _i64, _i64, c_float, c_int, c_float, _vp, _i64, _i64, _i64, c_int, c_int, c_int,
                                      _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "moma_nce_combine_packed": (c_int, [_vp, c_int, _vp, _vp, _i64, _i64, c_float, c_int, c_float,
                                        _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "moma_nce_fused_supported": (c_int, [_i64, _i64, _i64]),
    "moma_nce_fused_workspace_bytes": (c_size_t, [_i64, _i64, _i64]),
    "moma_nce_fused": (c_int, [_vp, _vp, _vp, _vp, _i64, _i64, _i64, c_float, c_int, c_float, _vp, c_size_t, _vp,
                               _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "moma_nce_fused_packed": (c_int, [_vp, _vp, _i64, _i64, _i64, c_float, _vp, c_size_t, _vp, _vp, _vp]),
    "moma_nce_logits": (c_int, [_vp, _vp, _vp, _i64, _i64, _i64, c_float, c_int, _vp, _vp]),
    "moma_nce_logits_qk": (c_int, [_vp, _vp, _i64, _i64, c_float, _vp, _vp]),
    "moma_attn_fwd": (c_int, [_vp, _vp, _vp, _vp, _vp, _i64, _i64, c_int, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "moma_debug_nce_tc": (c_int, [_vp, _vp, _i64, _i64, _i64, c_float, c_int, _vp, _vp, _vp, _vp, _vp, _vp]),
    "moma_debug_tc_error": (c_int, []),
    "moma_debug_launch_count": (ctypes.c_longlong, [c_int]),
    "moma_debug_flops": (ctypes.c_double, [c_int, c_int]),
    "moma_debug_set_pdl": (c_int, [c_int]),
    "moma_debug_probe_launch": (c_int, [c_int, c_int, c_int, c_int, c_int, _vp]),
    "moma_debug_gemm_tc": (c_int, [_vp, _i64, c_int, _vp, _vp, _i64, c_int, _vp, _vp, _i64, _i64, _i64, _i64, c_int, _vp, c_size_t, _vp]),
    "moma_debug_gemm_tc_workspace_bytes": (c_size_t, [_i64, _i64, _i64]),
    "moma_debug_gemm_tc_error": (c_int, []),
    "moma_attn_fwd_rows": (c_int, [_vp, _vp, _vp, _vp, _vp, _i64, _i64, c_int, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _vp]),
    "moma_attn_bwd_workspace_bytes": (c_size_t, [_i64, _i64, c_int]),
    "moma_attn_bwd": (c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, c_int,
                              _vp, _vp, _vp, _vp, _vp, _vp, c_size_t, _vp]),
    "moma_peer_ctrl_bytes": (c_size_t, []),
    "moma_peer_exchange": (c_int, [_vp, _i64, _i64, c_int, _vp, _i64, _i64, _i64, c_int, c_int, c_int, _vp, _vp]),
    "moma_linear_workspace_bytes": (c_size_t, [_i64, _i64, _i64]),
    "moma_linear_fwd": (c_int, [_vp, _vp, _vp, _i64, _i64, _i64, c_int, _vp, _vp, c_size_t, _vp]),
    "moma_linear_bwd": (c_int, [_vp, _vp, _vp, _vp, _i64, _i64, _i64, c_int, _vp, _vp, _vp, _vp, c_size_t, _vp]),
}

_lib = None


def lib_path() -> str:
    return os.environ.get("MOMA_B200_LIB", _build.LIB_PATH)


def load():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise RuntimeError(
            f"moma_b200: {path} is missing -- build it with `python -m moma_b200._build` "
            "(or __graft_entry__.build()); there is no CPU fallback")
    lib = ctypes.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    got = lib.moma_abi_version()
    if got != ABI_VERSION:
        raise RuntimeError(f"moma_b200: ABI version mismatch (library {got}, binding {ABI_VERSION})")
    _lib = lib
    return lib


def last_error() -> str:
    return load().moma_last_error().decode("utf-8", "replace")


def check(rc: int) -> None:
    if rc != 0:
        raise RuntimeError(f"moma_b200 error {rc}: {last_error()}")
