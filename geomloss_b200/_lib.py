"""ctypes binding of libb200ot.so — the only way the Python host side reaches the CUDA kernels.

There is deliberately no fallback: if the library is missing or a call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_float, c_int32, c_int64, c_void_p

_PKG = os.path.dirname(os.path.abspath(__file__))
# $B200OT_LIB: load a variant build instead (A/B kernel timing, tools/ab_ops.py) — never set in tests or the bench
LIB_PATH = os.environ.get("B200OT_LIB") or os.path.join(_PKG, "libb200ot.so")

_lib = None

_P = c_void_p  # device pointers and streams travel as void*


class B200OTError(RuntimeError):
    pass


_SIGNATURES = {
    "b200ot_version": (c_int32, []),
    "b200ot_strerror": (c_char_p, [c_int32]),
    "b200ot_last_cuda_error": (c_char_p, []),
    "b200ot_softmin_scratch_bytes": (c_int64, [c_int64, c_int64, c_int32]),
    "b200ot_softmin_fwd": (c_int32, [_P, _P, _P, _P, c_float, _P, _P, c_float, c_float, _P, _P, c_int64, c_int64,
                                     c_int32, c_int32, c_float, _P, c_int64, _P]),
    "b200ot_softmin_bwd_x": (c_int32, [_P, _P, _P, _P, c_float, _P, _P, _P, _P, c_int64, c_int64, c_int32, c_int32,
                                       c_float, _P, c_int64, _P]),
    "b200ot_packed_cols_floats": (c_int64, [c_int64, c_int32, c_int32]),
    "b200ot_softmin_pack": (c_int32, [_P, _P, _P, c_float, _P, c_int64, c_int32, c_int32, c_float, _P, _P]),
    "b200ot_softmin_num_splits": (c_int32, [c_int64, c_int64, c_int32]),
    "b200ot_softmin_partial": (c_int32, [_P, _P, _P, _P, c_int32, c_int64, c_int64, c_int32, c_int32, c_float, _P]),
    "b200ot_softmin_finalize": (c_int32, [_P, c_int32, _P, c_float, c_float, _P, _P, c_int64, c_float, _P]),
    "b200ot_ranges_shape": (None, [c_int32, ctypes.POINTER(c_int32), ctypes.POINTER(c_int32),
                                   ctypes.POINTER(c_int32)]),
    "b200ot_softmin_pack_gather": (c_int32, [_P, _P, _P, c_float, _P, _P, c_int64, c_int32, c_int32, c_float, _P, _P]),
    "b200ot_softmin_partial_ranges": (c_int32, [_P, _P, _P, _P, c_int64, _P, _P, c_int64, c_int32, c_int32, c_float,
                                                c_int32, _P]),
    "b200ot_softmin_merge": (c_int32, [_P, c_int32, _P, c_int64, _P]),
    "b200ot_softmin_bwd_partial": (c_int32, [_P, _P, _P, _P, _P, c_int32, c_int64, c_int64, c_int32, c_int32,
                                             c_float, _P]),
    "b200ot_softmin_bwd_partial_ranges": (c_int32, [_P, _P, _P, _P, _P, c_int64, _P, _P, c_int64, c_int32, c_int32,
                                                    c_float, c_int32, _P]),
    "b200ot_softmin_bwd_sums": (c_int32, [_P, _P, _P, _P, c_float, _P, _P, _P, c_int64, c_int64, c_int32, c_int32,
                                          c_float, _P, c_int64, _P]),
    "b200ot_rowsum_merge": (c_int32, [_P, c_int32, c_int32, _P, c_int64, _P]),
    "b200ot_softmin_bwd_finalize": (c_int32, [_P, c_int32, _P, _P, _P, _P, c_int64, c_int32, c_int32, c_float, _P]),
    "b200ot_kernel_conv_scratch_bytes": (c_int64, [c_int64, c_int64, c_int32]),
    "b200ot_kernel_conv_fwd": (c_int32, [_P, _P, _P, _P, _P, c_int64, c_int64, c_int32, c_int32, c_float, _P,
                                         c_int64, _P]),
    "b200ot_kernel_conv_bwd_x": (c_int32, [_P, _P, _P, _P, _P, _P, c_int64, c_int64, c_int32, c_int32, c_float, _P,
                                           c_int64, _P]),
    "b200ot_kernel_conv_fwd_bwd_x": (c_int32, [_P, _P, _P, _P, _P, _P, c_int64, c_int64, c_int32, c_int32, c_float,
                                               _P, c_int64, _P]),
    "b200ot_kernel_conv_pack_gather": (c_int32, [_P, _P, _P, _P, c_int64, c_int32, c_int32, c_float, _P, _P]),
    "b200ot_kernel_conv_partial_ranges": (c_int32, [_P, _P, _P, _P, c_int64, _P, _P, c_int64, c_int32, c_int32,
                                                    c_float, c_int32, c_int32, _P]),
    "b200ot_kernel_conv_finalize": (c_int32, [_P, c_int32, _P, c_int64, c_int32, _P]),
    "b200ot_kernel_conv_bwd_finalize": (c_int32, [_P, c_int32, _P, _P, _P, _P, c_int64, c_int32, c_int32, c_float,
                                                  _P]),
    "b200ot_sinkhorn_iteration_small": (c_int32, [_P] * 13 + [c_int64, c_int64, c_int64, c_int32, c_int32, c_float,
                                                  c_float, c_float, c_int32, _P]),
    "b200ot_sinkhorn_loop_small": (c_int32, [_P, _P, _P, _P, c_int32, ctypes.POINTER(ctypes.c_double), c_int32,
                                             ctypes.c_double, c_int32, _P, _P, ctypes.POINTER(c_int32), c_int64,
                                             c_int64, c_int64, c_int32, c_int32, _P]),
    "b200ot_sinkhorn_final_bwd_small": (c_int32, [_P] * 15 + [c_int64, c_int64, c_int64, c_int32, c_int32, c_float,
                                                  c_float, _P, c_int32, _P]),
    "b200ot_sinkhorn_cost_small": (c_int32, [_P] * 6 + [c_int64, c_int64, c_int64, c_float, c_float] + [_P] * 8),
    "b200ot_kernel_mmd_value_small": (c_int32, [_P] * 5 + [c_int64, c_int64, c_int64, _P, _P]),
    "b200ot_cloud_extent_scratch_bytes": (c_int64, []),
    "b200ot_cloud_extent": (c_int32, [_P, c_int64, _P, c_int64, c_int32, _P, _P, c_int64, _P]),
    "b200ot_kernel_mmd_small": (c_int32, [_P] * 8 + [c_int64, c_int64, c_int64, c_int32, c_int32, c_float, _P]),
    "b200ot_kernel_mmd_bwd_small": (c_int32, [_P] * 7 + [c_int64, c_int64, c_int64, c_int32, c_int32, c_float, _P]),
    "b200ot_softmin_grid": (c_int32, [_P, _P, c_float, _P, c_float, c_float, _P, c_int64, c_int32, c_int32, c_int32,
                                      c_float, _P]),
    "b200ot_ubench": (c_int32, [c_int32, c_int32, c_int32, _P, ctypes.POINTER(c_int32), _P]),
}


def exported_symbols():
    """Names every build of the library must export (mirrors include/b200ot.h)."""
    return sorted(_SIGNATURES)


def lib():
    """Load (once) and return the C-ABI library; raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise B200OTError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  geomloss_b200 has no CPU or eager fallback.")
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(code: int, what: str):
    if code != 0:
        L = lib()
        msg = L.b200ot_strerror(code).decode()
        if code == -3:
            msg += ": " + L.b200ot_last_cuda_error().decode()
        raise B200OTError(f"{what} failed: {msg}")
