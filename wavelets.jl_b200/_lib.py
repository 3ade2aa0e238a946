"""ctypes binding of libwavelets_b200.so (include/wavelets_b200.h).

The shared library is the product; there is no CPU or PyTorch fallback.  If it has not been built the
first call raises (build it with `python -c "import __graft_entry__ as g; g.build()"` or
`make -C wavelets.jl_b200/csrc`).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libwavelets_b200.so")

MAX_FILTER_LEN = 64
MAX_LIFT_STEPS = 16
MAX_LIFT_COEF = 8

OK, EDIMS, ELEVEL, EPOW2, EALIAS, ENOTCUBE, ETREE, EDTYPE, EARG, EWORKSPACE, ECUDA = range(11)
F32, F64, C64, C128 = range(4)
FLAG_STRICT_FP = 1
FLAG_FORCE_GENERIC = 2

# every symbol include/wavelets_b200.h declares (tests/test_host_logic.py::test_abi_exports_every_declared_symbol checks
# the .so exports all of them)
SYMBOLS = [
    "wb200_dwt_filter", "wb200_dwt_lifting", "wb200_wpt_filter", "wb200_wpt_lifting",
    "wb200_dwt_filter_host", "wb200_dwt_lifting_host", "wb200_workspace_bytes",
    "wb200_maxtransformlevels", "wb200_isvalidtree", "wb200_status_string",
    "wb200_last_error_string", "wb200_version", "wb200_launch_count",
    "wb200_profile_enable", "wb200_profile_collect",
    "wb200_modwt", "wb200_imodwt", "wb200_maxmodwttransformlevels",
    "wb200_trim_pool",
    "wb200_threshold", "wb200_threshold_biggest", "wb200_noisest", "wb200_denoise", "wb200_coefentropy", "wb200_bestbasistree",
]


class LiftStep(C.Structure):
    _fields_ = [("is_predict", C.c_int32), ("shift", C.c_int32), ("nc", C.c_int32),
                ("coef", C.c_double * MAX_LIFT_COEF)]


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: the CUDA library has not been built and wavelets_b200 has no fallback. "
            "Run `python -c \"import __graft_entry__ as g; g.build()\"`.")
    L = C.CDLL(LIB_PATH)
    i32, i64, vp, sz, u32, dbl = C.c_int32, C.c_int64, C.c_void_p, C.c_size_t, C.c_uint32, C.c_double
    p64, pd, pu8, pst = C.POINTER(C.c_int64), C.POINTER(C.c_double), C.POINTER(C.c_uint8), C.POINTER(LiftStep)
    L.wb200_dwt_filter.argtypes = [vp, vp, i32, p64, i64, pd, i32, i32, i32, i32, vp, sz, vp, u32]
    L.wb200_dwt_lifting.argtypes = [vp, vp, i32, p64, i64, pst, i32, dbl, dbl, i32, i32, i32, vp, sz, vp, u32]
    L.wb200_wpt_filter.argtypes = [vp, vp, i64, i64, pd, i32, pu8, i64, i32, i32, vp, sz, vp, u32]
    L.wb200_wpt_lifting.argtypes = [vp, vp, i64, i64, pst, i32, dbl, dbl, pu8, i64, i32, i32, vp, sz, vp, u32]
    L.wb200_dwt_filter_host.argtypes = [vp, vp, i32, p64, i64, pd, i32, i32, i32, i32, i32, u32]
    L.wb200_dwt_lifting_host.argtypes = [vp, vp, i32, p64, i64, pst, i32, dbl, dbl, i32, i32, i32, i32, u32]
    L.wb200_modwt.argtypes = [vp, vp, i64, i64, pd, i32, i32, i32, vp, sz, vp, u32]
    L.wb200_imodwt.argtypes = [vp, vp, i64, i64, pd, i32, i32, i32, vp, sz, vp, u32]
    L.wb200_modwt.restype = i32
    L.wb200_imodwt.restype = i32
    L.wb200_maxmodwttransformlevels.argtypes = [i64]
    L.wb200_maxmodwttransformlevels.restype = i32
    L.wb200_threshold.argtypes = [vp, i64, i32, C.c_double, i32, vp]
    L.wb200_threshold.restype = i32
    L.wb200_threshold_biggest.argtypes = [vp, i64, i64, i32, vp]
    L.wb200_threshold_biggest.restype = i32
    L.wb200_noisest.argtypes = [pd, vp, i32, C.POINTER(i64), i32, pd, i32, C.POINTER(LiftStep), i32, C.c_double, C.c_double, i32, i32, vp, u32]
    L.wb200_noisest.restype = i32
    L.wb200_denoise.argtypes = [vp, vp, i32, C.POINTER(i64), i32, pd, i32, C.POINTER(LiftStep), i32, C.c_double, C.c_double, i32,
                                i32, C.c_double, C.c_double, i32, C.POINTER(i32), i32, vp, u32]
    L.wb200_denoise.restype = i32
    L.wb200_coefentropy.argtypes = [pd, vp, i64, i32, dbl, i32, vp]
    L.wb200_coefentropy.restype = i32
    L.wb200_bestbasistree.argtypes = [pu8, pd, pd, vp, i64, i32, pd, i32, pst, i32, dbl, dbl, pu8, i64, i32, i32, vp, u32]
    L.wb200_bestbasistree.restype = i32
    L.wb200_workspace_bytes.argtypes = [i32, i32, p64, i64, i32, i32, u32]
    L.wb200_workspace_bytes.restype = sz
    L.wb200_maxtransformlevels.argtypes = [i64]
    L.wb200_isvalidtree.argtypes = [i64, pu8, i64]
    L.wb200_status_string.argtypes = [i32]
    L.wb200_status_string.restype = C.c_char_p
    L.wb200_last_error_string.restype = C.c_char_p
    L.wb200_launch_count.argtypes = [i32]
    L.wb200_launch_count.restype = i64
    L.wb200_trim_pool.argtypes = [i64]
    L.wb200_trim_pool.restype = i64
    L.wb200_profile_enable.argtypes = [i32]
    L.wb200_profile_enable.restype = None
    L.wb200_profile_collect.argtypes = [C.c_char_p, i64]
    L.wb200_profile_collect.restype = i64
    for name in ("wb200_dwt_filter", "wb200_dwt_lifting", "wb200_wpt_filter", "wb200_wpt_lifting",
                 "wb200_dwt_filter_host", "wb200_dwt_lifting_host", "wb200_maxtransformlevels",
                 "wb200_isvalidtree", "wb200_version"):
        getattr(L, name).restype = i32
    _lib = L
    return L


def make_steps(gls):
    """GLS descriptor -> (ctypes array of wb200_lift_step, n)."""
    arr = (LiftStep * max(1, len(gls.step)))()
    for i, s in enumerate(gls.step):
        arr[i].is_predict = 1 if s.steptype == "predict" else 0
        arr[i].shift = int(s.shift)
        arr[i].nc = len(s.coef)
        for k, c in enumerate(s.coef):
            arr[i].coef[k] = float(c)
    return arr, len(gls.step)


def dims_array(dims):
    return (C.c_int64 * 3)(*(list(dims) + [1] * (3 - len(dims))))
