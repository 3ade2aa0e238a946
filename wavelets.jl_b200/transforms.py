"""Transforms -- the reference's public transform API on CUDA tensors.

Host-side mirror of src/Transforms/transforms_main.jl:105-207: `dwt / idwt / dwt! / idwt! / wpt / iwpt /
wpt! / iwpt!` with the same argument meaning, defaults and error behaviour, plus the column-wise batch
forms `dwtc / idwtc` the reference advertises (README.md:6) but only stubs (transforms_main.jl:179-181).
Julia's `f!` is spelled `f_` here.  Every function is a thin call into libwavelets_b200.so through the C
ABI (include/wavelets_b200.h); there is no CPU or PyTorch fallback.

Arrays are torch CUDA tensors in Julia (column-major) layout: `x[i, j, k]` has strides (1, m, m*n).
Row-major inputs are accepted and re-laid out (one copy); results are always column-major tensors of
the same logical shape, so indices read exactly like the reference's.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from .util import iscube, isvalidtree, maketree, maxtransformlevels, sufficientpoweroftwo
from .wt import GLS, OrthoFilter

__all__ = ["modwt", "imodwt", "maxmodwttransformlevels", "dwt", "idwt", "dwt_", "idwt_", "wpt", "iwpt", "wpt_", "iwpt_", "dwtc", "idwtc",
           "dwt_oop_", "idwt_oop_", "ArgumentError", "DimensionMismatch", "set_strict_fp", "colmajor", "release_scratch"]


class ArgumentError(ValueError):
    """Julia ArgumentError."""


class DimensionMismatch(ValueError):
    """Julia DimensionMismatch."""


_STATUS_EXC = {
    _lib.EDIMS: DimensionMismatch, _lib.ELEVEL: ArgumentError, _lib.EPOW2: ArgumentError,
    _lib.EALIAS: ArgumentError, _lib.ENOTCUBE: ArgumentError, _lib.ETREE: ArgumentError,
    _lib.EDTYPE: TypeError, _lib.EARG: ArgumentError, _lib.EWORKSPACE: RuntimeError, _lib.ECUDA: RuntimeError,
}

_flags = 0


def set_strict_fp(on: bool) -> None:
    """strict = no FMA contraction: results bit-identical to the reference CPU path (slower)."""
    global _flags
    _flags = (_flags | _lib.FLAG_STRICT_FP) if on else (_flags & ~_lib.FLAG_STRICT_FP)


def release_scratch(keep_bytes: int = 0, device=None) -> int:
    """Give the library's cached device scratch (its private stream-ordered pool) back to the device, keeping at most
    `keep_bytes`; returns the bytes still reserved.  Calls made without a caller workspace re-reserve on demand."""
    dev = torch.cuda.current_device() if device is None else torch.device(device).index
    with torch.cuda.device(dev):
        torch.cuda.synchronize()
        left = int(_lib.lib().wb200_trim_pool(int(keep_bytes)))
    if left < 0:
        raise RuntimeError("wb200_trim_pool failed")
    return left


def _flags_value() -> int:
    return _flags


def _force_generic(on: bool) -> None:
    global _flags
    _flags = (_flags | _lib.FLAG_FORCE_GENERIC) if on else (_flags & ~_lib.FLAG_FORCE_GENERIC)


def _check(rc: int) -> None:
    if rc == _lib.OK:
        return
    L = _lib.lib()
    msg = L.wb200_status_string(rc).decode()
    detail = L.wb200_last_error_string().decode()
    if detail and detail != msg:
        msg = f"{msg} ({detail})"
    raise _STATUS_EXC.get(rc, RuntimeError)(msg)


_DTYPES = {torch.float32: _lib.F32, torch.float64: _lib.F64, torch.complex64: _lib.C64, torch.complex128: _lib.C128}


def _colmajor_strides(shape):
    st, acc = [], 1
    for s in shape:
        st.append(acc)
        acc *= int(s)
    return tuple(st)


def _is_colmajor(x: torch.Tensor) -> bool:
    if x.numel() == 0:
        return True
    exp = _colmajor_strides(x.shape)
    return all(s == 1 or st == e for s, st, e in zip(x.shape, x.stride(), exp))


def colmajor(x: torch.Tensor) -> torch.Tensor:
    """Return x (same logical shape) backed by column-major storage (copying only if needed)."""
    if _is_colmajor(x):
        return x
    rev = tuple(reversed(range(x.dim())))
    return x.permute(rev).contiguous().permute(rev)


def _similar(x: torch.Tensor) -> torch.Tensor:
    return torch.empty_strided(tuple(x.shape), _colmajor_strides(x.shape), dtype=x.dtype, device=x.device)


def _prep(x) -> torch.Tensor:
    if not isinstance(x, torch.Tensor):
        raise TypeError("wavelets_b200 transforms operate on torch CUDA tensors")
    if not x.is_cuda:
        raise TypeError("wavelets_b200 has no CPU path: move the tensor to a CUDA device (B200)")
    if not (x.dtype.is_floating_point or x.dtype.is_complex):
        x = x.to(torch.float64)          # Int -> float(x), transforms_main.jl:187-190
    if x.dtype not in _DTYPES:
        raise TypeError(f"unsupported element type {x.dtype}")
    return colmajor(x)


def _check_out(y, x: torch.Tensor, what: str) -> None:
    """The destination of an out-of-place call: a column-major CUDA tensor on x's device with x's element type and shape
    (the kernels write dense column-major elements of x's width through y's pointer)."""
    if not isinstance(y, torch.Tensor) or not y.is_cuda:
        raise TypeError(f"{what}: y must be a torch CUDA tensor")
    if y.device != x.device:
        raise TypeError(f"{what}: y and x must live on the same device")
    if y.dtype not in _DTYPES or y.dtype != x.dtype:
        raise TypeError(f"{what}: y must have x's element type ({x.dtype}), got {y.dtype}")
    if tuple(x.shape) != tuple(y.shape):
        raise DimensionMismatch("in and out array size must match")
    if not _is_colmajor(y):
        raise TypeError(f"{what}: y must be column-major (Julia layout); see colmajor()")


def _check_inplace(y, what: str) -> None:
    if not isinstance(y, torch.Tensor) or not y.is_cuda or y.dtype not in _DTYPES or not _is_colmajor(y):
        raise TypeError(f"{what}: y must be a column-major floating CUDA tensor (transformed in place)")


def _stream(x: torch.Tensor):
    return C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)


def _qmf(f: OrthoFilter):
    q = np.ascontiguousarray(f.qmf, dtype=np.float64)
    return q, q.ctypes.data_as(C.POINTER(C.c_double))


def _call_dwt(y: torch.Tensor, x: torch.Tensor, wt, L: int, fw: bool, ndim: int, batch: int) -> None:
    L_ = _lib.lib()
    dims = _lib.dims_array(list(x.shape[:ndim]))
    if not isinstance(wt, (OrthoFilter, GLS)):
        raise TypeError("wt must be an OrthoFilter or a GLS (see wavelet())")
    if x.numel() == 0:
        return
    with torch.cuda.device(x.device):
        if isinstance(wt, OrthoFilter):
            q, qp = _qmf(wt)
            rc = L_.wb200_dwt_filter(y.data_ptr(), x.data_ptr(), ndim, dims, batch, qp, len(q), int(L),
                                     1 if fw else 0, _DTYPES[x.dtype], None, 0, _stream(x), _flags)
        elif isinstance(wt, GLS):
            steps, ns = _lib.make_steps(wt)
            rc = L_.wb200_dwt_lifting(y.data_ptr(), x.data_ptr(), ndim, dims, batch, steps, ns,
                                      float(wt.norm1), float(wt.norm2), int(L), 1 if fw else 0,
                                      _DTYPES[x.dtype], None, 0, _stream(x), _flags)
        else:
            raise TypeError("wt must be an OrthoFilter or a GLS (see wavelet())")
    _check(rc)


def _xwt(x, wt, L, fw: bool):
    x = _prep(x)
    if not 1 <= x.dim() <= 3:
        raise DimensionMismatch("dwt supports 1-D, 2-D and 3-D arrays")
    if L is None:
        L = maxtransformlevels(x)
    y = _similar(x)
    _call_dwt(y, x, wt, L, fw, x.dim(), 1)   # lifting: x != y selects the allocating form (no extra copy)
    return y


def dwt(x, wt, L=None):
    """dwt(x, wt[, L=maxtransformlevels(x)]) -- transforms_main.jl:109-113 (filter), 119-124 (lifting)."""
    return _xwt(x, wt, L, True)


def idwt(x, wt, L=None):
    """idwt(x, wt[, L]) -- inverse of dwt(x, wt, L)."""
    return _xwt(x, wt, L, False)


def _xwt_bang(args, fw: bool):
    # dwt!(y, x, filter[, L])  |  dwt!(y, scheme[, L])      transforms_main.jl:114-117, 125-128
    if len(args) >= 2 and isinstance(args[1], GLS):
        y, scheme = args[0], args[1]
        L = args[2] if len(args) > 2 else None
        _check_inplace(y, "dwt_(y, scheme)")
        if L is None:
            L = maxtransformlevels(y)
        _call_dwt(y, y, scheme, L, fw, y.dim(), 1)
        return y
    if len(args) >= 3 and isinstance(args[2], OrthoFilter):
        y, x, filt = args[0], args[1], args[2]
        L = args[3] if len(args) > 3 else None
        x = _prep(x)
        _check_out(y, x, "dwt_(y, x, filter)")
        if L is None:
            L = maxtransformlevels(x)
        if y.data_ptr() == x.data_ptr():
            raise ArgumentError("in array is out array")
        _call_dwt(y, x, filt, L, fw, x.dim(), 1)
        return y
    raise TypeError("usage: dwt_(y, x, filter[, L]) or dwt_(y, scheme[, L])")


def dwt_(*args):
    """dwt!(y, x, wt::OrthoFilter[, L]) (out of place) / dwt!(y, wt::GLS[, L]) (in place)."""
    return _xwt_bang(args, True)


def idwt_(*args):
    """idwt!: the inverse of dwt!."""
    return _xwt_bang(args, False)


def dwt_oop_(y, x, wt, L=None):
    """dwt_oop!(y, x, wt, L) -- transforms_main.jl:193-207: out of place for both transform types."""
    if isinstance(wt, GLS):
        x = _prep(x)
        _check_out(y, x, "dwt_oop_(y, x, scheme)")
        if L is None:
            L = maxtransformlevels(x)
        _call_dwt(y, x, wt, L, True, x.dim(), 1)
        return y
    return dwt_(y, x, wt, L) if L is not None else dwt_(y, x, wt)


def idwt_oop_(y, x, wt, L=None):
    if isinstance(wt, GLS):
        x = _prep(x)
        _check_out(y, x, "idwt_oop_(y, x, scheme)")
        if L is None:
            L = maxtransformlevels(x)
        _call_dwt(y, x, wt, L, False, x.dim(), 1)
        return y
    return idwt_(y, x, wt, L) if L is not None else idwt_(y, x, wt)


# ---- column-wise batch forms --------------------------------------------------------------------------
def _xwtc(x, wt, L, fw: bool):
    x = _prep(x)
    if not 2 <= x.dim() <= 4:
        raise DimensionMismatch("dwtc expects (n, B), (m, n, B) or (m, n, d, B): the last dimension is the batch")
    nd = x.dim() - 1
    if L is None:
        L = min(maxtransformlevels(int(s)) for s in x.shape[:nd])
    y = _similar(x)
    _call_dwt(y, x, wt, L, fw, nd, int(x.shape[-1]))
    return y


def dwtc(x, wt, L=None):
    """Column-wise dwt: an independent transform of every slice along the LAST dimension."""
    return _xwtc(x, wt, L, True)


def idwtc(x, wt, L=None):
    return _xwtc(x, wt, L, False)


# ---- wavelet packets ----------------------------------------------------------------------------------
def _tree_arg(x, arg):
    n = int(x.shape[0])
    if arg is None:
        return maketree(n, maxtransformlevels(n), "full")
    if isinstance(arg, (int, np.integer)):
        return maketree(n, int(arg), "full")          # wpt(x, wt, L::Integer), transforms_main.jl:137-140
    return np.ascontiguousarray(np.asarray(arg), dtype=np.uint8)


def _call_wpt(y, x, wt, tree, fw: bool, batch: int = 1):
    L_ = _lib.lib()
    n = int(x.shape[0])
    tp = tree.ctypes.data_as(C.POINTER(C.c_uint8))
    with torch.cuda.device(x.device):
        if isinstance(wt, OrthoFilter):
            q, qp = _qmf(wt)
            rc = L_.wb200_wpt_filter(y.data_ptr(), x.data_ptr(), n, batch, qp, len(q), tp, len(tree),
                                     1 if fw else 0, _DTYPES[x.dtype], None, 0, _stream(x), _flags)
        elif isinstance(wt, GLS):
            steps, ns = _lib.make_steps(wt)
            rc = L_.wb200_wpt_lifting(y.data_ptr(), x.data_ptr(), n, batch, steps, ns, float(wt.norm1),
                                      float(wt.norm2), tp, len(tree), 1 if fw else 0, _DTYPES[x.dtype],
                                      None, 0, _stream(x), _flags)
        else:
            raise TypeError("wt must be an OrthoFilter or a GLS (see wavelet())")
    _check(rc)


def _xwpt(x, wt, tree, fw: bool):
    x = _prep(x)
    if x.dim() == 1:
        batch = 1
    elif x.dim() == 2:
        batch = int(x.shape[1])                        # column-wise batch (extension, as dwtc)
    else:
        raise DimensionMismatch("wpt expects a vector (or an (n, B) batch of columns)")
    tree = _tree_arg(x, tree)
    y = _similar(x)
    _call_wpt(y, x, wt, tree, fw, batch)
    return y


def wpt(x, wt, tree=None):
    """wpt(x, wt[, L | tree]) -- transforms_main.jl:137-146, 159-165. `tree`: BitVector as 0/1 array."""
    return _xwpt(x, wt, tree, True)


def iwpt(x, wt, tree=None):
    return _xwpt(x, wt, tree, False)


def _xwpt_bang(args, fw: bool):
    # wpt!(y, x, filter[, L | tree]) | wpt!(y, scheme[, L | tree])        transforms_main.jl:147-158, 166-175
    if len(args) >= 2 and isinstance(args[1], GLS):
        y, scheme = args[0], args[1]
        _check_inplace(y, "wpt_(y, scheme)")
        if y.dim() not in (1, 2):
            raise DimensionMismatch("wpt expects a vector (or an (n, B) batch of columns)")
        tree = _tree_arg(y, args[2] if len(args) > 2 else None)
        _call_wpt(y, y, scheme, tree, fw, 1 if y.dim() == 1 else int(y.shape[1]))
        return y
    if len(args) >= 3 and isinstance(args[2], OrthoFilter):
        y, x, filt = args[0], _prep(args[1]), args[2]
        _check_out(y, x, "wpt_(y, x, filter)")
        if x.dim() not in (1, 2):
            raise DimensionMismatch("wpt expects a vector (or an (n, B) batch of columns)")
        if y.data_ptr() == x.data_ptr():
            raise ArgumentError("in array is out array")
        tree = _tree_arg(x, args[3] if len(args) > 3 else None)
        _call_wpt(y, x, filt, tree, fw, 1 if x.dim() == 1 else int(x.shape[1]))
        return y
    raise TypeError("usage: wpt_(y, x, filter[, L|tree]) or wpt_(y, scheme[, L|tree])")


def wpt_(*args):
    return _xwpt_bang(args, True)


def iwpt_(*args):
    return _xwpt_bang(args, False)


# ---- maximal-overlap DWT (src/Transforms/transforms_maximal_overlap.jl) ------------------------------------
def maxmodwttransformlevels(x) -> int:
    """floor(log2(length(x))) -- src/Util/non_dyadic.jl:24-25."""
    n = int(x) if isinstance(x, (int, np.integer)) else int(x.shape[0])
    return int(np.floor(np.log2(n)))


def modwt(x, wt, L=None):
    """modwt(x, wt[, L]) -> n x (L+1) matrix [W_1 .. W_L V_L] (for an (n, B) batch: (n, L+1, B))."""
    if not isinstance(wt, OrthoFilter):
        raise TypeError("modwt takes an OrthoFilter")
    x = _prep(x)
    if x.dtype not in (torch.float32, torch.float64) or x.dim() not in (1, 2):
        raise TypeError("modwt takes a real vector (or an (n, B) batch of columns)")
    n = int(x.shape[0])
    B = 1 if x.dim() == 1 else int(x.shape[1])
    if L is None:
        L = maxmodwttransformlevels(n)
    if L > maxmodwttransformlevels(n):
        raise ArgumentError("Too many transform levels (length(x) < 2^L)")
    if L < 1:
        raise ArgumentError("L must be >= 1")
    shape = (n, L + 1) if x.dim() == 1 else (n, L + 1, B)
    y = torch.empty_strided(shape, _colmajor_strides(shape), dtype=x.dtype, device=x.device)
    if x.numel():
        q, qp = _qmf(wt)
        with torch.cuda.device(x.device):
            rc = _lib.lib().wb200_modwt(y.data_ptr(), x.data_ptr(), n, B, qp, len(q), int(L), _DTYPES[x.dtype], None, 0,
                                        _stream(x), _flags)
        _check(rc)
    return y


def imodwt(xw, wt):
    """imodwt(xw, wt): the inverse of modwt (xw: n x (L+1), or (n, L+1, B))."""
    if not isinstance(wt, OrthoFilter):
        raise TypeError("imodwt takes an OrthoFilter")
    xw = _prep(xw)
    if xw.dtype not in (torch.float32, torch.float64) or xw.dim() not in (2, 3):
        raise TypeError("imodwt takes an n x (L+1) real matrix (or an (n, L+1, B) batch)")
    n, ncols = int(xw.shape[0]), int(xw.shape[1])
    B = 1 if xw.dim() == 2 else int(xw.shape[2])
    shape = (n,) if xw.dim() == 2 else (n, B)
    x = torch.empty_strided(shape, _colmajor_strides(shape), dtype=xw.dtype, device=xw.device)
    if xw.numel():
        q, qp = _qmf(wt)
        with torch.cuda.device(xw.device):
            rc = _lib.lib().wb200_imodwt(x.data_ptr(), xw.data_ptr(), n, B, qp, len(q), ncols, _DTYPES[xw.dtype], None, 0,
                                         _stream(xw), _flags)
        _check(rc)
    return x
