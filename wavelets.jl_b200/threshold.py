"""Host mirror of the reference's Threshold module for the device path (src/Threshold/threshold_main.jl,
denoising.jl): threshold types, `threshold` / `threshold_` (threshold!), `VisuShrink`, `noisest`, `denoise`.

Everything runs through the C ABI (`wb200_threshold`, `wb200_noisest`, `wb200_denoise`) on CUDA tensors; there is no CPU
path.  `denoise` is a pure enqueue: with the default `estnoise` the noise level is estimated on the device and never
visits the host.  `BiggestTH` (m-term approximation) uses the same radix select as the noise estimate.  Not on the device path:
`matchingpursuit`, `bestbasistree`, entropy (SURVEY 8f rows 3+).
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np
import torch

from . import _lib
from .transforms import ArgumentError, _check, _colmajor_strides, _DTYPES, _flags_value, _is_colmajor, _prep, _stream
from .util import iscube, isvalidtree, maketree, maxtransformlevels
from . import wt as WT
from .wt import GLS, OrthoFilter, wavelet

__all__ = ["Entropy", "ShannonEntropy", "LogEnergyEntropy", "coefentropy", "bestbasistree", "THType", "HardTH", "SoftTH", "SemiSoftTH", "SteinTH", "BiggestTH", "PosTH", "NegTH", "DEFAULT_TH",
           "threshold", "threshold_", "DNFT", "VisuShrink", "denoise", "noisest", "DEFAULT_WAVELET"]


class THType:
    kind = None

    def __repr__(self):
        return type(self).__name__ + "()"


class HardTH(THType): kind = 0
class SoftTH(THType): kind = 1
class SemiSoftTH(THType): kind = 2
class SteinTH(THType): kind = 3
class NegTH(THType): kind = 4
class PosTH(THType): kind = 5
class BiggestTH(THType): kind = None


DEFAULT_TH = HardTH()


def _real(x):
    x = _prep(x)
    if x.dtype not in (torch.float32, torch.float64):
        raise TypeError("thresholding / denoising take real Float32 / Float64 arrays")
    return x


def threshold_(x, TH: THType, t=None):
    """threshold!(x, TH, t) / threshold!(x, TH) in place (threshold_main.jl:35-117)."""
    if not isinstance(TH, THType):
        raise TypeError("TH must be a threshold type (HardTH(), SoftTH(), ...)")
    if isinstance(TH, BiggestTH):           # threshold!(x, BiggestTH(), m::Int): the m-term approximation
        if t is None or int(t) != t:
            raise TypeError("BiggestTH takes an integer number of terms m")
        if int(t) < 0:
            raise AssertionError("m >= 0")
        if not isinstance(x, torch.Tensor) or not x.is_cuda or x.dtype not in (torch.float32, torch.float64):
            raise TypeError("threshold_ operates in place on a Float32 / Float64 CUDA tensor")
        if x.numel():
            if not (x.is_contiguous() or _is_colmajor(x)):
                raise TypeError("threshold_ needs dense storage")
            with torch.cuda.device(x.device):
                rc = _lib.lib().wb200_threshold_biggest(x.data_ptr(), x.numel(), int(t), _DTYPES[x.dtype], _stream(x))
            _check(rc)
        return x
    if isinstance(TH, (NegTH, PosTH)):
        if t is not None:
            raise TypeError(f"{TH!r} takes no threshold value")
        t = 0.0
    else:
        if t is None:
            raise TypeError(f"{TH!r} needs a threshold value")
        if not float(t) >= 0:
            raise AssertionError("t >= 0")
    if not isinstance(x, torch.Tensor) or not x.is_cuda or x.dtype not in (torch.float32, torch.float64):
        raise TypeError("threshold_ operates in place on a Float32 / Float64 CUDA tensor")
    if x.numel():
        if not (x.is_contiguous() or _is_colmajor(x)):
            raise TypeError("threshold_ needs dense storage")
        with torch.cuda.device(x.device):
            rc = _lib.lib().wb200_threshold(x.data_ptr(), x.numel(), TH.kind, float(t), _DTYPES[x.dtype], _stream(x))
        _check(rc)
    return x


def threshold(x, TH: THType, t=None):
    """threshold(x, TH, t): the copying form."""
    return threshold_(_real(x).clone(memory_format=torch.preserve_format), TH, t)


class DNFT:
    pass


class VisuShrink(DNFT):
    """VisuShrink(th, t) / VisuShrink(n): threshold type + threshold for unit noise level, sqrt(2 log n)."""

    def __init__(self, th, t=None):
        if t is None:                       # VisuShrink(n::Int)
            n = int(th)
            self.th, self.t = DEFAULT_TH, math.sqrt(2 * math.log(n))
        else:
            self.th, self.t = th, float(t)


DEFAULT_WAVELET = wavelet(WT.sym5, WT.Filter)


def _wt_args(wt):
    null_d = C.POINTER(C.c_double)()
    null_s = C.POINTER(_lib.LiftStep)()
    if wt is None:
        return 0, null_d, 0, null_s, 0, 0.0, 0.0, ()
    if isinstance(wt, OrthoFilter):
        q = np.ascontiguousarray(wt.qmf, dtype=np.float64)
        return 1, q.ctypes.data_as(C.POINTER(C.c_double)), len(q), null_s, 0, 0.0, 0.0, (q,)
    if isinstance(wt, GLS):
        steps, ns = _lib.make_steps(wt)
        return 2, null_d, 0, steps, ns, float(wt.norm1), float(wt.norm2), (steps,)
    raise TypeError("wt must be an OrthoFilter, a GLS or None")


def noisest(x, wt=DEFAULT_WAVELET, L: int = 1) -> float:
    """noisest(x, wt, L=1): MAD of y[detailrange(y, L)], y = dwt(x, wt, L), over 0.6745 (denoising.jl:94-101).
    Returns a Python float (the one call of this module that waits for the stream)."""
    x = _real(x)
    wk, qp, fl, st, ns, n1, n2, keep = _wt_args(wt)
    out = C.c_double(0.0)
    dims = _lib.dims_array(list(x.shape))
    with torch.cuda.device(x.device):
        rc = _lib.lib().wb200_noisest(C.byref(out), x.data_ptr(), x.dim(), dims, wk, qp, fl, st, ns, n1, n2, int(L), _DTYPES[x.dtype],
                                      _stream(x), _flags_value())
    _check(rc)
    return out.value


def denoise(x, wt=DEFAULT_WAVELET, L=None, dnt=None, estnoise=None, TI: bool = False, nspin=None):
    """denoise(x, wt; L=min(maxtransformlevels(x), 6), dnt=VisuShrink(size(x,1)), estnoise=noisest, TI=false,
    nspin=(8, ...)) -- denoising.jl:22-82.  `estnoise`: None = noisest on the device (no host round trip), or a
    callable (x, wt) -> float as in the reference."""
    x = _real(x)
    if x.dim() < 1 or x.dim() > 3:
        raise TypeError("denoise takes 1-D, 2-D or 3-D arrays")
    if not iscube(x):
        raise ArgumentError("array must be square/cube")
    if L is None:
        L = min(maxtransformlevels(x), 6)
    if dnt is None:
        dnt = VisuShrink(int(x.shape[0]))
    if not isinstance(dnt, VisuShrink):
        raise TypeError("dnt must be a VisuShrink")
    if isinstance(dnt.th, BiggestTH):       # threshold!(xt, BiggestTH(), sigma * t) has no method upstream either (m::Int)
        raise TypeError("VisuShrink needs a value threshold type; BiggestTH takes a term count")
    if nspin is None:
        nspin = tuple(8 for _ in range(x.dim()))
    sp = [int(nspin)] if isinstance(nspin, (int, np.integer)) else [int(v) for v in nspin]
    if len(sp) > x.dim() or any(v < 1 for v in sp):
        raise ArgumentError("nspin must hold one positive count per shifted dimension")
    spin = (C.c_int32 * 3)(*(sp + [1] * (3 - len(sp))))
    if TI and wt is None:
        raise RuntimeError("TI not supported with wt=nothing")
    sigma = float("nan") if estnoise is None else float(estnoise(x, wt))
    wk, qp, fl, st, ns, n1, n2, keep = _wt_args(wt)
    y = torch.empty_strided(tuple(x.shape), _colmajor_strides(x.shape), dtype=x.dtype, device=x.device)
    if x.numel():
        dims = _lib.dims_array(list(x.shape))
        with torch.cuda.device(x.device):
            rc = _lib.lib().wb200_denoise(y.data_ptr(), x.data_ptr(), x.dim(), dims, wk, qp, fl, st, ns, n1, n2, int(L),
                                          dnt.th.kind, float(dnt.t), sigma, 1 if TI else 0, spin, _DTYPES[x.dtype], _stream(x),
                                          _flags_value())
        _check(rc)
    return y


# ---- entropy / best basis (src/Threshold/entropy.jl) ----------------------------------------------------------------
class Entropy:
    kind = None

    def __repr__(self):
        return type(self).__name__ + "()"


class ShannonEntropy(Entropy): kind = 0       # Coifman-Wickerhauser
class LogEnergyEntropy(Entropy): kind = 1


def coefentropy(x, et: Entropy = None, nrm=None) -> float:
    """coefentropy(x, et, nrm = norm(x)): the additive entropy of the coefficients (entropy.jl:16-43)."""
    et = ShannonEntropy() if et is None else et
    if not isinstance(et, Entropy):
        raise TypeError("et must be ShannonEntropy() or LogEnergyEntropy()")
    x = _real(x)
    if nrm is not None and not float(nrm) >= 0:
        raise AssertionError("nrm >= 0")
    out = C.c_double(0.0)
    with torch.cuda.device(x.device):
        rc = _lib.lib().wb200_coefentropy(C.byref(out), x.data_ptr(), x.numel(), et.kind, float("nan") if nrm is None else float(nrm),
                                          _DTYPES[x.dtype], _stream(x))
    _check(rc)
    return out.value


def bestbasistree(y, wt, L=None, et: Entropy = None, return_entropies: bool = False):
    """bestbasistree(y, wt, L = maxtransformlevels(y), et = ShannonEntropy()) / bestbasistree(y, wt, tree, et): the best
    packet tree that is a subset of the input tree (entropy.jl:46-108).  Returns the tree as the uint8 array `maketree` gives
    (usable with `wpt` / `iwpt`)."""
    et = ShannonEntropy() if et is None else et
    if not isinstance(et, Entropy):
        raise TypeError("et must be ShannonEntropy() or LogEnergyEntropy()")
    y = _real(y)
    if y.dim() != 1:
        raise TypeError("bestbasistree takes a vector")
    n = int(y.shape[0])
    if L is None or isinstance(L, (int, np.integer)):
        tree = maketree(n, L, "full")
    else:
        tree = np.ascontiguousarray(np.asarray(L), dtype=np.uint8)
    if not isvalidtree(n, tree):
        raise ArgumentError("invalid tree")
    if wt is None:
        raise TypeError("bestbasistree needs a wavelet")
    wk, qp, fl, st, ns, n1, n2, keep = _wt_args(wt)
    Lmax = maxtransformlevels(n)
    best = np.zeros(len(tree), dtype=np.uint8)
    bf = np.zeros(len(tree), dtype=np.float64)
    af = np.zeros(1 << max(Lmax - 1, 0), dtype=np.float64)
    pu8, pd = C.POINTER(C.c_uint8), C.POINTER(C.c_double)
    with torch.cuda.device(y.device):
        rc = _lib.lib().wb200_bestbasistree(best.ctypes.data_as(pu8), bf.ctypes.data_as(pd), af.ctypes.data_as(pd), y.data_ptr(), n, wk, qp,
                                            fl, st, ns, n1, n2, tree.ctypes.data_as(pu8), len(tree), et.kind, _DTYPES[y.dtype], _stream(y),
                                            _flags_value())
    _check(rc)
    return (best, bf, af) if return_entropies else best
