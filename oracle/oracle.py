"""ctypes front-end of the CPU oracle (oracle/libwavelets_oracle.so).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / reference arm.  The product package never imports this.

Arrays are numpy, column-major (Julia layout): an array of Julia size (m, n, d) is
passed as a Fortran-ordered ndarray of shape (m, n, d).  A trailing batch
dimension (slices back to back) is supported by the *_batch functions.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libwavelets_oracle.so")

ORC_OK, ORC_EDIMS, ORC_ELEVEL, ORC_EPOW2, ORC_EALIAS, ORC_ENOTCUBE, ORC_ETREE = range(7)
ERRORS = {
    ORC_EDIMS: "in and out array size must match",
    ORC_ELEVEL: "L must be positive",
    ORC_EPOW2: "size must have a sufficient power of 2 factor",
    ORC_EALIAS: "in array is out array",
    ORC_ENOTCUBE: "array must be square/cube",
    ORC_ETREE: "invalid tree",
}
MAX_COEF = 8


class OracleError(ValueError):
    def __init__(self, code):
        super().__init__(ERRORS.get(code, f"oracle error {code}"))
        self.code = code


class Step(C.Structure):
    _fields_ = [("is_predict", C.c_int32), ("shift", C.c_int32), ("nc", C.c_int32),
                ("coef", C.c_double * MAX_COEF)]


def build(force: bool = False) -> str:
    """Compile the oracle with gcc (recipe: oracle/Makefile)."""
    src_m = max(os.path.getmtime(os.path.join(_HERE, f)) for f in ("wavelets_oracle.c", "oracle_impl.inc"))
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < src_m:
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB)
    return _lib


def _sfx(dtype):
    dtype = np.dtype(dtype)
    if dtype == np.float64:
        return "_f64", C.c_double
    if dtype == np.float32:
        return "_f32", C.c_float
    raise TypeError(f"oracle supports float32/float64, got {dtype}")


def make_steps(steps):
    """steps: iterable of objects/tuples (steptype|is_predict, coef, shift) -> ctypes array."""
    arr = (Step * len(steps))()
    for i, s in enumerate(steps):
        if hasattr(s, "steptype"):
            is_p, coef, shift = (s.steptype == "predict"), s.coef, s.shift
        else:
            is_p, coef, shift = s
        arr[i].is_predict = 1 if is_p else 0
        arr[i].shift = int(shift)
        arr[i].nc = len(coef)
        for k, c in enumerate(coef):
            arr[i].coef[k] = float(c)
    return arr


def _dims(shape):
    return (C.c_int64 * 3)(*(list(shape) + [1] * (3 - len(shape))))


def _check(rc):
    if rc != ORC_OK:
        raise OracleError(rc)


def _fcopy(x):
    x = np.asarray(x)
    return np.array(x, dtype=x.dtype, order="F", copy=True)


def dwt_filter(x, qmf, L, fw=True):
    """Reference `dwt(x, OrthoFilter, L)` / `idwt` (1-D/2-D/3-D), out of place."""
    x = _fcopy(x)
    sfx, ct = _sfx(x.dtype)
    y = np.empty_like(x, order="F")
    q = np.ascontiguousarray(qmf, dtype=np.float64)
    rc = getattr(lib(), "orc_dwt_filter" + sfx)(
        y.ctypes.data_as(C.c_void_p), x.ctypes.data_as(C.c_void_p), C.c_int(x.ndim), _dims(x.shape),
        q.ctypes.data_as(C.POINTER(C.c_double)), C.c_int(len(q)), C.c_int(int(L)), C.c_int(1 if fw else 0))
    _check(rc)
    return y


def dwt_lifting(x, steps, norm1, norm2, L, fw=True):
    """Reference `dwt(x, GLS, L)` / `idwt`: copy then in-place lifting transform."""
    y = _fcopy(x)
    sfx, ct = _sfx(y.dtype)
    st = make_steps(steps)
    rc = getattr(lib(), "orc_dwt_lifting" + sfx)(
        y.ctypes.data_as(C.c_void_p), C.c_int(y.ndim), _dims(y.shape), st, C.c_int(len(st)),
        C.c_double(norm1), C.c_double(norm2), C.c_int(int(L)), C.c_int(1 if fw else 0))
    _check(rc)
    return y


def wpt_filter(x, qmf, tree, fw=True):
    x = np.ascontiguousarray(x)
    sfx, ct = _sfx(x.dtype)
    y = np.empty_like(x)
    q = np.ascontiguousarray(qmf, dtype=np.float64)
    t = np.ascontiguousarray(tree, dtype=np.uint8)
    rc = getattr(lib(), "orc_wpt_filter" + sfx)(
        y.ctypes.data_as(C.c_void_p), x.ctypes.data_as(C.c_void_p), C.c_int64(x.shape[0]),
        q.ctypes.data_as(C.POINTER(C.c_double)), C.c_int(len(q)),
        t.ctypes.data_as(C.c_void_p), C.c_int64(len(t)), C.c_int(1 if fw else 0))
    _check(rc)
    return y


def wpt_lifting(x, steps, norm1, norm2, tree, fw=True):
    y = np.array(x, copy=True)
    sfx, ct = _sfx(y.dtype)
    st = make_steps(steps)
    t = np.ascontiguousarray(tree, dtype=np.uint8)
    rc = getattr(lib(), "orc_wpt_lifting" + sfx)(
        y.ctypes.data_as(C.c_void_p), C.c_int64(y.shape[0]), st, C.c_int(len(st)),
        C.c_double(norm1), C.c_double(norm2), t.ctypes.data_as(C.c_void_p), C.c_int64(len(t)),
        C.c_int(1 if fw else 0))
    _check(rc)
    return y


def dwt_filter_batch(x, ndim, qmf, L, fw=True, nthreads=0):
    """Independent transform of every slice along the last dimension."""
    x = _fcopy(x)
    assert x.ndim == ndim + 1
    sfx, ct = _sfx(x.dtype)
    y = np.empty_like(x, order="F")
    q = np.ascontiguousarray(qmf, dtype=np.float64)
    rc = getattr(lib(), "orc_batch_dwt_filter" + sfx)(
        y.ctypes.data_as(C.c_void_p), x.ctypes.data_as(C.c_void_p), C.c_int(ndim), _dims(x.shape[:ndim]),
        C.c_int64(x.shape[-1]), q.ctypes.data_as(C.POINTER(C.c_double)), C.c_int(len(q)),
        C.c_int(int(L)), C.c_int(1 if fw else 0), C.c_int(int(nthreads)))
    _check(rc)
    return y


def dwt_lifting_batch(x, ndim, steps, norm1, norm2, L, fw=True, nthreads=0):
    y = _fcopy(x)
    assert y.ndim == ndim + 1
    sfx, ct = _sfx(y.dtype)
    st = make_steps(steps)
    rc = getattr(lib(), "orc_batch_dwt_lifting" + sfx)(
        y.ctypes.data_as(C.c_void_p), C.c_int(ndim), _dims(y.shape[:ndim]), C.c_int64(y.shape[-1]),
        st, C.c_int(len(st)), C.c_double(norm1), C.c_double(norm2),
        C.c_int(int(L)), C.c_int(1 if fw else 0), C.c_int(int(nthreads)))
    _check(rc)
    return y


def maxtransformlevels(n: int) -> int:
    return int(lib().orc_maxtransformlevels(C.c_int64(int(n))))


def isvalidtree(n: int, tree) -> bool:
    t = np.ascontiguousarray(tree, dtype=np.uint8)
    return bool(lib().orc_isvalidtree(C.c_int64(int(n)), t.ctypes.data_as(C.c_void_p), C.c_int64(len(t))))


def modwt(x, qmf, L=None):
    """Reference `modwt(x, wt, L)`: returns the n x (L+1) matrix [W_1 .. W_L V_L]."""
    x = np.ascontiguousarray(x)
    sfx, ct = _sfx(x.dtype)
    n = x.shape[0]
    if L is None:
        L = int(np.floor(np.log2(n)))
    y = np.empty((n, L + 1), dtype=x.dtype, order="F")
    q = np.ascontiguousarray(qmf, dtype=np.float64)
    rc = getattr(lib(), "orc_modwt" + sfx)(y.ctypes.data_as(C.c_void_p), x.ctypes.data_as(C.c_void_p), C.c_int64(n),
                                            q.ctypes.data_as(C.POINTER(C.c_double)), C.c_int(len(q)), C.c_int(int(L)))
    _check(rc)
    return y


def imodwt(xw, qmf):
    xw = np.asfortranarray(xw)
    sfx, ct = _sfx(xw.dtype)
    n, ncols = xw.shape
    x = np.empty(n, dtype=xw.dtype)
    q = np.ascontiguousarray(qmf, dtype=np.float64)
    rc = getattr(lib(), "orc_imodwt" + sfx)(x.ctypes.data_as(C.c_void_p), xw.ctypes.data_as(C.c_void_p), C.c_int64(n),
                                             C.c_int(ncols), q.ctypes.data_as(C.POINTER(C.c_double)), C.c_int(len(q)))
    _check(rc)
    return x


# ---- Threshold / denoising ------------------------------------------------------------------
TH_KINDS = {"hard": 0, "soft": 1, "semisoft": 2, "stein": 3, "neg": 4, "pos": 5}


def _wt_args(wt):
    """wt: None | object with .qmf (OrthoFilter) | object with .step/.norm1/.norm2 (GLS) -> ctypes argument tuple"""
    null_d = C.POINTER(C.c_double)()
    if wt is None:
        return 0, null_d, 0, None, 0, 0.0, 0.0, ()
    if hasattr(wt, "qmf"):
        q = np.ascontiguousarray(wt.qmf, dtype=np.float64)
        return 1, q.ctypes.data_as(C.POINTER(C.c_double)), len(q), None, 0, 0.0, 0.0, (q,)
    st = make_steps(wt.step)
    return 2, null_d, 0, st, len(st), float(wt.norm1), float(wt.norm2), (st,)


def threshold(x, kind, t=0.0):
    """Reference `threshold(x, TH, t)`; kind in TH_KINDS."""
    y = _fcopy(x)
    sfx, ct = _sfx(y.dtype)
    getattr(lib(), "orc_threshold" + sfx)(y.ctypes.data_as(C.c_void_p), C.c_int64(y.size), C.c_int(TH_KINDS[kind]), C.c_double(float(t)))
    return y


def threshold_biggest(x, m):
    """Reference `threshold(x, BiggestTH(), m)` (ties at the cut dropped in index order)."""
    y = np.ascontiguousarray(np.array(x, copy=True).ravel(order="F"))
    sfx, ct = _sfx(y.dtype)
    getattr(lib(), "orc_threshold_biggest" + sfx)(y.ctypes.data_as(C.c_void_p), C.c_int64(y.size), C.c_int64(int(m)))
    return y.reshape(np.shape(x), order="F")


def noisest(x, wt, L=1):
    """Reference `noisest(x, wt, L)` (denoising.jl:94-101)."""
    x = _fcopy(x)
    sfx, ct = _sfx(x.dtype)
    wk, qp, fl, st, ns, n1, n2, keep = _wt_args(wt)
    sig = C.c_double(0.0)
    rc = getattr(lib(), "orc_noisest" + sfx)(C.byref(sig), x.ctypes.data_as(C.c_void_p), C.c_int(x.ndim), _dims(x.shape), C.c_int(wk),
                                             qp, C.c_int(fl), st, C.c_int(ns), C.c_double(n1), C.c_double(n2), C.c_int(int(L)))
    _check(rc)
    return sig.value


def denoise(x, wt, L, kind="hard", tfac=None, sigma=None, TI=False, nspin=8):
    """Reference `denoise(x, wt; L, dnt=VisuShrink(th, tfac), TI, nspin)`; sigma=None -> noisest."""
    x = _fcopy(x)
    sfx, ct = _sfx(x.dtype)
    if tfac is None:
        tfac = float(np.sqrt(2 * np.log(x.shape[0])))
    wk, qp, fl, st, ns, n1, n2, keep = _wt_args(wt)
    sp = [nspin] * x.ndim if isinstance(nspin, int) and x.ndim == 1 else ([nspin] + [1] * (x.ndim - 1) if isinstance(nspin, int) else list(nspin))
    spin = (C.c_int32 * 3)(*(sp + [1] * (3 - len(sp))))
    y = np.empty_like(x, order="F")
    rc = getattr(lib(), "orc_denoise" + sfx)(y.ctypes.data_as(C.c_void_p), x.ctypes.data_as(C.c_void_p), C.c_int(x.ndim), _dims(x.shape),
                                             C.c_int(wk), qp, C.c_int(fl), st, C.c_int(ns), C.c_double(n1), C.c_double(n2), C.c_int(int(L)),
                                             C.c_int(TH_KINDS[kind]), C.c_double(float(tfac)),
                                             C.c_double(float("nan") if sigma is None else float(sigma)), C.c_int(1 if TI else 0), spin)
    _check(rc)
    return y


# ---- best basis ---------------------------------------------------------------------------------
ET_KINDS = {"shannon": 0, "logenergy": 1}


def coefentropy(x, et="shannon", nrm=None):
    x = np.ascontiguousarray(x)
    sfx, ct = _sfx(x.dtype)
    if nrm is None:
        nrm = np.sqrt(np.sum(x.astype(np.float64) ** 2))
    f = getattr(lib(), "orc_coefentropy" + sfx)
    f.restype = ct
    return float(f(x.ctypes.data_as(C.c_void_p), C.c_int64(x.size), C.c_int(ET_KINDS[et]), ct(float(nrm))))


def bestbasistree(y, wt, tree, et="shannon"):
    """Reference `bestbasistree(y, wt, tree, et)` -> (besttree uint8, entr_bf, entr_af)."""
    y = np.ascontiguousarray(y)
    sfx, ct = _sfx(y.dtype)
    n = y.shape[0]
    t = np.ascontiguousarray(tree, dtype=np.uint8)
    best = t.copy()
    Lmax = maxtransformlevels(n)
    bf = np.zeros(len(t), dtype=y.dtype)
    af = np.zeros(1 << max(Lmax - 1, 0), dtype=y.dtype)
    wk, qp, fl, st, ns, n1, n2, keep = _wt_args(wt)
    rc = getattr(lib(), "orc_bestbasistree" + sfx)(best.ctypes.data_as(C.c_void_p), bf.ctypes.data_as(C.c_void_p), af.ctypes.data_as(C.c_void_p),
                                                   y.ctypes.data_as(C.c_void_p), C.c_int64(n), C.c_int(wk), qp, C.c_int(fl), st, C.c_int(ns),
                                                   C.c_double(n1), C.c_double(n2), t.ctypes.data_as(C.c_void_p), C.c_int64(len(t)),
                                                   C.c_int(ET_KINDS[et]))
    _check(rc)
    return best, bf, af
