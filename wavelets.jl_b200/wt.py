"""WT -- wavelet descriptors (host side).

Python mirror of the reference's `Wavelets.WT` module for the pieces the DWT hot
path consumes (src/WT/wt_main.jl): wavelet classes (`WT.db4`, `WT.sym8`,
`WT.cdf97` ...), the transform-type tags `WT.Filter` / `WT.Lifting`, the
`OrthoFilter` (qmf) and `GLS` (lifting step table) descriptors and the
`wavelet(class, type, boundary)` constructor (src/WT/wt_main.jl:262-264).

Descriptors stay on the host: only plain Float64 coefficient arrays cross the C
ABI (include/wavelets_b200.h), exactly as the Julia shim would pass
`wt.qmf` / `wt.step`.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Sequence, Tuple

import numpy as np

from ._filter_tables import FILTERS

__all__ = [
    "Filter", "Lifting", "Periodic", "WaveletClass", "OrthoFilter", "GLS", "LSStep",
    "wavelet", "daubechies", "Daubechies", "Coiflet", "Symlet", "Battle", "Haar",
    "Beylkin", "Vaidyanathan", "CDF",
]


# ---- transform / boundary tags (src/WT/wt_main.jl:23-49) -------------------------------------
class _Tag:
    def __init__(self, name):
        self._name = name

    def __repr__(self):
        return f"WT.{self._name}"


Filter = _Tag("Filter")
Lifting = _Tag("Lifting")
Periodic = _Tag("Periodic")
# Declared by the reference but never implemented by any transform:
padded = _Tag("padded")
NaivePer = _Tag("NaivePer")
SymBound = _Tag("SymBound")
DEFAULT_BOUNDARY = Periodic


# ---- wavelet classes (src/WT/wt_main.jl:52-128) ----------------------------------------------
@dataclass(frozen=True)
class WaveletClass:
    cls: str                 # "Daubechies", "Coiflet", ...
    namebase: str            # "db", "coif", ...
    moments: Tuple[int, ...] = ()
    ortho: bool = True

    @property
    def name(self) -> str:
        if self.namebase == "cdf":
            return f"cdf{self.moments[0]}/{self.moments[1]}"
        return self.namebase + ("".join(str(m) for m in self.moments))

    def __repr__(self):
        return f"WT.{self.name.replace('/', '')}"


def Daubechies(n: int) -> WaveletClass:
    return WaveletClass("Daubechies", "db", (int(n),))


def Coiflet(n: int) -> WaveletClass:
    return WaveletClass("Coiflet", "coif", (int(n),))


def Symlet(n: int) -> WaveletClass:
    return WaveletClass("Symlet", "sym", (int(n),))


def Battle(n: int) -> WaveletClass:
    return WaveletClass("Battle", "batt", (int(n),))


def Haar() -> WaveletClass:
    return WaveletClass("Haar", "haar")


def Beylkin() -> WaveletClass:
    return WaveletClass("Beylkin", "beyl")


def Vaidyanathan() -> WaveletClass:
    return WaveletClass("Vaidyanathan", "vaid")


def CDF(n1: int, n2: int) -> WaveletClass:
    return WaveletClass("CDF", "cdf", (int(n1), int(n2)), ortho=False)


haar = Haar()
beyl = Beylkin()
vaid = Vaidyanathan()
for _n in range(1, 11):
    globals()[f"db{_n}"] = Daubechies(_n)
for _n in (2, 4, 6, 8):
    globals()[f"coif{_n}"] = Coiflet(_n)
coif10 = Coiflet(10)  # Coiflet{10}() is constructible in the reference and has a FILTERS entry
for _n in range(4, 11):
    globals()[f"sym{_n}"] = Symlet(_n)
for _n in (2, 4, 6):
    globals()[f"batt{_n}"] = Battle(_n)
cdf97 = CDF(9, 7)


# ---- descriptors --------------------------------------------------------------------------------
@dataclass
class OrthoFilter:
    """Orthogonal filter-bank wavelet: `qmf` is the l2-normalised scaling filter h
    (src/WT/wt_main.jl:139-154)."""
    qmf: np.ndarray
    name: str
    boundary: object = Periodic

    def __len__(self):
        return int(self.qmf.shape[0])


@dataclass
class LSStep:
    """One lifting step (src/WT/wt_main.jl:195-209).  `steptype` is "predict" or "update";
    in the reference's convention Predict writes the first (even-sample) half."""
    steptype: str
    coef: List[float]
    shift: int


@dataclass
class GLS:
    """General lifting scheme (src/WT/wt_main.jl:224-229)."""
    step: List[LSStep]
    norm1: float
    norm2: float
    name: str
    boundary: object = Periodic


# lifting schemes, numeric content of WT.SCHEMES (src/WT/wt_main.jl:451-480)
def _schemes():
    a, b, c, d = 1.5861343420604, 0.05298011857291494, -0.882911075531393, -0.44350685204384654
    return {
        "cdf9/7": ([LSStep("update", [1.0 * a, 1.0 * a], 0),
                    LSStep("predict", [1.0 * b, 1.0 * b], 1),
                    LSStep("update", [1.0 * c, 1.0 * c], 0),
                    LSStep("predict", [1.0 * d, 1.0 * d], 1)],
                   1.1496043988603355, 0.8698644516247099),
        "haar": ([LSStep("predict", [-1.0], 0), LSStep("update", [0.5], 0)],
                 0.7071067811865475, 1.4142135623730951),
        "db1": ([LSStep("predict", [-1.0], 0), LSStep("update", [0.5], 0)],
                0.7071067811865475, 1.4142135623730951),
        "db2": ([LSStep("predict", [-1.7320508075688772], 0),
                 LSStep("update", [-0.0669872981077807, 0.4330127018922193], 1),
                 LSStep("predict", [1.0], -1)],
                0.5176380902050414, 1.9318516525781364),
    }


SCHEMES = _schemes()


# ---- Daubechies construction (src/WT/wt_main.jl:271-361) ----------------------------------------
def _vieta(roots: Sequence[complex]) -> np.ndarray:
    n = len(roots)
    C = np.zeros(n + 1, dtype=np.complex128)
    C[0] = 1
    for k in range(n):
        Ci = C[0]
        for i in range(k + 1):
            Cig = C[i + 1]
            C[i + 1] = Cig - roots[k] * Ci
            Ci = Cig
    return C


def daubechies(N: int) -> np.ndarray:
    """Daubechies scaling filter with N vanishing moments (2N taps) by the reference's recipe:
    truncated-binomial polynomial -> companion-matrix eigenvalues -> roots inside the unit circle
    -> Vieta -> l2 normalisation.  Last-ulp bits depend on the host LAPACK, which is
    why the C ABI takes the coefficients as an argument instead of baking its own."""
    assert N > 0
    C = np.array([math.comb(N - 1 + n, n) for n in range(N - 1, -1, -1)], dtype=np.float64)
    if N > 1:
        A = np.zeros((N - 1, N - 1))
        A[0, :] = -C[1:] / C[0]
        for i in range(N - 2):
            A[i + 1, i] = 1.0
        Y = np.linalg.eigvals(A)
    else:
        Y = np.zeros(0)
    Z = np.zeros(2 * N - 2, dtype=np.complex128)
    for i in range(N - 1):
        Yi = complex(Y[i])
        dd = 2 * np.sqrt(Yi * Yi - Yi)
        y2 = 1 - 2 * Yi
        Z[i] = y2 + dd
        Z[i + N - 1] = y2 - dd
    eps = np.finfo(np.float64).eps
    keep = [z for z in Z if abs(z) <= 1 + eps]
    R = [-1.0 + 0j] * N + keep
    HH = _vieta(R)
    HH = HH * (1 / np.linalg.norm(HH))
    return np.ascontiguousarray(HH.real)


def _ortho_filter(c: WaveletClass, boundary) -> OrthoFilter:
    name = c.name
    if c.cls == "Daubechies":
        q = daubechies(c.moments[0])
    else:
        if name not in FILTERS:
            raise ValueError("filter not found")          # ArgumentError("filter not found")
        q = np.array(FILTERS[name], dtype=np.float64)
    return OrthoFilter(q / np.linalg.norm(q), name, boundary)


def _gls(c: WaveletClass, boundary) -> GLS:
    name = c.name
    if name not in SCHEMES:
        raise ValueError("scheme not found")              # ArgumentError("scheme not found")
    steps, n1, n2 = SCHEMES[name]
    return GLS([LSStep(s.steptype, list(s.coef), s.shift) for s in steps], n1, n2, name, boundary)


def wavelet(c: WaveletClass, t=Filter, boundary=DEFAULT_BOUNDARY):
    """wavelet(c[, t=WT.Filter][, boundary=WT.Periodic]) -- src/WT/wt_main.jl:262-264.

    `wavelet(WT.cdf97)` raises like the reference (there is no CDF 9/7 filter pair)."""
    if not isinstance(c, WaveletClass):
        raise TypeError(f"no method wavelet({type(c).__name__}, ...)")  # MethodError in the reference
    if t is Filter:
        if not c.ortho:
            raise TypeError(f"no method wavelet({c!r}, WT.Filter): {c.name} exists only as a lifting scheme")
        return _ortho_filter(c, boundary)
    if t is Lifting:
        return _gls(c, boundary)
    raise TypeError("transform type must be WT.Filter or WT.Lifting")
