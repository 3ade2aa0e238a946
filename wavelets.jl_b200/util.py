"""Util -- index helpers of the reference's `Wavelets.Util` that the transform API needs on the host
(src/Util/non_dyadic.jl, src/Util/util_main.jl:21-27, 301-344).  Pure Python, no array arithmetic."""
from __future__ import annotations

import numpy as np

__all__ = ["maxtransformlevels", "sufficientpoweroftwo", "detailindex", "detailrange", "detailn",
           "maketree", "isvalidtree", "iscube", "ndyadicscales", "isdyadic"]


def _sizes(x):
    if isinstance(x, (int, np.integer)):
        return (int(x),)
    return tuple(int(s) for s in x.shape)


def sufficientpoweroftwo(x, L: int) -> bool:
    return all(n % (2 ** L) == 0 for n in _sizes(x))


def maxtransformlevels(x) -> int:
    """min over dims of the largest L with dim % 2^L == 0 (non_dyadic.jl:14-22)."""
    def one(n):
        if n <= 1:
            return 0
        tl = 0
        while n % (2 ** tl) == 0:
            tl += 1
        return tl - 1
    return min(one(n) for n in _sizes(x))


def detailn(n, l: int) -> int:
    n = _sizes(n)[0]
    return int(round(n / 2 ** l))


def detailindex(n, l: int, i: int) -> int:
    """1-based vector index of detail coefficient i at level l (non_dyadic.jl:5)."""
    n = _sizes(n)[0]
    return int(round(n / 2 ** l + i))


def detailrange(n, l: int) -> range:
    """1-based inclusive range of the level-l detail coefficients (non_dyadic.jl:8)."""
    n = _sizes(n)[0]
    return range(int(round(n / 2 ** l + 1)), int(round(n / 2 ** (l - 1))) + 1)


def ndyadicscales(n) -> int:
    return int(round(np.log2(_sizes(n)[0])))


def isdyadic(x) -> bool:
    return all(n == 2 ** int(round(np.log2(n))) for n in _sizes(x))


def iscube(x) -> bool:
    s = _sizes(x)
    return all(n == s[0] for n in s)


def maketree(n, L=None, s: str = "full") -> np.ndarray:
    """maketree(n, L, s) (util_main.jl:322-344) -> uint8 array, one byte per node (1-based heap order)."""
    if not isinstance(n, (int, np.integer)):
        n = _sizes(n)[0]
    ns = maxtransformlevels(int(n))
    if L is None:
        L = ns
    if not (0 <= L <= ns):
        raise AssertionError("0 <= L <= maxtransformlevels(n)")
    b = np.zeros(2 ** ns - 1, dtype=np.uint8)
    if s == "full":
        b[: 2 ** L - 1] = 1
    elif s == "dwt":
        for i in range(1, L + 1):
            b[2 ** (i - 1) - 1] = 1
    else:
        raise ValueError("uknown symbol")
    return b


def isvalidtree(x, b) -> bool:
    n = _sizes(x)[0]
    ns = maxtransformlevels(n)
    b = np.asarray(b)
    if len(b) != 2 ** ns - 1:
        return False
    for i in range(1, 1 << max(ns - 1, 0)):
        if not b[i - 1] and (b[2 * i - 1] or b[2 * i]):
            return False
    return True
