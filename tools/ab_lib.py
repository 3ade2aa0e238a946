#!/usr/bin/env python3
"""Interleaved A/B of two BUILDS of the library on one workload (both .so files are loaded into the same process and the
binding is re-pointed between measurements): six alternating rounds, medians reported, outputs of the two builds compared
bit for bit.  A sequential comparison is meaningless on a board whose clock drifts under its power cap.
    python tools/ab_lib.py KIND [libA.so libB.so]     KIND: filt1d | filt1d64 | filt1dbig | lift1d | wpt | fir3d | lift3d | lift2d | fir2d | modwt
    (defaults: wavelets.jl_b200/lib/libwavelets_b200_base.so  vs  wavelets.jl_b200/lib/libwavelets_b200.so)"""
import os, sys, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import wavelets_b200 as wb
from wavelets_b200 import _lib

kind = sys.argv[1]
libdir = os.path.join(ROOT, "wavelets.jl_b200", "lib")
libs = sys.argv[2:4] if len(sys.argv) >= 4 else [os.path.join(libdir, "libwavelets_b200_base.so"), os.path.join(libdir, "libwavelets_b200.so")]
handles = {}


def use(path):
    if path not in handles:
        _lib.LIB_PATH, _lib._lib = path, None
        handles[path] = _lib.lib()
    _lib.LIB_PATH, _lib._lib = path, handles[path]


dev = "cuda"
REPS = 5
if kind in ("filt1d", "filt1d64", "filt1dbig"):
    dt = torch.float64 if kind == "filt1d64" else torch.float32
    cols = {"filt1d": 2048, "filt1d64": 1024, "filt1dbig": 8192}[kind]       # filt1dbig: the bench's own batch, long enough to sit at the power cap
    if kind == "filt1dbig": REPS = 12
    wt = wb.wavelet(wb.WT.db4); x = torch.randn((cols, 1 << 20), device=dev, dtype=dt).t()
    mk = lambda: ((lambda: wb.dwtc(x, wt)), (lambda Y: wb.idwtc(Y, wt)))
elif kind == "lift1d":
    wt = wb.wavelet(wb.WT.cdf97, wb.WT.Lifting); x = torch.randn((2048, 1 << 20), device=dev).t()
    mk = lambda: ((lambda: wb.dwtc(x, wt)), (lambda Y: wb.idwtc(Y, wt)))
elif kind == "wpt":
    wt = wb.wavelet(wb.WT.sym8); x = torch.randn((1024, 1 << 16), device=dev).t()
    mk = lambda: ((lambda: wb.wpt(x, wt)), (lambda Y: wb.iwpt(Y, wt)))
elif kind in ("fir3d", "lift3d"):
    wt = wb.wavelet(wb.WT.db6) if kind == "fir3d" else wb.wavelet(wb.WT.cdf97, wb.WT.Lifting)
    x = torch.randn((512, 512, 512), device=dev).permute(2, 1, 0)
    mk = lambda: ((lambda: wb.dwt(x, wt, 3)), (lambda Y: wb.idwt(Y, wt, 3)))
elif kind == "modwt":
    wt = wb.wavelet(wb.WT.db4); x = torch.randn((64, 1 << 20), device=dev).t()
    mk = lambda: ((lambda: wb.modwt(x, wt, 10)), (lambda Y: wb.imodwt(Y, wt)))
else:
    wt = wb.wavelet(wb.WT.cdf97, wb.WT.Lifting) if kind == "lift2d" else wb.wavelet(wb.WT.db4)
    x = torch.randn((64, 4096, 4096), device=dev).permute(2, 1, 0)
    mk = lambda: ((lambda: wb.dwtc(x, wt, 8)), (lambda Y: wb.idwtc(Y, wt, 8)))
fwd, inv = mk()


def timeit(fn, reps=None):
    reps = reps or REPS
    fn(); fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


BIG = kind.endswith("big")       # 32 GiB operands: no room for copies of both builds' outputs (the small kinds compare them)
outs = {}
Y0 = None
for p in libs:
    use(p)
    Y = fwd(); X = inv(Y)
    if not BIG:
        print(f"{os.path.basename(p):36s} round trip max err {float((X - x).abs().max()):.3e}")
        outs[p] = (Y.clone(), X.clone())
    if Y0 is None: Y0 = Y
    del X
if not BIG:
    same = all(torch.equal(outs[libs[0]][i], outs[libs[1]][i]) for i in range(2))
    print("outputs of the two builds bit-identical:", same)
del outs, Y
res = {p: ([], []) for p in libs}
for rnd in range(6):
    for p in libs:
        use(p)
        res[p][0].append(timeit(fwd)); res[p][1].append(timeit(lambda: inv(Y0)))
for p, (f, i) in res.items():
    print(f"{kind:8s} {os.path.basename(p):36s} fwd {statistics.median(f):8.4f} ms   inv {statistics.median(i):8.4f} ms   pair {statistics.median(f) + statistics.median(i):8.4f} ms")
