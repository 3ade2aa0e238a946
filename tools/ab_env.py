#!/usr/bin/env python3
"""Generic interleaved A/B of environment knobs (read at dispatch) on one workload: the configurations alternate, six
measurements each, medians reported -- a sequential sweep is meaningless on a board whose clock drifts under its power cap.
    python tools/ab_env.py modwt|wpt|fir3d|lift2d|fir2d  'NAME=V[,NAME2=V2]' 'NAME=V' ...      ('' = defaults)"""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import wavelets_b200 as wb
kind = sys.argv[1]
cfgs = {"default": {}}
for a in sys.argv[2:]:
    if a:
        cfgs[a] = dict(kv.split("=") for kv in a.split(","))
keys = sorted({k for c in cfgs.values() for k in c})
dev = "cuda"
if kind.startswith("modwt"):
    Lm = int(kind[5:] or 10)
    wt = wb.wavelet(wb.WT.db4); x = torch.randn((64 if Lm <= 10 else 16, 1 << 20), device=dev).t()
    W = wb.modwt(x, wt, Lm)
    fwd, inv = (lambda: wb.modwt(x, wt, Lm)), (lambda: wb.imodwt(W, wt))
elif kind == "wpt":
    wt = wb.wavelet(wb.WT.sym8); x = torch.randn((1024, 1 << 16), device=dev).t()
    Y = wb.wpt(x, wt)
    fwd, inv = (lambda: wb.wpt(x, wt)), (lambda: wb.iwpt(Y, wt))
elif kind == "lift3d":
    wt = wb.wavelet(wb.WT.cdf97, wb.WT.Lifting); x = torch.randn((512, 512, 512), device=dev).permute(2, 1, 0)
    Y = wb.dwt(x, wt, 3)
    fwd, inv = (lambda: wb.dwt(x, wt, 3)), (lambda: wb.idwt(Y, wt, 3))
elif kind == "fir3d":
    wt = wb.wavelet(wb.WT.db6); x = torch.randn((512, 512, 512), device=dev).permute(2, 1, 0)
    Y = wb.dwt(x, wt, 3)
    fwd, inv = (lambda: wb.dwt(x, wt, 3)), (lambda: wb.idwt(Y, wt, 3))
else:
    wt = wb.wavelet(wb.WT.cdf97, wb.WT.Lifting) if kind == "lift2d" else wb.wavelet(wb.WT.db4)
    x = torch.randn((64, 4096, 4096), device=dev).permute(2, 1, 0)
    Y = wb.dwtc(x, wt, 8)
    fwd, inv = (lambda: wb.dwtc(x, wt, 8)), (lambda: wb.idwtc(Y, wt, 8))
def timeit(fn, reps=5):
    fn(); fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
res = {k: ([], []) for k in cfgs}
for rnd in range(6):
    for name, env in cfgs.items():
        for k in keys: os.environ.pop(k, None)
        os.environ.update(env)
        res[name][0].append(timeit(fwd)); res[name][1].append(timeit(inv))
def kernel_times(fn, reps=5):
    """per-kernel device ms per call from the library's own event hook (serialises the launches it brackets)"""
    import ctypes as C
    from wavelets_b200 import _lib
    L = _lib.lib()
    buf = C.create_string_buffer(1 << 14)
    L.wb200_profile_collect(buf, len(buf))
    L.wb200_profile_enable(1)
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    L.wb200_profile_enable(0)
    nb = L.wb200_profile_collect(buf, len(buf))
    return {ln.split()[0]: (int(ln.split()[1]) // reps, float(ln.split()[2]) / reps) for ln in buf.raw[:nb].decode().splitlines()}
if os.environ.get("AB_KERNELS"):
    for name, env in cfgs.items():
        for k in keys: os.environ.pop(k, None)
        os.environ.update(env)
        for lab, fn in (("fwd", fwd), ("inv", inv)):
            print(f"  [{name}] {lab}: " + ", ".join(f"{k} x{c} {ms:.4f} ms" for k, (c, ms) in kernel_times(fn).items()))
for name, (f, i) in res.items():
    print(f"{kind:6s} {name:44s} fwd {statistics.median(f):8.4f} ms   inv {statistics.median(i):8.4f} ms   pair {statistics.median(f) + statistics.median(i):8.4f} ms")
