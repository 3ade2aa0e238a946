#!/usr/bin/env python3
"""One dwt + idwt call of a chosen workload (for ncu / compute-sanitizer captures).
    python tools/run_once.py --kind filter1d --dtype f32 --batch 512
    python tools/run_once.py --kind lift2d --dtype f32 --batch 4 --n 4096 --levels 8"""
import argparse, ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import wavelets_b200 as wb
from wavelets_b200 import _lib

ap = argparse.ArgumentParser()
ap.add_argument("--kind", default="filter1d", choices=["filter1d", "lift1d", "lift2d", "filter2d", "wpt", "filter3d", "modwt"])
ap.add_argument("--dtype", default="f32")
ap.add_argument("--batch", type=int, default=512)
ap.add_argument("--n", type=int, default=0)
ap.add_argument("--levels", type=int, default=0)
ap.add_argument("--wavelet", default="")
ap.add_argument("--reps", type=int, default=1)
ap.add_argument("--strict", type=int, default=0)
a = ap.parse_args()
dev = torch.device("cuda:0")
tdt = torch.float32 if a.dtype == "f32" else torch.float64
wb.set_strict_fp(bool(a.strict))
if a.kind == "filter1d":
    n = a.n or (1 << 20)
    wt = wb.wavelet(getattr(wb.WT, a.wavelet or "db4"))
    x = torch.randn((a.batch, n), dtype=tdt, device=dev).t()
    for _ in range(a.reps):
        y = wb.dwtc(x, wt, a.levels or None); xr = wb.idwtc(y, wt, a.levels or None)
elif a.kind == "lift1d":
    n = a.n or (1 << 20)
    wt = wb.wavelet(wb.WT.cdf97, wb.WT.Lifting)
    x = torch.randn((a.batch, n), dtype=tdt, device=dev).t()
    for _ in range(a.reps):
        y = wb.dwtc(x, wt, a.levels or None); xr = wb.idwtc(y, wt, a.levels or None)
elif a.kind in ("lift2d", "filter2d"):
    n = a.n or 4096
    wt = wb.wavelet(wb.WT.cdf97, wb.WT.Lifting) if a.kind == "lift2d" else wb.wavelet(getattr(wb.WT, a.wavelet or "db4"))
    x = torch.randn((a.batch, n, n), dtype=tdt, device=dev).permute(2, 1, 0)
    for _ in range(a.reps):
        y = wb.dwtc(x, wt, a.levels or 8); xr = wb.idwtc(y, wt, a.levels or 8)
elif a.kind == "filter3d":
    n = a.n or 512
    wt = wb.wavelet(getattr(wb.WT, a.wavelet or "db6"))
    x = torch.randn((n, n, n), dtype=tdt, device=dev).permute(2, 1, 0)
    for _ in range(a.reps):
        y = wb.dwt(x, wt, a.levels or 3); xr = wb.idwt(y, wt, a.levels or 3)
elif a.kind == "modwt":
    n = a.n or (1 << 20)
    wt = wb.wavelet(getattr(wb.WT, a.wavelet or "db4"))
    x = torch.randn((a.batch, n), dtype=tdt, device=dev).t()
    for _ in range(a.reps):
        y = wb.modwt(x, wt, a.levels or 10); xr = wb.imodwt(y, wt)
else:
    n = a.n or (1 << 16)
    wt = wb.wavelet(getattr(wb.WT, a.wavelet or "sym8"))
    x = torch.randn((a.batch, n), dtype=tdt, device=dev).t()
    for _ in range(a.reps):
        y = wb.wpt(x, wt); xr = wb.iwpt(y, wt)
torch.cuda.synchronize()
print("max roundtrip err", float((xr - x).abs().max()))
