#!/usr/bin/env python3
"""Interleaved A/B of the L2 prefetch distance of the fused 1-D tile kernels (WB200_F1D_PREFETCH, in CTAs ahead).
    python tools/ab_prefetch.py [B] [f64] [lift]"""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import wavelets_b200 as wb
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
f64 = "f64" in sys.argv
lift = "lift" in sys.argv
wl = wb.wavelet(wb.WT.cdf97, wb.WT.Lifting) if lift else wb.wavelet(wb.WT.db4)
x = torch.randn((B, 1 << 20), device="cuda", dtype=torch.float64 if f64 else torch.float32).t()
y = wb.dwtc(x, wl)
gb = 2 * x.element_size() * B * (1 << 20) / 1e9
KEYS = ["WB200_LIFT1D_PREFETCH", "WB200_LIFT1D_PREFETCH_INV"] if lift else ["WB200_F1D_PREFETCH", "WB200_F1D_PREFETCH_INV"]
dists = [int(v) for v in os.environ.get("AB_DISTS", "0,296,592,888,1184,1480,1776").split(",")]
REPS = int(os.environ.get("AB_REPS", "5"))      # AB_REPS=12 with 8192 columns: long enough to sit at the power cap
def timeit(fn, reps=REPS):
    fn(); fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
res = {d: ([], []) for d in dists}
for rnd in range(6):
    for d in dists:
        for k in KEYS: os.environ[k] = str(d)
        res[d][0].append(timeit(lambda: wb.dwtc(x, wl))); res[d][1].append(timeit(lambda: wb.idwtc(y, wl)))
for d, (f, i) in res.items():
    mf, mi = statistics.median(f), statistics.median(i)
    print(f"prefetch {d:5d} CTAs ahead   fwd {mf:7.3f} ms {gb / mf * 1e3:6.0f} GB/s   inv {mi:7.3f} ms {gb / mi * 1e3:6.0f} GB/s   pair {mf + mi:7.3f} ms")
