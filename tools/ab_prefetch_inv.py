#!/usr/bin/env python3
"""Interleaved A/B of the INVERSE tile kernel's L2 prefetch restricted to its largest slices (WB200_F1D_PREFETCH_INV = distance,
WB200_F1D_PREFETCH_INV_LEVELS = deepest level prefetched).   python tools/ab_prefetch_inv.py [B]"""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import wavelets_b200 as wb
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
wl = wb.wavelet(wb.WT.db4)
x = torch.randn((B, 1 << 20), device="cuda").t()
y = wb.dwtc(x, wl)
gb = 2 * 4 * B * (1 << 20) / 1e9
cfgs = [(0, 8), (592, 1), (888, 1), (1184, 1), (888, 2), (296, 1)]
def timeit(fn, reps=5):
    fn(); fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
res = {c: [] for c in cfgs}
for rnd in range(6):
    for c in cfgs:
        os.environ["WB200_F1D_PREFETCH_INV"] = str(c[0]); os.environ["WB200_F1D_PREFETCH_INV_LEVELS"] = str(c[1])
        res[c].append(timeit(lambda: wb.idwtc(y, wl)))
for c, v in res.items():
    m = statistics.median(v)
    print(f"inverse prefetch {c[0]:5d} CTAs ahead, levels <= {c[1]}   {m:7.3f} ms {gb / m * 1e3:6.0f} GB/s")
