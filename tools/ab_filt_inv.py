#!/usr/bin/env python3
"""Interleaved A/B of tuning configurations of the 1-D filter kernels (the board's clock drifts under its power cap, so the
configurations alternate and every one is measured `rounds` times).   python tools/ab_filt_inv.py [fwd|inv] [B]"""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import wavelets_b200 as wb
direction = sys.argv[1] if len(sys.argv) > 1 else "inv"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
f64 = len(sys.argv) > 3 and sys.argv[3] == "f64"
wl = wb.wavelet(wb.WT.db4)
x = torch.randn((B, 1 << 20), device="cuda", dtype=torch.float64 if f64 else torch.float32).t()
y = wb.dwtc(x, wl)
b = 2 * x.element_size() * B * (1 << 20) / 1e9
sfx = "_INV" if direction == "inv" else ""
if f64:
    KEYS = ["WB200_TILE_F64", "WB200_F1D_NT" + sfx, "WB200_KMAX" + sfx, "WB200_TAILMAX_F64"]
    grid = ((2048, 96, 4, 2048), (4096, 96, 4, 4096), (2048, 64, 4, 2048), (2048, 128, 4, 2048), (2048, 96, 3, 2048), (2048, 96, 5, 2048), (4096, 128, 5, 4096), (1024, 64, 3, 1024))
else:
    KEYS = ["WB200_TILE_F32" + sfx, "WB200_F1D_NT" + sfx, "WB200_KMAX" + sfx, "WB200_TAILMAX_F32" + sfx]
    grid = ((4096, 96, 4, 4096), (4096, 64, 4, 4096), (4096, 128, 4, 4096), (4096, 96, 3, 4096), (4096, 96, 5, 4096), (4096, 96, 4, 2048), (4096, 96, 4, 8192), (8192, 192, 4, 4096), (2048, 96, 4, 4096))
cfgs = {"default": {}}
for tile, nt, k, tm in grid:
    cfgs[f"tile{tile}_nt{nt}_k{k}_tail{tm}"] = dict(zip(KEYS, map(str, (tile, nt, k, tm))))
fn = (lambda: wb.idwtc(y, wl)) if direction == "inv" else (lambda: wb.dwtc(x, wl))
def timeit(reps=5):
    fn(); fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
res = {k: [] for k in cfgs}
for rnd in range(6):
    for name, env in cfgs.items():
        for k in KEYS: os.environ.pop(k, None)
        os.environ.update(env)
        res[name].append(timeit())
for name, v in sorted(res.items(), key=lambda kv: statistics.median(kv[1])):
    print(f"{direction} {name:32s} median {statistics.median(v):7.3f} ms {b / statistics.median(v) * 1e3:6.0f} GB/s   min {min(v):7.3f} max {max(v):7.3f}")
