#!/usr/bin/env python3
"""Interleaved A/B of fused 1-D lifting configurations (environment knobs), medians of 6 alternating measurements."""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import wavelets_b200 as wb
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
wl = wb.wavelet(wb.WT.cdf97, wb.WT.Lifting)
x = torch.randn((B, 1 << 20), device="cuda").t()
y = wb.dwtc(x, wl)
b = 2 * 4 * B * (1 << 20) / 1e9
cfgs = {"default": {}, "inv_staged": {"WB200_LIFT1D_INV_STAGED": "1"},
        "inv_staged_tile4096": {"WB200_LIFT1D_INV_STAGED": "1", "WB200_LIFT1D_TILE_F32_INV": "4096"},
        "inv_tile4096": {"WB200_LIFT1D_TILE_F32_INV": "4096"},
        "inv_k3": {"WB200_LIFT1D_KMAX": "3"}, "inv_nt64": {"WB200_LIFT1D_NT_INV": "64"}, "inv_nt128": {"WB200_LIFT1D_NT_INV": "128"},
        "tailmax4096": {"WB200_LIFT1D_TAILMAX": "4096"}}
keys = sorted({k for c in cfgs.values() for k in c})
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
        res[name][0].append(timeit(lambda: wb.dwtc(x, wl)))
        res[name][1].append(timeit(lambda: wb.idwtc(y, wl)))
for name, (f, i) in sorted(res.items(), key=lambda kv: statistics.median(kv[1][1])):
    mf, mi = statistics.median(f), statistics.median(i)
    print(f"{name:24s} fwd {mf:7.3f} ms {b / mf * 1e3:6.0f} GB/s   inv {mi:7.3f} ms {b / mi * 1e3:6.0f} GB/s")
