#!/usr/bin/env python3
"""Interleaved A/B of the two tile configurations of the fused 2-D filter-bank level kernels (WB200_FIR_SMALL = 0 / 1)."""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import wavelets_b200 as wb
name = sys.argv[1] if len(sys.argv) > 1 else "db4"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
wt = wb.wavelet(getattr(wb.WT, name))
x = torch.randn((B, 4096, 4096), device="cuda").permute(2, 1, 0)
def timeit(reps=4):
    y = wb.dwtc(x, wt, 8); wb.idwtc(y, wt, 8)
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    torch.cuda.synchronize(); e[0].record()
    for _ in range(reps): y = wb.dwtc(x, wt, 8)
    e[1].record()
    for _ in range(reps): wb.idwtc(y, wt, 8)
    e[2].record(); torch.cuda.synchronize()
    return e[0].elapsed_time(e[1]) / reps, e[1].elapsed_time(e[2]) / reps
res = {"0": [], "1": []}
for rnd in range(5):
    for v in ("0", "1"):
        os.environ["WB200_FIR_SMALL"] = v
        res[v].append(timeit())
b = 2 * 4 * B * 4096 * 4096 / 1e9
for v, r in res.items():
    f = statistics.median(t[0] for t in r); i = statistics.median(t[1] for t in r)
    print(f"{name} x{B} FIR_SMALL={v}: fwd {f:.3f} ms {b / f * 1e3:.0f} GB/s  inv {i:.3f} ms {b / i * 1e3:.0f} GB/s  pair {2 * b / (f + i) * 1e3:.0f} GB/s")
