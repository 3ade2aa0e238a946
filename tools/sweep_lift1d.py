#!/usr/bin/env python3
"""Sweep the tuning knobs of the fused 1-D kernels (read from the environment at dispatch) in one process.
    python tools/sweep_lift1d.py [lift|filt] [B]"""
import itertools, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import wavelets_b200 as wb

kind = sys.argv[1] if len(sys.argv) > 1 else "lift"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
wl = wb.wavelet(wb.WT.cdf97, wb.WT.Lifting) if kind == "lift" else wb.wavelet(wb.WT.db4)
x = torch.randn((B, 1 << 20), device="cuda").t()
y = wb.dwtc(x, wl)
b = 2 * 4 * B * (1 << 20) / 1e9


def timeit(fn, reps=4):
    for _ in range(2):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


res = []
grid = itertools.product((2048, 4096, 8192, 16384), (0, 64, 96, 128, 192, 256, 384), (3, 4, 5, 6)) if kind == "lift" else \
    itertools.product((2048, 4096, 8192, 16384, 32768), (64, 96, 128, 192, 256, 384), (3, 4, 5, 6, 8))
for tile, nt, k in grid:
    if kind == "lift":
        P = "WB200_LIFT1D_"
        os.environ[P + "TILE_F32"] = str(tile); os.environ[P + "TILE_F32_INV"] = str(tile)
        os.environ[P + "NT"] = str(nt); os.environ[P + "NT_INV"] = str(nt); os.environ[P + "KMAX"] = str(k)
    else:
        os.environ["WB200_TILE_F32"] = str(tile); os.environ["WB200_TILE_F32_INV"] = str(tile)
        os.environ["WB200_F1D_NT"] = str(nt); os.environ["WB200_F1D_NT_INV"] = str(nt); os.environ["WB200_KMAX"] = str(k)
    f = timeit(lambda: wb.dwtc(x, wl)); i = timeit(lambda: wb.idwtc(y, wl))
    res.append((tile, nt, k, f, i))
    if os.environ.get("SWEEP_VERBOSE"):
        print(f"tile {tile:6d} nt {nt:4d} kmax {k}: fwd {f:7.3f} ms {b / f * 1e3:6.0f} GB/s   inv {i:7.3f} ms {b / i * 1e3:6.0f} GB/s", flush=True)
print("---- best forward")
for r in sorted(res, key=lambda r: r[3])[:8]:
    print(f"tile {r[0]:6d} nt {r[1]:4d} kmax {r[2]}: fwd {r[3]:7.3f} ms {b / r[3] * 1e3:6.0f} GB/s")
print("---- best inverse")
for r in sorted(res, key=lambda r: r[4])[:8]:
    print(f"tile {r[0]:6d} nt {r[1]:4d} kmax {r[2]}: inv {r[4]:7.3f} ms {b / r[4] * 1e3:6.0f} GB/s")
