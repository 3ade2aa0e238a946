#!/usr/bin/env python3
"""1-D lifting (cdf97, N = 2^20, L = 20, a batch of columns): per-kernel device times and whole-call event times."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import wavelets_b200 as wb
from wavelets_b200 import _lib
L = _lib.lib()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
dt = torch.float64 if (len(sys.argv) > 2 and sys.argv[2] == "f64") else torch.float32
wl = wb.wavelet(wb.WT.cdf97, wb.WT.Lifting)
x = torch.randn((B, 1 << 20), device="cuda", dtype=dt).t()
for _ in range(3):
    y = wb.dwtc(x, wl); xr = wb.idwtc(y, wl)
torch.cuda.synchronize()
L.wb200_profile_enable(1)
reps = 5
for _ in range(reps):
    y = wb.dwtc(x, wl); xr = wb.idwtc(y, wl)
torch.cuda.synchronize(); L.wb200_profile_enable(0)
buf = C.create_string_buffer(1 << 14); nb = L.wb200_profile_collect(buf, len(buf))
for ln in buf.raw[:nb].decode().splitlines():
    nm, c, ms = ln.split(); print(f"  {nm:28s} launches/rep {int(c) // reps:3d}  ms/rep {float(ms) / reps:8.4f}")
e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
e[0].record()
for _ in range(reps): y = wb.dwtc(x, wl)
e[1].record()
for _ in range(reps): xr = wb.idwtc(y, wl)
e[2].record(); torch.cuda.synchronize()
f, i = e[0].elapsed_time(e[1]) / reps, e[1].elapsed_time(e[2]) / reps
b = 2 * x.element_size() * B * (1 << 20) / 1e9
print(f"lift1d cdf97 2^20 x{B} {dt}: fwd {f:.3f} ms {b / f * 1e3:.0f} GB/s  inv {i:.3f} ms {b / i * 1e3:.0f} GB/s  rt err {float((xr - x).abs().max()):.2e}")
