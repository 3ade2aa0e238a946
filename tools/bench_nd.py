#!/usr/bin/env python3
"""Timing of the N-D filter-bank / wavelet-packet workloads of BASELINE.json configs[3..4] (CUDA events)."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import wavelets_b200 as wb
from wavelets_b200 import _lib
L = _lib.lib()

def timed(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

def kernels():
    buf = C.create_string_buffer(1 << 14); nb = L.wb200_profile_collect(buf, len(buf))
    return {ln.split()[0]: (int(ln.split()[1]), round(float(ln.split()[2]), 3)) for ln in buf.raw[:nb].decode().splitlines()}

def breakdown(fn):
    L.wb200_profile_enable(1); fn(); torch.cuda.synchronize(); L.wb200_profile_enable(0)
    print("    kernels:", kernels(), flush=True)

def report(name, samples, esz, ms_f, ms_i):
    gb = 2.0 * esz * samples
    print(f"{name}: fwd {ms_f:.3f} ms ({gb/ms_f/1e6:.0f} GB/s)  inv {ms_i:.3f} ms ({gb/ms_i/1e6:.0f} GB/s)  "
          f"pair {samples/(ms_f+ms_i)/1e3:.1f} Msamples/s  pair {2*gb/(ms_f+ms_i)/1e6:.0f} GB/s = {2*gb/(ms_f+ms_i)/1e6/6570:.3f} of HBM peak", flush=True)

dev = 'cuda'
# 3-D db6 512^3 f32, L=3 and L=9
wt = wb.wavelet(wb.WT.db6)
x = torch.randn((512, 512, 512), dtype=torch.float32, device=dev).permute(2, 1, 0)
for Lv in (3, 9):
    y = wb.dwt(x, wt, Lv)
    report(f"3-D db6 512^3 f32 L={Lv}", 512**3, 4, timed(lambda: wb.dwt(x, wt, Lv)), timed(lambda: wb.idwt(y, wt, Lv)))
    breakdown(lambda: wb.dwt(x, wt, Lv))
del x, y
# batched 2-D db4 filter 4096^2 x 16, L=8
wt = wb.wavelet(wb.WT.db4)
x = torch.randn((16, 4096, 4096), dtype=torch.float32, device=dev).permute(2, 1, 0)
y = wb.dwtc(x, wt, 8)
report("2-D db4 filter 4096^2 x16 f32 L=8", 16 * 4096**2, 4, timed(lambda: wb.dwtc(x, wt, 8)), timed(lambda: wb.idwtc(y, wt, 8)))
breakdown(lambda: wb.dwtc(x, wt, 8))
breakdown(lambda: wb.idwtc(y, wt, 8))
del x, y
# WPT sym8 full tree N=2^16, batch 1024
wt = wb.wavelet(wb.WT.sym8)
x = torch.randn((1024, 1 << 16), dtype=torch.float32, device=dev).t()
y = wb.wpt(x, wt)
report("WPT sym8 full tree 2^16 x1024 f32", 1024 * (1 << 16), 4, timed(lambda: wb.wpt(x, wt), 3), timed(lambda: wb.iwpt(y, wt), 3))
breakdown(lambda: wb.wpt(x, wt))
