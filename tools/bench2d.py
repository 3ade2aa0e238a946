#!/usr/bin/env python3
"""2-D workloads (4096^2 images, L=8): per-kernel device times through the library's profiling hook AND whole-call times
with CUDA events (the difference is allocation / launch overhead), for cdf97 lifting and an orthogonal filter bank.
usage: bench2d.py [B_f32 [B_f64]] [--sweep] [--wavelets cdf97,db4,...]"""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import wavelets_b200 as wb
from wavelets_b200 import _lib
L = _lib.lib()
args = [a for a in sys.argv[1:] if not a.startswith("--")]
Bf32 = int(args[0]) if len(args) > 0 else 16
Bf64 = int(args[1]) if len(args) > 1 else 8
sweep = "--sweep" in sys.argv
names = "cdf97,db4"
for i, a in enumerate(sys.argv):
    if a == "--wavelets":
        names = sys.argv[i + 1]
        args = [q for q in args if q != names]


def timed(fn, reps=5, warm_ms=100.0):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    per = max(e0.elapsed_time(e1), 1e-3)
    for _ in range(int(min(100, warm_ms / per))):
        fn()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def wavelet(name):
    if name == "cdf97":
        return wb.wavelet(wb.WT.cdf97, wb.WT.Lifting)
    return wb.wavelet(getattr(wb.WT, name))


for dt, B in ((torch.float32, Bf32), (torch.float64, Bf64)):
    if B <= 0:
        continue
    x = torch.randn((B, 4096, 4096), dtype=dt, device='cuda').permute(2, 1, 0)
    esz = x.element_size(); bytes_pass = 2 * esz * 4096 * 4096 * B
    for name in names.split(","):
        wl = wavelet(name)
        for _ in range(3):
            y = wb.dwtc(x, wl, 8); xr = wb.idwtc(y, wl, 8)
        torch.cuda.synchronize()
        L.wb200_profile_enable(1)
        for _ in range(5):
            y = wb.dwtc(x, wl, 8); xr = wb.idwtc(y, wl, 8)
        torch.cuda.synchronize(); L.wb200_profile_enable(0)
        buf = C.create_string_buffer(1 << 14); nb = L.wb200_profile_collect(buf, len(buf))
        tot = {}
        for ln in buf.raw[:nb].decode().splitlines():
            nm, c, ms = ln.split(); tot[nm] = (int(c), round(float(ms) / 5, 4))
        fwd = sum(v[1] for k, v in tot.items() if 'fwd' in k or 'forward' in k or 'analysis' in k)
        inv = sum(v[1] for k, v in tot.items() if 'inv' in k or 'synthesis' in k)
        print(dt, name, 'B', B, tot)
        print('  kernels only: fwd GB/s', round(bytes_pass / fwd / 1e6, 1), 'inv GB/s', round(bytes_pass / inv / 1e6, 1), 'pair GB/s',
              round(2 * bytes_pass / (fwd + inv) / 1e6, 1), 'rt', float((xr - x).abs().max()), flush=True)
        tf = timed(lambda: wb.dwtc(x, wl, 8)); ti = timed(lambda: wb.idwtc(y, wl, 8))
        tp = timed(lambda: wb.idwtc(wb.dwtc(x, wl, 8), wl, 8))
        print(f'  whole calls (events): fwd {tf:.3f} ms inv {ti:.3f} ms pair-loop {tp:.3f} ms -> pair GB/s {2 * bytes_pass / tp / 1e6:.1f}', flush=True)
        if sweep:
            prev = (0.0, 0.0)
            for lv in range(1, 9):
                tf = timed(lambda: wb.dwtc(x, wl, lv))
                yy = wb.dwtc(x, wl, lv)
                ti = timed(lambda: wb.idwtc(yy, wl, lv))
                print(f'  L={lv}: fwd {tf:.4f} ms (+{tf - prev[0]:.4f})  inv {ti:.4f} ms (+{ti - prev[1]:.4f})  pair GB/s {2 * bytes_pass / (tf + ti) / 1e6:.1f}', flush=True)
                prev = (tf, ti)
        del y, xr
    del x
    torch.cuda.empty_cache()
