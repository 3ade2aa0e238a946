#!/usr/bin/env python3
"""denoise (plain on 2^24 samples, translation-invariant on 1024^2 with 8 x 8 spins): per-kernel device times and call time."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import wavelets_b200 as wb
from wavelets_b200 import _lib
L = _lib.lib()


def prof(fn, reps=5):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    L.wb200_profile_enable(1)
    for _ in range(reps):
        fn()
    torch.cuda.synchronize(); L.wb200_profile_enable(0)
    buf = C.create_string_buffer(1 << 14); nb = L.wb200_profile_collect(buf, len(buf))
    for ln in buf.raw[:nb].decode().splitlines():
        nm, c, ms = ln.split(); print(f"    {nm:28s} launches {int(c) // reps:3d}  ms {float(ms) / reps:8.4f}")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(20):
        fn()
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps * 4):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (reps * 4)


x = torch.randn(1 << 24, device="cuda")
print("denoise sym5 2^24 f32 L=6: %.4f ms per call" % prof(lambda: wb.denoise(x)))
xi = torch.randn((1024, 1024), device="cuda")
print("denoise TI sym5 1024^2 f32 8x8 spins: %.4f ms per call" % prof(lambda: wb.denoise(xi, TI=True)))
