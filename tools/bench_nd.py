#!/usr/bin/env python3
"""3-D db6 512^3 (L=3) and WPT sym8 full tree: per-kernel device times (profiling hook) and whole-call event times."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import wavelets_b200 as wb
from wavelets_b200 import _lib
L = _lib.lib()


def hook(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    L.wb200_profile_enable(1)
    for _ in range(reps):
        fn()
    torch.cuda.synchronize(); L.wb200_profile_enable(0)
    buf = C.create_string_buffer(1 << 14); nb = L.wb200_profile_collect(buf, len(buf))
    tot = {}
    for ln in buf.raw[:nb].decode().splitlines():
        nm, c, ms = ln.split(); tot[nm] = (int(c) // reps, round(float(ms) / reps, 4))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(20):
        fn()
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return tot, e0.elapsed_time(e1) / reps


x3 = torch.randn((512, 512, 512), dtype=torch.float32, device='cuda').permute(2, 1, 0)
w6 = wb.wavelet(wb.WT.db6)
tot, ms = hook(lambda: wb.idwt(wb.dwt(x3, w6, 3), w6, 3))
print('3-D db6 512^3 L=3 pair', tot, f'{ms:.3f} ms -> {4 * 4 * 512**3 / ms / 1e6:.0f} GB/s', flush=True)
del x3
for B in (1024,):
    xp = torch.randn((B, 1 << 16), dtype=torch.float32, device='cuda').t()
    w8 = wb.wavelet(wb.WT.sym8)
    tot, ms = hook(lambda: wb.iwpt(wb.wpt(xp, w8), w8))
    print('WPT sym8 2^16 full tree B', B, tot, f'{ms:.3f} ms -> {4 * 4 * B * 65536 / ms / 1e6:.0f} GB/s', flush=True)
