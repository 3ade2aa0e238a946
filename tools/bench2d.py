#!/usr/bin/env python3
"""Per-kernel timing of the 2-D cdf97 lifting workload (4096^2, L=8) through the library's profiling hook."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import wavelets_b200 as wb
from wavelets_b200 import _lib
L = _lib.lib()
wl = wb.wavelet(wb.WT.cdf97, wb.WT.Lifting)
for dt, B in ((torch.float32, 16), (torch.float64, 8)):
    x = torch.randn((B, 4096, 4096), dtype=dt, device='cuda').permute(2, 1, 0)
    for _ in range(2):
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
    esz = x.element_size(); bytes_pass = 2 * esz * 4096 * 4096 * B
    fwd = sum(v[1] for k, v in tot.items() if 'fwd' in k or 'forward' in k or 'analysis' in k)
    inv = sum(v[1] for k, v in tot.items() if 'inv' in k or 'synthesis' in k)
    print(dt, tot)
    print('  fwd GB/s', round(bytes_pass / fwd / 1e6, 1), 'inv GB/s', round(bytes_pass / inv / 1e6, 1), 'pair GB/s',
          round(2 * bytes_pass / (fwd + inv) / 1e6, 1), 'rt', float((xr - x).abs().max()))
