#!/usr/bin/env python3
"""MODWT / IMODWT timing (db4): compulsory traffic = read n + write n*(L+1) elements forward, the reverse inverse."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import wavelets_b200 as wb
wt = wb.wavelet(wb.WT.db4)
for dt in (torch.float32, torch.float64):
    for n, B, L in ((1 << 20, 64, 10), (1 << 16, 1024, 10), (1 << 20, 64, 20)):
        x = torch.randn((B, n), dtype=dt, device='cuda').t()
        def timed(fn, reps=5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            fn(); torch.cuda.synchronize(); e0.record()
            for _ in range(reps): fn()
            e1.record(); torch.cuda.synchronize()
            return e0.elapsed_time(e1) / reps
        W = wb.modwt(x, wt, L)
        tf = timed(lambda: wb.modwt(x, wt, L)); ti = timed(lambda: wb.imodwt(W, wt))
        by = (L + 2) * n * B * x.element_size()
        rt = float((wb.imodwt(W, wt) - x).abs().max())
        print(f'{dt} n={n} B={B} L={L}: modwt {tf:.3f} ms {by / tf / 1e6:.0f} GB/s | imodwt {ti:.3f} ms {by / ti / 1e6:.0f} GB/s | '
              f'{n * B / (tf + ti) / 1e3:.0f} Msamples/s pair | rt {rt:.2e}', flush=True)
        del x, W
