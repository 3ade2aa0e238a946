#!/usr/bin/env python3
"""CPU mirror of the shared-memory layouts of k_wpt_sub_ana / k_wpt_sub_syn (csrc/fastpass.cu), checked against the
oracle's full packet tree.  Development aid for the index arithmetic; not part of the product or the tests."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import oracle as orc
import wavelets_b200 as wb


def ana(x, h, g, levels):
    m = len(x); mh = m // 2; F = len(h)
    cur = np.empty(m); cur[:mh] = x[0::2]; cur[mh:] = x[1::2]
    for l in range(levels):
        ml = m >> l; nh = ml // 2; hh = nh // 2; last = l == levels - 1
        out = np.full(m, np.nan)
        for j in range(m // ml):
            E = cur[j * nh:(j + 1) * nh]; O = cur[mh + j * nh: mh + (j + 1) * nh]
            xx = np.empty(ml); xx[0::2] = E; xx[1::2] = O
            for k in range(nh):
                a = sum(h[t] * xx[(2 * k + t) % ml] for t in range(F))
                d = sum(g[F - 1 - t] * xx[(2 * k + 2 - F + t) % ml] for t in range(F))
                if not last:
                    out[(k & 1) * mh + (2 * j) * hh + (k >> 1)] = a
                    out[(k & 1) * mh + (2 * j + 1) * hh + (k >> 1)] = d
                else:
                    out[j * ml + k] = a; out[j * ml + nh + k] = d
        assert not np.isnan(out).any()
        cur = out
    return cur


def syn(y, h, g, levels):
    m = len(y); mh = m // 2; F = len(h); Q = F // 2
    ml = m >> (levels - 1); nh = ml // 2
    cur = np.full(m, np.nan)
    for i in range(m):
        j, r = divmod(i, ml)
        cur[(0 if r < nh else mh - nh) + j * nh + r] = y[i]
    for l in range(levels - 1, -1, -1):
        ml = m >> l; nh = ml // 2
        out = np.full(m, np.nan)
        for j in range(m // ml):
            a = cur[j * nh:(j + 1) * nh]; d = cur[mh + j * nh: mh + (j + 1) * nh]
            o = (j & 1) * mh + (j >> 1) * ml
            for u in range(nh):
                x0 = sum(h[2 * t] * a[(u - t) % nh] for t in range(Q)) + sum(g[2 * t + 1] * d[(u + t) % nh] for t in range(Q))
                x1 = sum(h[2 * t + 1] * a[(u - t) % nh] for t in range(Q)) + sum(g[2 * t] * d[(u + t) % nh] for t in range(Q))
                out[o + 2 * u] = x0; out[o + 2 * u + 1] = x1
        assert not np.isnan(out).any()
        cur = out
    return cur


if __name__ == "__main__":
    for wn in ("haar", "db2", "db4", "sym8"):
        wt = wb.wavelet(getattr(wb.WT, wn)); q = np.asarray(wt.qmf, dtype=np.float64); F = len(q)
        h = q.copy(); g = np.array([(-1) ** i * q[i] for i in range(F)])
        for m, lv in ((64, 6), (64, 3), (48, 4), (8, 3), (4, 2), (2, 1), (96, 5)):
            x = np.random.default_rng(m + lv).standard_normal(m)
            tree = wb.maketree(m, lv, "full")
            ref = orc.wpt_filter(x, q, tree)
            e1 = np.abs(ana(x, h, g, lv) - ref).max()
            e2 = np.abs(syn(ref, h, g, lv) - x).max()
            print(wn, m, lv, f"{e1:.1e} {e2:.1e}")
