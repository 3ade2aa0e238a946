#!/usr/bin/env python3
"""CPU mirror of csrc/modwt.cu's group plan and tile indexing (plan_modwt, k_modwt_group, k_imodwt_group), checked
against the oracle.  Development aid for the index arithmetic (no GPU needed); not part of the product or the tests."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import oracle as orc

CAP_BYTES = 36864


def plan(n, L, F, esz):
    steps = []
    cap = CAP_BYTES // esz; rows_cap = cap // 32
    ok = 2 <= F <= 20 and F % 2 == 0
    j = 0
    if ok:
        g = dict(n=n, j0=0, rs=32, dmul=1, chunks=1)
        if n <= cap:
            g.update(K=L, H=0, M=(n + 31) // 32, periodic=1, NQ=n, ntiles=1)
        else:
            K = 0
            while K < L and (F - 1) * ((2 << K) - 1) <= cap // 9: K += 1
            H = ((F - 1) * ((1 << K) - 1) + 4 * K + 31) // 32
            g.update(K=K, H=H, M=rows_cap - H, periodic=0, NQ=32 * rows_cap)
            g['ntiles'] = (n + 32 * g['M'] - 1) // (32 * g['M'])
        if g['K'] > 0:
            steps.append(('g', g)); j = g['K']
    while j < L:
        s0 = 1 << j
        if (not ok) or s0 < 32 or n % s0 != 0:
            steps.append(('l', j + 1)); j += 1; continue
        P = n // s0
        g = dict(n=n, j0=j, rs=s0, dmul=32, chunks=s0 // 32)
        if P * 32 <= cap:
            g.update(K=L - j, H=0, M=P, periodic=1, NQ=P * 32, ntiles=s0 // 32)
        else:
            K = 0
            while K < L - j and (F - 1) * ((2 << K) - 1) <= rows_cap // 3: K += 1
            if K == 0:
                steps.append(('l', j + 1)); j += 1; continue
            H = (F - 1) * ((1 << K) - 1)
            g.update(K=K, H=H, M=rows_cap - H, periodic=0, NQ=32 * rows_cap)
            g['ntiles'] = ((n + s0 * g['M'] - 1) // (s0 * g['M'])) * g['chunks']
        steps.append(('g', g)); j += g['K']
    return steps


def fwd_group(g, vin, y, h, gg, v4=False):
    n, F = g['n'], len(h)
    lo = 0
    vout = np.full(n, np.nan)
    q = np.arange(g['NQ'])
    for tile in range(g['ntiles']):
        rc, hi = tile % g['chunks'], tile // g['chunks']
        base = hi * g['rs'] * g['M'] + rc * 32
        gi = base + (q & 31) + g['rs'] * ((q >> 5) - g['H'])
        buf = vin[gi % n].copy()
        lo = 0
        own = (q >= 32 * g['H']) & (gi < n)
        for i in range(g['K']):
            d = g['dmul'] << i
            qlo = 0 if g['periodic'] else (F - 1) * ((2 << i) - 1) * g['dmul']
            if v4:      # the 128-bit kernels round every level's valid start up to a chunk
                lo = 0 if g['periodic'] else (lo + (F - 1) * d + 3) & ~3
                qlo = lo
                assert qlo <= 32 * g['H'] or g['periodic']
            qq = q[qlo:]
            w = np.zeros(len(qq)); a = np.zeros(len(qq))
            for k in range(F):
                if g['periodic']:
                    idx = qq - (k * d) % g['NQ']; idx = np.where(idx < 0, idx + g['NQ'], idx)
                else:
                    idx = qq - k * d
                assert idx.min() >= 0
                w += h[k] * buf[idx]; a += gg[k] * buf[idx]
            nb = np.full(g['NQ'], np.nan); nb[qlo:] = a
            sel = own[qlo:]
            y[gi[qlo:][sel], g['j0'] + i] = w[sel]
            buf = nb
        assert not np.isnan(buf[own]).any()
        vout[gi[own]] = buf[own]
    assert not np.isnan(vout).any()
    return vout


def inv_group(g, vin, xw, h, gg, v4=False):
    n, F = g['n'], len(h)
    vout = np.full(n, np.nan)
    q = np.arange(g['NQ'])
    for tile in range(g['ntiles']):
        rc, hi = tile % g['chunks'], tile // g['chunks']
        base = hi * g['rs'] * g['M'] + rc * 32
        gi = base + (q & 31) + g['rs'] * (q >> 5)
        buf = vin[gi % n].copy()
        own_hi = g['NQ'] if g['periodic'] else 32 * g['M']
        hi_r = g['NQ']
        for i in range(g['K'] - 1, -1, -1):
            d = g['dmul'] << i
            qhi = g['NQ'] if g['periodic'] else g['NQ'] - (F - 1) * ((1 << g['K']) - (1 << i)) * g['dmul']
            qload = g['NQ'] if g['periodic'] else qhi + (F - 1) * d
            if v4 and not g['periodic']:
                hi_r = (hi_r - (F - 1) * d) & ~3
                qhi = hi_r; qload = (qhi + (F - 1) * d + 3) & ~3
                assert qhi >= own_hi
            assert qload <= g['NQ']
            wb = np.full(g['NQ'], np.nan); wb[:qload] = xw[gi[:qload] % n, g['j0'] + i]
            qq = q[:qhi]
            acc = np.zeros(qhi)
            for k in range(F):
                if g['periodic']:
                    idx = qq + (k * d) % g['NQ']; idx = np.where(idx >= g['NQ'], idx - g['NQ'], idx)
                else:
                    idx = qq + k * d
                acc += h[k] * wb[idx] + gg[k] * buf[idx]
            nb = np.full(g['NQ'], np.nan); nb[:qhi] = acc
            buf = nb
        sel = (q < own_hi) & (gi < n)
        assert not np.isnan(buf[sel]).any()
        vout[gi[sel]] = buf[sel]
    assert not np.isnan(vout).any()
    return vout


def run(n, L, qmf, esz=4, v4=False):
    F = len(qmf)
    gfil = np.array(qmf[::-1]) / np.sqrt(2); hfil = np.array([(-1) ** m * qmf[m] for m in range(F)]) / np.sqrt(2)
    x = np.cumsum(np.random.default_rng(n + L).standard_normal(n))
    ref = orc.modwt(x, np.asarray(qmf), L)
    steps = plan(n, L, F, esz)
    y = np.full((n, L + 1), np.nan)
    v = x
    for kind, g in steps:
        if kind == 'g':
            v = fwd_group(g, v, y, hfil, gfil, v4)
        else:
            j = g; s = 1 << (j - 1); t = np.arange(n)
            w = np.zeros(n); a = np.zeros(n)
            for k in range(F):
                w += hfil[k] * v[(t - k * s) % n]; a += gfil[k] * v[(t - k * s) % n]
            y[:, j - 1] = w; v = a
    y[:, L] = v
    e1 = np.abs(y - ref).max()
    v = ref[:, L]
    for kind, g in reversed(steps):
        if kind == 'g':
            v = inv_group(g, v, ref, hfil, gfil, v4)
        else:
            j = g; s = 1 << (j - 1); t = np.arange(n)
            acc = np.zeros(n)
            for k in range(F):
                acc += hfil[k] * ref[(t + k * s) % n, j - 1] + gfil[k] * v[(t + k * s) % n]
            v = acc
    e2 = np.abs(v - x).max()
    desc = [(k, (g['j0'], g['K'], g['periodic'], g['H'], g['M'], g['ntiles']) if k == 'g' else g) for k, g in steps]
    return e1, e2, desc


if __name__ == "__main__":
    import wavelets_b200 as wb
    for wn in ("haar", "db4", "db10"):
        qmf = list(wb.wavelet(getattr(wb.WT, wn)).qmf)
        for n, L, esz in ((129, 7, 4), (1000, 9, 8), (5000, 7, 8), (9216, 13, 4), (20000, 10, 4), (65536, 16, 4), (65536, 16, 8), (131072, 17, 4), (40960, 12, 8), (100000, 9, 4)):
            e1, e2, desc = run(n, L, qmf, esz)
            print(wn, n, L, esz, f"{e1:.2e} {e2:.2e}", desc)
            if n % 4 == 0:
                e1, e2, _ = run(n, L, qmf, esz, v4=True)
                print('   v4 rounding:', f"{e1:.2e} {e2:.2e}")
