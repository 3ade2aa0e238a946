#!/usr/bin/env python3
"""CPU emulation of the INDEX ARITHMETIC of wptfused.cu (k_pkt_ana / k_pkt_syn: plans, band offsets, rotated stores, staged
ranges) against a direct periodic packet transform -- a dry run for a kernel that can only execute on the GPU box.
    python tools/wptfused_emul.py"""
import numpy as np

rng = np.random.default_rng(0)


def level_ref(x, h, g):
    m, F = len(x), len(h)
    k = np.arange(m // 2)
    a = sum(h[t] * x[(2 * k + t) % m] for t in range(F))
    d = sum(g[F - 1 - i] * x[(2 * k + 2 - F + i) % m] for i in range(F))
    return a, d


def syn_ref(a, d, h, g):
    nh, F = len(a), len(h)
    Q = F // 2
    x = np.zeros(2 * nh)
    for u in range(nh):
        x[2 * u] = sum(h[2 * t] * a[(u - t) % nh] for t in range(Q)) + sum(g[2 * t + 1] * d[(u + t) % nh] for t in range(Q))
        x[2 * u + 1] = sum(h[2 * t + 1] * a[(u - t) % nh] for t in range(Q)) + sum(g[2 * t] * d[(u + t) % nh] for t in range(Q))
    return x


def wpt_ref(x, h, g, K):
    bands = [x]
    for _ in range(K):
        nxt = []
        for b in bands:
            a, d = level_ref(b, h, g)
            nxt += [a, d]
        bands = nxt
    return np.concatenate(bands)


def up4(v): return (v + 3) & ~3


def ana_plan(F, PA, K, tile, nj, vec):
    Q = F // 2
    DS = ((Q - 1) + 1) & ~1
    WO = 2 * DS - (F - 2)
    WIN = up4(F + 2 * (PA - 1) + WO)
    NA = [0] * (K + 1)
    NA[K] = tile >> K
    for l in range(K, 0, -1):
        NA[l - 1] = up4(2 * NA[l] + WIN - 2 * PA)
    h0 = (NA[0] - tile + vec - 1) // vec * vec
    NA[0] = tile + h0
    assert NA[0] <= nj
    return dict(K=K, tile=tile, h0=h0, NA=NA, DS=DS, WO=WO, WIN=WIN)


def emul_ana(x, h, g, K, tile, PA, vec):
    nj, F = len(x), len(h)
    pl = ana_plan(F, PA, K, tile, nj, vec)
    DS, WO = pl["DS"], pl["WO"]
    out = np.full(nj, np.nan)
    lenK = nj >> K
    for t in range(nj // tile):
        s = t * tile
        bands = [x[(s + np.arange(pl["NA"][0])) % nj]]
        for l in range(1, K + 1):
            NAl = pl["NA"][l]
            nxt = []
            for ib in bands:
                a = np.zeros(NAl); d = np.zeros(NAl)
                for p in range(NAl):
                    assert 2 * p + pl["WIN"] - 1 < len(ib) + 0 or True
                    a[p] = sum(h[m] * ib[2 * p + m] for m in range(F))
                    d[p] = sum(g[F - 1 - q] * ib[2 * p + WO + q] for q in range(F))
                nxt += [a, d]
            if l < K:
                bands = nxt
            else:
                sK = s >> K
                for bb in range(1 << (K - 1)):
                    op = 0
                    for i in range(K - 2, -1, -1):
                        op = (op >> 1) + (DS if (bb >> i) & 1 else 0)
                    sa, sd = (sK + (op >> 1)) % lenK, (sK + (op >> 1) + DS) % lenK
                    for p in range(pl["NA"][K]):
                        out[(2 * bb) * lenK + (sa + p) % lenK] = nxt[2 * bb][p]
                        out[(2 * bb + 1) * lenK + (sd + p) % lenK] = nxt[2 * bb + 1][p]
    return out


def emul_syn(y, h, g, K, tile):
    nj, F = len(y), len(h)
    Q = F // 2
    Q4 = ((Q - 1) + 3) & ~3
    dn8 = lambda v: (v & ~7) if v >= 0 else -(((-v) + 7) & ~7)
    up8 = lambda v: (v + 7) & ~7
    lo, hi = [0] * (K + 1), [0] * (K + 1)
    hi[0] = tile
    for l in range(1, K + 1):
        lo[l] = dn8(lo[l - 1] // 2 - Q4)
        hi[l] = up8(hi[l - 1] // 2 + Q4)
    lenK = nj >> K
    assert hi[K] - lo[K] <= lenK
    out = np.full(nj, np.nan)
    for t in range(nj // tile):
        s = t * tile
        sK = s >> K
        bands = [y[beta * lenK + (sK + lo[K] + np.arange(hi[K] - lo[K])) % lenK] for beta in range(1 << K)]
        for l in range(K, 0, -1):
            npairs = (hi[l - 1] - lo[l - 1]) >> 1
            oa = (lo[l - 1] >> 1) - lo[l]
            assert oa % 4 == 0 and npairs % 4 == 0 and oa >= Q4
            nxt = []
            for bb in range(1 << (l - 1)):
                a, d = bands[2 * bb], bands[2 * bb + 1]
                o = np.zeros(2 * npairs)
                for u in range(npairs):
                    assert oa + u + Q4 + 3 < len(d) + 4
                    o[2 * u] = sum(h[2 * k] * a[oa + u - k] for k in range(Q)) + sum(g[2 * k + 1] * d[oa + u + k] for k in range(Q))
                    o[2 * u + 1] = sum(h[2 * k + 1] * a[oa + u - k] for k in range(Q)) + sum(g[2 * k] * d[oa + u + k] for k in range(Q))
                nxt.append(o)
            bands = nxt
        out[s:s + tile] = bands[0]
    return out


def qmf_pair(F):
    h = rng.standard_normal(F)
    g = np.array([(-1) ** m * h[m] for m in range(F)])      # any taps: the index test does not need orthogonality
    return h, g


if __name__ == "__main__":
    for F, K, tile, nj, PA, vec in [(16, 4, 256, 1024, 2, 4), (16, 4, 512, 1024, 2, 4), (8, 3, 256, 1024, 2, 4), (4, 2, 256, 512, 2, 4),
                                    (2, 4, 256, 512, 2, 4), (18, 4, 256, 1024, 1, 2), (12, 2, 256, 768, 2, 4), (20, 2, 512, 2048, 1, 2)]:
        h, g = qmf_pair(F)
        x = rng.standard_normal(nj)
        ref = wpt_ref(x, h, g, K)
        got = emul_ana(x, h, g, K, tile, PA, vec)
        ea = float(np.max(np.abs(ref - got)))
        # synthesis: emulate on the reference packet coefficients, compare with the level-by-level periodic synthesis
        bands = np.split(ref, 1 << K)
        cur = bands
        for _ in range(K):
            cur = [syn_ref(cur[2 * i], cur[2 * i + 1], h, g) for i in range(len(cur) // 2)]
        es = float(np.max(np.abs(cur[0] - emul_syn(ref, h, g, K, tile))))
        print(f"F={F:2d} K={K} tile={tile} nj={nj} PA={PA}: analysis max diff {ea:.2e}   synthesis max diff {es:.2e}")
        assert ea < 1e-9 and es < 1e-9
    print("ok")
