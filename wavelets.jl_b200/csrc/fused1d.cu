// fused1d.cu -- fused multi-level 1-D filter-bank DWT for batches of contiguous columns (sm_100a).
//
// One transform direction is at most two launches, however many levels it has:
//
//   forward   stage A  k_ana_tiles : a CTA stages TILE input samples (+ halo) of one column into shared memory
//                                    with TMA bulk copies (cp.async.bulk + mbarrier; a second copy brings the
//                                    periodic wrap), then runs K analysis levels back to back out of shared
//                                    memory: every level's detail half goes straight to HBM with 64/128-bit
//                                    coalesced stores, the shrinking approximation stays on chip.
//             stage B  k_ana_tail  : the remaining n/2^K-sample approximation of a column fits one CTA's shared
//                                    memory; one CTA finishes all remaining levels there.
//   inverse   stage B' k_syn_tail  : levels L..K+1 of a column entirely in shared memory,
//             stage A' k_syn_tiles : a CTA stages its slice of a_K and of d_K..d_1 (TMA bulk copies, wrap pieces
//                                    included), synthesises K levels in shared memory and streams the TILE
//                                    output samples with 128-bit stores.
//
// So HBM traffic is ~ (1 + halo/TILE + 2^-K) x the compulsory 2*sizeof(T) bytes per sample instead of the
// 2x of a launch per level.  Decimation is fused with the filter; the detail band is produced
// from the same register window as the approximation (d[k + F/2 - 1] and a[k] read the same F inputs).
//
// Arithmetic is the reference's (filtdown!/filtup! closed forms) in the reference's summation
// order; STRICT keeps multiply and add separately rounded (bit-identical to the CPU path), otherwise FMA.
#include "fused1d_dev.cuh"
#include "thresh_dev.cuh"

#include <cstdlib>

namespace wb {

// Stage A.  The input line is a_{lvl0} of a column (src + col*src_stride, ncur = n0 >> lvl0 samples; lvl0 = 0: x).
// Details of local level l go to the d_{lvl0+l} band of y (y + col*n0 + (n0 >> (lvl0+l))); the level-K
// approximation goes to dst_a + col*dst_a_stride (y itself when no level remains, else the next stage's scratch).
template <typename T, int F, bool STRICT>
__global__ void __launch_bounds__(512)
k_ana_tiles(const T *__restrict__ src, int64_t src_stride, T *__restrict__ y, int64_t n0, int lvl0,
            T *__restrict__ dst_a, int64_t dst_a_stride,
            const __grid_constant__ Taps<T, F> c, const __grid_constant__ AnaPlan pl, int pf) {
    constexpr int PA = AnaPairs<T>::value;
    using G = FGeom<F, PA>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw);
    T *bufA = reinterpret_cast<T *>(smem_raw + 128);
    T *bufB = bufA + ((pl.tile + pl.h0 + 8 + 3) & ~3);
    const int64_t col = blockIdx.y;
    const int64_t s = (int64_t)blockIdx.x * pl.tile;
    const int64_t ncur = n0 >> lvl0;
    const T *xc = src + col * src_stride;
    T *yc = y + col * n0;

    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        const int count = pl.tile + pl.h0;
        mbar_expect_tx(bar, (uint32_t)(count * sizeof(T)));
        tma_load_wrapped<T>(bufA, xc, s, count, ncur, bar);
    }
    __syncthreads();       // barrier initialised before anyone polls it
    // Software prefetch into L2 of the tile a CTA `pf` launches further on will stage (CTAs start in linear block order), issued
    // in the shadow of this CTA's own TMA wait.  With ~8 small CTAs per SM only the ones still waiting for their tile have loads
    // in flight (27 % of the warp samples of profiles/r02h_fused1d_f32.md sit in that wait with DRAM at 78 %); the prefetch keeps
    // DRAM busy on their behalf and turns the later TMA load into an L2 hit.
    if (pf > 0 && threadIdx.x == 0) {
        const unsigned lin = blockIdx.y * gridDim.x + blockIdx.x + (unsigned)pf;      // (grid sizes are checked to fit 31 bits)
        const unsigned pcol = lin / gridDim.x;
        if (pcol < gridDim.y) {
            const int64_t ps = (int64_t)(lin - pcol * gridDim.x) * pl.tile;
            const int64_t room = ncur - ps;
            const int count = pl.tile + pl.h0;
            const int64_t cnt = room < (int64_t)count ? room : (int64_t)count;
            tma_prefetch_l2(src + (int64_t)pcol * src_stride + ps, (uint32_t)(cnt * sizeof(T)));
        }
    }
    mbar_wait(bar, 0);

    const T *in = bufA;
    T *out = bufB;
    for (int l = 1; l <= pl.K; ++l) {
        const int64_t nl = ncur >> l;              // length of the d band of this level (and of its approximation)
        const int64_t sl = s >> l;
        // d_j lives at y[n0/2^j .. n0/2^(j-1)); this tile's details start DS further (same register window as the
        // approximation) and wrap to the band start in the line's last tile
        T *dbase = yc + (n0 >> (lvl0 + l)) + sl + G::DS;
        const int64_t room = nl - sl - G::DS;                      // pairs before the wrap
        const int wrap_at = room < (int64_t)pl.ND[l] ? (int)room : 0x7fffffff;
        if (l < pl.K) {
            auto store_a = [&](int p, const T (&a)[PA]) {
                if constexpr (PA == 2) store2(out + p, a[0], a[1]); else out[p] = a[0];
            };
            ana_level<T, F, STRICT>(in, pl.NA[l], pl.ND[l], c, store_a, dbase, dbase - nl, wrap_at);
            __syncthreads();
            const T *t = in; in = out; out = const_cast<T *>(t);
        } else {
            T *dst = dst_a + col * dst_a_stride + sl;
            auto store_a = [&](int p, const T (&a)[PA]) {
                if constexpr (PA == 2) gstore2(dst + p, a[0], a[1]); else __stcs(dst + p, a[0]);
            };
            ana_level<T, F, STRICT>(in, pl.NA[l], pl.ND[l], c, store_a, dbase, dbase - nl, wrap_at);
        }
    }
}

// whole-line tail: the CTA owns a column of length m (<= what fits) and runs `levels` analysis levels in shared
// memory.  Levels of 64 samples and more reuse the vectorised tile routine on a line with its periodic wrap copied
// behind it; shorter levels (any m_l >= 2, including m_l < F) use a scalar closed form with modular indexing.
template <typename T, int F> struct TailGeom {
    static constexpr int PA = AnaPairs<T>::value;
    using G = FGeom<F, PA>;
    static constexpr int HW = (F + G::WO + 2 + 3) & ~3;      // wrap samples appended to a line (multiple of 4)
};

template <typename T, int F, bool STRICT>
__global__ void __launch_bounds__(256)
k_ana_tail(const T *__restrict__ src, int64_t src_stride, T *__restrict__ y, int64_t n, int m, int levels, int first_level,
           const __grid_constant__ Taps<T, F> c) {
    using fp = FP<STRICT>;
    using TG = TailGeom<T, F>;
    constexpr int PA = TG::PA;
    using G = typename TG::G;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    T *bufA = reinterpret_cast<T *>(smem_raw);
    T *bufB = bufA + ((m + TG::HW + 3) & ~3);
    const int64_t col = blockIdx.x;
    const T *sc = src + col * src_stride;
    T *yc = y + col * n;
    // the column arrives with one TMA bulk copy when it is 16-byte granular, else with plain loads
    __shared__ __align__(8) uint64_t tbar;
    const bool bulk = ((m * (int)sizeof(T)) % 16 == 0) && ((reinterpret_cast<uintptr_t>(sc) & 15) == 0);
    if (bulk) {
        if (threadIdx.x == 0) {
            mbar_init(&tbar, 1);
            mbar_expect_tx(&tbar, (uint32_t)(m * sizeof(T)));
            tma_bulk_g2s(bufA, sc, (uint32_t)(m * sizeof(T)), &tbar);
        }
        __syncthreads();
        mbar_wait(&tbar, 0);
    } else {
        for (int i = threadIdx.x; i < m; i += blockDim.x) bufA[i] = sc[i];
        __syncthreads();
    }
    for (int i = threadIdx.x; i < TG::HW; i += blockDim.x) bufA[m + i] = bufA[i % m];
    __syncthreads();
    T *in = bufA, *out = bufB;
    int ml = m;
    for (int l = 1; l <= levels; ++l) {
        const int nh = ml >> 1;
        const int lvl = first_level + l;             // absolute level number
        T *dband = yc + (n >> lvl);
        const bool last = (l == levels);
        if (ml >= 64 && (nh % PA) == 0) {
            // d[p + DS] wraps to the band start at pair nh - DS
            T *d0 = dband + G::DS, *d1 = dband + G::DS - nh;
            const int wrap_at = nh - G::DS;
            if (last) {
                auto store_a = [&](int p, const T (&a)[PA]) {
                    if constexpr (PA == 2) gstore2(yc + p, a[0], a[1]); else __stcs(yc + p, a[0]);
                };
                ana_level<T, F, STRICT>(in, nh, nh, c, store_a, d0, d1, wrap_at);
            } else {
                auto store_a = [&](int p, const T (&a)[PA]) {
                    if constexpr (PA == 2) store2(out + p, a[0], a[1]); else out[p] = a[0];
                };
                ana_level<T, F, STRICT>(in, nh, nh, c, store_a, d0, d1, wrap_at);
                __syncthreads();
                for (int i = threadIdx.x; i < TG::HW; i += blockDim.x) out[nh + i] = out[i % nh];   // periodic wrap
            }
        } else {
            for (int k = threadIdx.x; k < nh; k += blockDim.x) {
                int ia = 2 * k;                          // < ml
                T a = fp::mul(c.h[0], in[ia]);
#pragma unroll
                for (int q = 1; q < F; ++q) {
                    if (++ia == ml) ia = 0;
                    a = fp::mac(a, c.h[q], in[ia]);
                }
                int id = (2 * k + 2 - F) % ml;
                if (id < 0) id += ml;
                T d = fp::mul(c.g[F - 1], in[id]);
#pragma unroll
                for (int q = 1; q < F; ++q) {
                    if (++id == ml) id = 0;
                    d = fp::mac(d, c.g[F - 1 - q], in[id]);
                }
                dband[k] = d;
                if (last) yc[k] = a; else out[k] = a;
            }
        }
        __syncthreads();
        T *t = in; in = out; out = t;
        ml = nh;
    }
}

// ===================================================================================================
// INVERSE
// ===================================================================================================
// Stage A'.  Produces a_{lvl0} of a column (ncur = n0 >> lvl0 samples, to dst + col*dst_stride: y when lvl0 == 0)
// from a_{lvl0+K} (asrc + col*asrc_stride) and the detail bands d_{lvl0+K} .. d_{lvl0+1} of x.
template <typename T, int F, bool STRICT>
__global__ void __launch_bounds__(512)
k_syn_tiles(const T *__restrict__ asrc, int64_t asrc_stride, const T *__restrict__ x, int64_t n0, int lvl0,
            T *__restrict__ dst, int64_t dst_stride,
            const __grid_constant__ Taps<T, F> c, const __grid_constant__ SynPlan pl,
            const __grid_constant__ ThreshEpi epi, int thr_a, int pf, int pfl) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw);
    T *sm = reinterpret_cast<T *>(smem_raw + 128);
    const int64_t col = blockIdx.y;
    const int64_t s = (int64_t)blockIdx.x * pl.tile;
    const int64_t ncur = n0 >> lvl0;
    const T *xc = x + col * n0;
    const int K = pl.K;

    // one mbarrier per level (bar[l] covers d_l; bar[K] also the approximation): the coarse levels start while the
    // large fine-level detail slices are still in flight
    if (threadIdx.x == 0) {
        for (int l = 1; l <= K; ++l) mbar_init(bar + l, 1);
        for (int l = K; l >= 1; --l) {
            uint32_t bytes = (uint32_t)((pl.dhi[l] - pl.dlo[l]) * sizeof(T));
            if (l == K) bytes += (uint32_t)((pl.rhi[K] - pl.rlo[K]) * sizeof(T));
            mbar_expect_tx(bar + l, bytes);
            if (l == K)
                tma_load_wrapped<T>(sm + pl.aoff, asrc + col * asrc_stride, (s >> K) + pl.rlo[K], pl.rhi[K] - pl.rlo[K], ncur >> K, bar + l);
            tma_load_wrapped<T>(sm + pl.doff[l], xc + (n0 >> (lvl0 + l)), (s >> l) + pl.dlo[l], pl.dhi[l] - pl.dlo[l], ncur >> l, bar + l);
        }
    }
    __syncthreads();
    // L2 prefetch for the CTA `pf` launches further on (see k_ana_tiles), in the shadow of this CTA's own TMA waits.  The slices
    // of the coarse levels are small (1 KB at level 4 of a 4096-sample tile), so level l is prefetched by every 2^(l-1)-th tile
    // for 2^(l-1) tiles at once: about two 8 KB prefetch operations per CTA instead of five small ones.
    if (pf > 0 && threadIdx.x == 0) {
        const unsigned lin = blockIdx.y * gridDim.x + blockIdx.x + (unsigned)pf;
        const unsigned pcol = lin / gridDim.x;
        if (pcol < gridDim.y) {
            const unsigned pt = lin - pcol * gridDim.x;                 // tile index inside the column
            const T *pxc = x + (int64_t)pcol * n0;
            for (int l = 1; l <= K && l <= pfl; ++l) {      // pfl: deepest level whose slices are prefetched (WB200_F1D_PREFETCH_INV_LEVELS)
                const unsigned grp = 1u << ((l - 1) < 3 ? (l - 1) : 3);
                if (pt & (grp - 1)) continue;
                const int64_t len = ncur >> l;
                const int64_t lo = ((int64_t)pt * pl.tile) >> l;
                int64_t hi = lo + (int64_t)grp * (pl.tile >> l);
                if (hi > len) hi = len;
                if (hi > lo) tma_prefetch_l2(pxc + (n0 >> (lvl0 + l)) + lo, (uint32_t)((hi - lo) * sizeof(T)));
                if (l == K && hi > lo) tma_prefetch_l2(asrc + (int64_t)pcol * asrc_stride + lo, (uint32_t)((hi - lo) * sizeof(T)));
            }
        }
    }

    const T *abuf = sm + pl.aoff;
    const double tthr = epi.kind >= 0 ? (epi.sigma_dev ? __dmul_rn(*epi.sigma_dev, epi.tfac) : epi.t_host) : 0.0;
    for (int l = K; l >= 1; --l) {
        mbar_wait(bar + l, 0);
        if (epi.kind >= 0) {
            // threshold!(xt, TH, t) as a load epilogue (denoising.jl:44, 69): every staged coefficient of d_l (halo included:
            // it feeds this tile's outputs) and, where the approximation is x's own a_L, that slice too
            T *dst_d = sm + pl.doff[l];
            const int nd = pl.dhi[l] - pl.dlo[l];
            for (int i = threadIdx.x; i < nd; i += blockDim.x) dst_d[i] = thresh_apply<T>(dst_d[i], epi.kind, tthr);
            if (l == K && thr_a) {
                T *dst_a = sm + pl.aoff;
                const int na = pl.rhi[K] - pl.rlo[K];
                for (int i = threadIdx.x; i < na; i += blockDim.x) dst_a[i] = thresh_apply<T>(dst_a[i], epi.kind, tthr);
            }
            __syncthreads();
        }
        // produce the approximation one level up on [s_{l-1} + rlo[l-1], s_{l-1} + rhi[l-1])
        const int npairs = (pl.rhi[l - 1] - pl.rlo[l - 1]) >> 1;           // output pairs = values of u
        const int oa = (pl.rlo[l - 1] >> 1) - pl.rlo[l];                   // index of a[u_first] inside abuf
        const int od = (pl.rlo[l - 1] >> 1) - pl.dlo[l];
        const T *dbuf = sm + pl.doff[l];
        if (l > 1) {
            T *obuf = sm + (((l - 1) & 1) ? pl.poff : pl.qoff);
            auto so = [&](int ur, T o0, T o1, T o2, T o3) { store4(obuf + 2 * ur, o0, o1, o2, o3); };
            syn_level<T, F, STRICT>(abuf, dbuf, oa, od, npairs, c, so);
            __syncthreads();
            abuf = obuf;
        } else {
            T *o = dst + col * dst_stride + s;
            auto so = [&](int ur, T o0, T o1, T o2, T o3) { gstore4(o + 2 * ur, o0, o1, o2, o3); };
            syn_level<T, F, STRICT>(abuf, dbuf, oa, od, npairs, c, so);
        }
    }
}

// whole-line inverse tail: x[0:m) = [a_L | d_L | ... | d_{lv0+1}] of a column -> a_{lv0} (m samples).
// The bands are staged with gaps: every detail band is followed by its periodic wrap (QD samples) and every
// approximation buffer is preceded by its wrap (QAP samples), so that levels of 32+ samples run the vectorised
// tile routine; shorter levels use the scalar closed form.
template <typename T, int F> struct SynTailGeom {
    using G = FGeom<F>;
    static constexpr int QAP = (G::QA + 3) & ~3;          // left wrap slots in front of an approximation (multiple of 4)
    static constexpr int QDP = (G::QD + 3) & ~3;          // right wrap slots behind a detail band
    // element offset of detail band t (t = 0: d_L with mL samples, t: mL * 2^t samples) in the staged layout
    __host__ __device__ static int band_off(int mL, int t) {
        int off = QAP + ((mL + 3) & ~3);
        for (int u = 0; u < t; ++u) off += (((mL << u) + QDP + 3) & ~3);
        return off;
    }
};

template <typename T, int F, bool STRICT>
__global__ void __launch_bounds__(256)
k_syn_tail(const T *__restrict__ x, int64_t n, T *__restrict__ dst, int64_t dst_stride, int m, int levels,
           const __grid_constant__ Taps<T, F> c, const __grid_constant__ ThreshEpi epi) {
    using fp = FP<STRICT>;
    using SG = SynTailGeom<T, F>;
    using G = FGeom<F>;
    constexpr int Q = F / 2;
    constexpr int QAP = SG::QAP;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    T *stage = reinterpret_cast<T *>(smem_raw);             // [QAP | a_L][d_L | wrap][d_{L-1} | wrap] ...
    const int mL = m >> levels;
    const int stage_sz = SG::band_off(mL, levels);
    T *bufP = stage + ((stage_sz + 3) & ~3);                // approximations produced at levels-2, levels-4, ... (<= m/2)
    T *bufQ = bufP + ((QAP + (m >> 1) + 3) & ~3);           // ... at levels-3, levels-5, ...                 (<= m/4)
    const int64_t col = blockIdx.x;
    const T *xc = x + col * n;
    T *dc = dst + col * dst_stride;
    // ---- stage the coefficients: one TMA bulk copy per band that is 16-byte granular, plain loads for the small ones;
    //      then the periodic wraps from shared memory ----
    __shared__ __align__(8) uint64_t tbar;
    constexpr int V = 16 / (int)sizeof(T);
    const bool aligned = (reinterpret_cast<uintptr_t>(xc) & 15) == 0;
    if (threadIdx.x == 0) {
        mbar_init(&tbar, 1);
        uint32_t bytes = 0;
        if (aligned && mL % V == 0) bytes += (uint32_t)(mL * sizeof(T));
        for (int t = 0; t < levels; ++t) if (aligned && (mL << t) % V == 0) bytes += (uint32_t)((mL << t) * sizeof(T));
        mbar_expect_tx(&tbar, bytes);
        if (aligned && mL % V == 0) tma_bulk_g2s(stage + QAP, xc, (uint32_t)(mL * sizeof(T)), &tbar);
        for (int t = 0; t < levels; ++t)
            if (aligned && (mL << t) % V == 0)
                tma_bulk_g2s(stage + SG::band_off(mL, t), xc + (mL << t), (uint32_t)((mL << t) * sizeof(T)), &tbar);
    }
    if (!(aligned && mL % V == 0))
        for (int i = threadIdx.x; i < mL; i += blockDim.x) stage[QAP + i] = xc[i];
    for (int t = 0; t < levels; ++t) {
        const int sz = mL << t;
        if (!(aligned && sz % V == 0))
            for (int i = threadIdx.x; i < sz; i += blockDim.x) stage[SG::band_off(mL, t) + i] = xc[sz + i];
    }
    __syncthreads();
    mbar_wait(&tbar, 0);
    __syncthreads();
    if (epi.kind >= 0) {   // threshold!(xt, TH, t) as a load epilogue: a_L and every detail band of this column (before the wraps are copied)
        const double tthr = epi.sigma_dev ? __dmul_rn(*epi.sigma_dev, epi.tfac) : epi.t_host;
        for (int i = threadIdx.x; i < mL; i += blockDim.x) stage[QAP + i] = thresh_apply<T>(stage[QAP + i], epi.kind, tthr);
        for (int t = 0; t < levels; ++t) {
            const int sz = mL << t, off = SG::band_off(mL, t);
            for (int i = threadIdx.x; i < sz; i += blockDim.x) stage[off + i] = thresh_apply<T>(stage[off + i], epi.kind, tthr);
        }
        __syncthreads();
    }
    for (int i = threadIdx.x; i < QAP; i += blockDim.x) { int k = (mL - 1 - i) % mL; if (k < 0) k += mL; stage[QAP - 1 - i] = stage[QAP + k]; }
    for (int t = 0; t < levels; ++t) {
        const int sz = mL << t, off = SG::band_off(mL, t);
        for (int i = threadIdx.x; i < SG::QDP; i += blockDim.x) stage[off + sz + i] = stage[off + (i % sz)];
    }
    __syncthreads();
    int nh = mL;                                            // current approximation length
    T *a = stage;                                           // a[QAP + i] = approximation sample i
    for (int l = 0; l < levels; ++l) {
        const T *d = stage + SG::band_off(mL, l);           // d[i], i in [0, nh + QDP)
        const bool last = (l == levels - 1);
        T *out = ((levels - 2 - l) & 1) ? bufQ : bufP;      // out[QAP + i]
        if (nh >= 32 && (nh & 1) == 0) {
            if (last) {
                auto so = [&](int ur, T o0, T o1, T o2, T o3) { gstore4(dc + 2 * ur, o0, o1, o2, o3); };
                syn_level<T, F, STRICT>(a, d, QAP, 0, nh, c, so);
            } else {
                auto so = [&](int ur, T o0, T o1, T o2, T o3) { store4(out + QAP + 2 * ur, o0, o1, o2, o3); };
                syn_level<T, F, STRICT>(a, d, QAP, 0, nh, c, so);
                __syncthreads();
                for (int i = threadIdx.x; i < QAP; i += blockDim.x) out[QAP - 1 - i] = out[QAP + 2 * nh - 1 - i];   // left wrap
            }
        } else {
            for (int u = threadIdx.x; u < nh; u += blockDim.x) {
                int ia = (u - (Q - 1)) % nh;
                if (ia < 0) ia += nh;
                T rae = fp::mul(c.h[2 * (Q - 1)], a[QAP + ia]);
                T rao = fp::mul(c.h[2 * (Q - 1) + 1], a[QAP + ia]);
#pragma unroll
                for (int k = Q - 2; k >= 0; --k) {
                    if (++ia == nh) ia = 0;
                    rae = fp::mac(rae, c.h[2 * k], a[QAP + ia]);
                    rao = fp::mac(rao, c.h[2 * k + 1], a[QAP + ia]);
                }
                int id = u;
                T rde, rdo;      // fast mode: the detail terms continue the approximation's chain (no second FMUL, no final FADD)
                if constexpr (STRICT) { rde = fp::mul(c.g[1], d[id]); rdo = fp::mul(c.g[0], d[id]); }
                else { rde = fp::mac(rae, c.g[1], d[id]); rdo = fp::mac(rao, c.g[0], d[id]); }
#pragma unroll
                for (int k = 1; k < Q; ++k) {
                    if (++id == nh) id = 0;
                    rde = fp::mac(rde, c.g[2 * k + 1], d[id]);
                    rdo = fp::mac(rdo, c.g[2 * k], d[id]);
                }
                const T x0 = (STRICT ? fp::add(rae, rde) : rde), x1 = (STRICT ? fp::add(rao, rdo) : rdo);
                if (last) { dc[2 * u] = x0; dc[2 * u + 1] = x1; }
                else      { out[QAP + 2 * u] = x0; out[QAP + 2 * u + 1] = x1; }
            }
            if (!last) {
                __syncthreads();
                const int no = 2 * nh;
                for (int i = threadIdx.x; i < QAP; i += blockDim.x) { int k = (no - 1 - i) % no; if (k < 0) k += no; out[QAP - 1 - i] = out[QAP + k]; }
            }
        }
        __syncthreads();
        a = out;
        nh <<= 1;
    }
}

// ===================================================================================================
// host side: planning and dispatch
// ===================================================================================================
static int env_int(const char *name, int dflt) {
    const char *v = getenv(name);
    return (v && *v) ? atoi(v) : dflt;
}
static inline int64_t pow2_factor(int64_t n) { return n & (-n); }

// A transform of n samples / L levels is cut into tile stages (each fuses K_i levels out of shared-memory tiles,
// reading a_{lv} from HBM once and writing its details + a_{lv+K_i}) followed by one whole-line tail stage for
// the levels that remain once a column's approximation fits one CTA.
struct Stage { int K, tile, lv0; };
struct Fused1dCfg {
    bool ok = false;
    int nstages = 0;
    Stage st[8];
    int tail_levels = 0;   // levels left to the tail stage
    int tail_lv0 = 0;      // levels done before the tail
    int64_t m = 0;         // tail line length (n >> tail_lv0)
};

// Defaults from the round-2 interleaved A/B (tools/ab_filt_inv.py, profiles/r02_fused1d_ab.md): SMALL tiles, four levels per
// stage, 96-thread CTAs and a 16 KB whole-line tail.  A tile CTA is a chain of dependent levels (TMA wait, then K x
// (window loads -> FIR -> stores -> barrier)); many small resident CTAs hide each other's waits better than the 8192 /
// 16384-sample tiles of 256 threads and up to 8 levels that round 1 tuned one knob at a time (f32 +10 % forward, +16 % inverse;
// f64 +13 % / +22 %).
template <typename T> static int tail_max(bool fw) {
    if (sizeof(T) == 4) return fw ? env_int("WB200_TAILMAX_F32", 4096) : env_int("WB200_TAILMAX_F32_INV", env_int("WB200_TAILMAX_F32", 4096));
    return env_int("WB200_TAILMAX_F64", 2048);
}
template <typename T> static int tile_max(bool fw) {
    if (sizeof(T) == 4) return fw ? env_int("WB200_TILE_F32", 4096) : env_int("WB200_TILE_F32_INV", env_int("WB200_TILE_F32", 4096));
    return env_int("WB200_TILE_F64", 2048);
}
constexpr int KMAX_DEFAULT = 4;

// L2 prefetch distance of the tile kernels in CTAs (WB200_F1D_PREFETCH / WB200_F1D_PREFETCH_INV; 0 = off).  Interleaved A/B
// (tools/ab_prefetch.py, profiles/r02h_prefetch_ab.md): the FORWARD kernel -- one contiguous read stream per column -- gains
// 4-9 % with 600-1200 CTAs of lookahead (about one resident generation: 148 SMs x 6-8 CTAs); the INVERSE kernel, which reads
// five streams at power-of-two offsets, loses 7-13 % at every distance, so its default stays off.
static int tile_prefetch(bool fw) {
    const int v = fw ? env_int("WB200_F1D_PREFETCH", 888) : env_int("WB200_F1D_PREFETCH_INV", 0);
    return v < 0 ? 0 : v;
}
// threads per tile CTA (WB200_F1D_NT / WB200_F1D_NT_INV)
static int tile_threads(bool fw) {
    const int v = env_int(fw ? "WB200_F1D_NT" : "WB200_F1D_NT_INV", 96);
    return (v >= 32 && v <= 512) ? (v & ~31) : 96;
}

template <int F, int PA> static int ana_halo(int K, int (&H)[MAXK + 1]) {
    using G = FGeom<F, PA>;
    H[K] = 0;
    for (int l = K; l >= 1; --l) {
        const int a = 2 * H[l] + F - 2, b = F - 2 + G::WO;
        H[l - 1] = a > b ? a : b;
    }
    return H[0];
}

template <typename T, int F>
static Fused1dCfg plan_split(int64_t n, int L, bool fw) {
    Fused1dCfg c;
    const int tmax = tail_max<T>(fw);
    const int kcap = fw ? env_int("WB200_KMAX", KMAX_DEFAULT) : env_int("WB200_KMAX_INV", env_int("WB200_KMAX", KMAX_DEFAULT));
    const int halo_div = env_int("WB200_HALO_DIV", 4);        // accept at most tile/halo_div halo samples per tile
    int64_t cur = n;
    int lv = 0;
    while (cur > tmax && lv < L) {
        if (c.nstages == 8) return c;
        int64_t tile = tile_max<T>(fw);
        const int64_t p2 = pow2_factor(cur);
        while (tile > p2) tile >>= 1;                          // the tile must divide the line
        while (tile > cur / 2) tile >>= 1;                     // at least two tiles per line (a wrap piece never overlaps its tile)
        if (tile < 64 || (cur * (int64_t)sizeof(T)) % 16 != 0) return c;
        int need = 0;
        while ((cur >> need) > tmax) ++need;
        int K = need < (L - lv) ? need : (L - lv);
        if (K > kcap) K = kcap;
        if (K > MAXK) K = MAXK;
        int H[MAXK + 1];
        while (K >= 1 && (ana_halo<F, AnaPairs<T>::value>(K, H) > tile / halo_div || (tile >> K) < 4)) --K;
        if (K < 1) return c;
        c.st[c.nstages++] = Stage{K, (int)tile, lv};
        cur >>= K;
        lv += K;
    }
    if (lv < L && cur > tmax) return c;                        // cannot happen (loop exits only when one is false)
    c.tail_levels = L - lv;
    c.tail_lv0 = lv;
    c.m = cur;
    c.ok = true;
    return c;
}

template <int F, int PA> static void make_ana_plan(AnaPlan &pl, const Stage &sg, int vec) {
    pl.K = sg.K; pl.tile = sg.tile;
    int H[MAXK + 1];
    ana_halo<F, PA>(sg.K, H);
    pl.h0 = (H[0] + vec - 1) / vec * vec;
    pl.NA[0] = sg.tile + pl.h0; pl.ND[0] = sg.tile;
    for (int l = 1; l <= sg.K; ++l) { pl.ND[l] = sg.tile >> l; pl.NA[l] = pl.ND[l] + H[l]; }
}
template <typename T> static size_t ana_smem(const AnaPlan &pl) {
    const size_t a = ((size_t)pl.tile + pl.h0 + 8 + 3) & ~(size_t)3;
    const size_t b = ((size_t)pl.NA[1] + 8 + 3) & ~(size_t)3;
    return 128 + (a + b) * sizeof(T);
}

template <int F> static bool make_syn_plan(SynPlan &pl, const Stage &sg, int64_t ncur, size_t &smem_elems) {
    using G = FGeom<F>;
    // staged ranges are cut for the four-pair form of syn_level (16-byte windows): every range starts on a multiple of 8
    // samples of its level (so the half-rate index of the level above is a multiple of 4), with Q4 samples of a_l in front of
    // the first output pair and Q4 samples of d_l behind the last one; the two-pair form (Float64) needs no more than that
    auto dn8 = [](int v) { return (v >= 0) ? (v & ~7) : -(((-v) + 7) & ~7); };
    pl.K = sg.K; pl.tile = sg.tile;
    pl.rlo[0] = 0; pl.rhi[0] = sg.tile;
    pl.dlo[0] = pl.dhi[0] = pl.doff[0] = 0;
    for (int l = 1; l <= sg.K; ++l) {
        pl.rlo[l] = dn8(pl.rlo[l - 1] / 2 - G::Q4);   // rlo[l-1] is a multiple of 8, so /2 is a multiple of 4
        pl.rhi[l] = pl.rhi[l - 1] / 2;
        pl.dlo[l] = pl.rlo[l - 1] / 2;
        pl.dhi[l] = pl.rhi[l - 1] / 2 + G::Q4;
    }
    size_t off = 0;
    for (int l = 1; l <= sg.K; ++l) { pl.doff[l] = (int)off; off += (size_t)(pl.dhi[l] - pl.dlo[l]); }
    pl.aoff = (int)off; off += (size_t)(pl.rhi[sg.K] - pl.rlo[sg.K]);
    // ping-pong: poff holds the odd local levels a_1, a_3, ... ; qoff the even ones
    size_t psz = 0, qsz = 0;
    for (int l = 1; l < sg.K; ++l) {
        const size_t sz = (size_t)(pl.rhi[l] - pl.rlo[l]);
        if (l & 1) psz = sz > psz ? sz : psz; else qsz = sz > qsz ? sz : qsz;
    }
    pl.poff = (int)off; off += psz;
    pl.qoff = (int)off; off += qsz;
    smem_elems = off + 8;
    // every staged range must fit inside its (periodic) band: a bulk copy wraps at most once
    for (int l = 1; l <= sg.K; ++l)
        if ((int64_t)(pl.dhi[l] - pl.dlo[l]) > (ncur >> l)) return false;
    if ((int64_t)(pl.rhi[sg.K] - pl.rlo[sg.K]) > (ncur >> sg.K)) return false;
    return true;
}

template <typename T, int F>
static bool syn_plans_ok(const Fused1dCfg &cfg, int64_t n) {
    for (int i = 0; i < cfg.nstages; ++i) {
        SynPlan pl; size_t e;
        if (!make_syn_plan<F>(pl, cfg.st[i], n >> cfg.st[i].lv0, e)) return false;
        if (128 + e * sizeof(T) > 200 * 1024) return false;
    }
    return true;
}

// scratch layout: the approximation handed from stage i to the next stage (or to the tail) lives in buffer i & 1
template <typename T>
static size_t scratch_elems(const Fused1dCfg &cfg, int64_t n, int64_t B, size_t (&off)[2]) {
    size_t sz[2] = {0, 0};
    for (int i = 0; i < cfg.nstages; ++i) {
        const int lv1 = cfg.st[i].lv0 + cfg.st[i].K;
        const bool last_overall = (i == cfg.nstages - 1) && cfg.tail_levels == 0;
        if (last_overall) continue;
        const size_t e = (size_t)(n >> lv1) * (size_t)B;
        if (e > sz[i & 1]) sz[i & 1] = e;
    }
    off[0] = 0;
    off[1] = (sz[0] * sizeof(T) + 255) / 256 * 256 / sizeof(T);
    return off[1] + sz[1];
}

template <typename T, int F, bool STRICT>
static int32_t run_fused_1d(const PassOp<T> &op, T *y, const T *x, int64_t n, int64_t B, int L, bool fw,
                            void *workspace, size_t ws_bytes, cudaStream_t st) {
    const Fused1dCfg cfg = plan_split<T, F>(n, L, fw);
    if (!cfg.ok) return -1;
    if (cfg.nstages > 0 && (B > 65535 || !syn_plans_ok<T, F>(cfg, n))) return -1;
    Taps<T, F> taps;
    for (int m = 0; m < F; ++m) { taps.h[m] = op.fc.h[m]; taps.g[m] = op.fc.g[m]; }

    size_t soff[2];
    const size_t scratch_bytes = scratch_elems<T>(cfg, n, B, soff) * sizeof(T);
    T *scratch = nullptr;
    bool own_scratch = false;
    if (scratch_bytes) {
        if (workspace != nullptr) {
            if (ws_bytes < scratch_bytes) { set_error("workspace too small: %zu bytes given, %zu needed", ws_bytes, scratch_bytes); return WB200_EWORKSPACE; }
            scratch = (T *)workspace;
        } else {
            if (scratch_alloc((void **)&scratch, scratch_bytes, st) != cudaSuccess) { (void)cudaGetLastError(); set_error("cudaMallocAsync(fused scratch) failed"); return WB200_ECUDA; }
            own_scratch = true;
        }
    }
    int32_t rc = WB200_OK;
    auto finish = [&]() { if (own_scratch) cudaFreeAsync(scratch, st); return rc; };
    auto fail = [&](const char *what) { rc = WB200_ECUDA; set_error("%s", what); return finish(); };
    // buffer holding the approximation after stage i (i = -1: the input itself)
    auto abuf = [&](int i) -> T * { return scratch + soff[i & 1]; };

    if (fw) {
        for (int i = 0; i < cfg.nstages; ++i) {
            const Stage &sg = cfg.st[i];
            AnaPlan pl;
            make_ana_plan<F, AnaPairs<T>::value>(pl, sg, 16 / (int)sizeof(T));
            const size_t smem = ana_smem<T>(pl);
            auto kern = k_ana_tiles<T, F, STRICT>;
            if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { (void)cudaGetLastError(); return fail("cudaFuncSetAttribute(k_ana_tiles) failed"); }
            const int64_t ncur = n >> sg.lv0;
            const T *src = (i == 0) ? x : abuf(i - 1);
            const int64_t sstride = (i == 0) ? n : ncur;
            const bool last_overall = (i == cfg.nstages - 1) && cfg.tail_levels == 0;
            T *dsta = last_overall ? y : abuf(i);
            const int64_t dstride = last_overall ? n : (ncur >> sg.K);
            dim3 grid((unsigned)(ncur / sg.tile), (unsigned)B);
            {
                LaunchScope scope("fused_ana_tiles", st);
                const int pfv = ((uint64_t)grid.x * grid.y + (uint64_t)tile_prefetch(true) < 0x7fffffffULL) ? tile_prefetch(true) : 0;
                kern<<<grid, tile_threads(true), smem, st>>>(src, sstride, y, n, sg.lv0, dsta, dstride, taps, pl, pfv);
            }
            if (!check_launch("fused_ana_tiles")) { rc = WB200_ECUDA; return finish(); }
        }
        if (cfg.tail_levels > 0) {
            const int m = (int)cfg.m;
            constexpr int HW = TailGeom<T, F>::HW;
            const size_t smem = ((size_t)((m + HW + 3) & ~3) + (size_t)(((m >> 1) + HW + 3) & ~3) + 8) * sizeof(T);
            auto kern = k_ana_tail<T, F, STRICT>;
            if (smem > 232448) return fail("analysis tail does not fit shared memory (lower WB200_TAILMAX_*)");
            if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { (void)cudaGetLastError(); return fail("cudaFuncSetAttribute(k_ana_tail) failed"); }
            const T *src = cfg.nstages ? abuf(cfg.nstages - 1) : x;
            const int64_t sstride = cfg.nstages ? cfg.m : n;
            {
                LaunchScope scope("fused_ana_tail", st);
                kern<<<(unsigned)B, 256, smem, st>>>(src, sstride, y, n, m, cfg.tail_levels, cfg.tail_lv0, taps);
            }
            if (!check_launch("fused_ana_tail")) { rc = WB200_ECUDA; return finish(); }
        }
    } else {
        if (cfg.tail_levels > 0) {
            const int m = (int)cfg.m;
            using SG = SynTailGeom<T, F>;
            const size_t smem = ((size_t)((SG::band_off(m >> cfg.tail_levels, cfg.tail_levels) + 3) & ~3) +
                                 (size_t)((SG::QAP + (m >> 1) + 3) & ~3) + (size_t)((SG::QAP + (m >> 2) + 3) & ~3) + 16) * sizeof(T);
            auto kern = k_syn_tail<T, F, STRICT>;
            if (smem > 232448) return fail("synthesis tail does not fit shared memory (lower WB200_TAILMAX_*)");
            if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { (void)cudaGetLastError(); return fail("cudaFuncSetAttribute(k_syn_tail) failed"); }
            T *dst = cfg.nstages ? abuf(cfg.nstages - 1) : y;
            const int64_t dstride = cfg.nstages ? cfg.m : n;
            {
                LaunchScope scope("fused_syn_tail", st);
                kern<<<(unsigned)B, 256, smem, st>>>(x, n, dst, dstride, m, cfg.tail_levels, taps, op.epi);
            }
            if (!check_launch("fused_syn_tail")) { rc = WB200_ECUDA; return finish(); }
        }
        for (int i = cfg.nstages - 1; i >= 0; --i) {
            const Stage &sg = cfg.st[i];
            const int64_t ncur = n >> sg.lv0;
            SynPlan pl; size_t elems;
            if (!make_syn_plan<F>(pl, sg, ncur, elems)) return fail("internal: synthesis plan rejected after validation");
            const size_t smem = 128 + elems * sizeof(T);
            auto kern = k_syn_tiles<T, F, STRICT>;
            if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { (void)cudaGetLastError(); return fail("cudaFuncSetAttribute(k_syn_tiles) failed"); }
            const bool last_overall = (i == cfg.nstages - 1) && cfg.tail_levels == 0;
            const T *asrc = last_overall ? x : abuf(i);             // coarsest stage without a tail reads a_L from x
            const int64_t astride = last_overall ? n : (ncur >> sg.K);
            T *dst = (i == 0) ? y : abuf(i - 1);
            const int64_t dstride = (i == 0) ? n : ncur;
            dim3 grid((unsigned)(ncur / sg.tile), (unsigned)B);
            {
                LaunchScope scope("fused_syn_tiles", st);
                const int pfv = ((uint64_t)grid.x * grid.y + (uint64_t)tile_prefetch(false) < 0x7fffffffULL) ? tile_prefetch(false) : 0;
                kern<<<grid, tile_threads(false), smem, st>>>(asrc, astride, x, n, sg.lv0, dst, dstride, taps, pl, op.epi, last_overall ? 1 : 0, pfv, env_int("WB200_F1D_PREFETCH_INV_LEVELS", MAXK));
            }
            if (!check_launch("fused_syn_tiles")) { rc = WB200_ECUDA; return finish(); }
        }
    }
    return finish();
}

template <typename T, bool STRICT>
static int32_t dispatch_F(const PassOp<T> &op, T *y, const T *x, int64_t n, int64_t B, int L, bool fw,
                          void *ws, size_t wsb, cudaStream_t st) {
    switch (op.fc.F) {
#define WB_CASE(FF) case FF: return run_fused_1d<T, FF, STRICT>(op, y, x, n, B, L, fw, ws, wsb, st);
        WB_CASE(2) WB_CASE(4) WB_CASE(6) WB_CASE(8) WB_CASE(10) WB_CASE(12) WB_CASE(14) WB_CASE(16) WB_CASE(18) WB_CASE(20)
        WB_CASE(22) WB_CASE(24)                                  // db11, and the 24-tap coif8 / Vaidyanathan filters of wt_main.jl:372-436
#undef WB_CASE
    default: return -1;
    }
}

template <typename T>
int32_t fused_dwt(const PassOp<T> &op, T *y, const T *x, const ArrayGeom &g, int L, bool fw,
                  void *workspace, size_t ws_bytes, cudaStream_t st, uint32_t flags) {
    (void)flags;
    if (op.lifting || g.ndim != 1 || g.C != 1 || L < 1) return -1;
    if (env_int("WB200_DISABLE_FUSED", 0)) return -1;
    const int64_t n = g.dim[0];
    if (n > ((int64_t)1 << 30)) return -1;                       // tile-relative indices are 32-bit
    if (((uintptr_t)x | (uintptr_t)y) & 15) return -1;
    if ((n * (int64_t)sizeof(T)) % 16 != 0 && g.batch > 1) return -1;
    if (op.strict) return dispatch_F<T, true>(op, y, x, n, g.batch, L, fw, workspace, ws_bytes, st);
    return dispatch_F<T, false>(op, y, x, n, g.batch, L, fw, workspace, ws_bytes, st);
}

template <typename T> static size_t fused_ws_T(int64_t n, int64_t B, int L) {
    // the split depends on the filter length only through the halo; take the worst case over the supported lengths
    size_t need = 0;
    auto one = [&](const Fused1dCfg &c) {
        if (!c.ok) return;
        size_t off[2];
        const size_t b = scratch_elems<T>(c, n, B, off) * sizeof(T);
        need = b > need ? b : need;
    };
    for (int d = 0; d < 2; ++d) {
        const bool fw = d != 0;
        one(plan_split<T, 2>(n, L, fw)); one(plan_split<T, 4>(n, L, fw)); one(plan_split<T, 6>(n, L, fw)); one(plan_split<T, 8>(n, L, fw));
        one(plan_split<T, 10>(n, L, fw)); one(plan_split<T, 12>(n, L, fw)); one(plan_split<T, 14>(n, L, fw)); one(plan_split<T, 16>(n, L, fw));
        one(plan_split<T, 18>(n, L, fw)); one(plan_split<T, 20>(n, L, fw)); one(plan_split<T, 22>(n, L, fw)); one(plan_split<T, 24>(n, L, fw));
    }
    return need;
}

size_t fused_workspace_bytes(const ArrayGeom &g, int esize, int L, bool lifting, bool inplace, uint32_t flags) {
    (void)inplace;
    if (lifting || g.ndim != 1 || g.C != 1 || L < 1 || (flags & WB200_FLAG_FORCE_GENERIC)) return 0;
    const size_t need = (esize == 4) ? fused_ws_T<float>(g.dim[0], g.batch, L) : fused_ws_T<double>(g.dim[0], g.batch, L);
    return (need + 255) & ~(size_t)255;
}

template int32_t fused_dwt<float>(const PassOp<float> &, float *, const float *, const ArrayGeom &, int, bool, void *, size_t, cudaStream_t, uint32_t);
template int32_t fused_dwt<double>(const PassOp<double> &, double *, const double *, const ArrayGeom &, int, bool, void *, size_t, cudaStream_t, uint32_t);

} // namespace wb
