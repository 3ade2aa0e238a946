// wptfused.cu -- K consecutive FULL wavelet-packet levels of large nodes in ONE launch per direction (sm_100a).
//
// The packet sweeps above the on-chip subtrees (fastpass.cu: k_wpt_sub_*) used to be one HBM round trip per tree level
// (k_line_ana / k_line_syn).  Here a CTA stages one tile (+ halo) of a node with TMA bulk copies and runs K packet levels
// out of shared memory, keeping EVERY band (transforms_filter.jl:309-359: a full level splits each node [a | d] in place):
//
//   forward  k_pkt_ana : level l holds 2^l bands of NA[l] samples each (tile/2^l owned + the halo the remaining levels
//                        need); the detail band is produced from the same register window as the approximation, i.e. DS
//                        samples further on (cf. FGeom), so band beta of level l starts o_beta = o_parent/2 + (beta&1) DS
//                        samples after the tile's own position -- the offsets only matter when the level-K bands are
//                        stored (rotated, with the periodic wrap of the band).  K <= 1 + ctz(DS) keeps o_parent even.
//   inverse  k_pkt_syn : the 2^K level-K bands are staged on [lo[K], hi[K]) around the tile; level l-1 band b is
//                        synthesised from bands 2b (approximation) and 2b+1 (detail) of level l on a range that keeps Q4
//                        samples on both sides for the next level; the last level streams the tile with 128-bit stores.
//
// Arithmetic and summation order are those of ana_level / syn_level (fused1d_dev.cuh), i.e. the reference's; STRICT keeps
// products and sums separately rounded (bit-identical to the CPU path).
#include "fused1d_dev.cuh"

#include <cstdlib>
#include <cstring>

namespace wb {

constexpr int PKT_MAXK = 4;

struct PktAnaPlan {
    int K, tile, h0;
    int NA[PKT_MAXK + 1];      // valid samples per band at level l (index 0: staged input)
    int S[PKT_MAXK + 1];       // band stride in shared memory at level l
    int boff;                  // element offset of the second ping-pong buffer
};
struct PktSynPlan {
    int K, tile;
    int lo[PKT_MAXK + 1], hi[PKT_MAXK + 1];   // band range [lo, hi) relative to the tile position at level l (multiples of 8)
    int S[PKT_MAXK + 1];                       // band stride in shared memory at level l
    int qoff;                                  // element offset of the second ping-pong buffer
};

// even detail shift shared by Float32 (two pairs per thread-iteration) and Float64 (one)
template <int F> struct PktGeom {
    static constexpr int Q = F / 2;
    static constexpr int DS = ((Q - 1) + 1) & ~1;
    static constexpr int WO = 2 * DS - (F - 2);
    template <int PA> static constexpr int win() { return F + 2 * (PA - 1) + WO; }
};

// a[p .. p+PA) and d[p + DS .. p + DS + PA) of one band from the window at in[2p]
template <typename T, int F, bool STRICT, int PA>
__device__ __forceinline__ void pkt_pairs(const T *__restrict__ w2p, const Taps<T, F> &c, T (&a)[PA], T (&d)[PA]) {
    using fp = FP<STRICT>;
    using G = PktGeom<F>;
    constexpr int WIN = (G::template win<PA>() + 3) & ~3;
    T w[WIN];
    load_window<WIN>(w, w2p);
#pragma unroll
    for (int r = 0; r < PA; ++r) a[r] = fp::mul(c.h[0], w[2 * r]);
#pragma unroll
    for (int m = 1; m < F; ++m)
#pragma unroll
        for (int r = 0; r < PA; ++r) a[r] = fp::mac(a[r], c.h[m], w[2 * r + m]);
#pragma unroll
    for (int r = 0; r < PA; ++r) d[r] = fp::mul(c.g[F - 1], w[G::WO + 2 * r]);
#pragma unroll
    for (int q = 1; q < F; ++q)
#pragma unroll
        for (int r = 0; r < PA; ++r) d[r] = fp::mac(d[r], c.g[F - 1 - q], w[G::WO + 2 * r + q]);
}

// grid.x = ((b * nodes + q) * ntiles + t): tile t of node q (nj samples, period nj) of signal b
template <typename T, int F, bool STRICT>
__global__ void __launch_bounds__(256)
k_pkt_ana(const T *__restrict__ src, T *__restrict__ dst, int64_t n, int64_t nj, int64_t nodes, int ntiles,
          const __grid_constant__ Taps<T, F> c, const __grid_constant__ PktAnaPlan pl) {
    constexpr int PA = AnaPairs<T>::value;
    using G = PktGeom<F>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw);
    T *bufA = reinterpret_cast<T *>(smem_raw + 128);
    T *bufB = bufA + pl.boff;
    const int64_t bid = blockIdx.x;
    const int t = (int)(bid % ntiles);
    const int64_t qb = bid / ntiles;                 // b * nodes + q: nodes of a signal are back to back, so node qb starts at qb * nj
    (void)nodes; (void)n;
    const T *line = src + qb * nj;
    T *onode = dst + qb * nj;
    const int64_t s = (int64_t)t * pl.tile;
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        const int count = pl.tile + pl.h0;
        mbar_expect_tx(bar, (uint32_t)(count * sizeof(T)));
        tma_load_wrapped<T>(bufA, line, s, count, nj, bar);
    }
    __syncthreads();
    mbar_wait(bar, 0);

    const int K = pl.K;
    const int step = PA * blockDim.x;
    const T *in = bufA;
    T *out = bufB;
    // every level walks the flattened (band, pair group) index so that all threads stay busy when a band has fewer pair
    // groups than the CTA has threads (level 4 of a 4096-sample tile: 128 groups per band, 8 parent bands)
    const int nt = blockDim.x;
    for (int l = 1; l < K; ++l) {
        const int nb = 1 << (l - 1), cnt = pl.NA[l] / PA, Sp = pl.S[l - 1], Sl = pl.S[l];
        int bb = threadIdx.x / cnt, it = threadIdx.x - bb * cnt;
        while (bb < nb) {
            const int p = PA * it;
            const T *ib = in + bb * Sp;
            T *oa = out + (2 * bb) * Sl, *od = oa + Sl;
            T a[PA], d[PA];
            pkt_pairs<T, F, STRICT, PA>(ib + 2 * p, c, a, d);
            if constexpr (PA == 2) { store2(oa + p, a[0], a[1]); store2(od + p, d[0], d[1]); }
            else { oa[p] = a[0]; od[p] = d[0]; }
            it += nt;
            while (it >= cnt) { it -= cnt; ++bb; }
        }
        __syncthreads();
        const T *tt = in; in = out; out = const_cast<T *>(tt);
    }
    // last level: bands 2bb, 2bb+1 of level K go to HBM, each rotated by its accumulated detail shift
    {
        const int nb = 1 << (K - 1), cnt = pl.NA[K] / PA, Sp = pl.S[K - 1];
        const int lenK = (int)(nj >> K), sK = (int)(s >> K);      // 32-bit band arithmetic (nj <= 2^30); shifts are < lenK: one conditional wrap each
        int bb = threadIdx.x / cnt, it = threadIdx.x - bb * cnt;
        while (bb < nb) {
            int op = 0;                                   // offset of the parent band bb (level K-1): sum of its detail bits' shifts
            for (int i = K - 2; i >= 0; --i) op = (op >> 1) + (((bb >> i) & 1) ? G::DS : 0);
            // children: o = op/2 (approximation), op/2 + DS (detail); band beta lives at onode + beta*lenK, sample index mod lenK
            int sa = sK + (op >> 1);
            if (sa >= lenK) sa -= lenK;
            int sd = sa + G::DS;
            if (sd >= lenK) sd -= lenK;
            T *pa = onode + (int64_t)(2 * bb) * lenK, *pd = pa + lenK;
            const bool vec = PA == 1 || (((sa | sd) & 1) == 0);   // even starts: the pair stores stay 8-byte aligned and never straddle the wrap
            const int p = PA * it;
            T a[PA], d[PA];
            pkt_pairs<T, F, STRICT, PA>(in + bb * Sp + 2 * p, c, a, d);
            int ia = sa + p, id = sd + p;
            if (ia >= lenK) ia -= lenK;
            if (id >= lenK) id -= lenK;
            if constexpr (PA == 2) {
                if (vec) { gstore2(pa + ia, a[0], a[1]); gstore2(pd + id, d[0], d[1]); }
                else {
                    const int ia1 = ia + 1 >= lenK ? ia + 1 - lenK : ia + 1, id1 = id + 1 >= lenK ? id + 1 - lenK : id + 1;
                    __stcs(pa + ia, a[0]); __stcs(pa + ia1, a[1]);
                    __stcs(pd + id, d[0]); __stcs(pd + id1, d[1]);
                }
            } else { __stcs(pa + ia, a[0]); __stcs(pd + id, d[0]); }
            it += nt;
            while (it >= cnt) { it -= cnt; ++bb; }
        }
    }
}

template <typename T, int F, bool STRICT>
__global__ void __launch_bounds__(256)
k_pkt_syn(const T *__restrict__ src, T *__restrict__ dst, int64_t n, int64_t nj, int64_t nodes, int ntiles,
          const __grid_constant__ Taps<T, F> c, const __grid_constant__ PktSynPlan pl) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw);
    T *bufP = reinterpret_cast<T *>(smem_raw + 128);
    T *bufQ = bufP + pl.qoff;
    const int64_t bid = blockIdx.x;
    const int t = (int)(bid % ntiles);
    const int64_t qb = bid / ntiles;
    (void)nodes; (void)n;
    const T *inode = src + qb * nj;
    T *oline = dst + qb * nj;
    const int64_t s = (int64_t)t * pl.tile;
    const int K = pl.K;
    const int64_t lenK = nj >> K, sK = s >> K;
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        const int count = pl.hi[K] - pl.lo[K];
        mbar_expect_tx(bar, (uint32_t)((size_t)count * sizeof(T) << K));
        for (int beta = 0; beta < (1 << K); ++beta)
            tma_load_wrapped<T>(bufP + beta * pl.S[K], inode + (int64_t)beta * lenK, sK + pl.lo[K], count, lenK, bar);
    }
    __syncthreads();
    mbar_wait(bar, 0);

    const T *in = bufP;
    T *out = bufQ;
    using G = FGeom<F>;
    const int nt = blockDim.x;
    for (int l = K; l >= 1; --l) {
        const int nb = 1 << (l - 1), Sl = pl.S[l];
        const int npairs = (pl.hi[l - 1] - pl.lo[l - 1]) >> 1;
        const int oa = (pl.lo[l - 1] >> 1) - pl.lo[l];                     // index of the first output pair's a / d inside a band buffer
        // flattened (band, pair group) walk: a level-K band of a 4096-sample tile has 68 four-pair groups, far fewer than threads
        const bool quad = syn_quad_ok<T, F>(oa, oa, npairs);
        const int per = quad ? 4 : 2, cnt = npairs / per;
        const int back = quad ? G::Q4 : G::QA;
        const int So = l > 1 ? pl.S[l - 1] : 0;
        T *o0 = oline + s;
        int bb = threadIdx.x / cnt, it = threadIdx.x - bb * cnt;
        while (bb < nb) {
            const T *pa = in + (2 * bb) * Sl + oa - back, *pd = in + (2 * bb + 1) * Sl + oa;
            T *ob = out + bb * So;
            auto so = [&](int ur, T v0, T v1, T v2, T v3) {
                if (l > 1) store4(ob + 2 * ur, v0, v1, v2, v3); else gstore4(o0 + 2 * ur, v0, v1, v2, v3);
            };
            if constexpr (sizeof(T) == 4) {
                if (quad) syn_quad<T, F, STRICT>(pa, pd, 4 * it, c, so); else syn_duo<T, F, STRICT>(pa, pd, 2 * it, c, so);
            } else {
                syn_duo<T, F, STRICT>(pa, pd, 2 * it, c, so);
            }
            it += nt;
            while (it >= cnt) { it -= cnt; ++bb; }
        }
        if (l > 1) {
            __syncthreads();
            const T *tt = in; in = out; out = const_cast<T *>(tt);
        }
    }
}

// ===================================================================================================
// host side
// ===================================================================================================
static int env_int_pk(const char *name, int dflt) { const char *v = std::getenv(name); return (v && *v) ? std::atoi(v) : dflt; }

// most levels one launch can fuse for this filter: the detail shift must halve cleanly K-1 times
int wpt_fused_max_levels(int F) {
    if (F < 2 || F > 20 || (F & 1)) return 0;
    if (env_int_pk("WB200_DISABLE_WPTFUSED", 0)) return 0;
    const int Q = F / 2, DS = ((Q - 1) + 1) & ~1;
    int k = PKT_MAXK;
    if (DS > 0) { int z = 0; while (((DS >> z) & 1) == 0) ++z; k = 1 + z; }
    const int cap = env_int_pk("WB200_WPTFUSED_KMAX", PKT_MAXK);
    if (k > cap) k = cap;
    return k > PKT_MAXK ? PKT_MAXK : k;
}

template <typename T, int F> static bool make_pkt_ana_plan(PktAnaPlan &pl, int K, int tile, int64_t nj) {
    constexpr int PA = AnaPairs<T>::value;
    using G = PktGeom<F>;
    constexpr int WIN = (G::template win<PA>() + 3) & ~3;        // what pkt_pairs loads (16-byte granular)
    const int vec = 16 / (int)sizeof(T);
    pl.K = K; pl.tile = tile;
    if ((tile >> K) < 8 || (tile % (1 << K)) != 0) return false;
    pl.NA[K] = tile >> K;
    for (int l = K; l >= 1; --l) pl.NA[l - 1] = (2 * pl.NA[l] + WIN - 2 * PA + 3) & ~3;    // last window starts at 2 (NA[l] - PA)
    pl.h0 = (pl.NA[0] - tile + vec - 1) / vec * vec;
    if (pl.h0 < 0) return false;
    pl.NA[0] = tile + pl.h0;
    if ((int64_t)pl.NA[0] > nj) return false;                    // a bulk copy wraps at most once
    size_t a = 0, b = 0;
    for (int l = 0; l <= K; ++l) {
        pl.S[l] = ((pl.NA[l] + 3) & ~3) + 4;
        const size_t sz = (size_t)pl.S[l] << l;
        if (l < K) { if (l & 1) b = sz > b ? sz : b; else a = sz > a ? sz : a; }
    }
    pl.boff = (int)((a + 7) & ~(size_t)7);
    return (128 + (pl.boff + b + 8) * sizeof(T)) <= 160 * 1024;
}
template <typename T> static size_t pkt_ana_smem(const PktAnaPlan &pl) {
    size_t b = 0;
    for (int l = 1; l < pl.K; l += 2) { const size_t sz = (size_t)pl.S[l] << l; b = sz > b ? sz : b; }
    return 128 + ((size_t)pl.boff + b + 8) * sizeof(T);
}

template <typename T, int F> static bool make_pkt_syn_plan(PktSynPlan &pl, int K, int tile, int64_t nj, size_t &smem) {
    using G = FGeom<F>;
    auto dn8 = [](int v) { return (v >= 0) ? (v & ~7) : -(((-v) + 7) & ~7); };
    auto up8 = [](int v) { return (v + 7) & ~7; };
    pl.K = K; pl.tile = tile;
    if ((tile >> K) < 8 || (tile % (8 << K)) != 0) return false;
    pl.lo[0] = 0; pl.hi[0] = tile;
    for (int l = 1; l <= K; ++l) {
        pl.lo[l] = dn8(pl.lo[l - 1] / 2 - G::Q4);
        pl.hi[l] = up8(pl.hi[l - 1] / 2 + G::Q4);
    }
    if ((int64_t)(pl.hi[K] - pl.lo[K]) > (nj >> K)) return false;  // a bulk copy wraps at most once
    size_t p = 0, q = 0;
    for (int l = K; l >= 1; --l) {
        pl.S[l] = (pl.hi[l] - pl.lo[l]) + 4;
        const size_t sz = (size_t)pl.S[l] << l;
        if (((K - l) & 1) == 0) p = sz > p ? sz : p; else q = sz > q ? sz : q;
    }
    pl.S[0] = tile;
    pl.qoff = (int)((p + 7) & ~(size_t)7);
    smem = 128 + ((size_t)pl.qoff + q + 8) * sizeof(T);
    return smem <= 160 * 1024;
}

template <typename T> static int pkt_tile(int64_t nj) {
    int64_t tile = env_int_pk(sizeof(T) == 4 ? "WB200_WPTFUSED_TILE_F32" : "WB200_WPTFUSED_TILE_F64", sizeof(T) == 4 ? 4096 : 2048);
    const int64_t p2 = nj & (-nj);
    while (tile > p2) tile >>= 1;
    while (tile > nj / 2) tile >>= 1;                            // at least two tiles per node
    return (int)tile;
}

template <typename T, int F, bool STRICT>
static int wpt_fused_F(const T *S, T *D, int64_t n, int64_t nj, int K, int64_t nodes, int64_t B, const FilterCoefs<T> &fc,
                       bool fw, cudaStream_t st) {
    Taps<T, F> taps;
    for (int k = 0; k < F; ++k) { taps.h[k] = fc.h[k]; taps.g[k] = fc.g[k]; }
    const int tile = pkt_tile<T>(nj);
    if (tile < 256) return 0;
    const int64_t ntiles = nj / tile;
    const int64_t nblk = ntiles * nodes * B;
    if (nblk > 0x7fffffffLL || nblk <= 0) return 0;
    const int nthr = env_int_pk("WB200_WPTFUSED_NT", 128) & ~31;
    if (fw) {
        PktAnaPlan pl;
        std::memset(&pl, 0, sizeof(pl));
        if (!make_pkt_ana_plan<T, F>(pl, K, tile, nj)) return 0;
        const size_t smem = pkt_ana_smem<T>(pl);
        auto kern = k_pkt_ana<T, F, STRICT>;
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { (void)cudaGetLastError(); return 0; }
        LaunchScope scope("wpt_fused_levels_analysis", st);
        kern<<<(unsigned)nblk, (nthr >= 32 && nthr <= 256) ? nthr : 256, smem, st>>>(S, D, n, nj, nodes, (int)ntiles, taps, pl);
    } else {
        PktSynPlan pl;
        std::memset(&pl, 0, sizeof(pl));
        size_t smem = 0;
        if (!make_pkt_syn_plan<T, F>(pl, K, tile, nj, smem)) return 0;
        auto kern = k_pkt_syn<T, F, STRICT>;
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { (void)cudaGetLastError(); return 0; }
        LaunchScope scope("wpt_fused_levels_synthesis", st);
        kern<<<(unsigned)nblk, (nthr >= 32 && nthr <= 256) ? nthr : 256, smem, st>>>(S, D, n, nj, nodes, (int)ntiles, taps, pl);
    }
    return check_launch("wpt_fused_levels") ? 1 : -1;
}

template <typename T, int F> static bool pkt_plans_ok(int64_t nj, int K) {
    const int tile = pkt_tile<T>(nj);
    if (tile < 256) return false;
    PktAnaPlan pa; PktSynPlan ps; size_t smem;
    std::memset(&pa, 0, sizeof(pa)); std::memset(&ps, 0, sizeof(ps));
    return make_pkt_ana_plan<T, F>(pa, K, tile, nj) && make_pkt_syn_plan<T, F>(ps, K, tile, nj, smem);
}
// whether fast_wpt_fused_levels takes K levels below nodes of nj samples (both directions): the WPT driver plans its sweeps
// with this, so a planned sweep is never rejected at launch time
bool wpt_fused_ok(int esize, int F, int64_t nj, int K) {
    if (K < 2 || K > wpt_fused_max_levels(F)) return false;
    if (nj > ((int64_t)1 << 30) || (nj % ((int64_t)8 << K)) != 0) return false;
    switch (F) {
#define WB_CASE(FF) case FF: return esize == 4 ? pkt_plans_ok<float, FF>(nj, K) : pkt_plans_ok<double, FF>(nj, K);
        WB_CASE(2) WB_CASE(4) WB_CASE(6) WB_CASE(8) WB_CASE(10) WB_CASE(12) WB_CASE(14) WB_CASE(16) WB_CASE(18) WB_CASE(20)
#undef WB_CASE
    default: return false;
    }
}

// `nodes` full nodes of nj samples each per signal (level lv0 of the tree: nodes * nj == n), K full levels below them.
// 1 handled / 0 not covered (the caller falls back to one sweep per level) / -1 error.  S != D.
template <typename T>
int fast_wpt_fused_levels(const T *S, T *D, int64_t n, int64_t nj, int K, int64_t nodes, int64_t B,
                          const FilterCoefs<T> &fc, bool strict, bool fw, cudaStream_t st) {
    if (K < 2 || K > wpt_fused_max_levels(fc.F) || S == D) return 0;
    if (nj > ((int64_t)1 << 30) || (nj % ((int64_t)8 << K)) != 0) return 0;
    if ((((uintptr_t)S | (uintptr_t)D) & 15) != 0 || ((n * (int64_t)sizeof(T)) % 16) != 0) return 0;
    switch (fc.F) {
#define WB_CASE(FF) case FF: return strict ? wpt_fused_F<T, FF, true>(S, D, n, nj, K, nodes, B, fc, fw, st) : wpt_fused_F<T, FF, false>(S, D, n, nj, K, nodes, B, fc, fw, st);
        WB_CASE(2) WB_CASE(4) WB_CASE(6) WB_CASE(8) WB_CASE(10) WB_CASE(12) WB_CASE(14) WB_CASE(16) WB_CASE(18) WB_CASE(20)
#undef WB_CASE
    default: return 0;
    }
}

template int fast_wpt_fused_levels<float>(const float *, float *, int64_t, int64_t, int, int64_t, int64_t, const FilterCoefs<float> &, bool, bool, cudaStream_t);
template int fast_wpt_fused_levels<double>(const double *, double *, int64_t, int64_t, int, int64_t, int64_t, const FilterCoefs<double> &, bool, bool, cudaStream_t);

} // namespace wb
