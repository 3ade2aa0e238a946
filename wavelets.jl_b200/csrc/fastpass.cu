// fastpass.cu -- fast single-level filter-bank passes for the N-D and wavelet-packet drivers (sm_100a).
//
// The N-D drivers (api.cu run_nd) and the packet driver (run_wpt) work one level and one dimension at a time.
// Two kernel families replace the generic one-thread-per-output pass where the layout allows it:
//
//   line kernels  (k_line_ana / k_line_syn): the lines are CONTIGUOUS (dim-1 pass of an N-D level, packet nodes).
//       Same machinery as the fused 1-D tiles with K = 1: a CTA stages TILE (+halo) samples of one line with a TMA
//       bulk copy (second copy for the periodic wrap), computes approximation + detail from 128-bit shared-memory
//       windows and streams both halves out with 64/128-bit stores.
//   walk kernels  (k_walk_ana / k_walk_syn): the lines are STRIDED and another coordinate is contiguous (dim-2 /
//       dim-3 passes).  Threads run along the contiguous coordinate (every load and store is a coalesced 128-byte
//       row), each thread holds a short run of its line in registers (2*RK + F - 2 samples, all loads issued up
//       front) and emits RK approximation/detail pairs.  No shared memory, no transposes.
//
// Arithmetic: the closed forms of filtdown!/filtup! in the reference's summation order; the detail
// is produced from the same window as the approximation (d[k + F/2 - 1] and a[k] read the same F inputs).
#include "fused1d_dev.cuh"
#include <cstring>

namespace wb {

static int env_fast_enabled() {
    const char *v = getenv("WB200_DISABLE_FASTPASS");
    return (v && *v && atoi(v)) ? 0 : 1;
}

struct LineCoord { int64_t c0, c1, c2, c3; };
__device__ __forceinline__ LineCoord split_line(int64_t ln, const Extent &e) {
    LineCoord c;
    c.c0 = ln % e.n[0]; ln /= e.n[0];
    c.c1 = ln % e.n[1]; ln /= e.n[1];
    c.c2 = ln % e.n[2];
    c.c3 = ln / e.n[2];
    return c;
}

// ===================================================================================================
// contiguous lines: one level per launch out of TMA-staged tiles
// ===================================================================================================
template <typename T, int F, bool STRICT>
__global__ void __launch_bounds__(256)
k_line_ana(View<const T> src, View<T> dlo, View<T> dhi, Extent e, int tile, int h0, int ntiles,
           const __grid_constant__ Taps<T, F> c) {
    constexpr int PA = AnaPairs<T>::value;
    using G = FGeom<F, PA>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw);
    T *buf = reinterpret_cast<T *>(smem_raw + 128);
    const int64_t bid = blockIdx.x;
    const int64_t ln = bid / ntiles;
    const int t = (int)(bid - ln * ntiles);
    const LineCoord lc = split_line(ln, e);
    const T *x = src.line(lc.c0, lc.c1, lc.c2, lc.c3);
    T *lo = dlo.line(lc.c0, lc.c1, lc.c2, lc.c3);
    T *hi = dhi.line(lc.c0, lc.c1, lc.c2, lc.c3);
    const int64_t n = e.len, nh = n >> 1;
    const int64_t s = (int64_t)t * tile, sl = s >> 1;
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        mbar_expect_tx(bar, (uint32_t)((tile + h0) * sizeof(T)));
        tma_load_wrapped<T>(buf, x, s, tile + h0, n, bar);
    }
    __syncthreads();
    mbar_wait(bar, 0);
    const int ND = tile >> 1;
    T *dbase = hi + sl + G::DS;
    const int64_t room = nh - sl - G::DS;
    const int wrap_at = room < (int64_t)ND ? (int)room : 0x7fffffff;
    T *abase = lo + sl;
    auto store_a = [&](int p, const T (&a)[PA]) {
        if constexpr (PA == 2) gstore2(abase + p, a[0], a[1]); else __stcs(abase + p, a[0]);
    };
    ana_level<T, F, STRICT>(buf, ND, ND, c, store_a, dbase, dbase - nh, wrap_at);
}

template <typename T, int F, bool STRICT>
__global__ void __launch_bounds__(256)
k_line_syn(View<const T> slo, View<const T> shi, View<const T> salt, int64_t thr0, int64_t thr1, int64_t thr2, int64_t thr3,
           int has_alt, View<T> dst, Extent e, int ntiles, const __grid_constant__ Taps<T, F> c,
           const __grid_constant__ SynPlan pl) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw);
    T *sm = reinterpret_cast<T *>(smem_raw + 128);
    const int64_t bid = blockIdx.x;
    const int64_t ln = bid / ntiles;
    const int t = (int)(bid - ln * ntiles);
    const LineCoord lc = split_line(ln, e);
    const bool alt = has_alt && lc.c0 < thr0 && lc.c1 < thr1 && lc.c2 < thr2 && lc.c3 < thr3;
    const T *a = alt ? salt.line(lc.c0, lc.c1, lc.c2, lc.c3) : slo.line(lc.c0, lc.c1, lc.c2, lc.c3);
    const T *d = shi.line(lc.c0, lc.c1, lc.c2, lc.c3);
    T *o = dst.line(lc.c0, lc.c1, lc.c2, lc.c3);
    const int64_t nh = e.len >> 1;
    const int64_t s = (int64_t)t * pl.tile;
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        mbar_expect_tx(bar, (uint32_t)(((pl.rhi[1] - pl.rlo[1]) + (pl.dhi[1] - pl.dlo[1])) * sizeof(T)));
        tma_load_wrapped<T>(sm + pl.aoff, a, (s >> 1) + pl.rlo[1], pl.rhi[1] - pl.rlo[1], nh, bar);
        tma_load_wrapped<T>(sm + pl.doff[1], d, (s >> 1) + pl.dlo[1], pl.dhi[1] - pl.dlo[1], nh, bar);
    }
    __syncthreads();
    mbar_wait(bar, 0);
    const int npairs = pl.tile >> 1;
    const int oa = 0 - pl.rlo[1];
    const int od = 0 - pl.dlo[1];
    T *ob = o + s;
    auto so = [&](int ur, T o0, T o1, T o2, T o3) { gstore4(ob + 2 * ur, o0, o1, o2, o3); };
    syn_level<T, F, STRICT>(sm + pl.aoff, sm + pl.doff[1], oa, od, npairs, c, so);
}

// ===================================================================================================
// strided lines: register runs, coalesced across the contiguous coordinate
// ===================================================================================================
template <int F> struct WalkCfg { static constexpr int RK = F <= 12 ? 16 : 8; };

template <typename T, int F, bool STRICT>
__global__ void __launch_bounds__(128)
k_walk_ana(View<const T> src, View<T> dlo, View<T> dhi, Extent e, int nseg, const __grid_constant__ Taps<T, F> c) {
    using fp = FP<STRICT>;
    constexpr int RK = WalkCfg<F>::RK;
    constexpr int W = 2 * RK + F - 2;
    constexpr int DS = F / 2 - 1;
    const int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i0 >= e.n[0]) return;
    const int seg = blockIdx.y % nseg;
    const int64_t yy = blockIdx.y / nseg + (int64_t)blockIdx.z * (gridDim.y / nseg);
    const int64_t NY = e.n[1] * e.n[2] * e.n[3];
    if (yy >= NY) return;
    const int64_t c1 = yy % e.n[1], r = yy / e.n[1], c2 = r % e.n[2], c3 = r / e.n[2];
    const T *x = src.line(i0, c1, c2, c3);
    T *lo = dlo.line(i0, c1, c2, c3);
    T *hi = dhi.line(i0, c1, c2, c3);
    const int64_t n = e.len, nh = n >> 1;
    const int64_t k0 = (int64_t)seg * RK;
    T w[W];
    {
        int64_t idx = (2 * k0) % n;
#pragma unroll
        for (int m = 0; m < W; ++m) {
            w[m] = x[idx * src.ls];
            if (++idx == n) idx = 0;
        }
    }
#pragma unroll
    for (int k = 0; k < RK; ++k) {
        if (k0 + k < nh) {
            T a = fp::mul(c.h[0], w[2 * k]);
            T d = fp::mul(c.g[F - 1], w[2 * k]);
#pragma unroll
            for (int m = 1; m < F; ++m) {
                a = fp::mac(a, c.h[m], w[2 * k + m]);
                d = fp::mac(d, c.g[F - 1 - m], w[2 * k + m]);
            }
            lo[(k0 + k) * dlo.ls] = a;
            const int64_t kd = (k0 + k + DS) % nh;
            hi[kd * dhi.ls] = d;
        }
    }
}

template <typename T, int F, bool STRICT>
__global__ void __launch_bounds__(128)
k_walk_syn(View<const T> slo, View<const T> shi, View<const T> salt, int64_t thr0, int64_t thr1, int64_t thr2, int64_t thr3,
           int has_alt, View<T> dst, Extent e, int nseg, const __grid_constant__ Taps<T, F> c) {
    using fp = FP<STRICT>;
    constexpr int RK = WalkCfg<F>::RK;
    constexpr int Q = F / 2;
    constexpr int W = RK + Q - 1;
    const int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i0 >= e.n[0]) return;
    const int seg = blockIdx.y % nseg;
    const int64_t yy = blockIdx.y / nseg + (int64_t)blockIdx.z * (gridDim.y / nseg);
    const int64_t NY = e.n[1] * e.n[2] * e.n[3];
    if (yy >= NY) return;
    const int64_t c1 = yy % e.n[1], r = yy / e.n[1], c2 = r % e.n[2], c3 = r / e.n[2];
    const bool alt = has_alt && i0 < thr0 && c1 < thr1 && c2 < thr2 && c3 < thr3;
    const T *a;
    int64_t als;
    if (alt) { a = salt.line(i0, c1, c2, c3); als = salt.ls; } else { a = slo.line(i0, c1, c2, c3); als = slo.ls; }
    const T *d = shi.line(i0, c1, c2, c3);
    T *o = dst.line(i0, c1, c2, c3);
    const int64_t nh = e.len >> 1;
    const int64_t u0 = (int64_t)seg * RK;
    // wa[i] = a[(u0 - Q + 1 + i) mod nh], wd[i] = d[(u0 + i) mod nh]
    T wa[W], wd[W];
    {
        int64_t ia = (u0 - (Q - 1)) % nh;
        if (ia < 0) ia += nh;
        int64_t id = u0 % nh;
#pragma unroll
        for (int i = 0; i < W; ++i) {
            wa[i] = a[ia * als];
            wd[i] = d[id * shi.ls];
            if (++ia == nh) ia = 0;
            if (++id == nh) id = 0;
        }
    }
#pragma unroll
    for (int k = 0; k < RK; ++k) {
        if (u0 + k < nh) {
            // a[u - j] = wa[k + Q - 1 - j], d[u + j] = wd[k + j]
            T rae = fp::mul(c.h[2 * (Q - 1)], wa[k]);
            T rao = fp::mul(c.h[2 * (Q - 1) + 1], wa[k]);
#pragma unroll
            for (int j = Q - 2; j >= 0; --j) {
                rae = fp::mac(rae, c.h[2 * j], wa[k + Q - 1 - j]);
                rao = fp::mac(rao, c.h[2 * j + 1], wa[k + Q - 1 - j]);
            }
            T rde, rdo;      // fast mode: the detail terms continue the approximation's chain (no second FMUL, no final FADD)
            if constexpr (STRICT) { rde = fp::mul(c.g[1], wd[k]); rdo = fp::mul(c.g[0], wd[k]); }
            else { rde = fp::mac(rae, c.g[1], wd[k]); rdo = fp::mac(rao, c.g[0], wd[k]); }
#pragma unroll
            for (int j = 1; j < Q; ++j) {
                rde = fp::mac(rde, c.g[2 * j + 1], wd[k + j]);
                rdo = fp::mac(rdo, c.g[2 * j], wd[k + j]);
            }
            o[(2 * (u0 + k)) * dst.ls] = (STRICT ? fp::add(rae, rde) : rde);
            o[(2 * (u0 + k) + 1) * dst.ls] = (STRICT ? fp::add(rao, rdo) : rdo);
        }
    }
}

// ===================================================================================================
// wavelet-packet subtrees: a node of <= WPT_SUB_MAX samples is decomposed (or rebuilt) through ALL of its
// remaining full levels by one CTA in shared memory -- one launch instead of one per level, one HBM read and one
// write of the node.  Sub-node j of level l occupies [j*ml, (j+1)*ml) of the node's span and is replaced in place by
// [approximation | detail] (natural / Paley order, transforms_filter.jl:337-353).
// ===================================================================================================
static int env_int_fp(const char *name, int dflt) { const char *v = std::getenv(name); return (v && *v) ? std::atoi(v) : dflt; }
constexpr int WPT_SUB_MAX = 4096;        // default.  r02 sweep (sym8, 2^16, 1024 signals, wpt + iwpt): 4096 -> 2.40 ms, 8192 -> 2.35 ms
                                         // (one HBM sweep less, three resident CTAs instead of seven), 16384 -> 3.70 ms (one CTA per SM)
int wpt_subtree_max_samples(int esize) {
    const char *e = std::getenv("WB200_WPT_SUBMAX");
    int v = (e && *e) ? std::atoi(e) : WPT_SUB_MAX;
    (void)esize;
    if (v > 8192) v = 8192;
    int p = 2;
    while (p * 2 <= v) p *= 2;               // a power of two
    return p;
}

// The per-level work is FP32-bound (2 F multiply-adds per sample per level), so the shared-memory side has to stay out
// of the way: a thread computes FOUR consecutive output pairs from register windows filled with conflict-free 16-byte
// loads.  Analysis reads a sub-node through its two polyphase components -- stored SPLIT, [even samples | odd samples],
// by the level above (or by the global load) -- so that pair k needs E[k .. k+Q-1], O[k .. k+Q-1] (approximation) and
// E[k-Q+1 .. k], O[k-Q+1 .. k] (detail), Q = F/2: consecutive lanes read consecutive 16-byte chunks.  The periodic wrap
// is applied to chunk indices (sub-node half-lengths that are multiples of 4).  Sub-nodes of 4 or 2 samples are one
// register-resident node per thread with the wrap resolved at compile time; anything else takes the per-pair loop.
template <typename T> struct Vec4 { T v[4]; };
template <typename T> __device__ __forceinline__ void ld4(T (&w)[4], const T *p) {
    if constexpr (sizeof(T) == 4) { const float4 q = *reinterpret_cast<const float4 *>(p); w[0] = q.x; w[1] = q.y; w[2] = q.z; w[3] = q.w; }
    else { const double2 q0 = *reinterpret_cast<const double2 *>(p), q1 = *reinterpret_cast<const double2 *>(p + 2); w[0] = q0.x; w[1] = q0.y; w[2] = q1.x; w[3] = q1.y; }
}
template <typename T> __device__ __forceinline__ void st4(T *p, T a, T b, T c, T d) {
    if constexpr (sizeof(T) == 4) *reinterpret_cast<float4 *>(p) = make_float4(a, b, c, d);
    else { *reinterpret_cast<double2 *>(p) = make_double2(a, b); *reinterpret_cast<double2 *>(p + 2) = make_double2(c, d); }
}
template <typename T> __device__ __forceinline__ void st2(T *p, T a, T b) {
    if constexpr (sizeof(T) == 4) *reinterpret_cast<float2 *>(p) = make_float2(a, b);
    else *reinterpret_cast<double2 *>(p) = make_double2(a, b);
}

// a[k] = sum_t h[t] x[(2k+t) mod ML] (t increasing), d[k] = sum_i g[F-1-i] x[(2k+2-F+i) mod ML] (i increasing): one
// node of ML in {2, 4} samples in registers, x given as xe[] (even samples) and xo[] (odd samples)
template <typename T, int F, bool STRICT, int ML>
__device__ __forceinline__ void tiny_ana(const T (&xe)[ML / 2], const T (&xo)[ML / 2], T (&a)[ML / 2], T (&d)[ML / 2], const Taps<T, F> &c) {
    using fp = FP<STRICT>;
    constexpr int NH = ML / 2;
#pragma unroll
    for (int k = 0; k < NH; ++k) {
        T acc = fp::mul(c.h[0], xe[k % NH]);
#pragma unroll
        for (int t = 1; t < F; ++t) {
            const int i = (2 * k + t) % ML;
            acc = fp::mac(acc, c.h[t], (i & 1) ? xo[i >> 1] : xe[i >> 1]);
        }
        a[k] = acc;
        constexpr int BIG = ML * F;      // keeps the modulus argument non-negative
        const int i0 = (2 * k + 2 - F + BIG) % ML;
        T q = fp::mul(c.g[F - 1], (i0 & 1) ? xo[i0 >> 1] : xe[i0 >> 1]);
#pragma unroll
        for (int t = 1; t < F; ++t) {
            const int i = (2 * k + 2 - F + t + BIG) % ML;
            q = fp::mac(q, c.g[F - 1 - t], (i & 1) ? xo[i >> 1] : xe[i >> 1]);
        }
        d[k] = q;
    }
}
// x[2u] / x[2u+1] from the bands a[], d[] of NH in {1, 2} samples (generic_kernels.cu: k_filter_synthesis order)
template <typename T, int F, bool STRICT, int NH>
__device__ __forceinline__ void tiny_syn(const T (&a)[NH], const T (&d)[NH], T (&x)[2 * NH], const Taps<T, F> &c) {
    using fp = FP<STRICT>;
    constexpr int Q = F / 2;
    constexpr int BIG = NH * Q;
#pragma unroll
    for (int u = 0; u < NH; ++u) {
        T rae = fp::mul(c.h[2 * (Q - 1)], a[(u - (Q - 1) + BIG) % NH]);
        T rao = fp::mul(c.h[2 * (Q - 1) + 1], a[(u - (Q - 1) + BIG) % NH]);
#pragma unroll
        for (int t = Q - 2; t >= 0; --t) {
            rae = fp::mac(rae, c.h[2 * t], a[(u - t + BIG) % NH]);
            rao = fp::mac(rao, c.h[2 * t + 1], a[(u - t + BIG) % NH]);
        }
        T rde, rdo;      // fast mode: the detail terms continue the approximation's chain (no second FMUL, no final FADD)
        if constexpr (STRICT) { rde = fp::mul(c.g[1], d[u % NH]); rdo = fp::mul(c.g[0], d[u % NH]); }
        else { rde = fp::mac(rae, c.g[1], d[u % NH]); rdo = fp::mac(rao, c.g[0], d[u % NH]); }
#pragma unroll
        for (int t = 1; t < Q; ++t) {
            rde = fp::mac(rde, c.g[2 * t + 1], d[(u + t) % NH]);
            rdo = fp::mac(rdo, c.g[2 * t], d[(u + t) % NH]);
        }
        x[2 * u] = (STRICT ? fp::add(rae, rde) : rde);
        x[2 * u + 1] = (STRICT ? fp::add(rao, rdo) : rdo);
    }
}

// One analysis level of the sub-nodes of nh-sample components (nh a multiple of 4) held split in shared memory: a thread
// takes one 16-byte chunk of output pairs of one sub-node.  The detail pairs it produces are the ones DS = 4 CL further on
// (d[k + DS] reads x[2k + 2 DS + 2 - F ...]: with DS >= Q - 1 that is the same register window as a[k], cf. FGeom), so a
// thread loads CL + 1 chunks per component instead of 2 CL + 1 (sym8: 6 LDS.128 per 128 FMA instead of 10) and stores its
// detail chunk CL chunks further round the node.  P2 (the chunk count per component is a power of two -- every level of a
// dyadic signal): node / chunk indices and the periodic wrap are shifts and masks; otherwise divisions and
// compare-and-reset (the division form cost ~110 of the 260 instructions of an iteration).
template <typename T, int F, bool STRICT, bool P2>
__device__ __forceinline__ void wpt_sub_ana_chunks(const T *__restrict__ in, T *__restrict__ out, int m, int nh, bool last,
                                                   const Taps<T, F> &c) {
    using fp = FP<STRICT>;
    constexpr int Q = F / 2;
    constexpr int CL = (Q - 1 + 3) / 4;          // chunks of reach; DS = 4 CL
    constexpr int NCH = CL + 1;
    constexpr int DO = 4 * CL + 1 - Q;           // window offset of the detail taps (>= 0)
    const int mh = m >> 1, hh = nh >> 1, ml = nh << 1;
    const int cpn = nh >> 2;
    const int sh = P2 ? (31 - __clz(cpn)) : 0, mask = cpn - 1;
    for (int idx = threadIdx.x; idx < (m >> 3); idx += blockDim.x) {
        int j, c0;
        if constexpr (P2) { j = idx >> sh; c0 = idx & mask; } else { j = idx / cpn; c0 = idx - j * cpn; }
        const T *E = in + j * nh, *O = E + mh;
        T we[4 * NCH], wo[4 * NCH];
        int ci = c0;
#pragma unroll
        for (int i = 0; i < NCH; ++i) {
            const int cc = P2 ? ((c0 + i) & mask) : ci;
            T t4[4];
            ld4(t4, E + 4 * cc);
            we[4 * i] = t4[0]; we[4 * i + 1] = t4[1]; we[4 * i + 2] = t4[2]; we[4 * i + 3] = t4[3];
            ld4(t4, O + 4 * cc);
            wo[4 * i] = t4[0]; wo[4 * i + 1] = t4[1]; wo[4 * i + 2] = t4[2]; wo[4 * i + 3] = t4[3];
            if constexpr (!P2) { if (++ci == cpn) ci = 0; }
        }
        T av[4], dv[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            // a[k0 + r] = sum_t h[t] x[2 (k0 + r) + t]
            T acc = fp::mul(c.h[0], we[r]);
#pragma unroll
            for (int t = 1; t < F; ++t) acc = fp::mac(acc, c.h[t], (t & 1) ? wo[r + (t - 1) / 2] : we[r + t / 2]);
            av[r] = acc;
            // d[k0 + 4 CL + r] = sum_t g[F-1-t] x[2 (k0 + 4 CL + r) + 2 - F + t]
            T qd = fp::mul(c.g[F - 1], we[DO + r]);
#pragma unroll
            for (int t = 1; t < F; ++t) qd = fp::mac(qd, c.g[F - 1 - t], (t & 1) ? wo[DO + r + (t - 1) / 2] : we[DO + r + t / 2]);
            dv[r] = qd;
        }
        int cd;                                   // chunk the detail pairs belong to
        if constexpr (P2) cd = (c0 + CL) & mask; else cd = (c0 + CL) % cpn;
        const int k0 = 4 * c0, kd = 4 * cd;
        if (!last) {
            T *ea = out + (2 * j) * hh + (k0 >> 1), *ed = out + (2 * j + 1) * hh + (kd >> 1);
            st2(ea, av[0], av[2]); st2(ea + mh, av[1], av[3]);
            st2(ed, dv[0], dv[2]); st2(ed + mh, dv[1], dv[3]);
        } else {
            st4(out + j * ml + k0, av[0], av[1], av[2], av[3]);
            st4(out + j * ml + nh + kd, dv[0], dv[1], dv[2], dv[3]);
        }
    }
}

// One synthesis level of band-split sub-nodes (a_j at [j nh), d_j at m/2 + [j nh)); see wpt_sub_ana_chunks for P2.
template <typename T, int F, bool STRICT, bool P2>
__device__ __forceinline__ void wpt_sub_syn_chunks(const T *__restrict__ in, T *__restrict__ out, int m, int nh, const Taps<T, F> &c) {
    using fp = FP<STRICT>;
    constexpr int Q = F / 2;
    constexpr int CL = (Q - 1 + 3) / 4;
    const int mh = m >> 1, ml = nh << 1;
    const int cpn = nh >> 2;
    const int sh = P2 ? (31 - __clz(cpn)) : 0, mask = cpn - 1;
    for (int idx = threadIdx.x; idx < (m >> 3); idx += blockDim.x) {
        int j, c0;
        if constexpr (P2) { j = idx >> sh; c0 = idx & mask; } else { j = idx / cpn; c0 = idx - j * cpn; }
        const T *A = in + j * nh, *Dd = A + mh;
        T wa[4 * (CL + 1)], wd[4 * (CL + 1)];
        int ci = c0 - CL;
        if constexpr (!P2) { ci %= cpn; if (ci < 0) ci += cpn; }
#pragma unroll
        for (int i = 0; i <= CL; ++i) {
            const int cc = P2 ? ((ci + i) & mask) : ci;
            T t4[4];
            ld4(t4, A + 4 * cc);
            wa[4 * i] = t4[0]; wa[4 * i + 1] = t4[1]; wa[4 * i + 2] = t4[2]; wa[4 * i + 3] = t4[3];
            if constexpr (!P2) { if (++ci == cpn) ci = 0; }
        }
        ci = c0;
#pragma unroll
        for (int i = 0; i <= CL; ++i) {
            const int cc = P2 ? ((ci + i) & mask) : ci;
            T t4[4];
            ld4(t4, Dd + 4 * cc);
            wd[4 * i] = t4[0]; wd[4 * i + 1] = t4[1]; wd[4 * i + 2] = t4[2]; wd[4 * i + 3] = t4[3];
            if constexpr (!P2) { if (++ci == cpn) ci = 0; }
        }
        T xo[8];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            T rae = fp::mul(c.h[2 * (Q - 1)], wa[4 * CL + r - (Q - 1)]);
            T rao = fp::mul(c.h[2 * (Q - 1) + 1], wa[4 * CL + r - (Q - 1)]);
#pragma unroll
            for (int t = Q - 2; t >= 0; --t) {
                rae = fp::mac(rae, c.h[2 * t], wa[4 * CL + r - t]);
                rao = fp::mac(rao, c.h[2 * t + 1], wa[4 * CL + r - t]);
            }
            T rde, rdo;      // fast mode: the detail terms continue the approximation's chain (no second FMUL, no final FADD)
            if constexpr (STRICT) { rde = fp::mul(c.g[1], wd[r]); rdo = fp::mul(c.g[0], wd[r]); }
            else { rde = fp::mac(rae, c.g[1], wd[r]); rdo = fp::mac(rao, c.g[0], wd[r]); }
#pragma unroll
            for (int t = 1; t < Q; ++t) {
                rde = fp::mac(rde, c.g[2 * t + 1], wd[r + t]);
                rdo = fp::mac(rdo, c.g[2 * t], wd[r + t]);
            }
            xo[2 * r] = (STRICT ? fp::add(rae, rde) : rde);
            xo[2 * r + 1] = (STRICT ? fp::add(rao, rdo) : rdo);
        }
        T *o = out + (j & 1) * mh + (j >> 1) * ml + 8 * c0;
        st4(o, xo[0], xo[1], xo[2], xo[3]);
        st4(o + 4, xo[4], xo[5], xo[6], xo[7]);
    }
}

// FAST mode, trees that go all the way down (final nodes of one sample): the last four levels of a 16-sample node --
// node lengths 16, 8, 4, 2, every filter wrapped round the node several times -- are ONE fixed 16 x 16 linear map.  The host
// composes it in double precision from the same taps (leaf16_matrices); a thread then takes a whole node: 256 FMAs on
// constant-bank coefficients instead of 4 levels x 16 taps x 16 outputs = 1024 (sym8: a quarter of the subtree's arithmetic
// gone), no barriers, and the leaves go straight to HBM.  STRICT keeps the level-by-level reference order.
template <typename T> struct Leaf16 { T m[16][16]; };      // out[r] = sum_c m[r][c] in[c], both in natural order

// analysis: `in` holds the 16-sample nodes split (E_j at [8 j), O_j at m/2 + [8 j)); leaves of node j -> D[16 j ..]
template <typename T>
__device__ __forceinline__ void leaf16_ana(const T *__restrict__ in, T *__restrict__ D, int m, const Leaf16<T> &M) {
    const int mh = m >> 1;
    // one node per thread, NOT a loop: inside a loop the compiler hoists the 256 loop-invariant coefficients out of the constant
    // bank into registers and spills them; straight-line code reads them as FFMA constant operands.  (m / 16 <= blockDim.x: host)
    const int j = threadIdx.x;
    if (j < (m >> 4)) {
        T e[8], o[8];
        { T t4[4]; ld4(t4, in + 8 * j); e[0] = t4[0]; e[1] = t4[1]; e[2] = t4[2]; e[3] = t4[3];
          ld4(t4, in + 8 * j + 4); e[4] = t4[0]; e[5] = t4[1]; e[6] = t4[2]; e[7] = t4[3];
          ld4(t4, in + mh + 8 * j); o[0] = t4[0]; o[1] = t4[1]; o[2] = t4[2]; o[3] = t4[3];
          ld4(t4, in + mh + 8 * j + 4); o[4] = t4[0]; o[5] = t4[1]; o[6] = t4[2]; o[7] = t4[3]; }
#pragma unroll
        for (int r0 = 0; r0 < 16; r0 += 4) {
            T y[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                T acc = M.m[r0 + r][0] * e[0];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    if (i > 0) acc = fma(M.m[r0 + r][2 * i], e[i], acc);
                    acc = fma(M.m[r0 + r][2 * i + 1], o[i], acc);
                }
                y[r] = acc;
            }
            if constexpr (sizeof(T) == 4) __stcs(reinterpret_cast<float4 *>(D + 16 * j + r0), make_float4(y[0], y[1], y[2], y[3]));
            else { __stcs(reinterpret_cast<double2 *>(D + 16 * j + r0), make_double2(y[0], y[1])); __stcs(reinterpret_cast<double2 *>(D + 16 * j + r0 + 2), make_double2(y[2], y[3])); }
        }
    }
}
// synthesis: leaves of node j from S[16 j ..]; the rebuilt 16-sample node is band (j & 1) of its parent j >> 1 in the
// band-split layout of the next (32-sample) level, or -- a 16-sample subtree -- the output itself
template <typename T>
__device__ __forceinline__ void leaf16_syn(const T *__restrict__ S, T *__restrict__ out, int m, const Leaf16<T> &M) {
    const int mh = m >> 1;
    const int j = threadIdx.x;                   // one node per thread (see leaf16_ana)
    if (j < (m >> 4)) {
        T y[16];
#pragma unroll
        for (int r0 = 0; r0 < 16; r0 += 4) {
            if constexpr (sizeof(T) == 4) { const float4 v = __ldcs(reinterpret_cast<const float4 *>(S + 16 * j + r0)); y[r0] = v.x; y[r0 + 1] = v.y; y[r0 + 2] = v.z; y[r0 + 3] = v.w; }
            else { const double2 v0 = __ldcs(reinterpret_cast<const double2 *>(S + 16 * j + r0)), v1 = __ldcs(reinterpret_cast<const double2 *>(S + 16 * j + r0 + 2)); y[r0] = v0.x; y[r0 + 1] = v0.y; y[r0 + 2] = v1.x; y[r0 + 3] = v1.y; }
        }
        T *o = (m == 16) ? out : out + (j & 1) * mh + (j >> 1) * 16;
#pragma unroll
        for (int c0 = 0; c0 < 16; c0 += 4) {
            T x[4];
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) {
                T acc = M.m[c0 + cc][0] * y[0];
#pragma unroll
                for (int r = 1; r < 16; ++r) acc = fma(M.m[c0 + cc][r], y[r], acc);
                x[cc] = acc;
            }
            st4(o + c0, x[0], x[1], x[2], x[3]);
        }
    }
}

// Shared-memory layouts (both kernels ping-pong between two m-sample buffers):
//   analysis input of a level  : every sub-node SPLIT into its polyphase components, all even parts first --
//        node j (length ml, nh = ml/2): E_j at [j nh, (j+1) nh), O_j at m/2 + [j nh, (j+1) nh).  A warp's 16-byte loads are
//        then contiguous across sub-node boundaries (per-node [E|O] storage made every small-node level 2-way conflicted:
//        44 M bank conflicts per launch in profiles/r01c_wpt_sub_f32);
//   synthesis input of a level : all approximation bands first -- a_j at [j nh, ..), d_j at m/2 + [j nh, ..);
//   the last analysis level writes, and the first synthesis level reads, the natural packet order [a_j | d_j] per node.
template <typename T, int F, bool STRICT>
__global__ void __launch_bounds__(256, 5)
k_wpt_sub_ana(const T *__restrict__ S, T *__restrict__ D, int64_t n, int m, int levels, int64_t nodes,
              const __grid_constant__ Taps<T, F> c, int leaf16, const __grid_constant__ Leaf16<T> M16) {
    using fp = FP<STRICT>;
    constexpr int Q = F / 2;
    constexpr int CL = (Q - 1 + 3) / 4;          // chunks of left / right reach
    constexpr int NCH = 2 * CL + 1;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *in = reinterpret_cast<T *>(smem_raw);
    T *out = in + m;
    const int mh = m >> 1;
    const int64_t q = blockIdx.x % nodes, b = blockIdx.x / nodes;
    const int64_t base = b * n + q * (int64_t)m;
    for (int p = threadIdx.x; p < mh; p += blockDim.x) {      // split load of the root node
        in[p] = S[base + 2 * p];
        in[mh + p] = S[base + 2 * p + 1];
    }
    __syncthreads();
    const int lv_end = leaf16 ? levels - 4 : levels;          // leaf16: the 16-sample nodes finish in one matrix stage
    // (Tried and measured slower, r02: warp barriers between levels whose nodes a warp owns outright -- 0.58 -> 0.69 ms per 1024
    // signals; the block barrier per level stays.)
    for (int l = 0; l < lv_end; ++l) {
        const int ml = m >> l, nh = ml >> 1;
        const bool last = (l == levels - 1);
        const int hh = nh >> 1;                               // half length of a child (the next level's nh)
        // child cidx (2j: approximation, 2j+1: detail) of the next level: E' at cidx*hh, O' at mh + cidx*hh
        if (nh >= 4 && (nh & 3) == 0) {
            const int cpn = nh >> 2;                          // 16-byte chunks per component of a sub-node
            if ((cpn & (cpn - 1)) == 0) wpt_sub_ana_chunks<T, F, STRICT, true>(in, out, m, nh, last, c);
            else wpt_sub_ana_chunks<T, F, STRICT, false>(in, out, m, nh, last, c);
        } else if (ml == 4) {
            for (int j = threadIdx.x; j < (m >> 2); j += blockDim.x) {
                const T xe[2] = {in[2 * j], in[2 * j + 1]}, xo[2] = {in[mh + 2 * j], in[mh + 2 * j + 1]};
                T a[2], d[2];
                tiny_ana<T, F, STRICT, 4>(xe, xo, a, d, c);
                if (!last) { st2(out + 2 * j, a[0], d[0]); st2(out + mh + 2 * j, a[1], d[1]); }   // children 2j, 2j+1 of one sample pair each
                else st4(out + 4 * j, a[0], a[1], d[0], d[1]);
            }
        } else if (ml == 2) {                                 // (always the last level)
            for (int j = threadIdx.x; j < mh; j += blockDim.x) {
                const T xe[1] = {in[j]}, xo[1] = {in[mh + j]};
                T a[1], d[1];
                tiny_ana<T, F, STRICT, 2>(xe, xo, a, d, c);
                st2(out + 2 * j, a[0], d[0]);
            }
        } else {
            // any other sub-node length: one pair per thread, modular walk through the split components
            for (int idx = threadIdx.x; idx < mh; idx += blockDim.x) {
                const int j = idx / nh, k = idx - j * nh;
                const T *E = in + j * nh, *O = E + mh;
                int ia = 2 * k;
                T a = fp::mul(c.h[0], E[ia >> 1]);
#pragma unroll
                for (int t = 1; t < F; ++t) {
                    if (++ia == ml) ia = 0;
                    a = fp::mac(a, c.h[t], (ia & 1) ? O[ia >> 1] : E[ia >> 1]);
                }
                int id = (2 * k + 2 - F) % ml;
                if (id < 0) id += ml;
                T d = fp::mul(c.g[F - 1], (id & 1) ? O[id >> 1] : E[id >> 1]);
#pragma unroll
                for (int t = 1; t < F; ++t) {
                    if (++id == ml) id = 0;
                    d = fp::mac(d, c.g[F - 1 - t], (id & 1) ? O[id >> 1] : E[id >> 1]);
                }
                if (!last) {            // (nh is even here: another level follows)
                    out[(k & 1) * mh + (2 * j) * hh + (k >> 1)] = a;
                    out[(k & 1) * mh + (2 * j + 1) * hh + (k >> 1)] = d;
                } else {
                    out[j * ml + k] = a;
                    out[j * ml + nh + k] = d;
                }
            }
        }
        __syncthreads();
        T *t = in; in = out; out = t;
    }
    if (leaf16) { leaf16_ana<T>(in, D + base, m, M16); return; }
    for (int i = threadIdx.x; i < m; i += blockDim.x) D[base + i] = in[i];
}

template <typename T, int F, bool STRICT>
__global__ void __launch_bounds__(256, 5)
k_wpt_sub_syn(const T *__restrict__ S, T *__restrict__ D, int64_t n, int m, int levels, int64_t nodes,
              const __grid_constant__ Taps<T, F> c, int leaf16, const __grid_constant__ Leaf16<T> M16) {
    using fp = FP<STRICT>;
    constexpr int Q = F / 2;
    constexpr int CL = (Q - 1 + 3) / 4;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *in = reinterpret_cast<T *>(smem_raw);
    T *out = in + m;
    const int mh = m >> 1;
    const int64_t q = blockIdx.x % nodes, b = blockIdx.x / nodes;
    const int64_t base = b * n + q * (int64_t)m;
    if (leaf16) {
        leaf16_syn<T>(S + base, in, m, M16);      // leaves -> 16-sample nodes, already in the band-split layout of the 32-sample level
    } else {   // natural packet order -> band-split: node j of the deepest level: a_j to [j nh), d_j to mh + [j nh)
        const int ml = m >> (levels - 1), nh = ml >> 1;
        for (int i = threadIdx.x; i < m; i += blockDim.x) {
            const int j = i / ml, r = i - j * ml;
            in[(r < nh ? 0 : mh - nh) + j * nh + r] = S[base + i];
        }
    }
    __syncthreads();
    const int lv_first = leaf16 ? levels - 5 : levels - 1;
    for (int l = lv_first; l >= 0; --l) {
        const int ml = m >> l, nh = ml >> 1;
        // node j's output (ml samples) is band (j & 1) of its parent j >> 1 for the next, shallower level
        if (nh >= 4 && (nh & 3) == 0) {
            const int cpn = nh >> 2;
            if ((cpn & (cpn - 1)) == 0) wpt_sub_syn_chunks<T, F, STRICT, true>(in, out, m, nh, c);
            else wpt_sub_syn_chunks<T, F, STRICT, false>(in, out, m, nh, c);
        } else if (ml == 4) {
            for (int j = threadIdx.x; j < (m >> 2); j += blockDim.x) {
                const T a[2] = {in[2 * j], in[2 * j + 1]}, d[2] = {in[mh + 2 * j], in[mh + 2 * j + 1]};
                T x[4];
                tiny_syn<T, F, STRICT, 2>(a, d, x, c);
                st4(out + (j & 1) * mh + (j >> 1) * 4, x[0], x[1], x[2], x[3]);
            }
        } else if (ml == 2) {
            for (int j = threadIdx.x; j < mh; j += blockDim.x) {
                const T a[1] = {in[j]}, d[1] = {in[mh + j]};
                T x[2];
                tiny_syn<T, F, STRICT, 1>(a, d, x, c);
                st2(out + (j & 1) * mh + (j >> 1) * 2, x[0], x[1]);
            }
        } else {
            for (int idx = threadIdx.x; idx < mh; idx += blockDim.x) {
                const int j = idx / nh, u = idx - j * nh;
                const T *a = in + j * nh, *d = a + mh;
                int ia = (u - (Q - 1)) % nh;
                if (ia < 0) ia += nh;
                T rae = fp::mul(c.h[2 * (Q - 1)], a[ia]);
                T rao = fp::mul(c.h[2 * (Q - 1) + 1], a[ia]);
#pragma unroll
                for (int t = Q - 2; t >= 0; --t) {
                    if (++ia == nh) ia = 0;
                    rae = fp::mac(rae, c.h[2 * t], a[ia]);
                    rao = fp::mac(rao, c.h[2 * t + 1], a[ia]);
                }
                int id = u;
                T rde, rdo;      // fast mode: the detail terms continue the approximation's chain (no second FMUL, no final FADD)
                if constexpr (STRICT) { rde = fp::mul(c.g[1], d[id]); rdo = fp::mul(c.g[0], d[id]); }
                else { rde = fp::mac(rae, c.g[1], d[id]); rdo = fp::mac(rao, c.g[0], d[id]); }
#pragma unroll
                for (int t = 1; t < Q; ++t) {
                    if (++id == nh) id = 0;
                    rde = fp::mac(rde, c.g[2 * t + 1], d[id]);
                    rdo = fp::mac(rdo, c.g[2 * t], d[id]);
                }
                T *o = out + (j & 1) * mh + (j >> 1) * ml;
                o[2 * u] = (STRICT ? fp::add(rae, rde) : rde);
                o[2 * u + 1] = (STRICT ? fp::add(rao, rdo) : rdo);
            }
        }
        __syncthreads();
        T *t = in; in = out; out = t;
    }
    for (int i = threadIdx.x; i < m; i += blockDim.x) D[base + i] = in[i];
}

// the 16 x 16 maps of leaf16_ana / leaf16_syn: four periodic packet levels of a 16-sample node (natural order), composed in
// double precision from the taps as the kernels hold them (rounded to T)
template <typename T>
static void leaf16_matrices(const FilterCoefs<T> &fc, Leaf16<T> &A, Leaf16<T> &Sy) {
    const int F = fc.F, Q = F / 2;
    auto ana = [&](const double *x, int ml, double *y) {          // one level: y = [a | d]
        const int nh = ml / 2;
        for (int k = 0; k < nh; ++k) {
            double a = 0, d = 0;
            for (int t = 0; t < F; ++t) {
                a += (double)fc.h[t] * x[(2 * k + t) % ml];
                d += (double)fc.g[F - 1 - t] * x[(((2 * k + 2 - F + t) % ml) + ml) % ml];
            }
            y[k] = a; y[nh + k] = d;
        }
    };
    auto syn = [&](const double *a, const double *d, int nh, double *x) {
        for (int u = 0; u < nh; ++u) {
            double e = 0, o = 0;
            for (int t = 0; t < Q; ++t) {
                const double av = a[(((u - t) % nh) + nh) % nh], dv = d[(u + t) % nh];
                e += (double)fc.h[2 * t] * av + (double)fc.g[2 * t + 1] * dv;
                o += (double)fc.h[2 * t + 1] * av + (double)fc.g[2 * t] * dv;
            }
            x[2 * u] = e; x[2 * u + 1] = o;
        }
    };
    for (int c = 0; c < 16; ++c) {
        double cur[16] = {0}, nxt[16];
        cur[c] = 1.0;
        for (int ml = 16; ml >= 2; ml >>= 1) {                    // analysis of the unit vector e_c: column c of A
            for (int j = 0; j < 16 / ml; ++j) ana(cur + j * ml, ml, nxt + j * ml);
            for (int i = 0; i < 16; ++i) cur[i] = nxt[i];
        }
        for (int r = 0; r < 16; ++r) A.m[r][c] = (T)cur[r];
        double y[16] = {0};
        y[c] = 1.0;
        for (int ml = 2; ml <= 16; ml <<= 1) {                    // synthesis of the unit leaf vector e_c: column c of Sy
            for (int j = 0; j < 16 / ml; ++j) syn(y + j * ml, y + j * ml + ml / 2, ml / 2, nxt + j * ml);
            for (int i = 0; i < 16; ++i) y[i] = nxt[i];
        }
        for (int r = 0; r < 16; ++r) Sy.m[r][c] = (T)y[r];
    }
}

template <typename T, int F, bool STRICT>
static int wpt_sub_F(const T *S, T *D, int64_t n, int m, int levels, int64_t nodes, int64_t B, const FilterCoefs<T> &fc,
                     bool fw, cudaStream_t st) {
    Taps<T, F> taps;
    for (int k = 0; k < F; ++k) { taps.h[k] = fc.h[k]; taps.g[k] = fc.g[k]; }
    const int64_t nblk = nodes * B;
    if (nblk > 0x7fffffffLL) return 0;
    const size_t smem = (size_t)2 * m * sizeof(T);
    // 256 threads: two thread-iterations per level of a 4096-sample node, six CTAs per SM (r02 A/B: 8192-sample nodes with 512
    // threads 2.26 ms, 16384 with 1024 threads 2.07 ms against 2.08 ms per 1024 signals -- no gain, so the CTA stays small)
    int nthr = env_int_fp("WB200_WPT_SUB_NT", 256);
    if (nthr < 32 || nthr > 256) nthr = 256;
    nthr &= ~31;
    // fast mode, full-depth tree, 16-byte granular rows: the last four levels as one 16 x 16 map per node
    const bool leaf16 = !STRICT && levels >= 4 && ((int64_t)m >> levels) == 1 && (m >> 4) <= nthr && env_int_fp("WB200_WPT_LEAF16", 1) != 0 &&
                        ((reinterpret_cast<uintptr_t>(S) | reinterpret_cast<uintptr_t>(D)) & 15) == 0 && ((n * (int64_t)sizeof(T)) % 16) == 0;
    Leaf16<T> MA, MS;
    if (leaf16) leaf16_matrices<T>(fc, MA, MS); else { std::memset(&MA, 0, sizeof(MA)); std::memset(&MS, 0, sizeof(MS)); }
    if (fw) {
        auto kern = k_wpt_sub_ana<T, F, STRICT>;
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { (void)cudaGetLastError(); return 0; }
        LaunchScope scope("wpt_subtree_analysis", st);
        kern<<<(unsigned)nblk, nthr, smem, st>>>(S, D, n, m, levels, nodes, taps, leaf16 ? 1 : 0, MA);
    } else {
        auto kern = k_wpt_sub_syn<T, F, STRICT>;
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { (void)cudaGetLastError(); return 0; }
        LaunchScope scope("wpt_subtree_synthesis", st);
        kern<<<(unsigned)nblk, nthr, smem, st>>>(S, D, n, m, levels, nodes, taps, leaf16 ? 1 : 0, MS);
    }
    return check_launch("wpt_subtree") ? 1 : -1;
}

// nodes of m samples each (m <= WPT_SUB_MAX, m % 2^levels == 0), `levels` full levels; 1 handled / 0 not covered / -1 error
template <typename T>
int fast_wpt_subtree(const T *S, T *D, int64_t n, int64_t m, int levels, int64_t nodes, int64_t B,
                     const FilterCoefs<T> &fc, bool strict, bool fw, cudaStream_t st) {
    if (!env_fast_enabled() || m > wpt_subtree_max_samples((int)sizeof(T)) || m < 2 || levels < 1 || (m % ((int64_t)1 << levels)) != 0) return 0;
    switch (fc.F) {
#define WB_CASE(FF) case FF: return strict ? wpt_sub_F<T, FF, true>(S, D, n, (int)m, levels, nodes, B, fc, fw, st) : wpt_sub_F<T, FF, false>(S, D, n, (int)m, levels, nodes, B, fc, fw, st);
        WB_CASE(2) WB_CASE(4) WB_CASE(6) WB_CASE(8) WB_CASE(10) WB_CASE(12) WB_CASE(14) WB_CASE(16) WB_CASE(18) WB_CASE(20)
#undef WB_CASE
    default: return 0;
    }
}

// ===================================================================================================
// host dispatch
// ===================================================================================================
template <typename T> static bool aligned16_view(const View<T> &v, const Extent &e) {
    if ((reinterpret_cast<uintptr_t>(v.p) & 15) != 0) return false;
    for (int q = 0; q < 4; ++q)
        if (e.n[q] > 1 && ((v.s[q] * (int64_t)sizeof(T)) % 16) != 0) return false;
    return true;
}
static int pick_line_tile(int64_t len, int esize, bool whole_line_ok) {
    int64_t tile = esize == 4 ? 8192 : 4096;
    const int64_t p2 = len & (-len);
    while (tile > p2) tile >>= 1;
    // analysis may take the whole line as one tile (the wrap piece is the start of the same line); synthesis stages
    // ranges a little longer than half a tile and needs them to fit the half-length bands: at least two tiles per line
    while (tile > (whole_line_ok ? len : len / 2)) tile >>= 1;
    return (int)tile;
}

template <typename T, int F, bool STRICT>
static int fast_ana_F(const View<const T> &src, const View<T> &dlo, const View<T> &dhi, const Extent &e,
                      const FilterCoefs<T> &fc, cudaStream_t st) {
    Taps<T, F> taps;
    for (int m = 0; m < F; ++m) { taps.h[m] = fc.h[m]; taps.g[m] = fc.g[m]; }
    const int64_t NL = e.n[0] * e.n[1] * e.n[2] * e.n[3];
    if (NL <= 0 || e.len < 2) return 1;
    if (src.ls == 1 && dlo.ls == 1 && dhi.ls == 1) {
        // ---- contiguous lines ----
        constexpr int PA = AnaPairs<T>::value;
        using G = FGeom<F, PA>;
        const int tile = pick_line_tile(e.len, sizeof(T), true);
        const int vec = 16 / (int)sizeof(T);
        const int h0 = (F - 2 + G::WO + vec - 1) / vec * vec;
        if (tile < 64 || h0 > tile || (e.len * (int64_t)sizeof(T)) % 16 != 0 || e.len > ((int64_t)1 << 30)) return 0;
        if (!aligned16_view(src, e) || !aligned16_view(View<const T>{dlo.p, dlo.ls, {dlo.s[0], dlo.s[1], dlo.s[2], dlo.s[3]}}, e) ||
            !aligned16_view(View<const T>{dhi.p, dhi.ls, {dhi.s[0], dhi.s[1], dhi.s[2], dhi.s[3]}}, e)) return 0;
        const int64_t ntiles = e.len / tile;
        const int64_t nblk = ntiles * NL;
        if (nblk > 0x7fffffffLL) return 0;
        const size_t smem = 128 + (size_t)(tile + h0 + 8) * sizeof(T);
        auto kern = k_line_ana<T, F, STRICT>;
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { (void)cudaGetLastError(); return 0; }
        {
            LaunchScope scope("line_filter_analysis", st);
            const int nthr = tile >= 1024 ? 256 : (tile >= 512 ? 128 : 64);
            kern<<<(unsigned)nblk, nthr, smem, st>>>(src, dlo, dhi, e, tile, h0, (int)ntiles, taps);
        }
        return check_launch("line_filter_analysis") ? 1 : -1;
    }
    if (src.s[0] == 1 && dlo.s[0] == 1 && dhi.s[0] == 1 && e.n[0] >= 16) {
        // ---- strided lines, contiguous coordinate 0 ----
        constexpr int RK = WalkCfg<F>::RK;
        const int64_t nh = e.len / 2;
        const int64_t nseg = (nh + RK - 1) / RK;
        const int64_t NY = e.n[1] * e.n[2] * e.n[3];
        int64_t gy = nseg * NY, gz = 1;
        if (nseg > 65535) return 0;
        if (gy > 65535) { const int64_t per = 65535 / nseg; gz = (NY + per - 1) / per; gy = per * nseg; if (gz > 65535) return 0; }
        dim3 grid((unsigned)((e.n[0] + 127) / 128), (unsigned)gy, (unsigned)gz);
        {
            LaunchScope scope("walk_filter_analysis", st);
            k_walk_ana<T, F, STRICT><<<grid, 128, 0, st>>>(src, dlo, dhi, e, (int)nseg, taps);
        }
        return check_launch("walk_filter_analysis") ? 1 : -1;
    }
    return 0;
}

template <typename T, int F, bool STRICT>
static int fast_syn_F(const View<const T> &slo, const View<const T> &shi, const View<const T> &salt, const int64_t thr[4],
                      bool has_alt, const View<T> &dst, const Extent &e, const FilterCoefs<T> &fc, cudaStream_t st) {
    Taps<T, F> taps;
    for (int m = 0; m < F; ++m) { taps.h[m] = fc.h[m]; taps.g[m] = fc.g[m]; }
    const int64_t NL = e.n[0] * e.n[1] * e.n[2] * e.n[3];
    if (NL <= 0 || e.len < 2) return 1;
    if (slo.ls == 1 && shi.ls == 1 && dst.ls == 1 && (!has_alt || salt.ls == 1)) {
        using G = FGeom<F>;
        const int tile = pick_line_tile(e.len, sizeof(T), false);
        if (tile < 64 || (e.len * (int64_t)sizeof(T)) % 32 != 0 || e.len > ((int64_t)1 << 30)) return 0;
        if (!aligned16_view(slo, e) || !aligned16_view(shi, e) || (has_alt && !aligned16_view(salt, e)) ||
            !aligned16_view(View<const T>{dst.p, dst.ls, {dst.s[0], dst.s[1], dst.s[2], dst.s[3]}}, e)) return 0;
        SynPlan pl;      // ranges cut for the four-pair form of syn_level (fused1d_dev.cuh)
        memset(&pl, 0, sizeof(pl));
        pl.K = 1; pl.tile = tile;
        pl.rlo[0] = 0; pl.rhi[0] = tile;
        pl.rlo[1] = -G::Q4; pl.rhi[1] = tile / 2;
        pl.dlo[1] = 0; pl.dhi[1] = tile / 2 + G::Q4;
        pl.doff[1] = 0;
        pl.aoff = pl.dhi[1] - pl.dlo[1];
        const int64_t nh = e.len / 2;
        if ((int64_t)(pl.dhi[1] - pl.dlo[1]) > nh || (int64_t)(pl.rhi[1] - pl.rlo[1]) > nh) return 0;
        const size_t smem = 128 + (size_t)(pl.aoff + (pl.rhi[1] - pl.rlo[1]) + 8) * sizeof(T);
        const int64_t ntiles = e.len / tile;
        const int64_t nblk = ntiles * NL;
        if (nblk > 0x7fffffffLL) return 0;
        auto kern = k_line_syn<T, F, STRICT>;
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { (void)cudaGetLastError(); return 0; }
        {
            LaunchScope scope("line_filter_synthesis", st);
            const int nthr = tile >= 1024 ? 256 : (tile >= 512 ? 128 : 64);
            kern<<<(unsigned)nblk, nthr, smem, st>>>(slo, shi, salt, thr[0], thr[1], thr[2], thr[3], has_alt ? 1 : 0, dst, e, (int)ntiles, taps, pl);
        }
        return check_launch("line_filter_synthesis") ? 1 : -1;
    }
    if (slo.s[0] == 1 && shi.s[0] == 1 && dst.s[0] == 1 && (!has_alt || salt.s[0] == 1) && e.n[0] >= 16) {
        constexpr int RK = WalkCfg<F>::RK;
        const int64_t nh = e.len / 2;
        const int64_t nseg = (nh + RK - 1) / RK;
        const int64_t NY = e.n[1] * e.n[2] * e.n[3];
        int64_t gy = nseg * NY, gz = 1;
        if (nseg > 65535) return 0;
        if (gy > 65535) { const int64_t per = 65535 / nseg; gz = (NY + per - 1) / per; gy = per * nseg; if (gz > 65535) return 0; }
        dim3 grid((unsigned)((e.n[0] + 127) / 128), (unsigned)gy, (unsigned)gz);
        {
            LaunchScope scope("walk_filter_synthesis", st);
            k_walk_syn<T, F, STRICT><<<grid, 128, 0, st>>>(slo, shi, salt, thr[0], thr[1], thr[2], thr[3], has_alt ? 1 : 0, dst, e, (int)nseg, taps);
        }
        return check_launch("walk_filter_synthesis") ? 1 : -1;
    }
    return 0;
}


template <typename T>
int fast_filter_analysis(const View<const T> &src, const View<T> &dlo, const View<T> &dhi, const Extent &e,
                         const FilterCoefs<T> &fc, bool strict, cudaStream_t st) {
    if (!env_fast_enabled()) return 0;
    switch (fc.F) {
#define WB_CASE(FF) case FF: return strict ? fast_ana_F<T, FF, true>(src, dlo, dhi, e, fc, st) : fast_ana_F<T, FF, false>(src, dlo, dhi, e, fc, st);
        WB_CASE(2) WB_CASE(4) WB_CASE(6) WB_CASE(8) WB_CASE(10) WB_CASE(12) WB_CASE(14) WB_CASE(16) WB_CASE(18) WB_CASE(20)
#undef WB_CASE
    default: return 0;
    }
}
template <typename T>
int fast_filter_synthesis(const View<const T> &slo, const View<const T> &shi, const View<const T> &salt,
                          const int64_t thr[4], bool has_alt, const View<T> &dst, const Extent &e,
                          const FilterCoefs<T> &fc, bool strict, cudaStream_t st) {
    if (!env_fast_enabled()) return 0;
    switch (fc.F) {
#define WB_CASE(FF) case FF: return strict ? fast_syn_F<T, FF, true>(slo, shi, salt, thr, has_alt, dst, e, fc, st) : fast_syn_F<T, FF, false>(slo, shi, salt, thr, has_alt, dst, e, fc, st);
        WB_CASE(2) WB_CASE(4) WB_CASE(6) WB_CASE(8) WB_CASE(10) WB_CASE(12) WB_CASE(14) WB_CASE(16) WB_CASE(18) WB_CASE(20)
#undef WB_CASE
    default: return 0;
    }
}

#define WB_INST(T)                                                                                                    \
    template int fast_wpt_subtree<T>(const T *, T *, int64_t, int64_t, int, int64_t, int64_t, const FilterCoefs<T> &,  \
                                     bool, bool, cudaStream_t);                                                       \
    template int fast_filter_analysis<T>(const View<const T> &, const View<T> &, const View<T> &, const Extent &,      \
                                         const FilterCoefs<T> &, bool, cudaStream_t);                                 \
    template int fast_filter_synthesis<T>(const View<const T> &, const View<const T> &, const View<const T> &,        \
                                          const int64_t[4], bool, const View<T> &, const Extent &,                    \
                                          const FilterCoefs<T> &, bool, cudaStream_t);
WB_INST(float)
WB_INST(double)
#undef WB_INST

} // namespace wb
