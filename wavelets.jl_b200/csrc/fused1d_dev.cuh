// fused1d_dev.cuh -- device building blocks shared by the fused 1-D kernels (fused1d.cu) and the fast
// single-level passes (fastpass.cu): mbarrier / TMA bulk-copy wrappers, vector shared-memory helpers, the
// per-filter geometry, and the one-level analysis / synthesis routines that work out of shared memory.
#pragma once
#include "fused.cuh"

namespace wb {

// ---------------------------------------------------------------------------------------------------
// PTX helpers: mbarrier + TMA bulk copy (global -> shared)
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// 1-D bulk copy: 16-byte aligned src/dst, size a multiple of 16
__device__ __forceinline__ void tma_bulk_g2s(void *dst_smem, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// L2 prefetch of a contiguous global range (16-byte aligned, size a multiple of 16): no shared-memory destination, no barrier
__device__ __forceinline__ void tma_prefetch_l2(const void *src, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
// Copy `count` elements of the periodic line `line` (period n) starting at (possibly negative / overflowing)
// index `lo` into dst.  lo, count and n are multiples of the 16-byte vector, count <= n.  Returns bytes issued.
template <typename T>
__device__ __forceinline__ uint32_t tma_load_wrapped(T *dst, const T *line, int64_t lo, int count, int64_t n, uint64_t *bar) {
    if (lo < 0) lo += n;
    if (lo >= n) lo -= n;
    const int64_t first = (lo + count <= n) ? count : (n - lo);
    tma_bulk_g2s(dst, line + lo, (uint32_t)(first * sizeof(T)), bar);
    if (first < count) tma_bulk_g2s(dst + first, line, (uint32_t)((count - first) * sizeof(T)), bar);
    return (uint32_t)(count * sizeof(T));
}

// ---------------------------------------------------------------------------------------------------
// small vector helpers (float4 / double2 granularity = 16 bytes)
// ---------------------------------------------------------------------------------------------------
template <typename T> struct Vec;
template <> struct Vec<float> { using v16 = float4; using v8 = float2; static constexpr int N16 = 4; };
template <> struct Vec<double> { using v16 = double2; static constexpr int N16 = 2; };

template <int N> __device__ __forceinline__ void load_window(float (&w)[N], const float *p) {
    static_assert(N % 2 == 0, "window must be even");
    if constexpr (N % 4 == 0) {
#pragma unroll
        for (int i = 0; i < N / 4; ++i) {
            const float4 v = *reinterpret_cast<const float4 *>(p + 4 * i);
            w[4 * i] = v.x; w[4 * i + 1] = v.y; w[4 * i + 2] = v.z; w[4 * i + 3] = v.w;
        }
    } else {
#pragma unroll
        for (int i = 0; i < N / 2; ++i) {
            const float2 v = *reinterpret_cast<const float2 *>(p + 2 * i);
            w[2 * i] = v.x; w[2 * i + 1] = v.y;
        }
    }
}
template <int N> __device__ __forceinline__ void load_window(double (&w)[N], const double *p) {
    static_assert(N % 2 == 0, "window must be even");
#pragma unroll
    for (int i = 0; i < N / 2; ++i) {
        const double2 v = *reinterpret_cast<const double2 *>(p + 2 * i);
        w[2 * i] = v.x; w[2 * i + 1] = v.y;
    }
}
// 8-byte-aligned (float) / 16-byte-aligned (double) pair loads
template <int N> __device__ __forceinline__ void load_pairs(float (&w)[N], const float *p) {
#pragma unroll
    for (int i = 0; i < N / 2; ++i) {
        const float2 v = *reinterpret_cast<const float2 *>(p + 2 * i);
        w[2 * i] = v.x; w[2 * i + 1] = v.y;
    }
}
template <int N> __device__ __forceinline__ void load_pairs(double (&w)[N], const double *p) {
#pragma unroll
    for (int i = 0; i < N / 2; ++i) {
        const double2 v = *reinterpret_cast<const double2 *>(p + 2 * i);
        w[2 * i] = v.x; w[2 * i + 1] = v.y;
    }
}
__device__ __forceinline__ void store2(float *p, float a, float b) { *reinterpret_cast<float2 *>(p) = make_float2(a, b); }
__device__ __forceinline__ void store2(double *p, double a, double b) { *reinterpret_cast<double2 *>(p) = make_double2(a, b); }
__device__ __forceinline__ void store4(float *p, float a, float b, float c, float d) {
    *reinterpret_cast<float4 *>(p) = make_float4(a, b, c, d);
}
__device__ __forceinline__ void store4(double *p, double a, double b, double c, double d) {
    *reinterpret_cast<double2 *>(p) = make_double2(a, b);
    *reinterpret_cast<double2 *>(p + 2) = make_double2(c, d);
}
// streaming (evict-first) global stores: outputs are written once and not re-read by this kernel
__device__ __forceinline__ void gstore2(float *p, float a, float b) { __stcs(reinterpret_cast<float2 *>(p), make_float2(a, b)); }
__device__ __forceinline__ void gstore2(double *p, double a, double b) { __stcs(reinterpret_cast<double2 *>(p), make_double2(a, b)); }
__device__ __forceinline__ void gstore4(float *p, float a, float b, float c, float d) {
    __stcs(reinterpret_cast<float4 *>(p), make_float4(a, b, c, d));
}
__device__ __forceinline__ void gstore4(double *p, double a, double b, double c, double d) {
    __stcs(reinterpret_cast<double2 *>(p), make_double2(a, b));
    __stcs(reinterpret_cast<double2 *>(p + 2), make_double2(c, d));
}

// ---------------------------------------------------------------------------------------------------
// compile-time geometry of one even-length filter
// ---------------------------------------------------------------------------------------------------
// PA = output pairs per thread-iteration of the analysis level: 2 for 4-byte samples (16-byte window loads at
// 16-byte thread stride), 1 for 8-byte samples (two pairs would put the 16-byte loads at a 32-byte thread stride:
// a 2-way shared-memory bank conflict, measured at 129 M conflicts per launch in profiles/r01_fused1d_f64.md).
template <typename T> struct AnaPairs { static constexpr int value = sizeof(T) == 4 ? 2 : 1; };

template <int F, int PA = 2> struct FGeom {
    static_assert(F % 2 == 0 && F >= 2, "fused kernels take even filter lengths");
    static constexpr int Q = F / 2;
    // analysis: the detail computed from the window starting at 2j + WO is d[j + DS]; DS a multiple of PA keeps the
    // PA-wide detail stores aligned
    static constexpr int DS = PA == 2 ? (((Q - 1) + 1) & ~1) : (Q - 1);
    static constexpr int WO = 2 * DS - (F - 2);
    static constexpr int WIN = F + 2 * (PA - 1) + WO;  // inputs per thread-iteration (PA = 2: a multiple of 4)
    // synthesis: two output pairs (u, u+1) read a[u-QA .. u+2) and d[u .. u+QD)
    static constexpr int QA = ((Q - 1) + 1) & ~1;
    static constexpr int QD = ((Q + 1) + 1) & ~1;
    // 4-byte samples: FOUR output pairs (u .. u+3) from 16-byte windows a[u-Q4 .. u+4) and d[u .. u+4+Q4)
    static constexpr int Q4 = ((Q - 1) + 3) & ~3;
};

template <typename T, int F> struct Taps {
    T h[F];
    T g[F];
};

constexpr int MAXK = 8; // fused levels per tile kernel

// ===================================================================================================
// FORWARD
// ===================================================================================================
struct AnaPlan {
    int K;              // levels fused in stage A
    int tile;           // input samples per CTA
    int h0;             // staged halo samples (padded to a 16-byte multiple)
    int NA[MAXK + 1];   // approximation samples computed at level l (index 0: staged inputs)
    int ND[MAXK + 1];   // detail samples owned at level l
};

// one analysis level out of shared memory: `in` holds NAprev valid samples of a_{l-1}
//   a[j]      = sum_m h[m]     in[2j + m]                 (increasing m)
//   d[j + DS] = sum_p g[F-1-p] in[2j + WO + p]            (increasing input index)
// Approximations are produced for p < NA (halo included) and handed to store_a; details for p < ND go to d0 + p, or to
// d1 + p from pair `wrap_at` on (the band's periodic wrap inside the last tile of a line; wrap_at >= ND: no wrap).
// Three branch-free loops over one running index (details before the wrap, after it, approximation-only halo) keep the
// inner loop at loads + FMAs + two stores: the former single loop carried a predicated detail block and re-derived the wrap
// address per iteration (27 of its 59 instructions per iteration were neither loads, stores nor arithmetic).
template <typename T, int F, bool STRICT, typename SA>
__device__ __forceinline__ void ana_level(const T *__restrict__ in, int NA, int ND, const Taps<T, F> &c, SA store_a,
                                          T *__restrict__ d0, T *__restrict__ d1, int wrap_at) {
    using fp = FP<STRICT>;
    constexpr int PA = AnaPairs<T>::value;
    using G = FGeom<F, PA>;
    const int step = PA * blockDim.x;
    int p = PA * threadIdx.x;
    const int e1 = wrap_at < ND ? wrap_at : ND;
    const int nseg = wrap_at < ND ? 2 : 1;        // only the last tile of a line has a second (wrapped) detail segment
#pragma unroll 1
    for (int seg = 0; seg < nseg; ++seg) {
        T *__restrict__ dp = (seg ? d1 : d0) + p;      // running detail pointer: one 64-bit add per iteration
        const int e = seg ? ND : e1;
        for (; p < e; p += step, dp += step) {
            T w[G::WIN];
            load_window<G::WIN>(w, in + 2 * p);
            T a[PA], d[PA];
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
            store_a(p, a);
            if constexpr (PA == 2) gstore2(dp, d[0], d[1]); else __stcs(dp, d[0]);
        }
    }
    for (; p < NA; p += step) {
        T w[G::WIN];
        load_window<G::WIN>(w, in + 2 * p);
        T a[PA];
#pragma unroll
        for (int r = 0; r < PA; ++r) a[r] = fp::mul(c.h[0], w[2 * r]);
#pragma unroll
        for (int m = 1; m < F; ++m)
#pragma unroll
            for (int r = 0; r < PA; ++r) a[r] = fp::mac(a[r], c.h[m], w[2 * r + m]);
        store_a(p, a);
    }
}


// ===================================================================================================
// synthesis
// ===================================================================================================
struct SynPlan {
    int K;                 // levels fused in stage A'
    int tile;              // output samples per CTA
    int rlo[MAXK + 1];     // a_l needed on [s_l + rlo[l], s_l + rhi[l])   (multiples of 4; rlo[0] = 0, rhi[0] = tile)
    int rhi[MAXK + 1];
    int dlo[MAXK + 1];     // d_l staged on [s_l + dlo[l], s_l + dhi[l])
    int dhi[MAXK + 1];
    int doff[MAXK + 1];    // element offset of the d_l stage inside shared memory
    int aoff;              // element offset of the a_K stage
    int poff, qoff;        // ping-pong buffers for a_{K-1} .. a_1
};

// one synthesis level
//   x[2u]   = (sum_{i=u-Q+1..u} h[2(u-i)]   a[i]) + (sum_{i=u..u+Q-1} g[2(i-u)+1] d[i])
//   x[2u+1] = (sum_{i=u-Q+1..u} h[2(u-i)+1] a[i]) + (sum_{i=u..u+Q-1} g[2(i-u)]   d[i])
// `abuf[i]` = a_l[s_l + rlo_l + i], `dbuf[i]` = d_l[s_l + dlo_l + i]; outputs a_{l-1}[s_{l-1} + rlo_{l-1} + 2*ur ...]
// 4-byte samples with 16-byte-aligned windows (oa, od, npairs multiples of 4; Q4 staged samples in front of a[u_first] and
// behind d[u_last]): FOUR pairs per thread-iteration from 16-byte loads -- 4 LDS.128 per 64 FMA (db4) where the two-pair
// form below issues 12 LDS.64.  Otherwise two pairs per thread-iteration from 8-byte (Float64: 16-byte) loads.
// four output pairs ur .. ur+3 (4-byte samples): pa = abuf + oa - Q4, pd = dbuf + od, both 16-byte aligned, ur a multiple of 4
template <typename T, int F, bool STRICT, typename SO>
__device__ __forceinline__ void syn_quad(const T *__restrict__ pa, const T *__restrict__ pd, int ur, const Taps<T, F> &c, SO store_out) {
    using fp = FP<STRICT>;
    using G = FGeom<F>;
    constexpr int Q = G::Q, Q4 = G::Q4;
    T wa[Q4 + 4], wd[Q4 + 4];
    load_window<Q4 + 4>(wa, pa + ur);
    load_window<Q4 + 4>(wd, pd + ur);
    T o[8];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        // a[u' - k] = wa[Q4 + r - k], d[u' + k] = wd[r + k]
        T rae = fp::mul(c.h[2 * (Q - 1)], wa[Q4 + r - (Q - 1)]);
        T rao = fp::mul(c.h[2 * (Q - 1) + 1], wa[Q4 + r - (Q - 1)]);
#pragma unroll
        for (int k = Q - 2; k >= 0; --k) {
            rae = fp::mac(rae, c.h[2 * k], wa[Q4 + r - k]);
            rao = fp::mac(rao, c.h[2 * k + 1], wa[Q4 + r - k]);
        }
        T rde, rdo;      // fast mode: the detail terms continue the approximation's chain (no second FMUL, no final FADD)
        if constexpr (STRICT) { rde = fp::mul(c.g[1], wd[r]); rdo = fp::mul(c.g[0], wd[r]); }
        else { rde = fp::mac(rae, c.g[1], wd[r]); rdo = fp::mac(rao, c.g[0], wd[r]); }
#pragma unroll
        for (int k = 1; k < Q; ++k) {
            rde = fp::mac(rde, c.g[2 * k + 1], wd[r + k]);
            rdo = fp::mac(rdo, c.g[2 * k], wd[r + k]);
        }
        o[2 * r] = (STRICT ? fp::add(rae, rde) : rde);
        o[2 * r + 1] = (STRICT ? fp::add(rao, rdo) : rdo);
    }
    store_out(ur, o[0], o[1], o[2], o[3]);
    store_out(ur + 2, o[4], o[5], o[6], o[7]);
}
// two output pairs ur, ur+1: pa = abuf + oa - QA, pd = dbuf + od (8-byte aligned for 4-byte samples, 16 for 8-byte), ur even
template <typename T, int F, bool STRICT, typename SO>
__device__ __forceinline__ void syn_duo(const T *__restrict__ pa, const T *__restrict__ pd, int ur, const Taps<T, F> &c, SO store_out) {
    using fp = FP<STRICT>;
    using G = FGeom<F>;
    constexpr int Q = G::Q;
    T wa[G::QA + 2], wd[G::QD];
    load_pairs<G::QA + 2>(wa, pa + ur);
    load_pairs<G::QD>(wd, pd + ur);
    T o[4];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        // a[u' - k] = wa[QA + r - k], d[u' + k] = wd[r + k]
        T rae = fp::mul(c.h[2 * (Q - 1)], wa[G::QA + r - (Q - 1)]);
        T rao = fp::mul(c.h[2 * (Q - 1) + 1], wa[G::QA + r - (Q - 1)]);
#pragma unroll
        for (int k = Q - 2; k >= 0; --k) {
            rae = fp::mac(rae, c.h[2 * k], wa[G::QA + r - k]);
            rao = fp::mac(rao, c.h[2 * k + 1], wa[G::QA + r - k]);
        }
        T rde, rdo;      // fast mode: the detail terms continue the approximation's chain (no second FMUL, no final FADD)
        if constexpr (STRICT) { rde = fp::mul(c.g[1], wd[r]); rdo = fp::mul(c.g[0], wd[r]); }
        else { rde = fp::mac(rae, c.g[1], wd[r]); rdo = fp::mac(rao, c.g[0], wd[r]); }
#pragma unroll
        for (int k = 1; k < Q; ++k) {
            rde = fp::mac(rde, c.g[2 * k + 1], wd[r + k]);
            rdo = fp::mac(rdo, c.g[2 * k], wd[r + k]);
        }
        o[2 * r] = (STRICT ? fp::add(rae, rde) : rde);
        o[2 * r + 1] = (STRICT ? fp::add(rao, rdo) : rdo);
    }
    store_out(ur, o[0], o[1], o[2], o[3]);
}
// whether a level can run the four-pair form
template <typename T, int F> __device__ __forceinline__ bool syn_quad_ok(int oa, int od, int npairs) {
    return sizeof(T) == 4 && ((oa | od | npairs) & 3) == 0 && oa >= FGeom<F>::Q4;
}

template <typename T, int F, bool STRICT, typename SO>
__device__ __forceinline__ void syn_level(const T *__restrict__ abuf, const T *__restrict__ dbuf, int oa, int od, int npairs,
                                          const Taps<T, F> &c, SO store_out) {
    using G = FGeom<F>;
    if constexpr (sizeof(T) == 4) {
        if (syn_quad_ok<T, F>(oa, od, npairs)) {
            const T *pa = abuf + oa - G::Q4, *pd = dbuf + od;
            for (int ur = 4 * threadIdx.x; ur < npairs; ur += 4 * blockDim.x) syn_quad<T, F, STRICT>(pa, pd, ur, c, store_out);
            return;
        }
    }
    const T *pa = abuf + oa - G::QA, *pd = dbuf + od;
    for (int ur = 2 * threadIdx.x; ur < npairs; ur += 2 * blockDim.x) syn_duo<T, F, STRICT>(pa, pd, ur, c, store_out);
}


} // namespace wb
