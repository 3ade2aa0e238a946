// tile2d_shapes.cuh -- compile-time "shapes" the 2-D tile kernels specialise on: the lifting schemes (cdf 9/7, Haar, db2)
// and the orthogonal filter banks (even length F), each with its per-segment register transform.  A tile kernel
// (fused2d_tma.cuh) is written once against this interface:
//      Halo<S>::left()/right()        halo in polyphase pairs a segment needs on each side
//      CoefsOf<S, T>::type            the coefficient block passed as a __grid_constant__ kernel parameter
//      tile_transform<T, S, STRICT>   s[p], d[p] (pair p of a line segment)  ->  the level's outputs, in place
//      HasNorm<S>                     whether the level multiplies the two bands by (n1, n2) (lifting's normalize!)
#pragma once
#include "common.cuh"

namespace wb {

__device__ __forceinline__ int wrapi(int v, int n) { return v < 0 ? v + n : (v >= n ? v - n : v); }

// ---------------------------------------------------------------------------------------------------
// compile-time shapes of the lifting schemes the fused kernels specialise on (coefficients stay runtime)
// ---------------------------------------------------------------------------------------------------
#define WB_HD __host__ __device__ static constexpr
struct ShapeCdf97F { // U@0[2], P@1[2], U@0[2], P@1[2]      (WT.SCHEMES "cdf9/7", forward order)
    static constexpr int N = 4;
    WB_HD int pred(int i) { return (i & 1); }
    WB_HD int sh(int i) { return (i & 1); }
    WB_HD int nc(int) { return 2; }
};
struct ShapeCdf97I { // reversed: P@1, U@0, P@1, U@0
    static constexpr int N = 4;
    WB_HD int pred(int i) { return !(i & 1); }
    WB_HD int sh(int i) { return !(i & 1); }
    WB_HD int nc(int) { return 2; }
};
struct ShapeHaarF { // P@0[1], U@0[1]
    static constexpr int N = 2;
    WB_HD int pred(int i) { return i == 0; }
    WB_HD int sh(int) { return 0; }
    WB_HD int nc(int) { return 1; }
};
struct ShapeHaarI { // U@0, P@0
    static constexpr int N = 2;
    WB_HD int pred(int i) { return i == 1; }
    WB_HD int sh(int) { return 0; }
    WB_HD int nc(int) { return 1; }
};
struct ShapeDb2F { // P@0[1], U@1[2], P@-1[1]
    static constexpr int N = 3;
    WB_HD int pred(int i) { return i != 1; }
    WB_HD int sh(int i) { return i == 0 ? 0 : (i == 1 ? 1 : -1); }
    WB_HD int nc(int i) { return i == 1 ? 2 : 1; }
};
struct ShapeDb2I { // P@-1, U@1, P@0
    static constexpr int N = 3;
    WB_HD int pred(int i) { return i != 1; }
    WB_HD int sh(int i) { return i == 0 ? -1 : (i == 1 ? 1 : 0); }
    WB_HD int nc(int i) { return i == 1 ? 2 : 1; }
};
template <class S> struct Halo {
    WB_HD int left() { int h = 0; for (int i = 0; i < S::N; ++i) h += S::sh(i) > 0 ? S::sh(i) : 0; return h; }
    WB_HD int right() { int h = 0; for (int i = 0; i < S::N; ++i) h += (S::nc(i) - 1 - S::sh(i)) > 0 ? (S::nc(i) - 1 - S::sh(i)) : 0; return h; }
};
#undef WB_HD

template <class S, typename T> static bool shape_matches(const LiftScheme<T> &sc) {
    if (sc.nsteps != S::N) return false;
    for (int i = 0; i < S::N; ++i)
        if ((sc.is_predict[i] != 0) != (S::pred(i) != 0) || sc.shift[i] != S::sh(i) || sc.nc[i] != S::nc(i)) return false;
    return true;
}

// coefficients of one direction, trimmed to what the fused shapes need
template <typename T> struct LiftCoefs {
    T c[4][2];
    T n1, n2;
};

// ---------------------------------------------------------------------------------------------------
// all lifting steps of one line segment, in registers.  s[p], d[p] are the polyphase pair p of the segment;
// element 0 is global pair g0 (mod half).  Valid ranges shrink by each step's reach; the caller only consumes
// pairs [HL, NP-HR).
// ---------------------------------------------------------------------------------------------------
template <typename T, class S, bool STRICT, int NP>
__device__ __forceinline__ void lift_regs(T (&s)[NP], T (&d)[NP], const LiftCoefs<T> &lc, int g0, int half, bool edge) {
    using fp = FP<STRICT>;
    int lo_s = 0, hi_s = NP, lo_d = 0, hi_d = NP;
#pragma unroll
    for (int st = 0; st < S::N; ++st) {
        const int sh = S::sh(st), nc = S::nc(st);
        const bool pred = S::pred(st) != 0;
        const int left = sh > 0 ? sh : 0, right = (nc - 1 - sh) > 0 ? (nc - 1 - sh) : 0;
        int lo, hi;
        if (pred) { lo = lo_s > lo_d + left ? lo_s : lo_d + left; hi = hi_s < hi_d - right ? hi_s : hi_d - right; lo_s = lo; hi_s = hi; }
        else      { lo = lo_d > lo_s + left ? lo_d : lo_s + left; hi = hi_d < hi_s - right ? hi_d : hi_s - right; lo_d = lo; hi_d = hi; }
#pragma unroll
        for (int p = 0; p < NP; ++p) {
            if (p >= lo && p < hi) {
                T v = pred ? s[p] : d[p];
                // tap indices are compile-time constants after unrolling; the clamps only silence dead-code bounds
                const int i0 = (p - sh) < 0 ? 0 : ((p - sh) >= NP ? NP - 1 : (p - sh));
                const int i1 = (p + 1 - sh) < 0 ? 0 : ((p + 1 - sh) >= NP ? NP - 1 : (p + 1 - sh));
                const T t0 = pred ? d[i0] : s[i0];
                const T t1 = (nc > 1) ? (pred ? d[i1] : s[i1]) : T(0);
                if (nc == 1) {
                    v = fp::mac(v, lc.c[st][0], t0);
                } else if (STRICT) {
                    bool interior = true;
                    if (edge) {
                        int gi = g0 + p;
                        if (gi < 0) gi += half; else if (gi >= half) gi -= half;
                        interior = (gi >= left) && (gi <= half + sh - nc);
                    }
                    if (interior) v = fp::add(v, fp::mac(fp::mul(lc.c[st][0], t0), lc.c[st][1], t1));
                    else          v = fp::mac(fp::mac(v, lc.c[st][0], t0), lc.c[st][1], t1);
                } else {
                    v = fp::mac(fp::mac(v, lc.c[st][0], t0), lc.c[st][1], t1);
                }
                if (pred) s[p] = v; else d[p] = v;
            }
        }
    }
}


// ---------------------------------------------------------------------------------------------------
// orthogonal filter bank, F taps (even): one level on a segment held as polyphase pairs x[2p] = s[p], x[2p+1] = d[p].
//   analysis  (filtdown!, transforms_filter.jl:387-433):  a[p] = sum_m h[m] x[2p+m]        (increasing m)
//                                                          dd[p] = sum_m g[m] x[2p+1-m]      (increasing x index)
//   synthesis (filtup!, :467-541), s = approximation band, d = detail band:
//       x[2p]   = (h[F-2] a[p-H] + ... + h[0] a[p]) + (g[1] d[p] + g[3] d[p+1] + ... + g[F-1] d[p+H])
//       x[2p+1] = (h[F-1] a[p-H] + ... + h[1] a[p]) + (g[0] d[p] + g[2] d[p+1] + ... + g[F-2] d[p+H])
// with H = (F-2)/2; products accumulated in that order, the two band sums added last (same as the generic kernels).
// ---------------------------------------------------------------------------------------------------
template <int F_> struct ShapeFirA { static constexpr int F = F_; static_assert(F_ >= 2 && F_ % 2 == 0, "even filter length"); };
template <int F_> struct ShapeFirS { static constexpr int F = F_; static_assert(F_ >= 2 && F_ % 2 == 0, "even filter length"); };
template <int F> struct Halo<ShapeFirA<F>> {
    __host__ __device__ static constexpr int left() { return (F - 2) / 2; }
    __host__ __device__ static constexpr int right() { return (F - 2) / 2; }
};
template <int F> struct Halo<ShapeFirS<F>> {
    __host__ __device__ static constexpr int left() { return (F - 2) / 2; }
    __host__ __device__ static constexpr int right() { return (F - 2) / 2; }
};
template <typename T, int F> struct FirCoefs { T h[F]; T g[F]; };

template <class S> struct IsFir { static constexpr bool value = false; };
template <int F> struct IsFir<ShapeFirA<F>> { static constexpr bool value = true; };
template <int F> struct IsFir<ShapeFirS<F>> { static constexpr bool value = true; };
template <class S> struct HasNorm { static constexpr bool value = !IsFir<S>::value; };
template <class S, typename T> struct CoefsOf { using type = LiftCoefs<T>; };
template <int F, typename T> struct CoefsOf<ShapeFirA<F>, T> { using type = FirCoefs<T, F>; };
template <int F, typename T> struct CoefsOf<ShapeFirS<F>, T> { using type = FirCoefs<T, F>; };

template <typename T, int F, bool STRICT, int NP>
__device__ __forceinline__ void fir_ana_regs(T (&s)[NP], T (&d)[NP], const FirCoefs<T, F> &fc) {
    using fp = FP<STRICT>;
    constexpr int H = (F - 2) / 2;
    T oa[NP], od[NP];
#pragma unroll
    for (int p = H; p < NP - H; ++p) {
        // a[p]: x[2p + m], m = 0 .. F-1
        T a = fp::mul(fc.h[0], s[p]);
#pragma unroll
        for (int m = 1; m < F; ++m) a = fp::mac(a, fc.h[m], (m & 1) ? d[p + (m - 1) / 2] : s[p + m / 2]);
        // dd[p]: x[2p + 2 - F + i] with tap g[F-1-i], i = 0 .. F-1
        T q = fp::mul(fc.g[F - 1], s[p - H]);
#pragma unroll
        for (int i = 1; i < F; ++i) q = fp::mac(q, fc.g[F - 1 - i], (i & 1) ? d[p - H + (i - 1) / 2] : s[p - H + i / 2]);
        oa[p] = a;
        od[p] = q;
    }
#pragma unroll
    for (int p = H; p < NP - H; ++p) { s[p] = oa[p]; d[p] = od[p]; }
}

template <typename T, int F, bool STRICT, int NP>
__device__ __forceinline__ void fir_syn_regs(T (&s)[NP], T (&d)[NP], const FirCoefs<T, F> &fc) {
    using fp = FP<STRICT>;
    constexpr int H = (F - 2) / 2;
    T o0[NP], o1[NP];
#pragma unroll
    for (int p = H; p < NP - H; ++p) {
        // STRICT: the two band sums separately, added last (the reference's order); fast mode: the detail terms continue the
        // approximation's chain (no second FMUL, no final FADD)
        T ra = fp::mul(fc.h[F - 2], s[p - H]);
#pragma unroll
        for (int m = F - 4; m >= 0; m -= 2) ra = fp::mac(ra, fc.h[m], s[p - m / 2]);
        T rd = STRICT ? fp::mul(fc.g[1], d[p]) : fp::mac(ra, fc.g[1], d[p]);
#pragma unroll
        for (int m = 3; m < F; m += 2) rd = fp::mac(rd, fc.g[m], d[p + (m - 1) / 2]);
        o0[p] = STRICT ? fp::add(ra, rd) : rd;
        ra = fp::mul(fc.h[F - 1], s[p - H]);
#pragma unroll
        for (int m = F - 3; m >= 1; m -= 2) ra = fp::mac(ra, fc.h[m], s[p - (m - 1) / 2]);
        rd = STRICT ? fp::mul(fc.g[0], d[p]) : fp::mac(ra, fc.g[0], d[p]);
#pragma unroll
        for (int m = 2; m < F; m += 2) rd = fp::mac(rd, fc.g[m], d[p + m / 2]);
        o1[p] = STRICT ? fp::add(ra, rd) : rd;
    }
#pragma unroll
    for (int p = H; p < NP - H; ++p) { s[p] = o0[p]; d[p] = o1[p]; }
}

// one interface for the tile kernels
template <typename T, class S, bool STRICT, int NP> struct TileTransform {
    static __device__ __forceinline__ void run(T (&s)[NP], T (&d)[NP], const LiftCoefs<T> &lc, int g0, int half, bool edge) {
        lift_regs<T, S, STRICT, NP>(s, d, lc, g0, half, edge);
    }
};
template <typename T, int F, bool STRICT, int NP> struct TileTransform<T, ShapeFirA<F>, STRICT, NP> {
    static __device__ __forceinline__ void run(T (&s)[NP], T (&d)[NP], const FirCoefs<T, F> &fc, int, int, bool) {
        fir_ana_regs<T, F, STRICT, NP>(s, d, fc);
    }
};
template <typename T, int F, bool STRICT, int NP> struct TileTransform<T, ShapeFirS<F>, STRICT, NP> {
    static __device__ __forceinline__ void run(T (&s)[NP], T (&d)[NP], const FirCoefs<T, F> &fc, int, int, bool) {
        fir_syn_regs<T, F, STRICT, NP>(s, d, fc);
    }
};
// band factors (lifting: normalize!; filter banks: none)
template <class S, typename T, class Cf> __device__ __forceinline__ T band_n1(const Cf &c) { if constexpr (HasNorm<S>::value) return c.n1; else return T(1); }
template <class S, typename T, class Cf> __device__ __forceinline__ T band_n2(const Cf &c) { if constexpr (HasNorm<S>::value) return c.n2; else return T(1); }

} // namespace wb
