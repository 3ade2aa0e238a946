// modwt.cu -- maximal-overlap DWT (undecimated, a-trous): SURVEY 8(f) row 1, the first widening beyond the
// decimated path.  Replaces modwt / imodwt / modwt_step / imodwt_step of
// src/Transforms/transforms_maximal_overlap.jl (the reference's GPU extension hooks the same four,
// ext/WaveletsGPUExt/WaveletsGPUExt.jl:11).
//
//   modwt_step (level j, dilation s = 2^(j-1)):   W_j[t] = sum_n h[n] V[(t - n s) mod N],  V_j[t] = sum_n g[n] V[(t - n s) mod N]
//   imodwt_step:                                   V[t]  = sum_n ( h[n] W_j[(t + n s) mod N] + g[n] V_j[(t + n s) mod N] )
// with h = mirror(qmf)/sqrt(2), g = reverse(qmf)/sqrt(2) kept in Float64 (the reference never rounds the MODWT
// filters to the element type), accumulation in tap order.  One thread per output sample (both outputs of the
// forward step from the same taps), consecutive threads on consecutive samples: every tap is a coalesced row read
// that hits L1/L2 after the first touch; one launch per level (levels are sequentially dependent and the dilation
// makes tiles useless beyond the first levels).  HBM traffic per forward level: read N, write 2N.
//
// STRICT: a Float32 signal is filtered in Float64 with the accumulator rounded to Float32 after every tap, exactly
// like `w1[t] += h[n]*v[k]` on a Vector{Float32} does; non-strict Float32 uses Float32 taps and FMA.
#include "common.cuh"
#include <cmath>
#include <cstdlib>
#include <vector>

namespace wb {

struct ModwtTaps { double h[MAXF]; double g[MAXF]; float hf[MAXF]; float gf[MAXF]; int F; };

// Output-sample -> thread mapping.  Dilation s < 32: a CTA takes 256 consecutive samples (the F taps reach back
// (F-1) s <= a few hundred samples: L1 hits).  s >= 32: consecutive samples would make every tap of a CTA a different,
// never re-used 1 KiB segment (F-fold DRAM read amplification), so a CTA takes a tile of 32 consecutive PHASES (lanes:
// coalesced 128-byte rows) x M = 8 MK consecutive positions of the stride-s sequences; warp w walks positions
// [w MK, (w+1) MK), and the tile's (M + F - 1) input rows stay in L1 across its F taps.
struct ModwtMap { int64_t s; int64_t chunks; int tiled; int MK; };

template <class Fn>
__device__ __forceinline__ void modwt_for_each(int64_t n, const ModwtMap &mp, Fn fn) {
    if (!mp.tiled) {
        const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
        if (t < n) fn(t);
    } else {
        const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
        const int64_t tile = blockIdx.x, rc = tile % mp.chunks, hi = tile / mp.chunks;
        int64_t t = hi * (mp.s * 8 * mp.MK) + rc * 32 + lane + (int64_t)w * mp.MK * mp.s;
        for (int kk = 0; kk < mp.MK; ++kk, t += mp.s)
            if (t < n) fn(t);
    }
}
__device__ __forceinline__ int64_t wrap_dn(int64_t k, int64_t n) {   // k may be far below zero when s >= n
    if (k < 0) { k += n; if (k < 0) { k %= n; if (k < 0) k += n; } }
    return k;
}
__device__ __forceinline__ int64_t wrap_up(int64_t k, int64_t n) {
    if (k >= n) { k -= n; if (k >= n) k %= n; }
    return k;
}

template <typename T, bool STRICT>
__global__ void __launch_bounds__(256)
k_modwt_step(const T *__restrict__ v, T *__restrict__ v1, T *__restrict__ w1, int64_t n, int64_t sv, int64_t sv1, int64_t sw1,
             int64_t B, const __grid_constant__ ModwtMap mp, const __grid_constant__ ModwtTaps tp) {
    const int64_t s = mp.s;
    for (int64_t b = blockIdx.y; b < B; b += gridDim.y) {
        const T *vb = v + b * sv;
        T *wo = w1 + b * sw1, *vo = v1 + b * sv1;
        modwt_for_each(n, mp, [&](int64_t t) {
            int64_t k = t;
            if constexpr (sizeof(T) == 8) {
                double w = FP<STRICT>::mul(tp.h[0], (double)vb[k]), a = FP<STRICT>::mul(tp.g[0], (double)vb[k]);
                for (int m = 1; m < tp.F; ++m) {
                    k = wrap_dn(k - s, n);
                    const double x = vb[k];
                    w = FP<STRICT>::mac(w, tp.h[m], x);
                    a = FP<STRICT>::mac(a, tp.g[m], x);
                }
                wo[t] = w;
                vo[t] = a;
            } else if constexpr (STRICT) {
                float w = __double2float_rn(__dmul_rn(tp.h[0], (double)vb[k])), a = __double2float_rn(__dmul_rn(tp.g[0], (double)vb[k]));
                for (int m = 1; m < tp.F; ++m) {
                    k = wrap_dn(k - s, n);
                    const double x = (double)vb[k];
                    w = __double2float_rn(__dadd_rn((double)w, __dmul_rn(tp.h[m], x)));
                    a = __double2float_rn(__dadd_rn((double)a, __dmul_rn(tp.g[m], x)));
                }
                wo[t] = w;
                vo[t] = a;
            } else {
                float w = tp.hf[0] * vb[k], a = tp.gf[0] * vb[k];
                for (int m = 1; m < tp.F; ++m) {
                    k = wrap_dn(k - s, n);
                    const float x = vb[k];
                    w = fmaf(tp.hf[m], x, w);
                    a = fmaf(tp.gf[m], x, a);
                }
                wo[t] = w;
                vo[t] = a;
            }
        });
    }
}

template <typename T, bool STRICT>
__global__ void __launch_bounds__(256)
k_imodwt_step(const T *__restrict__ v, const T *__restrict__ w, T *__restrict__ v0, int64_t n, int64_t sv, int64_t sw, int64_t sv0,
              int64_t B, const __grid_constant__ ModwtMap mp, const __grid_constant__ ModwtTaps tp) {
    const int64_t s = mp.s;
    for (int64_t b = blockIdx.y; b < B; b += gridDim.y) {
        const T *vb = v + b * sv;
        const T *wb_ = w + b * sw;
        T *xo = v0 + b * sv0;
        modwt_for_each(n, mp, [&](int64_t t) {
            int64_t k = t;
            if constexpr (sizeof(T) == 8) {
                double acc = FP<STRICT>::add(FP<STRICT>::mul(tp.h[0], (double)wb_[k]), FP<STRICT>::mul(tp.g[0], (double)vb[k]));
                for (int m = 1; m < tp.F; ++m) {
                    k = wrap_up(k + s, n);
                    const double term = STRICT ? __dadd_rn(__dmul_rn(tp.h[m], (double)wb_[k]), __dmul_rn(tp.g[m], (double)vb[k]))
                                               : fma(tp.g[m], (double)vb[k], tp.h[m] * (double)wb_[k]);
                    acc = FP<STRICT>::add(acc, term);
                }
                xo[t] = acc;
            } else if constexpr (STRICT) {
                float acc = __double2float_rn(__dadd_rn(__dmul_rn(tp.h[0], (double)wb_[k]), __dmul_rn(tp.g[0], (double)vb[k])));
                for (int m = 1; m < tp.F; ++m) {
                    k = wrap_up(k + s, n);
                    const double term = __dadd_rn(__dmul_rn(tp.h[m], (double)wb_[k]), __dmul_rn(tp.g[m], (double)vb[k]));
                    acc = __double2float_rn(__dadd_rn((double)acc, term));
                }
                xo[t] = acc;
            } else {
                float acc = fmaf(tp.gf[0], vb[k], tp.hf[0] * wb_[k]);
                for (int m = 1; m < tp.F; ++m) {
                    k = wrap_up(k + s, n);
                    acc += fmaf(tp.gf[m], vb[k], tp.hf[m] * wb_[k]);
                }
                xo[t] = acc;
            }
        });
    }
}

static ModwtMap make_map(int64_t n, int j, unsigned &gridx) {
    ModwtMap mp;
    mp.s = (int64_t)1 << (j - 1);
    mp.tiled = mp.s >= 32 ? 1 : 0;
    if (!mp.tiled) { mp.MK = 1; mp.chunks = 1; gridx = (unsigned)((n + 255) / 256); return mp; }
    const int64_t per_phase = (n + mp.s - 1) / mp.s;              // positions of one stride-s sequence
    int64_t mk = (per_phase + 7) / 8;
    mp.MK = (int)(mk < 1 ? 1 : (mk > 8 ? 8 : mk));
    mp.chunks = mp.s / 32;
    const int64_t span = mp.s * 8 * mp.MK;
    gridx = (unsigned)(((n + span - 1) / span) * mp.chunks);
    return mp;
}

// ===================================================================================================
// fused multi-level kernels.  One launch takes K consecutive levels with the scaling coefficients resident in shared
// memory, so a group reads V once and writes its K detail rows and one V (forward), or reads them and writes one V
// (inverse) -- instead of a read + two writes per level.
//   tile = (H + M) rows x 32 lanes, flat index q = 32 row + lane.  Element (row, lane) is global sample
//   base + lane + rs (row - H)   (forward: halo rows first; inverse: owned rows first, halo after):
//     * first group ("flat"): rs = 32 -- the tile is a contiguous run of the signal, dilation 2^i is the flat offset 2^i;
//     * later groups ("phase", base dilation s0 = 2^j0 >= 32, n % s0 == 0): rs = s0 -- lanes are 32 consecutive phases
//       of the stride-s0 sequences (coalesced 128-byte rows), dilation s0 2^i is the flat offset 32 * 2^i.
//   halo mode: rows outside the owned range are recomputed per tile ((F-1)(2^K - 1) base steps of reach);
//   periodic mode: the tile holds whole periods (a short signal, or all n/s0 positions of its phases): no halo, taps
//   wrap inside the tile and ALL remaining levels run in the one launch.
// Arithmetic per output is the per-level kernels' (tap order, rounding), so strict mode stays bit-identical.
// ===================================================================================================
struct MGroup {
    int64_t n, rs, chunks, ntiles;
    int j0, K, dmul, H, M, periodic, NQ;
};
template <int F> struct MTaps { double h[F]; double g[F]; float hf[F]; float gf[F]; };

__device__ __forceinline__ int64_t gmod(int64_t g, int64_t n) {
    if (g < 0) { g += n; if (g < 0) { g %= n; if (g < 0) g += n; } }
    else if (g >= n) { g -= n; if (g >= n) g %= n; }
    return g;
}

template <typename T, int F, bool STRICT>
__global__ void __launch_bounds__(512)
k_modwt_group(const T *__restrict__ vin, int64_t svin, T *__restrict__ y, int64_t ys, T *__restrict__ vout, int64_t svout,
              int64_t B, const __grid_constant__ MGroup gp, const __grid_constant__ MTaps<F> tp) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int offs[F];
    T *const buf0 = reinterpret_cast<T *>(smem_raw);
    T *const buf1 = buf0 + ((gp.NQ + 31) & ~31);
    const int NT = blockDim.x, tid = threadIdx.x;
    const int64_t tile = blockIdx.x, rc = tile % gp.chunks, hi = tile / gp.chunks;
    const int64_t base = hi * (gp.rs * gp.M) + rc * 32;
    const int own_lo = 32 * gp.H;
    for (int64_t b = blockIdx.y; b < B; b += gridDim.y) {
        const T *vb = vin + b * svin;
        for (int q = tid; q < gp.NQ; q += NT) {
            const int64_t g = base + (q & 31) + gp.rs * ((q >> 5) - gp.H);
            buf0[q] = vb[gmod(g, gp.n)];
        }
        __syncthreads();
        for (int i = 0; i < gp.K; ++i) {
            const int d = gp.dmul << i;
            const T *src = (i & 1) ? buf1 : buf0;
            T *dst = (i & 1) ? buf0 : buf1;
            if (gp.periodic) {
                if (tid < F) offs[tid] = (int)(((int64_t)tid * d) % gp.NQ);
                __syncthreads();
            }
            const int qlo = gp.periodic ? 0 : (F - 1) * ((2 << i) - 1) * gp.dmul;
            T *wrow = y + b * ys + (int64_t)(gp.j0 + i) * gp.n;
            for (int q = qlo + tid; q < gp.NQ; q += NT) {
                T xs[F];
                xs[0] = src[q];
#pragma unroll
                for (int k = 1; k < F; ++k) {
                    int idx;
                    if (gp.periodic) { idx = q - offs[k]; if (idx < 0) idx += gp.NQ; }
                    else idx = q - k * d;
                    xs[k] = src[idx];
                }
                T w, a;
                if constexpr (sizeof(T) == 8) {
                    w = FP<STRICT>::mul(tp.h[0], xs[0]); a = FP<STRICT>::mul(tp.g[0], xs[0]);
#pragma unroll
                    for (int k = 1; k < F; ++k) { w = FP<STRICT>::mac(w, tp.h[k], xs[k]); a = FP<STRICT>::mac(a, tp.g[k], xs[k]); }
                } else if constexpr (STRICT) {
                    w = __double2float_rn(__dmul_rn(tp.h[0], (double)xs[0])); a = __double2float_rn(__dmul_rn(tp.g[0], (double)xs[0]));
#pragma unroll
                    for (int k = 1; k < F; ++k) {
                        w = __double2float_rn(__dadd_rn((double)w, __dmul_rn(tp.h[k], (double)xs[k])));
                        a = __double2float_rn(__dadd_rn((double)a, __dmul_rn(tp.g[k], (double)xs[k])));
                    }
                } else {
                    w = tp.hf[0] * xs[0]; a = tp.gf[0] * xs[0];
#pragma unroll
                    for (int k = 1; k < F; ++k) { w = fmaf(tp.hf[k], xs[k], w); a = fmaf(tp.gf[k], xs[k], a); }
                }
                dst[q] = a;
                if (q >= own_lo) {
                    const int64_t g = base + (q & 31) + gp.rs * ((q >> 5) - gp.H);
                    if (g < gp.n) wrow[g] = w;
                }
            }
            __syncthreads();
        }
        const T *fin = (gp.K & 1) ? buf1 : buf0;
        T *vo = vout + b * svout;
        for (int q = own_lo + tid; q < gp.NQ; q += NT) {
            const int64_t g = base + (q & 31) + gp.rs * ((q >> 5) - gp.H);
            if (g < gp.n) vo[g] = fin[q];
        }
        __syncthreads();
    }
}

template <typename T, int F, bool STRICT>
__global__ void __launch_bounds__(512)
k_imodwt_group(const T *__restrict__ vin, int64_t svin, const T *__restrict__ xw, int64_t ws, T *__restrict__ vout, int64_t svout,
               int64_t B, const __grid_constant__ MGroup gp, const __grid_constant__ MTaps<F> tp) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int offs[F];
    const int NQP = (gp.NQ + 31) & ~31;
    T *const buf0 = reinterpret_cast<T *>(smem_raw);
    T *const buf1 = buf0 + NQP;
    T *const wbuf = buf1 + NQP;
    const int NT = blockDim.x, tid = threadIdx.x;
    const int64_t tile = blockIdx.x, rc = tile % gp.chunks, hi = tile / gp.chunks;
    const int64_t base = hi * (gp.rs * gp.M) + rc * 32;      // owned rows first, the halo rows behind them
    const int own_hi = gp.periodic ? gp.NQ : 32 * gp.M;
    for (int64_t b = blockIdx.y; b < B; b += gridDim.y) {
        const T *vb = vin + b * svin;
        for (int q = tid; q < gp.NQ; q += NT) {
            const int64_t g = base + (q & 31) + gp.rs * (q >> 5);
            buf0[q] = vb[gmod(g, gp.n)];
        }
        int cur = 0;
        for (int i = gp.K - 1; i >= 0; --i) {
            const int d = gp.dmul << i;
            const T *wcol = xw + b * ws + (int64_t)(gp.j0 + i) * gp.n;
            const int qhi = gp.periodic ? gp.NQ : gp.NQ - (F - 1) * ((1 << gp.K) - (1 << i)) * gp.dmul;
            const int qload = gp.periodic ? gp.NQ : qhi + (F - 1) * d;       // W_j is only read below this flat index
            for (int q = tid; q < qload; q += NT) {
                const int64_t g = base + (q & 31) + gp.rs * (q >> 5);
                wbuf[q] = wcol[gmod(g, gp.n)];
            }
            if (gp.periodic && tid < F) offs[tid] = (int)(((int64_t)tid * d) % gp.NQ);
            __syncthreads();
            const T *src = cur ? buf1 : buf0;
            T *dst = cur ? buf0 : buf1;
            for (int q = tid; q < qhi; q += NT) {
                T xv[F], xd[F];
                xv[0] = src[q]; xd[0] = wbuf[q];
#pragma unroll
                for (int k = 1; k < F; ++k) {
                    int idx;
                    if (gp.periodic) { idx = q + offs[k]; if (idx >= gp.NQ) idx -= gp.NQ; }
                    else idx = q + k * d;
                    xv[k] = src[idx]; xd[k] = wbuf[idx];
                }
                T acc;
                if constexpr (sizeof(T) == 8) {
                    acc = FP<STRICT>::add(FP<STRICT>::mul(tp.h[0], xd[0]), FP<STRICT>::mul(tp.g[0], xv[0]));
#pragma unroll
                    for (int k = 1; k < F; ++k) {
                        const double term = STRICT ? __dadd_rn(__dmul_rn(tp.h[k], xd[k]), __dmul_rn(tp.g[k], xv[k]))
                                                   : fma(tp.g[k], (double)xv[k], tp.h[k] * (double)xd[k]);
                        acc = FP<STRICT>::add(acc, term);
                    }
                } else if constexpr (STRICT) {
                    acc = __double2float_rn(__dadd_rn(__dmul_rn(tp.h[0], (double)xd[0]), __dmul_rn(tp.g[0], (double)xv[0])));
#pragma unroll
                    for (int k = 1; k < F; ++k) {
                        const double term = __dadd_rn(__dmul_rn(tp.h[k], (double)xd[k]), __dmul_rn(tp.g[k], (double)xv[k]));
                        acc = __double2float_rn(__dadd_rn((double)acc, term));
                    }
                } else {
                    acc = fmaf(tp.gf[0], xv[0], tp.hf[0] * xd[0]);
#pragma unroll
                    for (int k = 1; k < F; ++k) acc += fmaf(tp.gf[k], xv[k], tp.hf[k] * xd[k]);
                }
                dst[q] = acc;
            }
            __syncthreads();
            cur ^= 1;
        }
        const T *fin = cur ? buf1 : buf0;
        T *vo = vout + b * svout;
        for (int q = tid; q < own_hi; q += NT) {
            const int64_t g = base + (q & 31) + gp.rs * (q >> 5);
            if (g < gp.n) vo[g] = fin[q];
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------------
// 128-bit editions of the two group kernels (n % 4 == 0, 16-byte aligned arrays): every loop moves FOUR consecutive
// flat elements.  The scalar kernels above issue one LDS per tap per output and were instruction-bound (82 % issue-slot
// utilisation, profiles/r01c_modwt_f32); here a tap of dilation d >= 4 is one 16-byte shared-memory load for four
// outputs, and the two finest dilations (1, 2) read one register window of 16-byte chunks.  Valid ranges are rounded to
// chunk boundaries per level (the flat plan reserves 4 K extra halo elements for that).
// ---------------------------------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ void v_ld4(T (&w)[4], const T *p) {
    if constexpr (sizeof(T) == 4) { const float4 q = *reinterpret_cast<const float4 *>(p); w[0] = q.x; w[1] = q.y; w[2] = q.z; w[3] = q.w; }
    else { const double2 q0 = *reinterpret_cast<const double2 *>(p), q1 = *reinterpret_cast<const double2 *>(p + 2); w[0] = q0.x; w[1] = q0.y; w[2] = q1.x; w[3] = q1.y; }
}
template <typename T> __device__ __forceinline__ void v_st4(T *p, const T (&w)[4]) {
    if constexpr (sizeof(T) == 4) *reinterpret_cast<float4 *>(p) = make_float4(w[0], w[1], w[2], w[3]);
    else { *reinterpret_cast<double2 *>(p) = make_double2(w[0], w[1]); *reinterpret_cast<double2 *>(p + 2) = make_double2(w[2], w[3]); }
}
// one forward tap on four outputs (the per-level kernels' arithmetic)
template <typename T, int F, bool STRICT>
__device__ __forceinline__ void fwd_tap(T (&w)[4], T (&a)[4], const T (&x)[4], const MTaps<F> &tp, int k, bool first) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        if constexpr (sizeof(T) == 8) {
            if (first) { w[r] = FP<STRICT>::mul(tp.h[k], x[r]); a[r] = FP<STRICT>::mul(tp.g[k], x[r]); }
            else       { w[r] = FP<STRICT>::mac(w[r], tp.h[k], x[r]); a[r] = FP<STRICT>::mac(a[r], tp.g[k], x[r]); }
        } else if constexpr (STRICT) {
            if (first) { w[r] = __double2float_rn(__dmul_rn(tp.h[k], (double)x[r])); a[r] = __double2float_rn(__dmul_rn(tp.g[k], (double)x[r])); }
            else {
                w[r] = __double2float_rn(__dadd_rn((double)w[r], __dmul_rn(tp.h[k], (double)x[r])));
                a[r] = __double2float_rn(__dadd_rn((double)a[r], __dmul_rn(tp.g[k], (double)x[r])));
            }
        } else {
            if (first) { w[r] = tp.hf[k] * x[r]; a[r] = tp.gf[k] * x[r]; }
            else       { w[r] = fmaf(tp.hf[k], x[r], w[r]); a[r] = fmaf(tp.gf[k], x[r], a[r]); }
        }
    }
}
// one inverse tap: acc (+)= h[k] w + g[k] v
template <typename T, int F, bool STRICT>
__device__ __forceinline__ void inv_tap(T (&acc)[4], const T (&xw_)[4], const T (&xv)[4], const MTaps<F> &tp, int k, bool first) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        if constexpr (sizeof(T) == 8) {
            if (first) acc[r] = FP<STRICT>::add(FP<STRICT>::mul(tp.h[k], xw_[r]), FP<STRICT>::mul(tp.g[k], xv[r]));
            else {
                const double term = STRICT ? __dadd_rn(__dmul_rn(tp.h[k], xw_[r]), __dmul_rn(tp.g[k], xv[r]))
                                           : fma(tp.g[k], (double)xv[r], tp.h[k] * (double)xw_[r]);
                acc[r] = FP<STRICT>::add(acc[r], term);
            }
        } else if constexpr (STRICT) {
            const double term = __dadd_rn(__dmul_rn(tp.h[k], (double)xw_[r]), __dmul_rn(tp.g[k], (double)xv[r]));
            if (first) acc[r] = __double2float_rn(term);
            else       acc[r] = __double2float_rn(__dadd_rn((double)acc[r], term));
        } else {
            if (first) acc[r] = fmaf(tp.gf[k], xv[r], tp.hf[k] * xw_[r]);
            else       acc[r] += fmaf(tp.gf[k], xv[r], tp.hf[k] * xw_[r]);
        }
    }
}

template <typename T, int F, bool STRICT, int D>      // the two finest dilations of a flat group: register window
__device__ __forceinline__ void fwd_window(T (&w)[4], T (&a)[4], const T *src, int c, int NCq, bool periodic, const MTaps<F> &tp) {
    constexpr int NCW = ((F - 1) * D + 3) / 4 + 1;
    T win[4 * NCW];
    int ci = c - (NCW - 1);
    if (periodic) { ci %= NCq; if (ci < 0) ci += NCq; }
#pragma unroll
    for (int i = 0; i < NCW; ++i) {
        T t4[4];
        v_ld4(t4, src + 4 * ci);
        win[4 * i] = t4[0]; win[4 * i + 1] = t4[1]; win[4 * i + 2] = t4[2]; win[4 * i + 3] = t4[3];
        ++ci;
        if (periodic && ci == NCq) ci = 0;
    }
#pragma unroll
    for (int k = 0; k < F; ++k) {
        const T x[4] = {win[4 * (NCW - 1) + 0 - k * D], win[4 * (NCW - 1) + 1 - k * D], win[4 * (NCW - 1) + 2 - k * D], win[4 * (NCW - 1) + 3 - k * D]};
        fwd_tap<T, F, STRICT>(w, a, x, tp, k, k == 0);
    }
}
template <typename T, int F, bool STRICT, int D>
__device__ __forceinline__ void inv_window(T (&acc)[4], const T *src, const T *wsrc, int c, int NCq, bool periodic, const MTaps<F> &tp) {
    constexpr int NCW = ((F - 1) * D + 3) / 4 + 1;
    T wv[4 * NCW], ww[4 * NCW];
    int ci = c;
#pragma unroll
    for (int i = 0; i < NCW; ++i) {
        T t4[4];
        v_ld4(t4, src + 4 * ci);
        wv[4 * i] = t4[0]; wv[4 * i + 1] = t4[1]; wv[4 * i + 2] = t4[2]; wv[4 * i + 3] = t4[3];
        v_ld4(t4, wsrc + 4 * ci);
        ww[4 * i] = t4[0]; ww[4 * i + 1] = t4[1]; ww[4 * i + 2] = t4[2]; ww[4 * i + 3] = t4[3];
        ++ci;
        if (periodic && ci == NCq) ci = 0;
    }
#pragma unroll
    for (int k = 0; k < F; ++k) {
        const T xv[4] = {wv[k * D], wv[k * D + 1], wv[k * D + 2], wv[k * D + 3]};
        const T xd[4] = {ww[k * D], ww[k * D + 1], ww[k * D + 2], ww[k * D + 3]};
        inv_tap<T, F, STRICT>(acc, xd, xv, tp, k, k == 0);
    }
}

template <typename T, int F, bool STRICT>
__global__ void __launch_bounds__(512)
k_modwt_group_v4(const T *__restrict__ vin, int64_t svin, T *__restrict__ y, int64_t ys, T *__restrict__ vout, int64_t svout,
                 int64_t B, const __grid_constant__ MGroup gp, const __grid_constant__ MTaps<F> tp) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int offs[F];
    T *const buf0 = reinterpret_cast<T *>(smem_raw);
    T *const buf1 = buf0 + ((gp.NQ + 31) & ~31);
    const int NT = blockDim.x, tid = threadIdx.x;
    const int NCq = gp.NQ >> 2;
    const int64_t tile = blockIdx.x, rc = tile % gp.chunks, hi = tile / gp.chunks;
    const int64_t base = hi * (gp.rs * gp.M) + rc * 32;
    const int own_c = 8 * gp.H;                              // first owned chunk
    for (int64_t b = blockIdx.y; b < B; b += gridDim.y) {
        const T *vb = vin + b * svin;
        for (int c = tid; c < NCq; c += NT) {
            const int q = 4 * c;
            const int64_t g = gmod(base + (q & 31) + gp.rs * ((q >> 5) - gp.H), gp.n);     // chunk-aligned, never straddles n
            T t4[4];
            v_ld4(t4, vb + g);
            v_st4(buf0 + q, t4);
        }
        __syncthreads();
        int lo = 0;
        for (int i = 0; i < gp.K; ++i) {
            const int d = gp.dmul << i;
            const T *src = (i & 1) ? buf1 : buf0;
            T *dst = (i & 1) ? buf0 : buf1;
            if (gp.periodic) {
                if (tid < F) offs[tid] = (int)(((int64_t)tid * d) % gp.NQ);
                __syncthreads();
            } else lo = (lo + (F - 1) * d + 3) & ~3;
            T *wrow = y + b * ys + (int64_t)(gp.j0 + i) * gp.n;
            for (int c = (lo >> 2) + tid; c < NCq; c += NT) {
                const int q = 4 * c;
                T w[4], a[4];
                if (d == 1) fwd_window<T, F, STRICT, 1>(w, a, src, c, NCq, gp.periodic != 0, tp);
                else if (d == 2) fwd_window<T, F, STRICT, 2>(w, a, src, c, NCq, gp.periodic != 0, tp);
                else {
#pragma unroll
                    for (int k = 0; k < F; ++k) {
                        int idx;
                        if (gp.periodic) { idx = q - offs[k]; if (idx < 0) idx += gp.NQ; }
                        else idx = q - k * d;
                        T x[4];
                        v_ld4(x, src + idx);
                        fwd_tap<T, F, STRICT>(w, a, x, tp, k, k == 0);
                    }
                }
                v_st4(dst + q, a);
                if (c >= own_c) {
                    const int64_t g = base + (q & 31) + gp.rs * ((q >> 5) - gp.H);
                    if (g < gp.n) v_st4(wrow + g, w);
                }
            }
            __syncthreads();
        }
        const T *fin = (gp.K & 1) ? buf1 : buf0;
        T *vo = vout + b * svout;
        for (int c = own_c + tid; c < NCq; c += NT) {
            const int q = 4 * c;
            const int64_t g = base + (q & 31) + gp.rs * ((q >> 5) - gp.H);
            if (g < gp.n) { T t4[4]; v_ld4(t4, fin + q); v_st4(vo + g, t4); }
        }
        __syncthreads();
    }
}

template <typename T, int F, bool STRICT>
__global__ void __launch_bounds__(512)
k_imodwt_group_v4(const T *__restrict__ vin, int64_t svin, const T *__restrict__ xw, int64_t ws, T *__restrict__ vout, int64_t svout,
                  int64_t B, const __grid_constant__ MGroup gp, const __grid_constant__ MTaps<F> tp) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int offs[F];
    const int NQP = (gp.NQ + 31) & ~31;
    T *const buf0 = reinterpret_cast<T *>(smem_raw);
    T *const buf1 = buf0 + NQP;
    T *const wbuf = buf1 + NQP;
    const int NT = blockDim.x, tid = threadIdx.x;
    const int NCq = gp.NQ >> 2;
    const int64_t tile = blockIdx.x, rc = tile % gp.chunks, hi = tile / gp.chunks;
    const int64_t base = hi * (gp.rs * gp.M) + rc * 32;
    const int own_c = gp.periodic ? NCq : 8 * gp.M;          // owned chunks: [0, own_c)
    for (int64_t b = blockIdx.y; b < B; b += gridDim.y) {
        const T *vb = vin + b * svin;
        for (int c = tid; c < NCq; c += NT) {
            const int q = 4 * c;
            const int64_t g = gmod(base + (q & 31) + gp.rs * (q >> 5), gp.n);
            T t4[4];
            v_ld4(t4, vb + g);
            v_st4(buf0 + q, t4);
        }
        int cur = 0, qhi = gp.NQ;
        for (int i = gp.K - 1; i >= 0; --i) {
            const int d = gp.dmul << i;
            const T *wcol = xw + b * ws + (int64_t)(gp.j0 + i) * gp.n;
            int qload = gp.NQ;
            if (!gp.periodic) { qhi = (qhi - (F - 1) * d) & ~3; qload = (qhi + (F - 1) * d + 3) & ~3; }
            for (int c = tid; c < (qload >> 2); c += NT) {
                const int q = 4 * c;
                const int64_t g = gmod(base + (q & 31) + gp.rs * (q >> 5), gp.n);
                T t4[4];
                v_ld4(t4, wcol + g);
                v_st4(wbuf + q, t4);
            }
            if (gp.periodic && tid < F) offs[tid] = (int)(((int64_t)tid * d) % gp.NQ);
            __syncthreads();
            const T *src = cur ? buf1 : buf0;
            T *dst = cur ? buf0 : buf1;
            for (int c = tid; c < (qhi >> 2); c += NT) {
                const int q = 4 * c;
                T acc[4];
                if (d == 1) inv_window<T, F, STRICT, 1>(acc, src, wbuf, c, NCq, gp.periodic != 0, tp);
                else if (d == 2) inv_window<T, F, STRICT, 2>(acc, src, wbuf, c, NCq, gp.periodic != 0, tp);
                else {
#pragma unroll
                    for (int k = 0; k < F; ++k) {
                        int idx;
                        if (gp.periodic) { idx = q + offs[k]; if (idx >= gp.NQ) idx -= gp.NQ; }
                        else idx = q + k * d;
                        T xv[4], xd[4];
                        v_ld4(xv, src + idx);
                        v_ld4(xd, wbuf + idx);
                        inv_tap<T, F, STRICT>(acc, xd, xv, tp, k, k == 0);
                    }
                }
                v_st4(dst + q, acc);
            }
            __syncthreads();
            cur ^= 1;
        }
        const T *fin = cur ? buf1 : buf0;
        T *vo = vout + b * svout;
        for (int c = tid; c < own_c; c += NT) {
            const int q = 4 * c;
            const int64_t g = base + (q & 31) + gp.rs * (q >> 5);
            if (g < gp.n) { T t4[4]; v_ld4(t4, fin + q); v_st4(vo + g, t4); }
        }
        __syncthreads();
    }
}

// ---- host: the group plan ---------------------------------------------------------------------------
struct MStep { bool fused; int level; MGroup g; };      // fused group, or one level (1-based) through the per-level kernel
constexpr int MODWT_CAP_BYTES = 36864;                    // one shared-memory buffer of a forward tile (two per CTA)
constexpr int MODWT_INV_CAP_BYTES = 16384;                // inverse: three per CTA; small tiles keep more CTAs per SM resident, which hides the
                                                          // per-level W loads better (L = 10: 36 K 2.23, 24 K 1.95, 16 K 1.81, 12 K 2.19 ms;
                                                          // 16 K also wins at L = 5 and L = 20: interleaved A/B, tools/ab_env.py, round 2)

static int modwt_cap_bytes(bool fw) {      // shared-memory bytes of one tile buffer (tuning knob; the plan adapts to it)
    const char *e = std::getenv(fw ? "WB200_MODWT_CAP" : "WB200_MODWT_INV_CAP");
    int v = e ? std::atoi(e) : (fw ? MODWT_CAP_BYTES : MODWT_INV_CAP_BYTES);
    if (v < 8192) v = 8192;
    if (v > (fw ? 110 : 72) * 1024) v = (fw ? 110 : 72) * 1024;
    return v & ~1023;
}
static void plan_modwt(std::vector<MStep> &steps, int64_t n, int L, int F, int esz, bool fw) {
    steps.clear();
    const bool can_fuse = (F >= 2 && F <= 20 && (F & 1) == 0) && std::getenv("WB200_DISABLE_MODWT_FUSED") == nullptr;
    const int cap = modwt_cap_bytes(fw) / esz, rows_cap = cap / 32;
    int j = 0;
    if (can_fuse) {
        MGroup g{};
        g.n = n; g.j0 = 0; g.rs = 32; g.dmul = 1; g.chunks = 1;
        if (n <= cap) {
            g.K = L; g.H = 0; g.M = (int)((n + 31) / 32); g.periodic = 1; g.NQ = (int)n; g.ntiles = 1;
        } else {
            int K = 0;
            while (K < L && (int64_t)(F - 1) * ((2 << K) - 1) <= cap / 9) ++K;
            g.K = K; g.H = ((F - 1) * ((1 << K) - 1) + 4 * K + 31) / 32; g.M = rows_cap - g.H; g.periodic = 0;   // (+4K: chunk rounding of the 128-bit kernels)
            g.NQ = 32 * rows_cap; g.ntiles = (n + 32 * (int64_t)g.M - 1) / (32 * (int64_t)g.M);
        }
        if (g.K > 0) { steps.push_back({true, 0, g}); j = g.K; }
    }
    while (j < L) {
        const int64_t s0 = (int64_t)1 << j;
        if (!can_fuse || s0 < 32 || n % s0 != 0) { steps.push_back({false, j + 1, MGroup{}}); ++j; continue; }
        const int64_t P = n / s0;
        MGroup g{};
        g.n = n; g.j0 = j; g.rs = s0; g.dmul = 32; g.chunks = s0 / 32;
        if (P * 32 <= cap) {
            g.K = L - j; g.H = 0; g.M = (int)P; g.periodic = 1; g.NQ = (int)P * 32; g.ntiles = g.chunks;
        } else {
            int K = 0;
            while (K < L - j && (F - 1) * ((2 << K) - 1) <= rows_cap / 3) ++K;
            if (K == 0) { steps.push_back({false, j + 1, MGroup{}}); ++j; continue; }
            g.K = K; g.H = (F - 1) * ((1 << K) - 1); g.M = rows_cap - g.H; g.periodic = 0; g.NQ = 32 * rows_cap;
            g.ntiles = ((n + s0 * g.M - 1) / (s0 * g.M)) * g.chunks;
        }
        if (g.ntiles > 0x7fffffffLL) { steps.push_back({false, j + 1, MGroup{}}); ++j; continue; }
        steps.push_back({true, 0, g});
        j += g.K;
    }
}

template <int F> static void fill_mtaps(MTaps<F> &m, const ModwtTaps &tp) {
    for (int k = 0; k < F; ++k) { m.h[k] = tp.h[k]; m.g[k] = tp.g[k]; m.hf[k] = tp.hf[k]; m.gf[k] = tp.gf[k]; }
}
template <typename T, int F>
static bool launch_group(bool fw, bool strict, const MGroup &g, const T *vin, int64_t svin, T *yw, int64_t ysw, T *vout, int64_t svout,
                         int64_t B, const ModwtTaps &tp, cudaStream_t st) {
    MTaps<F> mt; fill_mtaps<F>(mt, tp);
    const size_t nqp = (size_t)((g.NQ + 31) & ~31);
    const size_t smem = (fw ? 2 : 3) * nqp * sizeof(T);
    dim3 grid((unsigned)g.ntiles, (unsigned)(B < 65535 ? B : 65535)), block(512);
#define WB_GO(KERN, NAME)                                                                                           \
    {                                                                                                               \
        auto kern = KERN;                                                                                           \
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {    \
            (void)cudaGetLastError(); set_error("cudaFuncSetAttribute(" NAME ") failed"); return false; }            \
        LaunchScope scope(NAME, st);                                                                                \
        kern<<<grid, block, smem, st>>>(vin, svin, yw, ysw, vout, svout, B, g, mt);                                  \
    }
    auto al = [](const void *q, int64_t stride) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0 && (stride & 3) == 0; };
    // (Float64 keeps the scalar kernels: four doubles per thread are two 16-byte accesses 32 bytes apart across lanes --
    //  2-way bank conflicts -- and measured 1.7x slower than one double per lane)
    const bool v4 = sizeof(T) == 4 && (g.n & 3) == 0 && (g.NQ & 3) == 0 && al(vin, svin) && al(yw, ysw) && al(vout, svout) &&
                    std::getenv("WB200_MODWT_SCALAR") == nullptr;
    if (v4) {
        if (fw) { if (strict) WB_GO((k_modwt_group_v4<T, F, true>), "modwt_group") else WB_GO((k_modwt_group_v4<T, F, false>), "modwt_group") }
        else    { if (strict) WB_GO((k_imodwt_group_v4<T, F, true>), "imodwt_group") else WB_GO((k_imodwt_group_v4<T, F, false>), "imodwt_group") }
    } else {
        if (fw) { if (strict) WB_GO((k_modwt_group<T, F, true>), "modwt_group") else WB_GO((k_modwt_group<T, F, false>), "modwt_group") }
        else    { if (strict) WB_GO((k_imodwt_group<T, F, true>), "imodwt_group") else WB_GO((k_imodwt_group<T, F, false>), "imodwt_group") }
    }
#undef WB_GO
    return check_launch(fw ? "modwt_group" : "imodwt_group");
}
template <typename T>
static bool launch_group_F(bool fw, bool strict, const MGroup &g, const T *vin, int64_t svin, T *yw, int64_t ysw, T *vout, int64_t svout,
                           int64_t B, const ModwtTaps &tp, cudaStream_t st) {
    switch (tp.F) {
#define WB_CASE(FF) case FF: return launch_group<T, FF>(fw, strict, g, vin, svin, yw, ysw, vout, svout, B, tp, st);
        WB_CASE(2) WB_CASE(4) WB_CASE(6) WB_CASE(8) WB_CASE(10) WB_CASE(12) WB_CASE(14) WB_CASE(16) WB_CASE(18) WB_CASE(20)
#undef WB_CASE
    default: set_error("internal: fused MODWT planned for filter length %d", tp.F); return false;
    }
}

static void make_modwt_taps(ModwtTaps &tp, const double *qmf, int flen) {
    tp.F = flen;
    const double r2 = std::sqrt(2.0);
    for (int m = 0; m < MAXF; ++m) { tp.h[m] = tp.g[m] = 0.0; tp.hf[m] = tp.gf[m] = 0.f; }
    for (int m = 0; m < flen; ++m) {
        tp.g[flen - 1 - m] = qmf[m] / r2;                        // scfilter = reverse(qmf)
        tp.h[m] = ((m % 2 == 0) ? qmf[m] : -qmf[m]) / r2;       // dcfilter = mirror(qmf)
    }
    for (int m = 0; m < flen; ++m) { tp.hf[m] = (float)tp.h[m]; tp.gf[m] = (float)tp.g[m]; }
}
static int maxmodwtlevels(int64_t n) { int l = 0; while (((int64_t)1 << (l + 1)) <= n) ++l; return l; }

template <typename T>
static int32_t run_modwt(T *y, const T *x, int64_t n, int64_t B, const ModwtTaps &tp, int L, bool strict, T *scratch,
                         cudaStream_t st) {
    // every step (fused group or single level) hands V on; it lands alternately in y's last column and in the scratch so
    // that the last step leaves V_L in y[:, L+1]
    const int64_t ys = n * (L + 1);
    std::vector<MStep> steps;
    plan_modwt(steps, n, L, tp.F, (int)sizeof(T), true);
    const int ns = (int)steps.size();
    dim3 grid(1, (unsigned)(B < 65535 ? B : 65535)), block(256);
    const T *src = x; int64_t ssrc = n;
    for (int i = 0; i < ns; ++i) {
        const bool to_y = ((ns - 1 - i) % 2 == 0);
        T *vdst = to_y ? (y + (int64_t)L * n) : scratch;
        const int64_t svd = to_y ? ys : n;
        if (steps[i].fused) {
            if (!launch_group_F<T>(true, strict, steps[i].g, src, ssrc, y, ys, vdst, svd, B, tp, st)) return WB200_ECUDA;
        } else {
            const int j = steps[i].level;
            const ModwtMap mp = make_map(n, j, grid.x);
            {
                LaunchScope scope("modwt_step", st);
                if (strict) k_modwt_step<T, true><<<grid, block, 0, st>>>(src, vdst, y + (int64_t)(j - 1) * n, n, ssrc, svd, ys, B, mp, tp);
                else        k_modwt_step<T, false><<<grid, block, 0, st>>>(src, vdst, y + (int64_t)(j - 1) * n, n, ssrc, svd, ys, B, mp, tp);
            }
            if (!check_launch("modwt_step")) return WB200_ECUDA;
        }
        src = vdst; ssrc = svd;
    }
    return WB200_OK;
}
template <typename T>
static int32_t run_imodwt(T *xo, const T *xw, int64_t n, int64_t B, const ModwtTaps &tp, int ncols, bool strict, T *scratch,
                          cudaStream_t st) {
    const int64_t ws = n * ncols;
    if (ncols == 1) return cudaMemcpy2DAsync(xo, n * sizeof(T), xw, ws * sizeof(T), n * sizeof(T), B, cudaMemcpyDeviceToDevice, st) == cudaSuccess ? WB200_OK : WB200_ECUDA;
    if (ncols - 1 > 62) { set_error("imodwt: %d columns", ncols); return WB200_EDIMS; }
    std::vector<MStep> steps;
    plan_modwt(steps, n, ncols - 1, tp.F, (int)sizeof(T), false);      // a plan of its own (three buffers per tile), walked backwards
    const int ns = (int)steps.size();
    dim3 grid(1, (unsigned)(B < 65535 ? B : 65535)), block(256);
    const T *v = xw + (int64_t)(ncols - 1) * n; int64_t sv = ws;
    for (int i = ns - 1; i >= 0; --i) {
        const bool to_x = (i % 2 == 0);
        T *dst = to_x ? xo : scratch;
        if (steps[i].fused) {
            if (!launch_group_F<T>(false, strict, steps[i].g, v, sv, const_cast<T *>(xw), ws, dst, n, B, tp, st)) return WB200_ECUDA;
        } else {
            const int j = steps[i].level;
            const ModwtMap mp = make_map(n, j, grid.x);
            {
                LaunchScope scope("imodwt_step", st);
                if (strict) k_imodwt_step<T, true><<<grid, block, 0, st>>>(v, xw + (int64_t)(j - 1) * n, dst, n, sv, ws, n, B, mp, tp);
                else        k_imodwt_step<T, false><<<grid, block, 0, st>>>(v, xw + (int64_t)(j - 1) * n, dst, n, sv, ws, n, B, mp, tp);
            }
            if (!check_launch("imodwt_step")) return WB200_ECUDA;
        }
        v = dst; sv = n;
    }
    return WB200_OK;
}

} // namespace wb

using namespace wb;

static int32_t modwt_common(void *out, const void *in, int64_t n, int64_t batch, const double *qmf, int32_t flen, int32_t Lc,
                            bool fw, int32_t dtype, void *workspace, size_t ws_bytes, void *stream, uint32_t flags) {
    if (dtype != WB200_F32 && dtype != WB200_F64) { set_error("modwt supports Float32/Float64"); return WB200_EDTYPE; }
    if (qmf == nullptr || flen < 2 || flen > WB200_MAX_FILTER_LEN || out == nullptr || in == nullptr) { set_error("bad filter or null pointer"); return WB200_EARG; }
    if (n < 1 || batch < 0) { set_error("n = %lld, batch = %lld", (long long)n, (long long)batch); return WB200_EDIMS; }
    if (fw) {
        if (Lc > maxmodwtlevels(n)) { set_error("Too many transform levels (length(x) < 2^L)"); return WB200_ELEVEL; }
        if (Lc < 1) { set_error("L must be >= 1"); return WB200_ELEVEL; }
    } else if (Lc < 1) { set_error("xw needs at least one column"); return WB200_EDIMS; }
    if (batch == 0) return WB200_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t esz = dtype == WB200_F64 ? 8 : 4;
    const size_t need = (size_t)n * (size_t)batch * esz;
    void *scratch = workspace;
    bool own = false;
    if (workspace == nullptr) {
        if (scratch_alloc(&scratch, need, st) != cudaSuccess) { (void)cudaGetLastError(); set_error("cudaMallocAsync(modwt scratch) failed"); return WB200_ECUDA; }
        own = true;
    } else if (ws_bytes < need) { set_error("workspace too small: %zu bytes given, %zu needed", ws_bytes, need); return WB200_EWORKSPACE; }
    ModwtTaps tp;
    make_modwt_taps(tp, qmf, flen);
    const bool strict = (flags & WB200_FLAG_STRICT_FP) != 0;
    int32_t rc;
    if (dtype == WB200_F64) rc = fw ? run_modwt<double>((double *)out, (const double *)in, n, batch, tp, Lc, strict, (double *)scratch, st)
                                    : run_imodwt<double>((double *)out, (const double *)in, n, batch, tp, Lc, strict, (double *)scratch, st);
    else                    rc = fw ? run_modwt<float>((float *)out, (const float *)in, n, batch, tp, Lc, strict, (float *)scratch, st)
                                    : run_imodwt<float>((float *)out, (const float *)in, n, batch, tp, Lc, strict, (float *)scratch, st);
    if (own) cudaFreeAsync(scratch, st);
    return rc;
}

extern "C" int32_t wb200_modwt(void *y, const void *x, int64_t n, int64_t batch, const double *qmf, int32_t flen, int32_t L,
                               int32_t dtype, void *workspace, size_t workspace_bytes, void *stream, uint32_t flags) {
    return modwt_common(y, x, n, batch, qmf, flen, L, true, dtype, workspace, workspace_bytes, stream, flags);
}
extern "C" int32_t wb200_imodwt(void *x, const void *xw, int64_t n, int64_t batch, const double *qmf, int32_t flen, int32_t ncols,
                                int32_t dtype, void *workspace, size_t workspace_bytes, void *stream, uint32_t flags) {
    return modwt_common(x, xw, n, batch, qmf, flen, ncols, false, dtype, workspace, workspace_bytes, stream, flags);
}
extern "C" int32_t wb200_maxmodwttransformlevels(int64_t n) { return maxmodwtlevels(n); }
