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
    // V_j lands alternately in y's last column and in the scratch so that V_L ends in y[:, L+1]
    const int64_t ys = n * (L + 1);
    dim3 grid(1, (unsigned)(B < 65535 ? B : 65535)), block(256);
    const T *src = x; int64_t ssrc = n;
    for (int j = 1; j <= L; ++j) {
        const ModwtMap mp = make_map(n, j, grid.x);
        const bool to_y = ((L - j) % 2 == 0);
        T *vdst = to_y ? (y + (int64_t)L * n) : scratch;
        const int64_t svd = to_y ? ys : n;
        {
            LaunchScope scope("modwt_step", st);
            if (strict) k_modwt_step<T, true><<<grid, block, 0, st>>>(src, vdst, y + (int64_t)(j - 1) * n, n, ssrc, svd, ys, B, mp, tp);
            else        k_modwt_step<T, false><<<grid, block, 0, st>>>(src, vdst, y + (int64_t)(j - 1) * n, n, ssrc, svd, ys, B, mp, tp);
        }
        if (!check_launch("modwt_step")) return WB200_ECUDA;
        src = vdst; ssrc = svd;
    }
    return WB200_OK;
}
template <typename T>
static int32_t run_imodwt(T *xo, const T *xw, int64_t n, int64_t B, const ModwtTaps &tp, int ncols, bool strict, T *scratch,
                          cudaStream_t st) {
    const int64_t ws = n * ncols;
    if (ncols == 1) return cudaMemcpy2DAsync(xo, n * sizeof(T), xw, ws * sizeof(T), n * sizeof(T), B, cudaMemcpyDeviceToDevice, st) == cudaSuccess ? WB200_OK : WB200_ECUDA;
    dim3 grid(1, (unsigned)(B < 65535 ? B : 65535)), block(256);
    const T *v = xw + (int64_t)(ncols - 1) * n; int64_t sv = ws;
    for (int j = ncols - 1; j >= 1; --j) {
        if (j > 62) { set_error("imodwt: %d columns", ncols); return WB200_EDIMS; }
        const ModwtMap mp = make_map(n, j, grid.x);
        const bool to_x = ((j - 1) % 2 == 0);
        T *dst = to_x ? xo : scratch;
        {
            LaunchScope scope("imodwt_step", st);
            if (strict) k_imodwt_step<T, true><<<grid, block, 0, st>>>(v, xw + (int64_t)(j - 1) * n, dst, n, sv, ws, n, B, mp, tp);
            else        k_imodwt_step<T, false><<<grid, block, 0, st>>>(v, xw + (int64_t)(j - 1) * n, dst, n, sv, ws, n, B, mp, tp);
        }
        if (!check_launch("imodwt_step")) return WB200_ECUDA;
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
        keep_pool_memory();
        if (cudaMallocAsync(&scratch, need, st) != cudaSuccess) { (void)cudaGetLastError(); set_error("cudaMallocAsync(modwt scratch) failed"); return WB200_ECUDA; }
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
