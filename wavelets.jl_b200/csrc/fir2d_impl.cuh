// fir2d_impl.cuh -- one-launch-per-level 2-D orthogonal filter-bank transform (even F <= 20) on the tensor-map TMA tile
// kernels of fused2d_tma.cuh: the reference's level is a dim-2 pass over every line plus a dim-1 pass over every line
// (src/Transforms/transforms_filter.jl:165-183), each through a strided copy; here a CTA stages a TI x TJ tile (+ the
// F-2 halo samples per side, periodic wrap patched on border tiles), runs both passes on register-resident segments
// and writes the four quadrants, so a level reads its input once and writes its output once.
// Included by fir2d_f32.cu / fir2d_f64.cu (one translation unit per element type: 40 kernel instantiations each).
#pragma once
#include "fused.cuh"
#include "tile2d_shapes.cuh"
#include "fused2d_tma.cuh"

#include <cstdlib>
#include <type_traits>

namespace wb {

// tile / segment sizes: long filters and Float64 keep the per-thread window (2 S + 2 F - 4 samples) small enough for two
// resident CTAs per SM
template <typename T, int F> struct FirTile {
    static constexpr bool SMALL = (sizeof(T) == 8) || (F >= 14);
    // Float32 filters of 8, 10 and 12 taps are built in both configurations (WB200_FIR_SMALL = 0 / 1 overrides the default)
    static constexpr bool BOTH = (sizeof(T) == 4) && (F == 8 || F == 10 || F == 12);
};
template <typename T, int F, bool FW, bool SMALL>
using FirCfg = Cfg3<T, std::conditional_t<FW, ShapeFirA<F>, ShapeFirS<F>>, 128, SMALL ? 32 : 64, SMALL ? 8 : 16, SMALL ? 8 : 16>;

template <typename T, int F, bool STRICT, bool FW, bool SMALL>
static int32_t fir_launch_cfg(const T *a, int64_t lda, int64_t bsa, const T *xd, int64_t ldx, int64_t bsx,
                              T *o1, int64_t ld1, int64_t bs1, T *o2, int64_t ld2, int64_t bs2,
                              int n, int64_t B, const FirCoefs<T, F> &fc, cudaStream_t st) {
    using S = std::conditional_t<FW, ShapeFirA<F>, ShapeFirS<F>>;
    using C3 = FirCfg<T, F, FW, SMALL>;
    const int nx = n / C3::TI, ny = n / C3::TJ;
    if constexpr (FW) {
        TensorMap tm;
        if (!make_tensor_map<T>(tm, a, n, n, B, lda, bsa, C3::PI, C3::RJ)) { set_error("fir2d: tensor map rejected"); return WB200_ECUDA; }
        auto kern = k_lift2d_fwd_tma<T, S, STRICT, C3>;
        const size_t smem = C3::SMEM_F;
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
            (void)cudaGetLastError(); set_error("cudaFuncSetAttribute(fir2d_fwd) failed"); return WB200_ECUDA;
        }
        {
            LaunchScope scope("fused_fir2d_fwd", st);
            kern<<<dim3((unsigned)nx, (unsigned)ny, (unsigned)B), C3::NT, smem, st>>>(tm, a, lda, bsa, o1, ld1, bs1, o2, ld2, bs2, n, fc);
        }
        return check_launch("fused_fir2d_fwd") ? WB200_OK : WB200_ECUDA;
    } else {
        TensorMap tml, tmx;
        const int nh = n / 2;
        if (!make_tensor_map<T>(tml, a, nh, nh, B, lda, bsa, C3::PC, C3::JQ) ||
            !make_tensor_map<T>(tmx, xd, n, n, B, ldx, bsx, C3::PC, C3::JQ)) { set_error("fir2d: tensor map rejected"); return WB200_ECUDA; }
        auto kern = k_lift2d_inv_tma<T, S, STRICT, C3>;
        const size_t smem = C3::SMEM_I;
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
            (void)cudaGetLastError(); set_error("cudaFuncSetAttribute(fir2d_inv) failed"); return WB200_ECUDA;
        }
        {
            LaunchScope scope("fused_fir2d_inv", st);
            kern<<<dim3((unsigned)nx, (unsigned)ny, (unsigned)B), C3::NT, smem, st>>>(tml, tmx, a, lda, bsa, xd, ldx, bsx, o1, ld1, bs1, n, fc);
        }
        return check_launch("fused_fir2d_inv") ? WB200_OK : WB200_ECUDA;
    }
}

template <typename T, int F, bool STRICT, bool FW>
static int32_t fir_launch_level(const T *a, int64_t lda, int64_t bsa, const T *xd, int64_t ldx, int64_t bsx,
                                T *o1, int64_t ld1, int64_t bs1, T *o2, int64_t ld2, int64_t bs2,
                                int n, int64_t B, const FirCoefs<T, F> &fc, cudaStream_t st) {
    if constexpr (FirTile<T, F>::BOTH) {
        const char *e = std::getenv("WB200_FIR_SMALL");
        // r02 interleaved A/B (tools/ab_fir2d.py, 4096^2 x 64): db4 synthesis 3835 vs 3612 GB/s on the small tile, analysis
        // 3385 vs 3542; db6 loses on the small tile in both directions
        const bool small = e ? (std::atoi(e) != 0) : ((F == 8 && !FW) ? true : FirTile<T, F>::SMALL);
        if (small) return fir_launch_cfg<T, F, STRICT, FW, true>(a, lda, bsa, xd, ldx, bsx, o1, ld1, bs1, o2, ld2, bs2, n, B, fc, st);
        return fir_launch_cfg<T, F, STRICT, FW, false>(a, lda, bsa, xd, ldx, bsx, o1, ld1, bs1, o2, ld2, bs2, n, B, fc, st);
    } else {
        return fir_launch_cfg<T, F, STRICT, FW, FirTile<T, F>::SMALL>(a, lda, bsa, xd, ldx, bsx, o1, ld1, bs1, o2, ld2, bs2, n, B, fc, st);
    }
}

// same level walk as run2d (fused2d.cu): the approximation ping-pongs through two compact scratch buffers
template <typename T, int F, bool STRICT>
static int32_t fir_run2d(const PassOp<T> &op, T *y, const T *x, const T *ll_src, int64_t ll_ld, int64_t ll_bs,
                         const ArrayGeom &g, int Lf, bool fw, void *scratch, cudaStream_t st, bool ll_to_scratch) {
    const int64_t N = g.dim[0], B = g.batch;
    const int64_t bsN = N * N;
    FirCoefs<T, F> fc;
    for (int m = 0; m < F; ++m) { fc.h[m] = op.fc.h[m]; fc.g[m] = op.fc.g[m]; }
    T *buf[2];
    buf[0] = (T *)scratch;
    const size_t b0 = (((size_t)(N / 2) * (N / 2) * B * sizeof(T)) + 255) & ~(size_t)255;
    buf[1] = (T *)((char *)scratch + b0);
    if (fw) {
        for (int l = 1; l <= Lf; ++l) {
            const int n = (int)(N >> (l - 1));
            const T *src = (l == 1) ? x : buf[(l - 2) & 1];
            const int64_t lds = (l == 1) ? N : n, bss = (l == 1) ? bsN : (int64_t)n * n;
            T *llo; int64_t ldl, bsl;
            if (l == Lf && !ll_to_scratch) { llo = y; ldl = N; bsl = bsN; }
            else         { llo = buf[(l - 1) & 1]; ldl = n / 2; bsl = (int64_t)(n / 2) * (n / 2); }
            int32_t rc = fir_launch_level<T, F, STRICT, true>(src, lds, bss, nullptr, 0, 0, llo, ldl, bsl, y, N, bsN, n, B, fc, st);
            if (rc != WB200_OK) return rc;
        }
    } else {
        for (int l = Lf; l >= 1; --l) {
            const int n = (int)(N >> (l - 1));
            const T *lls; int64_t ldl, bsl;
            if (l == Lf) { lls = ll_src; ldl = ll_ld; bsl = ll_bs; }
            else         { lls = buf[(l - 1) & 1]; ldl = n / 2; bsl = (int64_t)(n / 2) * (n / 2); }
            T *dst; int64_t ldd, bsd;
            if (l == 1) { dst = y; ldd = N; bsd = bsN; }
            else        { dst = buf[(l - 2) & 1]; ldd = n; bsd = (int64_t)n * n; }
            int32_t rc = fir_launch_level<T, F, STRICT, false>(lls, ldl, bsl, x, N, bsN, dst, ldd, bsd, nullptr, 0, 0, n, B, fc, st);
            if (rc != WB200_OK) return rc;
        }
    }
    return WB200_OK;
}

template <typename T>
int fir2d_tile_edge(int F) {      // largest tile edge of the configuration serving this filter (0: not served)
    if (F < 2 || F > 20 || (F & 1)) return 0;
    return 128;
}
template <typename T>
bool fir2d_available() { return get_encode_tiled() != nullptr; }

template <typename T>
int32_t fir2d_run(const PassOp<T> &op, T *y, const T *x, const T *ll_src, int64_t ll_ld, int64_t ll_bs,
                  const ArrayGeom &g, int Lf, bool fw, void *scratch, cudaStream_t st, bool ll_to_scratch) {
#define WB_FIR(F)                                                                                                   \
    case F: return op.strict ? fir_run2d<T, F, true>(op, y, x, ll_src, ll_ld, ll_bs, g, Lf, fw, scratch, st, ll_to_scratch) \
                             : fir_run2d<T, F, false>(op, y, x, ll_src, ll_ld, ll_bs, g, Lf, fw, scratch, st, ll_to_scratch);
    switch (op.fc.F) {
        WB_FIR(2) WB_FIR(4) WB_FIR(6) WB_FIR(8) WB_FIR(10) WB_FIR(12) WB_FIR(14) WB_FIR(16) WB_FIR(18) WB_FIR(20)
    default: break;
    }
#undef WB_FIR
    set_error("internal: fir2d_run called for an unsupported filter length %d", op.fc.F);
    return WB200_EARG;
}

template <typename T, int F>
static int32_t fir_level_f(const PassOp<T> &op, bool fw, const T *a, int64_t lda, int64_t bsa, const T *xd, int64_t ldx, int64_t bsx,
                           T *o1, int64_t ld1, int64_t bs1, T *o2, int64_t ld2, int64_t bs2, int n, int64_t B, cudaStream_t st) {
    FirCoefs<T, F> fc;
    for (int m = 0; m < F; ++m) { fc.h[m] = op.fc.h[m]; fc.g[m] = op.fc.g[m]; }
    if (fw) return op.strict ? fir_launch_level<T, F, true, true>(a, lda, bsa, xd, ldx, bsx, o1, ld1, bs1, o2, ld2, bs2, n, B, fc, st)
                             : fir_launch_level<T, F, false, true>(a, lda, bsa, xd, ldx, bsx, o1, ld1, bs1, o2, ld2, bs2, n, B, fc, st);
    return op.strict ? fir_launch_level<T, F, true, false>(a, lda, bsa, xd, ldx, bsx, o1, ld1, bs1, o2, ld2, bs2, n, B, fc, st)
                     : fir_launch_level<T, F, false, false>(a, lda, bsa, xd, ldx, bsx, o1, ld1, bs1, o2, ld2, bs2, n, B, fc, st);
}
template <typename T>
int32_t fir2d_level(const PassOp<T> &op, bool fw, const T *a, int64_t lda, int64_t bsa, const T *xd, int64_t ldx, int64_t bsx,
                    T *o1, int64_t ld1, int64_t bs1, T *o2, int64_t ld2, int64_t bs2, int n, int64_t B, cudaStream_t st) {
#define WB_FIR(F) case F: return fir_level_f<T, F>(op, fw, a, lda, bsa, xd, ldx, bsx, o1, ld1, bs1, o2, ld2, bs2, n, B, st);
    switch (op.fc.F) {
        WB_FIR(2) WB_FIR(4) WB_FIR(6) WB_FIR(8) WB_FIR(10) WB_FIR(12) WB_FIR(14) WB_FIR(16) WB_FIR(18) WB_FIR(20)
    default: break;
    }
#undef WB_FIR
    set_error("internal: fir2d_level called for an unsupported filter length %d", op.fc.F);
    return WB200_EARG;
}

} // namespace wb
