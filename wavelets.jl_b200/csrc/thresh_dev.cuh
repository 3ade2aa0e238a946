// thresh_dev.cuh -- threshold!(x, TH, t) per element (src/Threshold/threshold_main.jl:35-117), shared by the elementwise
// kernel (threshold.cu) and by the synthesis kernels that apply it as a load epilogue (fused1d.cu): the comparisons and
// products with t are done in double and rounded to T on the store, as Julia's promotion does (t = sigma * dnt.t is Float64).
#pragma once
#include "common.cuh"

namespace wb {

template <typename T> __device__ __forceinline__ double sgn_of(T v) { return (v > 0) ? 1.0 : ((v < 0) ? -1.0 : (double)v); }

template <typename T> __device__ __forceinline__ T thresh_apply(T v, int kind, double t) {
    T o = v;
    switch (kind) {
    case WB200_TH_HARD: if (fabs((double)v) <= t) o = 0; break;
    case WB200_TH_SOFT: { const double sh = __dsub_rn(fabs((double)v), t); o = (sh < 0) ? (T)0 : (T)__dmul_rn(sgn_of(v), sh); } break;
    case WB200_TH_SEMISOFT:
        if ((double)v <= __dmul_rn(2.0, t)) {          // (sic) x[i], not abs(x[i])
            const double sh = __dsub_rn(fabs((double)v), t);
            if (sh < 0) o = 0;
            else if (__dsub_rn(sh, t) < 0) o = (T)__dmul_rn(__dmul_rn(sgn_of(v), sh), 2.0);
        }
        break;
    case WB200_TH_STEIN: {
        T vv;
        if constexpr (sizeof(T) == 4) vv = __fmul_rn(v, v); else vv = __dmul_rn(v, v);
        const double sh = __dsub_rn(1.0, __ddiv_rn(__dmul_rn(t, t), (double)vv));
        o = (sh < 0) ? (T)0 : (T)__dmul_rn((double)v, sh);
    } break;
    case WB200_TH_NEG: if (v < 0) o = 0; break;
    case WB200_TH_POS: if (v > 0) o = 0; break;
    default: break;
    }
    return o;
}

} // namespace wb
