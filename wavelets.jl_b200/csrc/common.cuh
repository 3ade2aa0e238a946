// common.cuh -- shared device/host definitions for libwavelets_b200 (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include "../../include/wavelets_b200.h"

namespace wb {

constexpr int MAXF = WB200_MAX_FILTER_LEN;
constexpr int MAXSTEPS = WB200_MAX_LIFT_STEPS;
constexpr int MAXCOEF = WB200_MAX_LIFT_COEF;

// ---------------------------------------------------------------------------------------------------
// Floating-point policy.
//   STRICT = true : every product and every sum is rounded separately (__fmul_rn/__fadd_rn are never
//                   contracted by nvcc), in the reference's operation order -> bit-identical to the
//                   Julia CPU path (which never fuses a*b+c).
//   STRICT = false: same order, a*b+c contracted to one FMA (one rounding less per tap).  One more liberty in the
//                   synthesis kernels: the reference adds the approximation-band sum and the detail-band sum of an output
//                   last; fast mode lets the detail terms continue the approximation's accumulator (one chain: no second
//                   FMUL, no final FADD -- 32 instead of 36 floating-point instructions per two db4 output pairs).
// ---------------------------------------------------------------------------------------------------
template <bool STRICT> struct FP {
    static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
    static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
    static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
    static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
    // acc + a*b
    static __device__ __forceinline__ float mac(float acc, float a, float b) {
        if constexpr (STRICT) return __fadd_rn(acc, __fmul_rn(a, b));
        else return __fmaf_rn(a, b, acc);
    }
    static __device__ __forceinline__ double mac(double acc, double a, double b) {
        if constexpr (STRICT) return __dadd_rn(acc, __dmul_rn(a, b));
        else return __fma_rn(a, b, acc);
    }
};

// A strided view of equal-length lines inside an array.  Element (line coords c0..c3, position k) lives at
//   p[c0*s[0] + c1*s[1] + c2*s[2] + c3*s[3] + k*ls].
template <typename T> struct View {
    T *p;
    int64_t ls;
    int64_t s[4];
    __host__ __device__ __forceinline__ T *line(int64_t c0, int64_t c1, int64_t c2, int64_t c3) const {
        return p + c0 * s[0] + c1 * s[1] + c2 * s[2] + c3 * s[3];
    }
};

// Extents of the line set: `len` samples per line, n[0..3] lines along each outer coordinate.
// Coordinate 0 is the one that is contiguous in memory when the lines themselves are strided.
struct Extent {
    int64_t len;
    int64_t n[4];
};

// filter taps rounded to T (makereverseqmfpair): h[m] and g[m] = (-1)^m h[m]
template <typename T> struct FilterCoefs {
    T h[MAXF];
    T g[MAXF];
    int F;
};

// lifting scheme after makescheme(T, scheme, fw): steps already ordered / signed for the direction.
template <typename T> struct LiftScheme {
    int nsteps;
    int is_predict[MAXSTEPS];
    int shift[MAXSTEPS];
    int nc[MAXSTEPS];
    T coef[MAXSTEPS][MAXCOEF];
    T norm1, norm2;
    int halo_l, halo_r; // pairs of halo needed on each side for a windowed tile (sum of per-step reaches)
};

// ---------------------------------------------------------------------------------------------------
// error plumbing (thread-local detail string + launch counter), defined in api.cu
// ---------------------------------------------------------------------------------------------------
void set_error(const char *fmt, ...);
cudaError_t scratch_alloc(void **p, size_t bytes, cudaStream_t st);   // library-private stream-ordered pool (api.cu); free with cudaFreeAsync
bool check_launch(const char *what); // cudaGetLastError after a launch; records the error
// Every kernel launch sits inside a LaunchScope: it bumps the per-thread launch counter and, when profiling is
// enabled (wb200_profile_enable), brackets the launch with CUDA events on the launching stream so bench.py can
// read per-kernel device times from inside its timed region (wb200_profile_collect).
struct LaunchScope {
    const char *name;
    cudaStream_t st;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    LaunchScope(const char *name, cudaStream_t st);
    ~LaunchScope();
};

// threshold!(x, TH, t) folded into the LOADS of an inverse transform ("the threshold is an elementwise
// epilogue that can ride in the idwt load"): kind < 0 = none; t = sigma_dev ? *sigma_dev * tfac : t_host.  Honoured by the
// fused 1-D filter synthesis kernels only (idwt_filter_thresholded, api.cu, reports when it could not take a call).
struct ThreshEpi {
    int kind = -1;
    double t_host = 0.0, tfac = 1.0;
    const double *sigma_dev = nullptr;
};

// inverse 1-D filter-bank transform with the threshold applied to every coefficient as it is staged (api.cu): WB200_OK when
// the fused kernels took the call, -1 when the shape is not theirs (the caller then thresholds in its own pass)
int32_t idwt_filter_thresholded(void *y, const void *x, int64_t n, int64_t batch, const double *qmf, int32_t flen, int32_t L,
                                int32_t dtype, const ThreshEpi &epi, cudaStream_t st, uint32_t flags);

template <typename T> constexpr int dtype_of();
template <> constexpr int dtype_of<float>() { return WB200_F32; }
template <> constexpr int dtype_of<double>() { return WB200_F64; }

// ---------------------------------------------------------------------------------------------------
// generic (any length / filter / stride) one-level passes: generic_kernels.cu
// `active` (optional, device): one byte per value of line coordinate 1; lines with active[c1] == 0 are
// carried through unchanged (wavelet-packet leaves).
// ---------------------------------------------------------------------------------------------------
template <typename T>
bool launch_filter_analysis(const View<const T> &src, const View<T> &dlo, const View<T> &dhi,
                            const Extent &e, const FilterCoefs<T> &fc, bool strict, cudaStream_t st,
                            const uint8_t *active = nullptr);
// slo/shi: approximation / detail halves; lines whose coordinates are all below thr[] read the
// approximation half from `salt` instead (the LL... corner produced by the previous inverse level).
template <typename T>
bool launch_filter_synthesis(const View<const T> &slo, const View<const T> &shi, const View<const T> &salt,
                             const int64_t thr[4], bool has_alt, const View<T> &dst,
                             const Extent &e, const FilterCoefs<T> &fc, bool strict, cudaStream_t st,
                             const uint8_t *active = nullptr);
template <typename T>
bool launch_lifting_analysis(const View<const T> &src, const View<T> &dlo, const View<T> &dhi,
                             const Extent &e, const LiftScheme<T> &sc, bool strict, cudaStream_t st,
                             const uint8_t *active = nullptr);
template <typename T>
bool launch_lifting_synthesis(const View<const T> &slo, const View<const T> &shi, const View<const T> &salt,
                              const int64_t thr[4], bool has_alt, const View<T> &dst,
                              const Extent &e, const LiftScheme<T> &sc, bool strict, cudaStream_t st,
                              const uint8_t *active = nullptr);
// plain strided copy of whole lines (WPT leaves, L == 0)
template <typename T>
bool launch_copy_lines(const View<const T> &src, const View<T> &dst, const Extent &e, cudaStream_t st);

// ---------------------------------------------------------------------------------------------------
// one-level pass abstraction shared by the filter and lifting drivers
// ---------------------------------------------------------------------------------------------------
// fast single-level filter passes (fastpass.cu): 1 = handled, 0 = layout not covered (use the generic pass), -1 = error
template <typename T>
int fast_filter_analysis(const View<const T> &src, const View<T> &dlo, const View<T> &dhi, const Extent &e,
                         const FilterCoefs<T> &fc, bool strict, cudaStream_t st);
template <typename T>
int fast_filter_synthesis(const View<const T> &slo, const View<const T> &shi, const View<const T> &salt,
                          const int64_t thr[4], bool has_alt, const View<T> &dst, const Extent &e,
                          const FilterCoefs<T> &fc, bool strict, cudaStream_t st);

// whole packet subtrees in shared memory (fastpass.cu): every node of m samples goes through `levels` full levels
template <typename T>
int fast_wpt_subtree(const T *S, T *D, int64_t n, int64_t m, int levels, int64_t nodes, int64_t B,
                     const FilterCoefs<T> &fc, bool strict, bool fw, cudaStream_t st);

int wpt_subtree_max_samples(int esize);   // largest packet node the on-chip subtree kernels take (fastpass.cu)

// K consecutive full packet levels of large nodes in one launch (wptfused.cu): `nodes` nodes of nj samples per signal
template <typename T>
int fast_wpt_fused_levels(const T *S, T *D, int64_t n, int64_t nj, int K, int64_t nodes, int64_t B,
                          const FilterCoefs<T> &fc, bool strict, bool fw, cudaStream_t st);
int wpt_fused_max_levels(int F);          // most levels one such launch fuses for an F-tap filter (0: not covered)
bool wpt_fused_ok(int esize, int F, int64_t nj, int K);   // the kernels' own acceptance test, for planning

template <typename T> struct PassOp {
    ThreshEpi epi;
    bool lifting;
    bool strict;
    cudaStream_t st;
    FilterCoefs<T> fc;
    LiftScheme<T> sc;
    bool generic_only = false;
    bool analysis(const View<const T> &src, const View<T> &dlo, const View<T> &dhi, const Extent &e,
                  const uint8_t *active = nullptr) const {
        if (!lifting && !generic_only && active == nullptr) {
            const int r = fast_filter_analysis<T>(src, dlo, dhi, e, fc, strict, st);
            if (r != 0) return r > 0;
        }
        return lifting ? launch_lifting_analysis<T>(src, dlo, dhi, e, sc, strict, st, active)
                       : launch_filter_analysis<T>(src, dlo, dhi, e, fc, strict, st, active);
    }
    bool synthesis(const View<const T> &slo, const View<const T> &shi, const View<const T> &salt,
                   const int64_t thr[4], bool has_alt, const View<T> &dst, const Extent &e,
                   const uint8_t *active = nullptr) const {
        if (!lifting && !generic_only && active == nullptr) {
            const int r = fast_filter_synthesis<T>(slo, shi, salt, thr, has_alt, dst, e, fc, strict, st);
            if (r != 0) return r > 0;
        }
        return lifting ? launch_lifting_synthesis<T>(slo, shi, salt, thr, has_alt, dst, e, sc, strict, st, active)
                       : launch_filter_synthesis<T>(slo, shi, salt, thr, has_alt, dst, e, fc, strict, st, active);
    }
};

// Geometry of a column-major array (C, d1, d2, d3) x batch; C = 2 for complex (interleaved re/im), else 1.
struct ArrayGeom {
    int64_t C;
    int64_t dim[3];    // d1, d2, d3 (unused trailing dims = 1)
    int64_t batch;
    int ndim;
    int64_t stride(int ax) const { // element stride of array axis ax (0 = component, 1..3 = dims)
        int64_t s = 1;
        if (ax >= 1) s *= C;
        for (int a = 1; a < ax; ++a) s *= dim[a - 1];
        return s;
    }
    int64_t slice() const { return C * dim[0] * dim[1] * dim[2]; }
    int64_t total() const { return slice() * batch; }
};


} // namespace wb
