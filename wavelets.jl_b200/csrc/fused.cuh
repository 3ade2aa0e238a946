// fused.cuh -- entry points of the fused sm_100a kernels (fused1d.cu, fused2d.cu).
#pragma once
#include "common.cuh"

namespace wb {

// Try to run the whole multi-level transform with the fused kernels.
// Returns a wb200_status (>= 0) when the call was handled, or -1 when the shape / wavelet is not covered
// and the generic per-pass drivers must take it.
template <typename T>
int32_t fused_dwt(const PassOp<T> &op, T *y, const T *x, const ArrayGeom &g, int L, bool fw,
                  void *workspace, size_t ws_bytes, cudaStream_t st, uint32_t flags);

// Device scratch the fused path needs for this shape (0 when it does not apply).
size_t fused_workspace_bytes(const ArrayGeom &g, int esize, int L, bool lifting, bool inplace, uint32_t flags);

// ---- fused multi-level 1-D lifting for batches of contiguous columns (lift1d.cu); same return contract as fused_dwt ----
template <typename T>
int32_t fused_lift1d(const PassOp<T> &op, T *y, const T *x, const ArrayGeom &g, int L, bool fw,
                     void *workspace, size_t ws_bytes, cudaStream_t st);

// ---- 2-D lifting levels (fused2d.cu) ----------------------------------------------------------------------
// Number of leading levels (1..Lf) the fused 2-D lifting kernels take for this call (0: none).
template <typename T> int fused2d_levels(const PassOp<T> &op, const ArrayGeom &g, int L, bool fw);
// Scratch for the approximation ping-pong of Lf fused levels.
template <typename T> size_t fused2d_scratch_bytes(const ArrayGeom &g, int Lf);
// forward: levels 1..Lf of x; detail quadrants land in y, the level-Lf approximation in y's leading corner.
// inverse: levels Lf..1; the level-Lf approximation is read from ll_src (leading dimension ll_ld, batch stride
// ll_bs: x's corner when no coarser level was inverted before, else where the generic remainder left it), details
// from x; the result fills y.   x must not alias y; ll_src must not alias y when Lf == 1.
template <typename T>
int32_t fused2d_run(const PassOp<T> &op, T *y, const T *x, const T *ll_src, int64_t ll_ld, int64_t ll_bs,
                    const ArrayGeom &g, int Lf, bool fw, void *scratch, cudaStream_t st, bool ll_to_scratch = false);
// (forward, ll_to_scratch: the level-Lf approximation stays in the scratch ping-pong buffer (Lf-1)&1, compact,
//  for the pyramid-tail kernel instead of going to y's corner)

// Pyramid tail: all `levels` remaining levels of an nt x nt (nt <= 128) approximation per image in one launch.
// forward: src (plain nt x nt) -> dst = the corner's Mallat pyramid; inverse: src = pyramid -> dst plain nt x nt.
template <typename T> bool fused2d_tail_ok(const PassOp<T> &op, const ArrayGeom &g, int64_t nt, int levels);
template <typename T>
int32_t fused2d_tail(const PassOp<T> &op, const T *src, int64_t ld_s, int64_t bs_s, T *dst, int64_t ld_d, int64_t bs_d,
                     int nt, int levels, int64_t B, bool fw, cudaStream_t st);

// ---- 2-D filter-bank levels (fir2d_f32.cu / fir2d_f64.cu): the same contract as fused2d_run, for OrthoFilter calls ----
template <typename T> int fir2d_tile_edge(int F);
template <typename T> bool fir2d_available();
template <typename T>
int32_t fir2d_run(const PassOp<T> &op, T *y, const T *x, const T *ll_src, int64_t ll_ld, int64_t ll_bs,
                  const ArrayGeom &g, int Lf, bool fw, void *scratch, cudaStream_t st, bool ll_to_scratch);
// One level on B images of n x n (the fused kernels' own pointer contract).  forward: a = source (lda, bsa) -> o1 = the
// (approx, approx) quadrant, o2 = the array whose other three quadrants are written.  inverse: a = (approx, approx)
// quadrant source, xd = the array holding the other three quadrants -> o1 = merged n x n output.
template <typename T>
int32_t fir2d_level(const PassOp<T> &op, bool fw, const T *a, int64_t lda, int64_t bsa, const T *xd, int64_t ldx, int64_t bsx,
                    T *o1, int64_t ld1, int64_t bs1, T *o2, int64_t ld2, int64_t bs2, int n, int64_t B, cudaStream_t st);

// ---- 3-D lifting levels in two passes (lift3d.cu): a register walk along dim 3 + the 2-D lifting level kernel on the planes ----
template <typename T> int lift3d_levels(const PassOp<T> &op, const ArrayGeom &g, int L, bool fw);
template <typename T>
int32_t lift3d_run(const PassOp<T> &op, T *y, const T *x, const ArrayGeom &g, int L, int Lf, bool fw, T *W, cudaStream_t st);
// one 2-D lifting level on B images of n x n with the fused tile kernels (fused2d.cu); pointer contract of fir2d_level
template <typename T>
int32_t lift2d_level(const PassOp<T> &op, bool fw, const T *a, int64_t lda, int64_t bsa, const T *xd, int64_t ldx, int64_t bsx,
                     T *o1, int64_t ld1, int64_t bs1, T *o2, int64_t ld2, int64_t bs2, int n, int64_t B, cudaStream_t st);

// ---- one-pass 3-D filter-bank levels (fir3d_f32.cu / fir3d_f64.cu, fir3d_impl.cuh) ----
// Number of leading levels the marching kernels take (every dimension of the level's corner whole tiles), their LLL ping-pong
// scratch, and the level walk (same contract as fused2d_run: forward leaves the level-Lf approximation in y's corner; inverse
// reads it from ll_src with the given row / plane / batch strides -- never y itself when Lf == 1).
template <typename T> int fir3d_levels(const PassOp<T> &op, const ArrayGeom &g, int L, bool fw);
template <typename T> size_t fir3d_scratch_bytes(const ArrayGeom &g, int Lf);
template <typename T>
int32_t fir3d_run(const PassOp<T> &op, T *y, const T *x, const T *ll_src, int64_t ll_ld, int64_t ll_ps, int64_t ll_bs,
                  const ArrayGeom &g, int Lf, bool fw, void *scratch, cudaStream_t st);

} // namespace wb
