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

} // namespace wb
