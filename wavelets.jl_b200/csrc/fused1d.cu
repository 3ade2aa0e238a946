// fused1d.cu -- placeholder until the fused kernels land (generic passes handle everything).
#include "fused.cuh"
namespace wb {
template <typename T>
int32_t fused_dwt(const PassOp<T> &, T *, const T *, const ArrayGeom &, int, bool, void *, size_t, cudaStream_t, uint32_t) { return -1; }
size_t fused_workspace_bytes(const ArrayGeom &, int, int, bool, bool, uint32_t) { return 0; }
template int32_t fused_dwt<float>(const PassOp<float> &, float *, const float *, const ArrayGeom &, int, bool, void *, size_t, cudaStream_t, uint32_t);
template int32_t fused_dwt<double>(const PassOp<double> &, double *, const double *, const ArrayGeom &, int, bool, void *, size_t, cudaStream_t, uint32_t);
}
