// fir2d_f32.cu -- float instantiations of the fused 2-D filter-bank level kernels (fir2d_impl.cuh)
#include "fir2d_impl.cuh"
namespace wb {
template int fir2d_tile_edge<float>(int);
template bool fir2d_available<float>();
template int32_t fir2d_run<float>(const PassOp<float> &, float *, const float *, const float *, int64_t, int64_t, const ArrayGeom &, int, bool, void *, cudaStream_t, bool);
template int32_t fir2d_level<float>(const PassOp<float> &, bool, const float *, int64_t, int64_t, const float *, int64_t, int64_t, float *, int64_t, int64_t, float *, int64_t, int64_t, int, int64_t, cudaStream_t);
} // namespace wb
