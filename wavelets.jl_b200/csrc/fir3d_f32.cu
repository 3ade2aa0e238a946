// fir3d_f32.cu -- float instantiations of the one-pass 3-D filter-bank level kernels (fir3d_impl.cuh)
#include "fir3d_impl.cuh"
namespace wb {
template int fir3d_levels<float>(const PassOp<float> &, const ArrayGeom &, int, bool);
template size_t fir3d_scratch_bytes<float>(const ArrayGeom &, int);
template int32_t fir3d_run<float>(const PassOp<float> &, float *, const float *, const float *, int64_t, int64_t, int64_t, const ArrayGeom &, int, bool, void *, cudaStream_t);
} // namespace wb
