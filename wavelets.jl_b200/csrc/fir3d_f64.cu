// fir3d_f64.cu -- double instantiations of the one-pass 3-D filter-bank level kernels (fir3d_impl.cuh)
#include "fir3d_impl.cuh"
namespace wb {
template int fir3d_levels<double>(const PassOp<double> &, const ArrayGeom &, int, bool);
template size_t fir3d_scratch_bytes<double>(const ArrayGeom &, int);
template int32_t fir3d_run<double>(const PassOp<double> &, double *, const double *, const double *, int64_t, int64_t, int64_t, const ArrayGeom &, int, bool, void *, cudaStream_t);
} // namespace wb
