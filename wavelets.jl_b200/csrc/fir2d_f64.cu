// fir2d_f64.cu -- double instantiations of the fused 2-D filter-bank level kernels (fir2d_impl.cuh)
#include "fir2d_impl.cuh"
namespace wb {
template int fir2d_tile_edge<double>(int);
template bool fir2d_available<double>();
template int32_t fir2d_run<double>(const PassOp<double> &, double *, const double *, const double *, int64_t, int64_t, const ArrayGeom &, int, bool, void *, cudaStream_t, bool);
template int32_t fir2d_level<double>(const PassOp<double> &, bool, const double *, int64_t, int64_t, const double *, int64_t, int64_t, double *, int64_t, int64_t, double *, int64_t, int64_t, int, int64_t, cudaStream_t);
} // namespace wb
