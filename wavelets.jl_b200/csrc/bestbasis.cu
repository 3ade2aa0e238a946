// bestbasis.cu -- SURVEY 8(f) row 3: coefentropy and bestbasistree (src/Threshold/entropy.jl:16-129) on the device.
// The full packet tree is walked level by level with the batched one-level transform (the nodes of a level are equal-length
// arrays back to back: `batch` = number of nodes), and a reduction kernel takes every node's additive entropy before it is
// split.  Only the 2^Lmax - 1 + 2^(Lmax-1) entropies cross to the host, where the reference's bottom-up comparison builds the
// tree.  Per-coefficient terms are computed in T as upstream (s = (x / nrm)^2, -s log s or -log s); the sums are accumulated
// in double in a fixed order (upstream: sequentially in T), so entropies agree to rounding, not bit for bit.
#include "common.cuh"
#include <cmath>
#include <vector>

namespace wb {

template <typename T> __device__ __forceinline__ double ent_term(T x, int et, T nrm) {
    T q, s;
    if constexpr (sizeof(T) == 4) { q = __fdiv_rn(x, nrm); s = __fmul_rn(q, q); } else { q = __ddiv_rn(x, nrm); s = __dmul_rn(q, q); }
    if (s == 0) return -0.0;
    if constexpr (sizeof(T) == 4) return (double)((et == 0) ? -s * logf(s) : -logf(s));
    else return (et == 0) ? -s * log(s) : -log(s);
}
// fixed-order block reduction of one double per thread (256 threads)
__device__ __forceinline__ double block_sum(double v, double *sh) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0;
    if (threadIdx.x == 0) for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sh[w];
    __syncthreads();
    return t;       // valid in thread 0
}
// partial sums of squares: block b covers a fixed slice; part[b]
template <typename T>
__global__ void __launch_bounds__(256) k_sumsq(const T *__restrict__ x, int64_t n, double *__restrict__ part) {
    __shared__ double sh[8];
    const int64_t per = (n + gridDim.x - 1) / gridDim.x, lo = (int64_t)blockIdx.x * per, hi = lo + per < n ? lo + per : n;
    double acc = 0;
    for (int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) acc += (double)x[i] * (double)x[i];
    const double t = block_sum(acc, sh);
    if (threadIdx.x == 0) part[blockIdx.x] = t;
}
template <typename T>
__global__ void k_norm_finish(const double *__restrict__ part, int np, T *__restrict__ nrm_out, double nrm_given) {
    if (threadIdx.x == 0) {
        if (nrm_given == nrm_given) { *nrm_out = (T)nrm_given; return; }
        double t = 0;
        for (int i = 0; i < np; ++i) t += part[i];
        *nrm_out = (T)sqrt(t);
    }
}
// one CTA per node: out[node] = sum of the node's entropy terms
template <typename T>
__global__ void __launch_bounds__(256) k_node_entropy(const T *__restrict__ x, int64_t nj, int et, const T *__restrict__ nrm_p, double *__restrict__ out) {
    __shared__ double sh[8];
    const T nrm = *nrm_p;
    const T *xn = x + (int64_t)blockIdx.x * nj;
    double acc = 0;
    if (nrm != 0)
        for (int64_t i = threadIdx.x; i < nj; i += blockDim.x) acc += ent_term<T>(xn[i], et, nrm);
    const double t = block_sum(acc, sh);
    if (threadIdx.x == 0) out[blockIdx.x] = t;
}

} // namespace wb

using namespace wb;

namespace {
constexpr int NPART = 256;

template <typename T>
int32_t entropies(std::vector<double> &bf, std::vector<double> &af, const T *y, int64_t n, int Lmax, int32_t wkind, const double *qmf,
                  int32_t flen, const wb200_lift_step *steps, int32_t nsteps, double norm1, double norm2, int32_t et, int32_t dtype,
                  cudaStream_t st, uint32_t flags) {
    const int64_t ntree = ((int64_t)1 << Lmax) - 1, n_af = (int64_t)1 << (Lmax - 1);
    const size_t arr = (((size_t)n * sizeof(T)) + 255) & ~(size_t)255;
    const size_t ent_bytes = (size_t)(ntree + n_af) * sizeof(double);
    char *pool = nullptr;
    if (scratch_alloc((void **)&pool, 2 * arr + ent_bytes + NPART * sizeof(double) + 256, st) != cudaSuccess) {
        (void)cudaGetLastError(); set_error("cudaMallocAsync(bestbasis scratch) failed"); return WB200_ECUDA;
    }
    T *xa = (T *)pool, *xb = (T *)(pool + arr);
    double *ent = (double *)(pool + 2 * arr), *part = ent + ntree + n_af;
    T *nrm = (T *)(part + NPART);
    int32_t rc = WB200_OK;
    if (cudaMemcpyAsync(xa, y, sizeof(T) * (size_t)n, cudaMemcpyDeviceToDevice, st) != cudaSuccess) { (void)cudaGetLastError(); rc = WB200_ECUDA; }
    if (rc == WB200_OK) {
        { LaunchScope scope("sumsq", st); k_sumsq<T><<<NPART, 256, 0, st>>>(y, n, part); }
        { LaunchScope scope("norm_finish", st); k_norm_finish<T><<<1, 32, 0, st>>>(part, NPART, nrm, std::nan("")); }
    }
    int64_t k = 0;
    for (int lv = 0; lv < Lmax && rc == WB200_OK; ++lv) {
        const int64_t nodes = (int64_t)1 << lv, nj = n >> lv;
        { LaunchScope scope("node_entropy", st); k_node_entropy<T><<<(unsigned)nodes, 256, 0, st>>>(xa, nj, et, nrm, ent + k); }
        k += nodes;
        const int64_t d1[3] = {nj, 1, 1};
        if (wkind == 1) rc = wb200_dwt_filter(xb, xa, 1, d1, nodes, qmf, flen, 1, 1, dtype, nullptr, 0, (void *)st, flags);
        else            rc = wb200_dwt_lifting(xb, xa, 1, d1, nodes, steps, nsteps, norm1, norm2, 1, 1, dtype, nullptr, 0, (void *)st, flags);
        T *t = xa; xa = xb; xb = t;
    }
    if (rc == WB200_OK) {
        LaunchScope scope("node_entropy", st);
        k_node_entropy<T><<<(unsigned)n_af, 256, 0, st>>>(xa, n / n_af, et, nrm, ent + ntree);
    }
    if (rc == WB200_OK && !check_launch("bestbasis")) rc = WB200_ECUDA;
    if (rc == WB200_OK) {
        std::vector<double> h((size_t)(ntree + n_af));
        if (cudaMemcpyAsync(h.data(), ent, ent_bytes, cudaMemcpyDeviceToHost, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess) {
            (void)cudaGetLastError(); set_error("bestbasistree: copy of the entropies failed"); rc = WB200_ECUDA;
        } else {
            bf.assign(h.begin(), h.begin() + ntree);
            af.assign(h.begin() + ntree, h.end());
        }
    }
    cudaFreeAsync(pool, st);
    return rc;
}

double bestsub(const std::vector<double> &bf, const std::vector<double> &af, int64_t i) {      // bestsubtree_entropy, 1-based
    const int64_t nbf = (int64_t)bf.size(), naf = (int64_t)af.size();
    double sum;
    if (nbf < (i << 1)) sum = af[(size_t)(i - naf)];
    else sum = bestsub(bf, af, i << 1) + bestsub(bf, af, (i << 1) + 1);
    return bf[(size_t)(i - 1)] < sum ? bf[(size_t)(i - 1)] : sum;
}
} // namespace

extern "C" int32_t wb200_coefentropy(double *out, const void *x, int64_t count, int32_t et, double nrm, int32_t dtype, void *stream) {
    if (dtype != WB200_F32 && dtype != WB200_F64) { set_error("coefentropy supports Float32/Float64"); return WB200_EDTYPE; }
    if (out == nullptr || count < 0 || (count > 0 && x == nullptr) || et < 0 || et > 1) { set_error("bad argument"); return WB200_EARG; }
    if (nrm == nrm && nrm < 0) { set_error("nrm must be >= 0"); return WB200_EARG; }                  // @assert nrm >= 0
    *out = 0.0;
    if (count == 0) return WB200_OK;
    cudaStream_t st = (cudaStream_t)stream;
    char *pool = nullptr;
    if (scratch_alloc((void **)&pool, (NPART + 2) * sizeof(double) + 64, st) != cudaSuccess) { (void)cudaGetLastError(); set_error("cudaMallocAsync failed"); return WB200_ECUDA; }
    double *part = (double *)pool, *res = part + NPART;
    void *nrmp = (void *)(res + 1);
    {
        LaunchScope scope("coefentropy", st);
        if (dtype == WB200_F64) {
            k_sumsq<double><<<NPART, 256, 0, st>>>((const double *)x, count, part);
            k_norm_finish<double><<<1, 32, 0, st>>>(part, NPART, (double *)nrmp, nrm);
            k_node_entropy<double><<<1, 256, 0, st>>>((const double *)x, count, et, (const double *)nrmp, res);
        } else {
            k_sumsq<float><<<NPART, 256, 0, st>>>((const float *)x, count, part);
            k_norm_finish<float><<<1, 32, 0, st>>>(part, NPART, (float *)nrmp, nrm);
            k_node_entropy<float><<<1, 256, 0, st>>>((const float *)x, count, et, (const float *)nrmp, res);
        }
    }
    int32_t rc = check_launch("coefentropy") ? WB200_OK : WB200_ECUDA;
    if (rc == WB200_OK && (cudaMemcpyAsync(out, res, sizeof(double), cudaMemcpyDeviceToHost, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess)) {
        (void)cudaGetLastError(); rc = WB200_ECUDA;
    }
    cudaFreeAsync(pool, st);
    return rc;
}

extern "C" int32_t wb200_bestbasistree(uint8_t *besttree, double *entr_bf, double *entr_af, const void *y, int64_t n, int32_t wkind,
                                       const double *qmf, int32_t flen, const wb200_lift_step *steps, int32_t nsteps, double norm1,
                                       double norm2, const uint8_t *tree, int64_t ntree, int32_t et, int32_t dtype, void *stream,
                                       uint32_t flags) {
    if (dtype != WB200_F32 && dtype != WB200_F64) { set_error("bestbasistree supports Float32/Float64"); return WB200_EDTYPE; }
    if (besttree == nullptr || y == nullptr || tree == nullptr || et < 0 || et > 1) { set_error("bad argument"); return WB200_EARG; }
    if ((wkind == 1 && (qmf == nullptr || flen < 2)) || (wkind == 2 && (steps == nullptr || nsteps < 1)) || (wkind != 1 && wkind != 2)) {
        set_error("bad wavelet description"); return WB200_EARG;
    }
    if (n < 2 || !wb200_isvalidtree(n, tree, ntree)) { set_error("invalid tree"); return WB200_ETREE; }
    const int Lmax = wb200_maxtransformlevels(n);
    std::vector<double> bf, af;
    cudaStream_t st = (cudaStream_t)stream;
    int32_t rc;
    if (dtype == WB200_F64) rc = entropies<double>(bf, af, (const double *)y, n, Lmax, wkind, qmf, flen, steps, nsteps, norm1, norm2, et, dtype, st, flags);
    else                    rc = entropies<float>(bf, af, (const float *)y, n, Lmax, wkind, qmf, flen, steps, nsteps, norm1, norm2, et, dtype, st, flags);
    if (rc != WB200_OK) return rc;
    const double tie_rel = (dtype == WB200_F64) ? 1e-13 : 1e-6;
    for (int64_t i = 1; i <= ntree; ++i) {                  // entropy.jl:93-105
        if ((i > 1 && !besttree[(i >> 1) - 1]) || !tree[i - 1]) besttree[i - 1] = 0;
        else {
            // Upstream decides `entr_bf[i] <= bestsubtree_entropy(...)` on sums accumulated sequentially in T; here the sums
            // are block-reduced in double, so two entropies that are EQUAL upstream (exact ties: a node and its children
            // carrying the same cost, e.g. all-zero nodes or Haar on piecewise-constant data) may differ here in the last
            // bits.  Ties are resolved the way upstream's `<=` resolves them -- the node is NOT split -- by comparing with
            // a relative slack of a few rounding errors of T-precision sums.
            const double a = bf[(size_t)(i - 1)], b = bestsub(bf, af, i);
            const double slack = tie_rel * std::fmax(std::fabs(a), std::fabs(b));
            besttree[i - 1] = (a <= b + slack) ? 0 : 1;
        }
    }
    if (entr_bf) for (int64_t i = 0; i < ntree; ++i) entr_bf[i] = bf[(size_t)i];
    if (entr_af) for (size_t i = 0; i < af.size(); ++i) entr_af[i] = af[i];
    return WB200_OK;
}
