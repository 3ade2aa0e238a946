// threshold.cu -- SURVEY 8(f) row 2: the main CALLER of the transform path, kept on the device end to end.
//   threshold!(x, TH, t)                 src/Threshold/threshold_main.jl:35-117   -> k_threshold (elementwise)
//   noisest(x, wt) = mad!(dr) / 0.6745   src/Threshold/denoising.jl:88-106        -> level-1 dwt + device MAD (radix select)
//   denoise(x, wt; L, dnt, TI, nspin)    src/Threshold/denoising.jl:22-82         -> wb200_denoise
// Nothing here synchronises the stream: the noise level stays in device memory and the threshold kernel reads it, so a
// denoise call is a pure enqueue like the transforms (only wb200_noisest, which must hand a number back, waits).
// Arithmetic follows Julia's promotion rules: t = sigma * dnt.t is Float64, so comparisons and products with t are done in
// double and rounded to T on the store; medians are exact order statistics (middle(a, b) = a/2 + b/2 in T).
#include "common.cuh"
#include <cmath>
#include <cstdlib>

namespace wb {

// ---- threshold!(x, TH, t) -------------------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ double sgn_of(T v) { return (v > 0) ? 1.0 : ((v < 0) ? -1.0 : (double)v); }

template <typename T>
__global__ void __launch_bounds__(256)
k_threshold(T *__restrict__ x, int64_t n, int kind, double t_host, const double *__restrict__ sigma_dev, double tfac) {
    const double t = sigma_dev ? __dmul_rn(*sigma_dev, tfac) : t_host;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const T v = x[i];
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
        x[i] = o;
    }
}

// ---- exact order statistics by radix select (no sort, no host round trip) --------------------------------------
// Two ranks are selected together (the two middle elements of an even-length array).  State lives in device memory:
//   sel[s].prefix / mask : the bits of the answer fixed so far;  sel[s].rank : rank among the still matching elements
struct SelState { unsigned long long prefix, mask; long long rank; };
struct SelBuf { SelState st[2]; unsigned int hist[2][256]; };

template <typename T> struct Key;
template <> struct Key<float> {
    using U = unsigned int; static constexpr int BITS = 32;
    static __device__ __forceinline__ unsigned long long of(float v) { const unsigned int u = __float_as_uint(v); return (u & 0x80000000u) ? (unsigned int)~u : (u | 0x80000000u); }
    static __device__ __forceinline__ float back(unsigned long long k) { const unsigned int u = (unsigned int)k; return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u); }
};
template <> struct Key<double> {
    using U = unsigned long long; static constexpr int BITS = 64;
    static __device__ __forceinline__ unsigned long long of(double v) { const unsigned long long u = (unsigned long long)__double_as_longlong(v); return (u >> 63) ? ~u : (u | 0x8000000000000000ull); }
    static __device__ __forceinline__ double back(unsigned long long k) { return __longlong_as_double((long long)((k >> 63) ? (k & 0x7fffffffffffffffull) : ~k)); }
};

__global__ void k_sel_init(SelBuf *sb, long long r0, long long r1) {
    if (threadIdx.x < 2) { sb->st[threadIdx.x].prefix = 0; sb->st[threadIdx.x].mask = 0; sb->st[threadIdx.x].rank = threadIdx.x ? r1 : r0; }
    for (int i = threadIdx.x; i < 512; i += blockDim.x) (&sb->hist[0][0])[i] = 0;
}
template <typename T, bool ABS>
__global__ void __launch_bounds__(256)
k_sel_hist(const T *__restrict__ v, int64_t m, SelBuf *sb, int shift) {
    __shared__ unsigned int h[2][256];
    for (int i = threadIdx.x; i < 512; i += blockDim.x) (&h[0][0])[i] = 0;
    __syncthreads();
    const unsigned long long p0 = sb->st[0].prefix, m0 = sb->st[0].mask, p1 = sb->st[1].prefix, m1 = sb->st[1].mask;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (int64_t)gridDim.x * blockDim.x) {
        const unsigned long long k = Key<T>::of(ABS ? (v[i] < 0 ? -v[i] : v[i]) : v[i]);
        const unsigned int d = (unsigned int)(k >> shift) & 255u;
        if ((k & m0) == p0) atomicAdd(&h[0][d], 1u);
        if ((k & m1) == p1) atomicAdd(&h[1][d], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 512; i += blockDim.x) { const unsigned int c = (&h[0][0])[i]; if (c) atomicAdd(&(&sb->hist[0][0])[i], c); }
}
__global__ void k_sel_pick(SelBuf *sb, int shift) {
    const int s = threadIdx.x;
    if (s < 2) {
        long long r = sb->st[s].rank;
        int d = 0;
        for (; d < 255; ++d) { const long long c = sb->hist[s][d]; if (r < c) break; r -= c; }
        sb->st[s].rank = r;
        sb->st[s].prefix |= (unsigned long long)d << shift;
        sb->st[s].mask |= 0xffull << shift;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 512; i += blockDim.x) (&sb->hist[0][0])[i] = 0;
}
// median = middle(v[k0], v[k1]) = v0/2 + v1/2 in T;  mode 0: store it (T) for the deviation pass;  mode 1: sigma = mad / 0.6745
template <typename T>
__global__ void k_sel_finish(const SelBuf *sb, T *med_out, double *sigma_out, int mode) {
    if (threadIdx.x == 0) {
        const T a = Key<T>::back(sb->st[0].prefix), b = Key<T>::back(sb->st[1].prefix);
        T med;
        if constexpr (sizeof(T) == 4) med = __fadd_rn(__fmul_rn(a, 0.5f), __fmul_rn(b, 0.5f)); else med = __dadd_rn(__dmul_rn(a, 0.5), __dmul_rn(b, 0.5));
        if (mode == 0) *med_out = med;
        else *sigma_out = __ddiv_rn((double)med, 0.6745);
    }
}
template <typename T>
__global__ void __launch_bounds__(256) k_absdev(T *__restrict__ v, int64_t m, const T *__restrict__ med) {
    const T c = *med;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (int64_t)gridDim.x * blockDim.x) {
        T d;
        if constexpr (sizeof(T) == 4) d = __fsub_rn(v[i], c); else d = __dsub_rn(v[i], c);
        v[i] = d < 0 ? -d : d;
    }
}

// ---- translation-invariant cycle spinning (Base.circshift: out[(i + s) mod n] = in[i] per dimension) ---------------
// Translation-invariant denoising runs its nspin^d shifted copies as ONE batch: all shifted copies are laid out back to back
// (k_spin_scatter), transformed / thresholded / inverted by the batched entry points, and summed back in spin order with the
// inverse shifts (k_spin_gather_add) -- the same per-element arithmetic and the same accumulation order as the reference's
// loop over spins (denoising.jl:36-64), in ~6 launches per chunk of spins instead of ~6 per spin.
struct SpinPlan { int64_t d[3]; int32_t ns[3]; int32_t ndim; };
__device__ __forceinline__ void spin_shift(const SpinPlan &p, int64_t spin, int64_t (&s)[3]) {
    int64_t r = spin;
    for (int a = 0; a < 3; ++a) {
        int64_t v = 0;
        if (a < p.ndim) { if (p.ndim == 1) v = spin; else { v = r % p.ns[a]; r /= p.ns[a]; } }
        s[a] = v % p.d[a];
    }
}
template <typename T>
__global__ void __launch_bounds__(256) k_spin_scatter(T *__restrict__ Z, const T *__restrict__ x, const __grid_constant__ SpinPlan p, int64_t s0, int cnt) {
    const int64_t tot = p.d[0] * p.d[1] * p.d[2];
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < tot * cnt; idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t e = idx % tot, sp = idx / tot;
        int64_t s[3];
        spin_shift(p, s0 + sp, s);
        const int64_t i = e % p.d[0], r = e / p.d[0], j = r % p.d[1], k = r / p.d[1];
        int64_t oi = i + s[0]; if (oi >= p.d[0]) oi -= p.d[0];
        int64_t oj = j + s[1]; if (oj >= p.d[1]) oj -= p.d[1];
        int64_t ok = k + s[2]; if (ok >= p.d[2]) ok -= p.d[2];
        Z[sp * tot + (ok * p.d[1] + oj) * p.d[0] + oi] = x[e];           // z = circshift(x, shift)
    }
}
template <typename T>
__global__ void __launch_bounds__(256) k_spin_gather_add(T *__restrict__ y, const T *__restrict__ Z, const __grid_constant__ SpinPlan p, int64_t s0, int cnt) {
    const int64_t tot = p.d[0] * p.d[1] * p.d[2];
    for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < tot; o += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = o % p.d[0], r = o / p.d[0], j = r % p.d[1], k = r / p.d[1];
        T acc = y[o];
        for (int sp = 0; sp < cnt; ++sp) {                                 // arrayadd!(y, circshift(z, -shift)), spins in order
            int64_t s[3];
            spin_shift(p, s0 + sp, s);
            int64_t si = i + s[0]; if (si >= p.d[0]) si -= p.d[0];
            int64_t sj = j + s[1]; if (sj >= p.d[1]) sj -= p.d[1];
            int64_t sk = k + s[2]; if (sk >= p.d[2]) sk -= p.d[2];
            const T w = Z[(int64_t)sp * tot + (sk * p.d[1] + sj) * p.d[0] + si];
            if constexpr (sizeof(T) == 4) acc = __fadd_rn(acc, w); else acc = __dadd_rn(acc, w);
        }
        y[o] = acc;
    }
}
template <typename T>
__global__ void __launch_bounds__(256) k_scale(T *__restrict__ y, int64_t n, double s) {   // rmul!(y, 1 / pns)
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        y[i] = (T)__dmul_rn((double)y[i], s);
}

// ---- threshold!(x, BiggestTH(), m): keep the m entries of largest magnitude (threshold_main.jl:21-33) ----------------------
// v = the (n-m)-th smallest |x| (0-based) by radix select; everything below it goes, and of the entries EQUAL to it the first
// `rank` in index order go too (rank = the select state's remaining rank = how many ties sort before the kept ones).  The
// reference orders ties by an unstable QuickSort, i.e. arbitrarily; index order is the oracle's (stable) choice.
template <typename T>
__global__ void __launch_bounds__(256) k_biggest_apply(T *__restrict__ x, int64_t n, const SelBuf *__restrict__ sb) {
    const T v = Key<T>::back(sb->st[0].prefix);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const T a = x[i] < 0 ? -x[i] : x[i];
        if (a < v) x[i] = 0;
    }
}
template <typename T>
__global__ void __launch_bounds__(1024) k_biggest_ties(T *__restrict__ x, int64_t n, const SelBuf *__restrict__ sb) {
    // one CTA walks the array in index order and zeroes the first `rank` entries whose magnitude equals v (rare path)
    __shared__ int wsum[32];
    __shared__ long long base_s;
    const long long todo = sb->st[0].rank;
    if (todo <= 0) return;
    const T v = Key<T>::back(sb->st[0].prefix);
    if (threadIdx.x == 0) base_s = 0;
    __syncthreads();
    for (int64_t c0 = 0; c0 < n; c0 += blockDim.x) {
        const int64_t i = c0 + threadIdx.x;
        int tie = 0;
        if (i < n) { const T a = x[i] < 0 ? -x[i] : x[i]; tie = (a == v) ? 1 : 0; }
        int incl = tie;
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if ((threadIdx.x & 31) >= o) incl += t; }
        if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = incl;
        __syncthreads();
        int woff = 0;
        for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) woff += wsum[w];
        const long long ord = base_s + woff + incl - tie;                  // ties before this entry
        if (tie && ord < todo) x[i] = 0;
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) base_s += woff + incl;
        __syncthreads();
        if (base_s >= todo) break;
    }
}

static unsigned grid_for_n(int64_t n) {
    int64_t b = (n + 255) / 256;
    if (b < 1) b = 1;
    if (b > 148 * 16) b = 148 * 16;
    return (unsigned)b;
}

// device MAD of m contiguous values (destroyed) -> *sigma_dev = mad / 0.6745
template <typename T>
static bool device_mad(T *v, int64_t m, SelBuf *sb, T *med, double *sigma_dev, cudaStream_t st) {
    const unsigned g = grid_for_n(m);
    for (int round = 0; round < 2; ++round) {
        {
            LaunchScope scope("select_init", st);
            k_sel_init<<<1, 256, 0, st>>>(sb, (long long)((m - 1) / 2), (long long)(m / 2));
        }
        for (int shift = Key<T>::BITS - 8; shift >= 0; shift -= 8) {
            { LaunchScope scope("select_hist", st); k_sel_hist<T, false><<<g, 256, 0, st>>>(v, m, sb, shift); }
            { LaunchScope scope("select_pick", st); k_sel_pick<<<1, 256, 0, st>>>(sb, shift); }
        }
        { LaunchScope scope("select_finish", st); k_sel_finish<T><<<1, 32, 0, st>>>(sb, med, sigma_dev, round); }
        if (round == 0) { LaunchScope scope("mad_absdev", st); k_absdev<T><<<g, 256, 0, st>>>(v, m, med); }
    }
    return check_launch("device_mad");
}

} // namespace wb

using namespace wb;

namespace {
struct WtArgs { int32_t wkind; const double *qmf; int32_t flen; const wb200_lift_step *steps; int32_t nsteps; double norm1, norm2; };

// one transform through the library's own entry points (x != y: the out-of-place forms)
int32_t xform(void *y, const void *x, int32_t ndim, const int64_t *dims, const WtArgs &w, int32_t L, int32_t fw, int32_t dtype,
              void *stream, uint32_t flags, int64_t batch = 1) {
    if (w.wkind == 1) return wb200_dwt_filter(y, x, ndim, dims, batch, w.qmf, w.flen, L, fw, dtype, nullptr, 0, stream, flags);
    return wb200_dwt_lifting(y, x, ndim, dims, batch, w.steps, w.nsteps, w.norm1, w.norm2, L, fw, dtype, nullptr, 0, stream, flags);
}
bool check_wt(const WtArgs &w) {
    if (w.wkind == 0) return true;
    if (w.wkind == 1) return w.qmf != nullptr && w.flen >= 2;
    if (w.wkind == 2) return w.steps != nullptr && w.nsteps >= 1;
    return false;
}

// noisest on the device: *sigma_dev = mad(y[n1/2 : n1]) / 0.6745, y = level-1 transform of x (or x itself).  tmp: tot elements.
template <typename T>
int32_t noisest_dev(double *sigma_dev, const T *x, int32_t ndim, const int64_t *dims, int64_t tot, const WtArgs &w, T *tmp,
                    SelBuf *sb, T *med, int32_t dtype, cudaStream_t st, uint32_t flags, int32_t L = 1) {
    // detailrange(size(y, 1), L) = round(n1/2^L + 1) : round(n1/2^(L-1)) (src/Util/non_dyadic.jl:9), as LINEAR indices into y;
    // round(Int, .): ties to even
    if (L < 1 || L > 60) { set_error("L must be positive"); return WB200_ELEVEL; }
    const int64_t n1 = dims[0], lo = (int64_t)std::nearbyint((double)n1 / (double)((int64_t)1 << L) + 1) - 1,
                  hi = (int64_t)std::nearbyint((double)n1 / (double)((int64_t)1 << (L - 1))), m = hi - lo;
    if (m < 1 || hi > tot) { set_error("noisest: no detail coefficients at level %d (size(x,1) = %lld)", (int)L, (long long)n1); return WB200_EDIMS; }
    if (w.wkind == 0) {
        if (cudaMemcpyAsync(tmp, x + lo, sizeof(T) * (size_t)m, cudaMemcpyDeviceToDevice, st) != cudaSuccess) { (void)cudaGetLastError(); return WB200_ECUDA; }
    } else {
        const int32_t rc = xform(tmp, x, ndim, dims, w, L, 1, dtype, (void *)st, flags);
        if (rc != WB200_OK) return rc;
        if (lo > 0 && cudaMemcpyAsync(tmp, tmp + lo, sizeof(T) * (size_t)m, cudaMemcpyDeviceToDevice, st) != cudaSuccess) { (void)cudaGetLastError(); return WB200_ECUDA; }
    }
    // (the copy ranges [lo, lo+m) and [0, m) never overlap: m = hi - lo <= lo for every L >= 1)
    return device_mad<T>(tmp, m, sb, med, sigma_dev, st) ? WB200_OK : WB200_ECUDA;
}

template <typename T>
int32_t denoise_t(T *y, const T *x, int32_t ndim, const int64_t *dims, const WtArgs &w, int32_t L, int32_t th_kind, double tfac,
                  double sigma, int32_t TI, const int32_t *nspin, int32_t dtype, cudaStream_t st, uint32_t flags) {
    int64_t tot = 1;
    for (int a = 0; a < ndim; ++a) tot *= dims[a];
    for (int a = 1; a < ndim; ++a) if (dims[a] != dims[0]) { set_error("array must be square/cube"); return WB200_ENOTCUBE; }
    if (tot == 0) return WB200_OK;
    if (TI && w.wkind == 0) { set_error("TI not supported with wt=nothing"); return WB200_EARG; }
    const bool est = sigma != sigma;
    // scratch: TI needs three arrays (z, xt, w), plain needs one; the selection state and two scalars ride behind them
    int64_t pns = 1, chunk = 1;
    if (TI) {
        for (int a = 0; a < ndim; ++a) pns *= nspin[a] > 0 ? nspin[a] : 1;
        // spins per batch: two batch buffers of `chunk` shifted copies each, at most ~1 GiB together
        const char *e = std::getenv("WB200_DENOISE_CHUNK_MB");
        const size_t budget = (size_t)(e ? std::atoll(e) : 512) << 20;
        chunk = (int64_t)(budget / ((size_t)tot * sizeof(T)));
        if (chunk < 1) chunk = 1;
        if (chunk > pns) chunk = pns;
        if (chunk > 4096) chunk = 4096;
    }
    const size_t arr = (((size_t)tot * (size_t)chunk * sizeof(T)) + 255) & ~(size_t)255;
    const int narr = TI ? 2 : 1;
    char *pool = nullptr;
    if (scratch_alloc((void **)&pool, narr * arr + 4096, st) != cudaSuccess) { (void)cudaGetLastError(); set_error("cudaMallocAsync(denoise scratch) failed"); return WB200_ECUDA; }
    T *a0 = (T *)pool, *a1 = (T *)(pool + (narr > 1 ? arr : 0));
    SelBuf *sb = (SelBuf *)(pool + narr * arr);
    T *med = (T *)(pool + narr * arr + 3072);
    double *sig = (double *)(pool + narr * arr + 3072 + 64);
    int32_t rc = WB200_OK;
    const unsigned g = grid_for_n(tot);
    if (est) rc = noisest_dev<T>(sig, x, ndim, dims, tot, w, a0, sb, med, dtype, st, flags);
    auto thr = [&](T *c) {
        LaunchScope scope("threshold", st);
        k_threshold<T><<<g, 256, 0, st>>>(c, tot, th_kind, sigma * tfac, est ? sig : nullptr, tfac);
    };
    if (rc == WB200_OK && !TI) {
        if (w.wkind == 0) {
            if (y != x && cudaMemcpyAsync(y, x, sizeof(T) * (size_t)tot, cudaMemcpyDeviceToDevice, st) != cudaSuccess) { (void)cudaGetLastError(); rc = WB200_ECUDA; }
            if (rc == WB200_OK) thr(y);
        } else {
            rc = xform(a0, x, ndim, dims, w, L, 1, dtype, (void *)st, flags);
            if (rc == WB200_OK) { thr(a0); rc = xform(y, a0, ndim, dims, w, L, 0, dtype, (void *)st, flags); }
        }
    } else if (rc == WB200_OK) {
        SpinPlan sp{};
        sp.ndim = ndim;
        for (int a = 0; a < 3; ++a) { sp.d[a] = a < ndim ? dims[a] : 1; sp.ns[a] = (a < ndim && nspin[a] > 0) ? nspin[a] : 1; }
        if (cudaMemsetAsync(y, 0, sizeof(T) * (size_t)tot, st) != cudaSuccess) { (void)cudaGetLastError(); rc = WB200_ECUDA; }
        for (int64_t s0 = 0; s0 < pns && rc == WB200_OK; s0 += chunk) {
            const int cnt = (int)((pns - s0 < chunk) ? (pns - s0) : chunk);
            const unsigned gb = grid_for_n(tot * cnt);
            { LaunchScope scope("spin_scatter", st); k_spin_scatter<T><<<gb, 256, 0, st>>>(a0, x, sp, s0, cnt); }
            rc = xform(a1, a0, ndim, dims, w, L, 1, dtype, (void *)st, flags, cnt);
            if (rc != WB200_OK) break;
            {
                LaunchScope scope("threshold", st);
                k_threshold<T><<<gb, 256, 0, st>>>(a1, tot * cnt, th_kind, sigma * tfac, est ? sig : nullptr, tfac);
            }
            rc = xform(a0, a1, ndim, dims, w, L, 0, dtype, (void *)st, flags, cnt);
            if (rc != WB200_OK) break;
            { LaunchScope scope("spin_gather_add", st); k_spin_gather_add<T><<<g, 256, 0, st>>>(y, a0, sp, s0, cnt); }
        }
        if (rc == WB200_OK) { LaunchScope scope("scale", st); k_scale<T><<<g, 256, 0, st>>>(y, tot, 1.0 / (double)pns); }
    }
    if (rc == WB200_OK && !check_launch("denoise")) rc = WB200_ECUDA;
    cudaFreeAsync(pool, st);
    return rc;
}
} // namespace

extern "C" int32_t wb200_threshold(void *x, int64_t count, int32_t kind, double t, int32_t dtype, void *stream) {
    if (dtype != WB200_F32 && dtype != WB200_F64) { set_error("threshold supports Float32/Float64"); return WB200_EDTYPE; }
    if (kind < WB200_TH_HARD || kind > WB200_TH_POS) { set_error("unknown threshold kind %d", kind); return WB200_EARG; }
    if (count < 0 || (count > 0 && x == nullptr)) { set_error("bad array"); return WB200_EARG; }
    if ((kind <= WB200_TH_STEIN) && !(t >= 0)) { set_error("threshold must be >= 0"); return WB200_EARG; }     // @assert t >= 0
    if (count == 0) return WB200_OK;
    cudaStream_t st = (cudaStream_t)stream;
    {
        LaunchScope scope("threshold", st);
        if (dtype == WB200_F64) k_threshold<double><<<grid_for_n(count), 256, 0, st>>>((double *)x, count, kind, t, nullptr, 1.0);
        else                    k_threshold<float><<<grid_for_n(count), 256, 0, st>>>((float *)x, count, kind, t, nullptr, 1.0);
    }
    return check_launch("threshold") ? WB200_OK : WB200_ECUDA;
}

extern "C" int32_t wb200_threshold_biggest(void *x, int64_t count, int64_t m, int32_t dtype, void *stream) {
    if (dtype != WB200_F32 && dtype != WB200_F64) { set_error("threshold supports Float32/Float64"); return WB200_EDTYPE; }
    if (count < 0 || (count > 0 && x == nullptr) || m < 0) { set_error("bad argument (m >= 0)"); return WB200_EARG; }   // @assert m >= 0
    if (count == 0 || m >= count) return WB200_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t esz = dtype == WB200_F64 ? 8 : 4;
    if (m == 0) return cudaMemsetAsync(x, 0, (size_t)count * esz, st) == cudaSuccess ? WB200_OK : WB200_ECUDA;
    SelBuf *sb = nullptr;
    if (scratch_alloc((void **)&sb, sizeof(SelBuf), st) != cudaSuccess) { (void)cudaGetLastError(); set_error("cudaMallocAsync(select state) failed"); return WB200_ECUDA; }
    const unsigned g = grid_for_n(count);
    const long long k = (long long)(count - m);
    { LaunchScope scope("select_init", st); k_sel_init<<<1, 256, 0, st>>>(sb, k, k); }
    for (int shift = (int)esz * 8 - 8; shift >= 0; shift -= 8) {
        {
            LaunchScope scope("select_hist", st);
            if (dtype == WB200_F64) k_sel_hist<double, true><<<g, 256, 0, st>>>((const double *)x, count, sb, shift);
            else                    k_sel_hist<float, true><<<g, 256, 0, st>>>((const float *)x, count, sb, shift);
        }
        { LaunchScope scope("select_pick", st); k_sel_pick<<<1, 256, 0, st>>>(sb, shift); }
    }
    {
        LaunchScope scope("biggest_apply", st);
        if (dtype == WB200_F64) { k_biggest_ties<double><<<1, 1024, 0, st>>>((double *)x, count, sb); k_biggest_apply<double><<<g, 256, 0, st>>>((double *)x, count, sb); }
        else                    { k_biggest_ties<float><<<1, 1024, 0, st>>>((float *)x, count, sb); k_biggest_apply<float><<<g, 256, 0, st>>>((float *)x, count, sb); }
    }
    const bool ok = check_launch("threshold_biggest");
    cudaFreeAsync(sb, st);
    return ok ? WB200_OK : WB200_ECUDA;
}

extern "C" int32_t wb200_noisest(double *sigma_out, const void *x, int32_t ndim, const int64_t *dims, int32_t wkind,
                                 const double *qmf, int32_t flen, const wb200_lift_step *steps, int32_t nsteps, double norm1,
                                 double norm2, int32_t L, int32_t dtype, void *stream, uint32_t flags) {
    if (dtype != WB200_F32 && dtype != WB200_F64) { set_error("noisest supports Float32/Float64"); return WB200_EDTYPE; }
    if (sigma_out == nullptr || x == nullptr || dims == nullptr || ndim < 1 || ndim > 3) { set_error("bad argument"); return WB200_EARG; }
    const WtArgs w{wkind, qmf, flen, steps, nsteps, norm1, norm2};
    if (!check_wt(w)) { set_error("bad wavelet description"); return WB200_EARG; }
    int64_t tot = 1;
    for (int a = 0; a < ndim; ++a) { if (dims[a] < 1) { set_error("dims[%d] = %lld", a, (long long)dims[a]); return WB200_EDIMS; } tot *= dims[a]; }
    cudaStream_t st = (cudaStream_t)stream;
    const size_t esz = dtype == WB200_F64 ? 8 : 4;
    const size_t arr = (((size_t)tot * esz) + 255) & ~(size_t)255;
    char *pool = nullptr;
    if (scratch_alloc((void **)&pool, arr + 4096, st) != cudaSuccess) { (void)cudaGetLastError(); set_error("cudaMallocAsync(noisest scratch) failed"); return WB200_ECUDA; }
    SelBuf *sb = (SelBuf *)(pool + arr);
    double *sig = (double *)(pool + arr + 3072 + 64);
    int32_t rc;
    if (dtype == WB200_F64) rc = noisest_dev<double>(sig, (const double *)x, ndim, dims, tot, w, (double *)pool, sb, (double *)(pool + arr + 3072), dtype, st, flags, L);
    else                    rc = noisest_dev<float>(sig, (const float *)x, ndim, dims, tot, w, (float *)pool, sb, (float *)(pool + arr + 3072), dtype, st, flags, L);
    if (rc == WB200_OK) {
        if (cudaMemcpyAsync(sigma_out, sig, sizeof(double), cudaMemcpyDeviceToHost, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess) {
            (void)cudaGetLastError(); set_error("noisest: copy of the result failed"); rc = WB200_ECUDA;
        }
    }
    cudaFreeAsync(pool, st);
    return rc;
}

extern "C" int32_t wb200_denoise(void *y, const void *x, int32_t ndim, const int64_t *dims, int32_t wkind, const double *qmf,
                                 int32_t flen, const wb200_lift_step *steps, int32_t nsteps, double norm1, double norm2, int32_t L,
                                 int32_t th_kind, double tfac, double sigma, int32_t TI, const int32_t *nspin, int32_t dtype,
                                 void *stream, uint32_t flags) {
    if (dtype != WB200_F32 && dtype != WB200_F64) { set_error("denoise supports Float32/Float64"); return WB200_EDTYPE; }
    if (y == nullptr || x == nullptr || dims == nullptr || ndim < 1 || ndim > 3 || (TI && nspin == nullptr)) { set_error("bad argument"); return WB200_EARG; }
    if (th_kind < WB200_TH_HARD || th_kind > WB200_TH_POS) { set_error("unknown threshold kind %d", th_kind); return WB200_EARG; }
    const WtArgs w{wkind, qmf, flen, steps, nsteps, norm1, norm2};
    if (!check_wt(w)) { set_error("bad wavelet description"); return WB200_EARG; }
    for (int a = 0; a < ndim; ++a) if (dims[a] < 1) { set_error("dims[%d] = %lld", a, (long long)dims[a]); return WB200_EDIMS; }
    if (y == x && w.wkind != 0) { set_error("denoise: y must not alias x"); return WB200_EALIAS; }
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == WB200_F64) return denoise_t<double>((double *)y, (const double *)x, ndim, dims, w, L, th_kind, tfac, sigma, TI, nspin, dtype, st, flags);
    return denoise_t<float>((float *)y, (const float *)x, ndim, dims, w, L, th_kind, tfac, sigma, TI, nspin, dtype, st, flags);
}
