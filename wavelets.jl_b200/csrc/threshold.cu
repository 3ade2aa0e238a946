// threshold.cu -- SURVEY 8(f) row 2: the main CALLER of the transform path, kept on the device end to end.
//   threshold!(x, TH, t)                 src/Threshold/threshold_main.jl:35-117   -> k_threshold (elementwise)
//   noisest(x, wt) = mad!(dr) / 0.6745   src/Threshold/denoising.jl:88-106        -> level-1 dwt + device MAD (radix select)
//   denoise(x, wt; L, dnt, TI, nspin)    src/Threshold/denoising.jl:22-82         -> wb200_denoise
// Nothing here synchronises the stream: the noise level stays in device memory and the threshold kernel reads it, so a
// denoise call is a pure enqueue like the transforms (only wb200_noisest, which must hand a number back, waits).
// Arithmetic follows Julia's promotion rules: t = sigma * dnt.t is Float64, so comparisons and products with t are done in
// double and rounded to T on the store; medians are exact order statistics (middle(a, b) = a/2 + b/2 in T).
#include "common.cuh"
#include "thresh_dev.cuh"
#include <cmath>
#include <cstdlib>

namespace wb {

// ---- threshold!(x, TH, t) -------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
k_threshold(T *__restrict__ x, int64_t n, int kind, double t_host, const double *__restrict__ sigma_dev, double tfac) {
    const double t = sigma_dev ? __dmul_rn(*sigma_dev, tfac) : t_host;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        x[i] = thresh_apply<T>(x[i], kind, t);
}

// ---- exact order statistics by radix select (no sort, no host round trip) --------------------------------------
// Two ranks are selected together (the two middle elements of an even-length array).  State lives in device memory:
//   sel[s].prefix / mask : the bits of the answer fixed so far;  sel[s].rank : rank among the still matching elements
struct SelState { unsigned long long prefix, mask; long long rank; };
constexpr int SEL_BINS = 2048;             // 11-bit digits: three passes over 32-bit keys, six over 64-bit keys
struct SelBuf { SelState st[2]; unsigned int done, pad; unsigned int hist[2][SEL_BINS]; };
constexpr size_t SELBUF_BYTES = (sizeof(SelBuf) + 255) & ~(size_t)255;

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
// digit `pass` (0 = most significant) of a BITS-wide key: bits [shift, shift + width)
template <typename T> static inline int sel_passes() { return (Key<T>::BITS + 10) / 11; }
template <typename T> static inline void sel_digit(int pass, int &shift, int &width) {
    const int hi = Key<T>::BITS - 11 * pass;
    const int lo = hi - 11 > 0 ? hi - 11 : 0;
    shift = lo; width = hi - lo;
}

__global__ void k_sel_init(SelBuf *sb, long long r0, long long r1) {
    if (threadIdx.x < 2) { sb->st[threadIdx.x].prefix = 0; sb->st[threadIdx.x].mask = 0; sb->st[threadIdx.x].rank = threadIdx.x ? r1 : r0; }
    if (threadIdx.x == 0) { sb->done = 0; sb->pad = 0; }
    for (int i = threadIdx.x; i < 2 * SEL_BINS; i += blockDim.x) (&sb->hist[0][0])[i] = 0;
}
// One pass of the select in ONE launch: every CTA histograms its share of the still-matching keys (lanes that hit the same
// bin are aggregated with match_any before the shared-memory atomic: the leading digit of N(0,1) data lands in a handful of
// bins), adds it to the global histogram and takes a ticket; the LAST CTA picks the bin that holds each rank, narrows the
// state, clears the histogram for the next pass and -- after the final pass (fin >= 0) -- forms the median
// (middle(a, b) = a/2 + b/2 in T): fin 0 stores it for the deviation pass, fin 1 stores sigma = mad / 0.6745.
// (The first edition made 8 + 8 + 2 launches per statistic: 0.28 of the 0.355 ms of a 2^24-sample denoise.)
template <typename T, bool ABS>
__global__ void __launch_bounds__(256)
k_sel_pass(const T *__restrict__ v, int64_t m, SelBuf *sb, int shift, int width, int fin, T *med_out, double *sigma_out) {
    __shared__ unsigned int h[2][SEL_BINS];
    __shared__ long long scan[256];
    __shared__ unsigned int ticket;
    for (int i = threadIdx.x; i < 2 * SEL_BINS; i += blockDim.x) (&h[0][0])[i] = 0;
    __syncthreads();
    const unsigned long long p0 = sb->st[0].prefix, m0 = sb->st[0].mask, p1 = sb->st[1].prefix, m1 = sb->st[1].mask;
    const unsigned int dmask = (1u << width) - 1u;
    const int lane = threadIdx.x & 31;
    // (warp-aggregating the increments with match_any was measured: 54 us per pass against 21 us for plain shared-memory
    //  atomics on 8 M keys -- the 11-bit leading digit already spreads N(0,1) data over enough bins)
    (void)lane;
    const bool same = (p0 == p1) && (m0 == m1);                            // both ranks still in the same bin: count once
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (int64_t)gridDim.x * blockDim.x) {
        const unsigned long long k = Key<T>::of(ABS ? (v[i] < 0 ? -v[i] : v[i]) : v[i]);
        const unsigned int d = (unsigned int)(k >> shift) & dmask;
        if ((k & m0) == p0) atomicAdd(&h[0][d], 1u);
        if (!same && (k & m1) == p1) atomicAdd(&h[1][d], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * SEL_BINS; i += blockDim.x) {
        const unsigned int c = (same && i >= SEL_BINS) ? h[0][i - SEL_BINS] : (&h[0][0])[i];
        if (c) atomicAdd(&(&sb->hist[0][0])[i], c);
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) ticket = atomicAdd(&sb->done, 1u);
    __syncthreads();
    if (ticket != gridDim.x - 1) return;
    // ---- last CTA: pick the bins, narrow the state, reset ----
    __threadfence();
    constexpr int PER = SEL_BINS / 256;
    for (int s = 0; s < 2; ++s) {
        unsigned int c[PER];
        long long sum = 0;
#pragma unroll
        for (int q = 0; q < PER; ++q) { c[q] = __ldcg(&sb->hist[s][threadIdx.x * PER + q]); sum += c[q]; }
        scan[threadIdx.x] = sum;
        __syncthreads();
        for (int o = 1; o < 256; o <<= 1) {                                // inclusive scan of the 256 partial sums
            const long long t = (threadIdx.x >= (unsigned)o) ? scan[threadIdx.x - o] : 0;
            __syncthreads();
            scan[threadIdx.x] += t;
            __syncthreads();
        }
        const long long incl = scan[threadIdx.x], excl = incl - sum;
        const long long r = sb->st[s].rank;
        const long long total = scan[255];
        // the thread whose bins hold rank r (the last non-empty thread takes a rank beyond the total, as the serial scan did)
        const bool mine = (r >= excl && r < incl) || (r >= total && threadIdx.x == 255);
        __syncthreads();
        if (mine) {
            long long rr = r - excl;
            int q = 0;
            for (; q < PER - 1; ++q) { if (rr < (long long)c[q]) break; rr -= c[q]; }
            const unsigned long long d = (unsigned long long)(threadIdx.x * PER + q);
            sb->st[s].rank = rr;
            sb->st[s].prefix |= d << shift;
            sb->st[s].mask |= (unsigned long long)dmask << shift;
        }
        __syncthreads();
    }
    for (int i = threadIdx.x; i < 2 * SEL_BINS; i += blockDim.x) (&sb->hist[0][0])[i] = 0;
    if (threadIdx.x == 0) {
        sb->done = 0;
        if (fin >= 0) {
            const T a = Key<T>::back(sb->st[0].prefix), b = Key<T>::back(sb->st[1].prefix);
            T med;
            if constexpr (sizeof(T) == 4) med = __fadd_rn(__fmul_rn(a, 0.5f), __fmul_rn(b, 0.5f)); else med = __dadd_rn(__dmul_rn(a, 0.5), __dmul_rn(b, 0.5));
            if (fin == 0) *med_out = med;
            else *sigma_out = __ddiv_rn((double)med, 0.6745);
        }
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
// Both kernels work ROW-wise (a row = d[0] contiguous samples at fixed (j, k)): the shift of a spin and the wrapped source
// / destination row are resolved once per row with the 64-bit divisions, threads then move along the row with 32-bit
// indices -- the first edition resolved three 64-bit divisions per ELEMENT and ran at 0.37 TB/s (0.69 + 0.50 ms of a 1.95 ms
// TI call on 1024^2 x 64 spins).
constexpr int SPIN_CHUNK = 2048;           // samples of a row handled by one CTA
template <typename T>
__global__ void __launch_bounds__(256) k_spin_scatter(T *__restrict__ Z, const T *__restrict__ x, const __grid_constant__ SpinPlan p, int64_t s0, int cnt, int rpc) {
    const int d0 = (int)p.d[0];
    const int nchunk = (d0 + SPIN_CHUNK - 1) / SPIN_CHUNK;
    const int64_t rows = p.d[1] * p.d[2], tot = rows * d0;
    const int64_t blk = blockIdx.x;
    const int chunk = (int)(blk % nchunk);
    const int sp = blockIdx.y;
    if (sp >= cnt) return;
    int64_t s[3];
    spin_shift(p, s0 + sp, s);
    const int sh = (int)s[0];
    const int i1 = min(d0, (chunk + 1) * SPIN_CHUNK);
    for (int64_t row = (blk / nchunk) * rpc, rend = min(rows, row + rpc); row < rend; ++row) {    // rpc short rows per CTA
        const int64_t j = row % p.d[1], k = row / p.d[1];
        int64_t oj = j + s[1]; if (oj >= p.d[1]) oj -= p.d[1];
        int64_t ok = k + s[2]; if (ok >= p.d[2]) ok -= p.d[2];
        const T *src = x + row * d0;
        T *dst = Z + (int64_t)sp * tot + (ok * p.d[1] + oj) * d0;
        for (int i = chunk * SPIN_CHUNK + threadIdx.x; i < i1; i += blockDim.x) {
            int oi = i + sh; if (oi >= d0) oi -= d0;
            dst[oi] = src[i];                                              // z = circshift(x, shift)
        }
    }
}
template <typename T>
__global__ void __launch_bounds__(256) k_spin_gather_add(T *__restrict__ y, const T *__restrict__ Z, const __grid_constant__ SpinPlan p, int64_t s0, int cnt) {
    extern __shared__ int64_t spin_tab[];                                  // [cnt][2]: row offset of the shifted source row, dim-1 shift
    const int d0 = (int)p.d[0];
    const int nchunk = (d0 + SPIN_CHUNK - 1) / SPIN_CHUNK;
    const int64_t rows = p.d[1] * p.d[2], tot = rows * d0;
    const int64_t blk = blockIdx.x;
    const int chunk = (int)(blk % nchunk);
    const int64_t row = blk / nchunk;
    if (row >= rows) return;
    const int64_t j = row % p.d[1], k = row / p.d[1];
    for (int sp = threadIdx.x; sp < cnt; sp += blockDim.x) {
        int64_t s[3];
        spin_shift(p, s0 + sp, s);
        int64_t sj = j + s[1]; if (sj >= p.d[1]) sj -= p.d[1];
        int64_t sk = k + s[2]; if (sk >= p.d[2]) sk -= p.d[2];
        spin_tab[2 * sp] = (int64_t)sp * tot + (sk * p.d[1] + sj) * d0;
        spin_tab[2 * sp + 1] = s[0];
    }
    __syncthreads();
    T *yr = y + row * d0;
    const int i1 = min(d0, (chunk + 1) * SPIN_CHUNK);
    for (int i = chunk * SPIN_CHUNK + threadIdx.x; i < i1; i += blockDim.x) {
        T acc = yr[i];
        for (int sp = 0; sp < cnt; ++sp) {                                 // arrayadd!(y, circshift(z, -shift)), spins in order
            int si = i + (int)spin_tab[2 * sp + 1]; if (si >= d0) si -= d0;
            const T w = Z[spin_tab[2 * sp] + si];
            if constexpr (sizeof(T) == 4) acc = __fadd_rn(acc, w); else acc = __dadd_rn(acc, w);
        }
        yr[i] = acc;
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
    const unsigned gs = g < 148 * 6 ? g : 148 * 6;                         // every CTA adds up to 2 x 2048 bins to the global histogram
    for (int round = 0; round < 2; ++round) {
        {
            LaunchScope scope("select_init", st);
            k_sel_init<<<1, 256, 0, st>>>(sb, (long long)((m - 1) / 2), (long long)(m / 2));
        }
        const int np = sel_passes<T>();
        for (int pass = 0; pass < np; ++pass) {
            int shift, width;
            sel_digit<T>(pass, shift, width);
            LaunchScope scope("select_pass", st);
            k_sel_pass<T, false><<<gs, 256, 0, st>>>(v, m, sb, shift, width, pass == np - 1 ? round : -1, med, sigma_dev);
        }
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
        if (chunk > 2048) chunk = 2048;          // (the gather kernel keeps a 16-byte table entry per spin in shared memory)
    }
    const size_t arr = (((size_t)tot * (size_t)chunk * sizeof(T)) + 255) & ~(size_t)255;
    const int narr = TI ? 2 : 1;
    char *pool = nullptr;
    if (scratch_alloc((void **)&pool, narr * arr + SELBUF_BYTES + 256, st) != cudaSuccess) { (void)cudaGetLastError(); set_error("cudaMallocAsync(denoise scratch) failed"); return WB200_ECUDA; }
    T *a0 = (T *)pool, *a1 = (T *)(pool + (narr > 1 ? arr : 0));
    SelBuf *sb = (SelBuf *)(pool + narr * arr);
    T *med = (T *)(pool + narr * arr + SELBUF_BYTES);
    double *sig = (double *)(pool + narr * arr + SELBUF_BYTES + 64);
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
            if (rc == WB200_OK) {
                // threshold as an epilogue of the inverse transform's loads where the fused 1-D synthesis kernels take the
                // call (one pass over the coefficients less); otherwise its own elementwise pass
                int32_t r2 = -1;
                if (w.wkind == 1 && ndim == 1) {
                    ThreshEpi e;
                    e.kind = th_kind; e.t_host = sigma * tfac; e.tfac = tfac; e.sigma_dev = est ? sig : nullptr;
                    r2 = idwt_filter_thresholded(y, a0, dims[0], 1, w.qmf, w.flen, L, dtype, e, st, flags);
                }
                if (r2 == -1) { thr(a0); rc = xform(y, a0, ndim, dims, w, L, 0, dtype, (void *)st, flags); }
                else rc = r2;
            }
        }
    } else if (rc == WB200_OK) {
        SpinPlan sp{};
        sp.ndim = ndim;
        for (int a = 0; a < 3; ++a) { sp.d[a] = a < ndim ? dims[a] : 1; sp.ns[a] = (a < ndim && nspin[a] > 0) ? nspin[a] : 1; }
        if (cudaMemsetAsync(y, 0, sizeof(T) * (size_t)tot, st) != cudaSuccess) { (void)cudaGetLastError(); rc = WB200_ECUDA; }
        for (int64_t s0 = 0; s0 < pns && rc == WB200_OK; s0 += chunk) {
            const int cnt = (int)((pns - s0 < chunk) ? (pns - s0) : chunk);
            const unsigned gb = grid_for_n(tot * cnt);
            const int64_t rowblocks = sp.d[1] * sp.d[2] * ((sp.d[0] + SPIN_CHUNK - 1) / SPIN_CHUNK);
            if (rowblocks > 0x7fffffffLL || sp.d[0] > 0x3fffffffLL) { set_error("denoise: array too large for the spin kernels"); rc = WB200_EDIMS; break; }
            {   // short rows: several per CTA (a 1024-sample row is one iteration of a 256-thread CTA)
                const int rpc = (int)std::max<int64_t>(1, std::min<int64_t>(16, 8192 / sp.d[0]));
                const int64_t nch = (sp.d[0] + SPIN_CHUNK - 1) / SPIN_CHUNK;
                const int64_t sblocks = ((sp.d[1] * sp.d[2] + rpc - 1) / rpc) * nch;
                LaunchScope scope("spin_scatter", st);
                k_spin_scatter<T><<<dim3((unsigned)sblocks, (unsigned)cnt), 256, 0, st>>>(a0, x, sp, s0, cnt, rpc);
            }
            rc = xform(a1, a0, ndim, dims, w, L, 1, dtype, (void *)st, flags, cnt);
            if (rc != WB200_OK) break;
            {
                LaunchScope scope("threshold", st);
                k_threshold<T><<<gb, 256, 0, st>>>(a1, tot * cnt, th_kind, sigma * tfac, est ? sig : nullptr, tfac);
            }
            rc = xform(a0, a1, ndim, dims, w, L, 0, dtype, (void *)st, flags, cnt);
            if (rc != WB200_OK) break;
            { LaunchScope scope("spin_gather_add", st); k_spin_gather_add<T><<<(unsigned)rowblocks, 256, (size_t)cnt * 16, st>>>(y, a0, sp, s0, cnt); }
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
    const int np = dtype == WB200_F64 ? sel_passes<double>() : sel_passes<float>();
    for (int pass = 0; pass < np; ++pass) {
        int shift, width;
        LaunchScope scope("select_pass", st);
        if (dtype == WB200_F64) { sel_digit<double>(pass, shift, width); k_sel_pass<double, true><<<g, 256, 0, st>>>((const double *)x, count, sb, shift, width, -1, nullptr, nullptr); }
        else                    { sel_digit<float>(pass, shift, width);  k_sel_pass<float, true><<<g, 256, 0, st>>>((const float *)x, count, sb, shift, width, -1, nullptr, nullptr); }
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
    if (scratch_alloc((void **)&pool, arr + SELBUF_BYTES + 256, st) != cudaSuccess) { (void)cudaGetLastError(); set_error("cudaMallocAsync(noisest scratch) failed"); return WB200_ECUDA; }
    SelBuf *sb = (SelBuf *)(pool + arr);
    double *sig = (double *)(pool + arr + SELBUF_BYTES + 64);
    int32_t rc;
    if (dtype == WB200_F64) rc = noisest_dev<double>(sig, (const double *)x, ndim, dims, tot, w, (double *)pool, sb, (double *)(pool + arr + SELBUF_BYTES), dtype, st, flags, L);
    else                    rc = noisest_dev<float>(sig, (const float *)x, ndim, dims, tot, w, (float *)pool, sb, (float *)(pool + arr + SELBUF_BYTES), dtype, st, flags, L);
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
