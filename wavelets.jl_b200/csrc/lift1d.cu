// lift1d.cu -- fused multi-level 1-D LIFTING transform for batches of contiguous columns (sm_100a): the lifting
// counterpart of fused1d.cu (north_star: "the lifting predict/update unsafe_dwt1level! loop, for 1-D ... and column-wise
// batches").
//
// The reference's 1-D lifting level (src/Transforms/transforms_lifting.jl:82-122) is split -> one pass per predict /
// update step -> normalize, all in place, i.e. (2 + nsteps) sweeps over the line per level.  Here
//
//   forward   k_lift1d_ana : a CTA stages TILE samples of one column (+ the cumulative halo of K levels, periodic wrap
//                            included) with TMA bulk copies and runs K levels out of shared memory.  A thread takes a
//                            segment of SEG pairs plus halo as 16-byte loads of the INTERLEAVED line (the split is the
//                            register de-interleave), applies every predict / update step in registers (lift_regs), scales
//                            by the two norms, streams the detail half to its band in HBM and keeps the approximation
//                            on chip for the next level.  The segment stride is an odd number of 16-byte vectors, so the
//                            window loads are bank-conflict free.
//   inverse   k_lift1d_syn : mirror image: the a_K slice and the d_K .. d_1 slices arrive by TMA, each level normalises,
//                            runs the reversed steps in registers and writes the merged (interleaved) line with 16-byte
//                            stores -- to shared memory, or to HBM at level 1.
//
// Levels beyond K (the n/2^K-sample approximations: 1/2^K of the data) run on the generic one-level lifting passes.
// Arithmetic order is the reference's (lift_inbounds! / lift_perboundary!): STRICT keeps
// x + ((c0*a + c1*b)) for interior elements and ((x + c0*a) + c1*b) for elements whose taps wrap around the line end.
#include "fused1d_dev.cuh"
#include "tile2d_shapes.cuh"

#include <cstdlib>

namespace wb {
namespace l1 {

constexpr int MAXK1 = 8;

template <typename T> struct Geo {
    static constexpr int V = 16 / (int)sizeof(T);
    // forward: interleaved input, thread stride 2*SEG_A elements = odd number of vectors (f32: 12 floats = 3 vectors; f64: 10 doubles = 5)
    static constexpr int SEG_A = sizeof(T) == 4 ? 6 : 5;
    // inverse: planar inputs, thread stride SEG_S elements = odd number of vectors (f32: 12 floats = 3; f64: 6 doubles = 3)
    static constexpr int SEG_S = sizeof(T) == 4 ? 12 : 6;
    static constexpr int G = 8 / (int)sizeof(T);      // elements per 8-byte store
};
// halo in pairs per level, rounded up to even (keeps every staged range 16-byte granular)
template <class S> struct HaloE {
    static constexpr int L_ = Halo<S>::left(), R_ = Halo<S>::right();
    static constexpr int value = ((L_ > R_ ? L_ : R_) + 1) & ~1;
};

struct AnaPlanL {
    int K, tile;
    int E[MAXK1 + 1];      // halo (elements of a_l) kept on each side of the tile's slice of a_l; E[K] = 0
};
struct SynPlanL {
    int K, tile;
    int doff[MAXK1 + 1];   // element offset of the staged d_l slice (content: pairs [s_l - 8, s_l + T_l + 8))
    int aoff;              // staged a_K slice (same range convention)
    int poff, qoff;        // ping-pong buffers for a_{K-1} .. a_1 (content [s_l - 8, s_l + T_l + 8))
    int woff;              // per-warp output staging (32 * 2 * SEG elements per warp), or -1: direct 16-byte stores
};

template <typename T, int N> __device__ __forceinline__ void ldw(T (&w)[N], const T *p) {
    constexpr int V = 16 / (int)sizeof(T);
    static_assert(N % V == 0, "window must be whole vectors");
#pragma unroll
    for (int i = 0; i < N / V; ++i) {
        if constexpr (sizeof(T) == 4) {
            const float4 v = *reinterpret_cast<const float4 *>(p + 4 * i);
            w[4 * i] = v.x; w[4 * i + 1] = v.y; w[4 * i + 2] = v.z; w[4 * i + 3] = v.w;
        } else {
            const double2 v = *reinterpret_cast<const double2 *>(p + 2 * i);
            w[2 * i] = v.x; w[2 * i + 1] = v.y;
        }
    }
}
// 8-byte pieces: two floats or one double
__device__ __forceinline__ void st8(float *p, const float *v) { *reinterpret_cast<float2 *>(p) = make_float2(v[0], v[1]); }
__device__ __forceinline__ void st8(double *p, const double *v) { *p = v[0]; }
__device__ __forceinline__ void st8_cs(float *p, const float *v) { __stcs(reinterpret_cast<float2 *>(p), make_float2(v[0], v[1])); }
__device__ __forceinline__ void st8_cs(double *p, const double *v) { __stcs(p, v[0]); }
// 16-byte pieces
__device__ __forceinline__ void st16v(float *p, const float *v) { *reinterpret_cast<float4 *>(p) = make_float4(v[0], v[1], v[2], v[3]); }
__device__ __forceinline__ void st16v(double *p, const double *v) { *reinterpret_cast<double2 *>(p) = make_double2(v[0], v[1]); }
__device__ __forceinline__ void st16v_cs(float *p, const float *v) { __stcs(reinterpret_cast<float4 *>(p), make_float4(v[0], v[1], v[2], v[3])); }
__device__ __forceinline__ void st16v_cs(double *p, const double *v) { __stcs(reinterpret_cast<double2 *>(p), make_double2(v[0], v[1])); }

__device__ __forceinline__ int wrapm(int v, int n) { v %= n; return v < 0 ? v + n : v; }

// ===================================================================================================
// forward
// ===================================================================================================
// The input line is a_{lvl0} of a column (src + col*src_stride, ncur = n0 >> lvl0 samples; lvl0 = 0: x itself).  Details of
// local level l go to the d_{lvl0+l} band of y (y + col*n0 + (n0 >> (lvl0+l))); the level-K approximation goes to
// dst_a + col*dst_a_stride (y itself when no level remains, else the next stage's scratch).
template <typename T, class S, bool STRICT>
__global__ void __launch_bounds__(512)
k_lift1d_ana(const T *__restrict__ src, int64_t src_stride, T *__restrict__ y, int64_t n0, int lvl0,
             T *__restrict__ dst_a, int64_t dst_a_stride, const __grid_constant__ LiftCoefs<T> lc,
             const __grid_constant__ AnaPlanL pl, int pf) {
    using fp = FP<STRICT>;
    constexpr int SEG = Geo<T>::SEG_A, G = Geo<T>::G;
    constexpr int HM = HaloE<S>::value, NP = SEG + 2 * HM;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw);
    T *bufA = reinterpret_cast<T *>(smem_raw + 128);
    T *bufB = bufA + ((pl.tile + 2 * pl.E[0] + 2 * NP + 3) & ~3);
    const int64_t col = blockIdx.y;
    const int64_t s = (int64_t)blockIdx.x * pl.tile;
    const T *xc = src + col * src_stride;
    T *yc = y + col * n0;
    const int64_t ncur = n0 >> lvl0;
    const bool edge = STRICT && (blockIdx.x == 0 || blockIdx.x == gridDim.x - 1);

    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        const int count = pl.tile + 2 * pl.E[0];
        mbar_expect_tx(bar, (uint32_t)(count * sizeof(T)));
        tma_load_wrapped<T>(bufA, xc, s - pl.E[0], count, ncur, bar);
    }
    __syncthreads();
    if (pf > 0 && threadIdx.x == 0) {      // L2 prefetch of the tile the CTA `pf` launches further on will stage (fused1d.cu: k_ana_tiles)
        const unsigned lin = blockIdx.y * gridDim.x + blockIdx.x + (unsigned)pf;
        const unsigned pcol = lin / gridDim.x;
        if (pcol < gridDim.y)
            tma_prefetch_l2(src + (int64_t)pcol * src_stride + (int64_t)(lin - pcol * gridDim.x) * pl.tile, (uint32_t)(pl.tile * sizeof(T)));
    }
    mbar_wait(bar, 0);

    // per-warp staging of the detail outputs: a lane's segment is SEG consecutive coefficients, so direct stores would touch
    // every 32-byte sector of the warp's 32*SEG-element run three times (ncu: 3x excessive L2 sectors, profiles/r02c_lift1d_f32.md);
    // the warp parks the run in shared memory and writes it out as whole 16-byte pieces, lane after lane
    constexpr int V = Geo<T>::V;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    T *wst = bufB + ((((pl.tile >> 1) + 2 * pl.E[1] + 2 * NP + 3) & ~3)) + warp * (32 * SEG);
    const T *in = bufA;
    T *out = bufB;
    for (int l = 1; l <= pl.K; ++l) {
        const int Tl = pl.tile >> l;                    // owned pairs of this level
        const int El = pl.E[l];
        const int ncomp = Tl + 2 * El;                  // pairs computed: [s_l - E_l, s_l + T_l + E_l)
        const int half = (int)(ncur >> l);              // pairs of the whole line at this level
        const int sl = (int)(s >> l);
        T *dband = yc + (n0 >> (lvl0 + l)) + sl;        // d_{lvl0+l}[s_l ...]
        const bool last = (l == pl.K);
        T *dsta = dst_a + col * dst_a_stride + sl;
        // `in` starts HM pairs before the first computed pair
        const int g00 = sl - El - HM;
        for (int qb = warp * 32; qb * SEG < ncomp; qb += blockDim.x) {      // warp-uniform trip count
            const int q = qb + lane;
            if (q * SEG < ncomp) {
                T w[2 * NP];
                ldw<T, 2 * NP>(w, in + 2 * q * SEG);
                T sv[NP], dv[NP];
#pragma unroll
                for (int pp = 0; pp < NP; ++pp) { sv[pp] = w[2 * pp]; dv[pp] = w[2 * pp + 1]; }
                int g0 = g00 + q * SEG;
                if (edge) g0 = wrapm(g0, half);
                lift_regs<T, S, STRICT, NP>(sv, dv, lc, g0, half, edge);
#pragma unroll
                for (int pp = 0; pp < SEG; pp += G) {
                    const int p = q * SEG + pp;         // computed-range index of this piece
                    T os[G], od[G];
#pragma unroll
                    for (int e = 0; e < G; ++e) { os[e] = fp::mul(sv[HM + pp + e], lc.n1); od[e] = fp::mul(dv[HM + pp + e], lc.n2); }
                    st8(wst + lane * SEG + pp, od);
                    if (p < ncomp) {
                        const int po = p - El;          // owned-range index
                        if (!last) st8(out + p, os);
                        else if (po >= 0 && po < Tl) st8_cs(dsta + po, os);
                    }
                }
            }
            __syncwarp();
#pragma unroll
            for (int i = lane; i < 32 * SEG / V; i += 32) {
                const int p = qb * SEG + i * V;
                const int po = p - El;
                if (p < ncomp && po >= 0 && po < Tl) st16v_cs(dband + po, wst + i * V);
            }
            __syncwarp();
        }
        if (!last) {
            __syncthreads();
            const T *t = in; in = out; out = const_cast<T *>(t);
        }
    }
}

// ===================================================================================================
// inverse
// ===================================================================================================
// Produces a_{lvl0} of a column (ncur = n0 >> lvl0 samples, to dst + col*dst_stride: y when lvl0 == 0) from a_{lvl0+K}
// (asrc + col*asrc_stride) and the detail bands d_{lvl0+K} .. d_{lvl0+1} of x.
template <typename T, class S, bool STRICT>
__global__ void __launch_bounds__(512)
k_lift1d_syn(const T *__restrict__ asrc, int64_t asrc_stride, const T *__restrict__ x, int64_t n0, int lvl0,
             T *__restrict__ dst, int64_t dst_stride, const __grid_constant__ LiftCoefs<T> lc,
             const __grid_constant__ SynPlanL pl, int pf) {
    using fp = FP<STRICT>;
    constexpr int SEG = Geo<T>::SEG_S, V = Geo<T>::V;
    constexpr int HM = HaloE<S>::value, NP = SEG + 2 * HM, NW = SEG + 8;
    static_assert(NW % V == 0 && HM <= 4, "window of SEG + 8 elements from an aligned start");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw);
    T *sm = reinterpret_cast<T *>(smem_raw + 128);
    const int64_t col = blockIdx.y;
    const int64_t s = (int64_t)blockIdx.x * pl.tile;
    const T *xc = x + col * n0;
    const int K = pl.K;
    const int64_t ncur = n0 >> lvl0;
    const bool edge = STRICT && (blockIdx.x == 0 || blockIdx.x == gridDim.x - 1);

    // one mbarrier per level (bar[l] covers d_l; bar[K] also the approximation): coarse levels start while the large
    // fine-level detail slices are still in flight
    if (threadIdx.x == 0) {
        for (int l = 1; l <= K; ++l) mbar_init(bar + l, 1);
        for (int l = K; l >= 1; --l) {
            const int cnt = (pl.tile >> l) + 16;
            mbar_expect_tx(bar + l, (uint32_t)((l == K ? 2 : 1) * cnt * sizeof(T)));
            if (l == K) tma_load_wrapped<T>(sm + pl.aoff, asrc + col * asrc_stride, (s >> K) - 8, cnt, ncur >> K, bar + l);
            tma_load_wrapped<T>(sm + pl.doff[l], xc + (n0 >> (lvl0 + l)), (s >> l) - 8, cnt, ncur >> l, bar + l);
        }
    }
    __syncthreads();
    if (pf > 0 && threadIdx.x == 0) {      // L2 prefetch for the CTA `pf` launches further on; level l by every 2^(l-1)-th tile (fused1d.cu: k_syn_tiles)
        const unsigned lin = blockIdx.y * gridDim.x + blockIdx.x + (unsigned)pf;
        const unsigned pcol = lin / gridDim.x;
        if (pcol < gridDim.y) {
            const unsigned pt = lin - pcol * gridDim.x;
            const T *pxc = x + (int64_t)pcol * n0;
            for (int l = 1; l <= K; ++l) {
                const unsigned grp = 1u << ((l - 1) < 3 ? (l - 1) : 3);
                if (pt & (grp - 1)) continue;
                const int64_t len = ncur >> l;
                const int64_t lo = ((int64_t)pt * pl.tile) >> l;
                int64_t hi = lo + (int64_t)grp * (pl.tile >> l);
                if (hi > len) hi = len;
                if (hi > lo) tma_prefetch_l2(pxc + (n0 >> (lvl0 + l)) + lo, (uint32_t)((hi - lo) * sizeof(T)));
                if (l == K && hi > lo) tma_prefetch_l2(asrc + (int64_t)pcol * asrc_stride + lo, (uint32_t)((hi - lo) * sizeof(T)));
            }
        }
    }

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    T *wst = sm + (pl.woff >= 0 ? pl.woff : 0) + warp * (32 * 2 * SEG);
    const T *abuf = sm + pl.aoff;
    for (int l = K; l >= 1; --l) {
        mbar_wait(bar + l, 0);
        const int Tl = pl.tile >> l;
        const int half = (int)(ncur >> l);
        const int sl = (int)(s >> l);
        // computed pairs: [ulo, uhi) relative to s_l; buffers hold pairs [-8, T_l + 8)
        const int ulo = (l == 1) ? 0 : -4, uhi = (l == 1) ? Tl : Tl + 4;
        const int ncomp = uhi - ulo;
        const int woff = ulo + 4;                       // buffer index of pair (ulo - 4)
        const T *dbuf = sm + pl.doff[l];
        T *obuf = sm + (((l - 1) & 1) ? pl.poff : pl.qoff);
        T *og = dst + col * dst_stride + s;
        for (int qb = warp * 32; qb * SEG < ncomp; qb += blockDim.x) {      // warp-uniform trip count
            const int q = qb + lane;
            if (q * SEG < ncomp) {
                T wa[NW], wd[NW];
                ldw<T, NW>(wa, abuf + woff + q * SEG);
                ldw<T, NW>(wd, dbuf + woff + q * SEG);
                T sv[NP], dv[NP];
#pragma unroll
                for (int pp = 0; pp < NP; ++pp) { sv[pp] = fp::mul(wa[pp + 4 - HM], lc.n1); dv[pp] = fp::mul(wd[pp + 4 - HM], lc.n2); }
                int g0 = sl + ulo + q * SEG - HM;
                if (edge) g0 = wrapm(g0, half);
                lift_regs<T, S, STRICT, NP>(sv, dv, lc, g0, half, edge);
#pragma unroll
                for (int pp = 0; pp < SEG; pp += V / 2) {
                    const int u = q * SEG + pp;         // computed-range index of the first pair of this 16-byte piece
                    T o[V];
#pragma unroll
                    for (int e = 0; e < V / 2; ++e) { o[2 * e] = sv[HM + pp + e]; o[2 * e + 1] = dv[HM + pp + e]; }
                    if (l > 1) { if (u < ncomp) st16v(obuf + 2 * u, o); }    // content starts at sample 2 * ulo = -8
                    else if (pl.woff < 0) { if (u < ncomp) st16v_cs(og + 2 * u, o); }
                    else st16v(wst + lane * 2 * SEG + 2 * pp, o);
                }
            }
            if (l == 1 && pl.woff >= 0) {   // optional lane-contiguous copy-out through a per-warp staging run (WB200_LIFT1D_INV_STAGED)
                __syncwarp();
                for (int i = lane; i < 32 * 2 * SEG / V; i += 32) {
                    const int idx = 2 * qb * SEG + i * V;                   // sample index inside the tile
                    if (idx < 2 * ncomp) st16v_cs(og + idx, wst + i * V);
                }
                __syncwarp();
            }
        }
        if (l > 1) {
            __syncthreads();
            abuf = obuf;
        }
    }
}

// ===================================================================================================
// whole-line tail: the CTA owns a column of m samples (the approximation left by the tile stages) and runs every remaining
// level in shared memory, in place on the dyadic lattice (tail level r works on the samples whose indices are multiples of
// 2^(r-1)); any lifting scheme (runtime step table), any m >= 2 including lines shorter than a step's reach.
// ===================================================================================================
template <typename T, bool STRICT, bool FW>
__device__ __forceinline__ void tail_level(T *A, int nl, int st, const LiftScheme<T> &sc) {
    using fp = FP<STRICT>;
    const int half = nl >> 1;
    if (!FW) {   // normalize! precedes the steps on the inverse path
        for (int k = threadIdx.x; k < nl; k += blockDim.x) A[k * st] = fp::mul(A[k * st], (k & 1) ? sc.norm2 : sc.norm1);
        __syncthreads();
    }
    for (int sx = 0; sx < sc.nsteps; ++sx) {
        const int sh = sc.shift[sx], nc = sc.nc[sx];
        const bool pred = sc.is_predict[sx] != 0;
        const int left = sh > 0 ? sh : 0;
        const int opar = pred ? 1 : 0;
        for (int p = threadIdx.x; p < half; p += blockDim.x) {
            T *tg = A + (2 * p + (pred ? 0 : 1)) * st;
            T v = *tg;
            const bool interior = (p >= left) && (p <= half + sh - nc) && (nc <= 3);
            if (interior && nc > 1) {
                T acc = fp::mul(sc.coef[sx][0], A[(2 * (p - sh) + opar) * st]);
                for (int k = 1; k < nc; ++k) acc = fp::mac(acc, sc.coef[sx][k], A[(2 * (p + k - sh) + opar) * st]);
                v = fp::add(v, acc);
            } else {
                for (int k = 0; k < nc; ++k) {
                    int q = (p + k - sh) % half;
                    if (q < 0) q += half;
                    v = fp::mac(v, sc.coef[sx][k], A[(2 * q + opar) * st]);
                }
            }
            *tg = v;
        }
        __syncthreads();
    }
    if (FW) {
        for (int k = threadIdx.x; k < nl; k += blockDim.x) A[k * st] = fp::mul(A[k * st], (k & 1) ? sc.norm2 : sc.norm1);
        __syncthreads();
    }
}
// lattice index i of an m-sample line after `levels` levels -> offset inside the column's [a | d_L | ... ] layout
__device__ __forceinline__ int64_t tail_pos(int i, int levels, int64_t n, int lv0) {
    const int tz = i ? __ffs(i) - 1 : 31;
    if (tz >= levels) return i >> levels;                               // approximation sample
    return (n >> (lv0 + tz + 1)) + (i >> (tz + 1));                     // detail of level lv0 + tz + 1
}
template <typename T, bool STRICT>
__global__ void __launch_bounds__(256)
k_lift1d_tail_fwd(const T *__restrict__ src, int64_t src_stride, T *__restrict__ y, int64_t n, int m, int levels, int lv0,
                  const __grid_constant__ LiftScheme<T> sc) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *A = reinterpret_cast<T *>(smem_raw);
    const T *sp = src + (int64_t)blockIdx.x * src_stride;
    T *yc = y + (int64_t)blockIdx.x * n;
    for (int i = threadIdx.x; i < m; i += blockDim.x) A[i] = sp[i];
    __syncthreads();
    for (int r = 1; r <= levels; ++r) tail_level<T, STRICT, true>(A, m >> (r - 1), 1 << (r - 1), sc);
    for (int i = threadIdx.x; i < m; i += blockDim.x) yc[tail_pos(i, levels, n, lv0)] = A[i];
}
template <typename T, bool STRICT>
__global__ void __launch_bounds__(256)
k_lift1d_tail_inv(const T *__restrict__ x, int64_t n, T *__restrict__ dst, int64_t dst_stride, int m, int levels, int lv0,
                  const __grid_constant__ LiftScheme<T> sc) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *A = reinterpret_cast<T *>(smem_raw);
    const T *xc = x + (int64_t)blockIdx.x * n;
    T *dp = dst + (int64_t)blockIdx.x * dst_stride;
    for (int i = threadIdx.x; i < m; i += blockDim.x) A[i] = xc[tail_pos(i, levels, n, lv0)];
    __syncthreads();
    for (int r = levels; r >= 1; --r) tail_level<T, STRICT, false>(A, m >> (r - 1), 1 << (r - 1), sc);
    for (int i = threadIdx.x; i < m; i += blockDim.x) dp[i] = A[i];
}

// ===================================================================================================
// host side
// ===================================================================================================
static int env_l(const char *name, int dflt) {
    const char *v = std::getenv(name);
    return (v && *v) ? std::atoi(v) : dflt;
}
template <typename T> static void fill_lc(LiftCoefs<T> &lc, const LiftScheme<T> &sc) {
    for (int i = 0; i < 4; ++i)
        for (int k = 0; k < 2; ++k) lc.c[i][k] = (i < sc.nsteps && k < sc.nc[i]) ? sc.coef[i][k] : T(0);
    lc.n1 = sc.norm1; lc.n2 = sc.norm2;
}
template <typename T> static int shape_of(const LiftScheme<T> &sc, bool fw) {
    if (fw) {
        if (shape_matches<ShapeCdf97F>(sc)) return 1;
        if (shape_matches<ShapeHaarF>(sc)) return 2;
        if (shape_matches<ShapeDb2F>(sc)) return 3;
    } else {
        if (shape_matches<ShapeCdf97I>(sc)) return 1;
        if (shape_matches<ShapeHaarI>(sc)) return 2;
        if (shape_matches<ShapeDb2I>(sc)) return 3;
    }
    return 0;
}

struct StageL { int K, tile, lv0; };
struct PlanL {
    bool ok = false;
    int nstages = 0;
    StageL st[6];
    int lv = 0;            // levels done by the tile stages
    int64_t m = 0;         // line length left (n >> lv)
    bool tail = false;     // the remaining L - lv levels run in the whole-line tail kernel (else: generic passes)
};
// One tile stage: K fused levels out of a TILE-sample tile.  The cumulative halo 2*HM*(2^K - 1) per side stays below tile/8
// (forward), the deepest level keeps at least 16 pairs per tile, a line holds at least two tiles (a wrapped TMA piece never
// overlaps its tile).
template <typename T> static bool plan_stage(int64_t cur, int levels, int HM, bool fw, int &K, int &tile_out) {
    // defaults from the r02 sweep (tools/sweep_lift1d.py, profiles/r02_lift1d_sweep.md): small tiles, four levels per stage --
    // a CTA is a chain of dependent levels, so many small resident CTAs hide each other's barriers and TMA waits
    int64_t tile = fw ? env_l(sizeof(T) == 4 ? "WB200_LIFT1D_TILE_F32" : "WB200_LIFT1D_TILE_F64", sizeof(T) == 4 ? 4096 : 2048)
                      : env_l(sizeof(T) == 4 ? "WB200_LIFT1D_TILE_F32_INV" : "WB200_LIFT1D_TILE_F64_INV", sizeof(T) == 4 ? 4096 : 2048);   // interleaved A/B (tools/ab_lift1d.py): 4096 beats 2048 by 11 %
    const int64_t p2 = cur & (-cur);
    while (tile > p2) tile >>= 1;
    while (tile > cur / 2) tile >>= 1;
    if (tile < 256 || (cur * (int64_t)sizeof(T)) % 16 != 0) return false;
    const int kmax = env_l("WB200_LIFT1D_KMAX", 4);
    K = levels < kmax ? levels : kmax;
    if (K > MAXK1) K = MAXK1;
    while (K >= 1 && ((tile >> K) < 16 || (fw && 2 * HM * ((1 << K) - 1) > tile / 8))) --K;
    if (K < 1) return false;
    tile_out = (int)tile;
    return true;
}
template <typename T> static PlanL plan_l(int64_t n, int L, int HM, bool fw) {
    PlanL p;
    const int64_t tailmax = env_l("WB200_LIFT1D_TAILMAX", 2048);
    int64_t cur = n;
    int lv = 0;
    while (lv < L && cur > tailmax && p.nstages < 6) {
        int K, tile;
        if (!plan_stage<T>(cur, L - lv, HM, fw, K, tile)) break;
        p.st[p.nstages++] = StageL{K, tile, lv};
        cur >>= K;
        lv += K;
    }
    p.lv = lv; p.m = cur;
    p.tail = (lv < L) && (size_t)cur * sizeof(T) <= 96 * 1024 && cur >= 2;
    if (p.nstages == 0) {                                      // short line: the whole transform in the tail kernel (one launch)
        p.ok = p.tail && n >= 256;
        return p;
    }
    if (lv < L && p.st[0].K < 2) return p;                    // scratch of the later stages must fit the generic workspace plan
    p.ok = true;
    return p;
}
// L2 prefetch distance of the tile kernels in CTAs (WB200_LIFT1D_PREFETCH / WB200_LIFT1D_PREFETCH_INV; 0 = off), dropped when the grid
// does not fit 31 bits.  Same outcome as for the filter kernels (profiles/r02h_prefetch_ab.md): forward +7 % at ~900 CTAs ahead,
// inverse -14 % at every distance, so only the forward kernel prefetches by default.
static int lift_prefetch(bool fw, dim3 grid) {
    int v = fw ? env_l("WB200_LIFT1D_PREFETCH", 888) : env_l("WB200_LIFT1D_PREFETCH_INV", 0);
    if (v < 0 || (uint64_t)grid.x * grid.y + (uint64_t)v >= 0x7fffffffULL) v = 0;
    return v;
}
// threads per CTA: the level-1 segment count of a tile spread over whole rounds (a 256-thread CTA left a third of its lanes
// idle on the 342 level-1 segments of an 8192-sample synthesis tile)
static int block_for(int segments, const char *envname) {
    const int forced = env_l(envname, 0);
    if (forced >= 32 && forced <= 512) return forced & ~31;
    int rounds = (segments + 95) / 96;                     // ~96 threads per CTA (r02 sweep)
    if (rounds < 1) rounds = 1;
    int nt = ((segments + rounds - 1) / rounds + 31) & ~31;
    if (nt < 64) nt = 64;
    if (nt > 512) nt = 512;
    return nt;
}
template <typename T> static void make_ana(AnaPlanL &pl, const StageL &sg, int HM) {
    pl.K = sg.K; pl.tile = sg.tile;
    pl.E[sg.K] = 0;
    for (int l = sg.K; l >= 1; --l) pl.E[l - 1] = 2 * (pl.E[l] + HM);
}
template <typename T, int NPA> static size_t ana_smem(const AnaPlanL &pl, int nt) {
    const size_t a = ((size_t)pl.tile + 2 * pl.E[0] + 2 * NPA + 3) & ~(size_t)3;
    const size_t b = ((size_t)(pl.tile >> 1) + 2 * pl.E[1] + 2 * NPA + 3) & ~(size_t)3;
    return 128 + (a + b + (size_t)nt * Geo<T>::SEG_A) * sizeof(T);
}
template <typename T> static size_t make_syn(SynPlanL &pl, const StageL &sg, int nt) {
    constexpr int SEG = Geo<T>::SEG_S;
    pl.K = sg.K; pl.tile = sg.tile;
    size_t off = 0;
    auto span = [&](int l) { return (size_t)(((sg.tile >> l) + 16 + SEG + 8 + 3) & ~3); };   // content + one segment of slack for the last window
    pl.doff[0] = 0;
    for (int l = 1; l <= sg.K; ++l) { pl.doff[l] = (int)off; off += span(l); }
    pl.aoff = (int)off; off += span(sg.K);
    size_t psz = 0, qsz = 0;
    for (int l = 1; l < sg.K; ++l) {     // a_l (1 <= l < K) lives in poff (l odd) or qoff (l even)
        const size_t sz = span(l);
        if (l & 1) psz = sz > psz ? sz : psz; else qsz = sz > qsz ? sz : qsz;
    }
    pl.poff = (int)off; off += psz;
    pl.qoff = (int)off; off += qsz;
    pl.woff = -1;
    if (env_l("WB200_LIFT1D_INV_STAGED", 0)) { pl.woff = (int)off; off += (size_t)nt * 2 * SEG; }
    return 128 + off * sizeof(T);
}

// scratch layout (elements of T): the approximation handed from stage i to the next consumer lives in buffer i & 1; the
// generic remainder (only when the tail kernel cannot take the line) ping-pongs through two more buffers
template <typename T> struct ScratchL { size_t a[2] = {0, 0}, g1 = 0, g0 = 0, total = 0; };
template <typename T> static ScratchL<T> scratch_l(const PlanL &p, int64_t n, int64_t B, int L) {
    ScratchL<T> sc;
    auto al = [](size_t e) { return (e * sizeof(T) + 255) / 256 * 256 / sizeof(T); };
    size_t sz[2] = {0, 0};
    for (int i = 0; i < p.nstages; ++i) {
        const int lv1 = p.st[i].lv0 + p.st[i].K;
        if (i == p.nstages - 1 && lv1 == L) continue;          // the last stage writes y
        const size_t e = (size_t)(n >> lv1) * (size_t)B;
        if (e > sz[i & 1]) sz[i & 1] = e;
    }
    sc.a[0] = 0;
    sc.a[1] = al(sz[0]);
    size_t off = sc.a[1] + al(sz[1]);
    if (p.lv < L && !p.tail) {
        const size_t m = (size_t)p.m * (size_t)B;
        sc.g1 = off; off += al(m / 2);
        sc.g0 = off; off += al(m / 4);
    }
    sc.total = off;
    return sc;
}

// generic one-level passes for a remainder the tail kernel cannot hold: levels lv+1 .. L on the compact approximations
template <typename T>
static int32_t remainder_fw(const PassOp<T> &op, T *y, int64_t n, int64_t B, const T *aK, int K, int L, T *s1, T *s0) {
    auto lines = [&](T *p, int64_t len, int64_t bstride, View<T> &v, Extent &e) {
        e.len = len; e.n[0] = 1; e.n[1] = 1; e.n[2] = 1; e.n[3] = B;
        v.p = p; v.ls = 1; v.s[0] = 1; v.s[1] = 0; v.s[2] = 0; v.s[3] = bstride;
    };
    T *buf[2] = {s0, s1};          // a_{K+r} lives in buf[r & 1]: s1 (m/2 per column) for odd r, s0 (m/4) for even r
    for (int l = K + 1; l <= L; ++l) {
        const int64_t nin = n >> (l - 1), nout = n >> l;
        const int r = l - K;
        View<T> src, dlo, dhi; Extent e, e2;
        if (l == K + 1) lines(const_cast<T *>(aK), nin, nin, src, e);
        else            lines(buf[(r - 1) & 1], nin, nin, src, e);
        lines(y + nout, nout, n, dhi, e2);
        if (l == L) lines(y, nout, n, dlo, e2);
        else        lines(buf[r & 1], nout, nout, dlo, e2);
        View<const T> cs; cs.p = src.p; cs.ls = src.ls; for (int q = 0; q < 4; ++q) cs.s[q] = src.s[q];
        if (!op.analysis(cs, dlo, dhi, e)) return WB200_ECUDA;
    }
    return WB200_OK;
}
template <typename T>
static int32_t remainder_inv(const PassOp<T> &op, const T *x, int64_t n, int64_t B, T *aK, int K, int L, T *s1, T *s0) {
    auto lines = [&](T *p, int64_t len, int64_t bstride, View<T> &v, Extent &e) {
        e.len = len; e.n[0] = 1; e.n[1] = 1; e.n[2] = 1; e.n[3] = B;
        v.p = p; v.ls = 1; v.s[0] = 1; v.s[1] = 0; v.s[2] = 0; v.s[3] = bstride;
    };
    auto cv = [](const View<T> &v) { View<const T> c; c.p = v.p; c.ls = v.ls; for (int q = 0; q < 4; ++q) c.s[q] = v.s[q]; return c; };
    T *buf[2] = {s0, s1};          // a_{K+r} lives in buf[r & 1]
    const int64_t thr[4] = {0, 0, 0, 0};
    for (int l = L; l >= K + 1; --l) {
        const int64_t nout = n >> (l - 1), nin = n >> l;
        const int r = l - K;
        View<T> slo, shi, dst; Extent e, e2;
        if (l == L) lines(const_cast<T *>(x), nin, n, slo, e2);
        else        lines(buf[r & 1], nin, nin, slo, e2);
        lines(const_cast<T *>(x) + nin, nin, n, shi, e2);
        if (l == K + 1) lines(aK, nout, nout, dst, e);
        else            lines(buf[(r - 1) & 1], nout, nout, dst, e);
        if (!op.synthesis(cv(slo), cv(shi), cv(slo), thr, false, dst, e)) return WB200_ECUDA;
    }
    return WB200_OK;
}

template <typename T, class SF, class SI_, bool STRICT>
static int32_t run_l1(const PassOp<T> &op, T *y, const T *x, int64_t n, int64_t B, int L, bool fw, const PlanL &p,
                      T *scratch, const ScratchL<T> &so, cudaStream_t st) {
    LiftCoefs<T> lc;
    fill_lc<T>(lc, op.sc);
    auto abuf = [&](int i) -> T * { return scratch + so.a[i & 1]; };      // approximation after stage i
    const int ns = p.nstages;
    const bool rest = p.lv < L;
    // a_{p.lv}: compact (p.m samples per column) behind the tile stages; with no stage at all it is the array itself
    T *alast = rest ? (ns ? abuf(ns - 1) : (fw ? const_cast<T *>(x) : y)) : nullptr;
    const int64_t alast_stride = ns ? p.m : n;
    const size_t tail_smem = (size_t)p.m * sizeof(T);
    if (fw) {
        constexpr int HM = HaloE<SF>::value;
        for (int i = 0; i < ns; ++i) {
            const StageL &sg = p.st[i];
            AnaPlanL pl;
            make_ana<T>(pl, sg, HM);
            const int nt = block_for(((sg.tile >> 1) + 2 * pl.E[1] + Geo<T>::SEG_A - 1) / Geo<T>::SEG_A, "WB200_LIFT1D_NT");
            const size_t smem = ana_smem<T, Geo<T>::SEG_A + 2 * HM>(pl, nt);
            auto kern = k_lift1d_ana<T, SF, STRICT>;
            if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
                (void)cudaGetLastError(); set_error("cudaFuncSetAttribute(k_lift1d_ana) failed"); return WB200_ECUDA;
            }
            const int64_t ncur = n >> sg.lv0;
            const T *src = (i == 0) ? x : abuf(i - 1);
            const int64_t sstride = (i == 0) ? n : ncur;
            const bool to_y = (i == ns - 1) && !rest;
            T *dsta = to_y ? y : abuf(i);
            const int64_t dstride = to_y ? n : (ncur >> sg.K);
            dim3 grid((unsigned)(ncur / sg.tile), (unsigned)B);
            {
                LaunchScope scope("fused_lift1d_ana", st);
                kern<<<grid, nt, smem, st>>>(src, sstride, y, n, sg.lv0, dsta, dstride, lc, pl, lift_prefetch(true, grid));
            }
            if (!check_launch("fused_lift1d_ana")) return WB200_ECUDA;
        }
        if (!rest) return WB200_OK;
        if (p.tail) {
            auto kern = k_lift1d_tail_fwd<T, STRICT>;
            if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tail_smem) != cudaSuccess) {
                (void)cudaGetLastError(); set_error("cudaFuncSetAttribute(k_lift1d_tail_fwd) failed"); return WB200_ECUDA;
            }
            {
                LaunchScope scope("fused_lift1d_tail_fwd", st);
                kern<<<(unsigned)B, 256, tail_smem, st>>>(alast, alast_stride, y, n, (int)p.m, L - p.lv, p.lv, op.sc);
            }
            return check_launch("fused_lift1d_tail_fwd") ? WB200_OK : WB200_ECUDA;
        }
        return remainder_fw<T>(op, y, n, B, alast, p.lv, L, scratch + so.g1, scratch + so.g0);
    }
    if (rest) {
        if (p.tail) {
            auto kern = k_lift1d_tail_inv<T, STRICT>;
            if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tail_smem) != cudaSuccess) {
                (void)cudaGetLastError(); set_error("cudaFuncSetAttribute(k_lift1d_tail_inv) failed"); return WB200_ECUDA;
            }
            {
                LaunchScope scope("fused_lift1d_tail_inv", st);
                kern<<<(unsigned)B, 256, tail_smem, st>>>(x, n, alast, alast_stride, (int)p.m, L - p.lv, p.lv, op.sc);
            }
            if (!check_launch("fused_lift1d_tail_inv")) return WB200_ECUDA;
        } else {
            const int32_t rc = remainder_inv<T>(op, x, n, B, alast, p.lv, L, scratch + so.g1, scratch + so.g0);
            if (rc != WB200_OK) return rc;
        }
    }
    for (int i = ns - 1; i >= 0; --i) {
        const StageL &sg = p.st[i];
        SynPlanL pl;
        const int nt = block_for(((sg.tile >> 1) + Geo<T>::SEG_S - 1) / Geo<T>::SEG_S, "WB200_LIFT1D_NT_INV");
        const size_t smem = make_syn<T>(pl, sg, nt);
        auto kern = k_lift1d_syn<T, SI_, STRICT>;
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
            (void)cudaGetLastError(); set_error("cudaFuncSetAttribute(k_lift1d_syn) failed"); return WB200_ECUDA;
        }
        const int64_t ncur = n >> sg.lv0;
        const bool from_x = (i == ns - 1) && !rest;                        // the coarsest stage without a remainder reads a_L from x
        const T *asrc = from_x ? x : abuf(i);
        const int64_t astride = from_x ? n : (ncur >> sg.K);
        T *dst = (i == 0) ? y : abuf(i - 1);
        const int64_t dstride = (i == 0) ? n : ncur;
        dim3 grid((unsigned)(ncur / sg.tile), (unsigned)B);
        {
            LaunchScope scope("fused_lift1d_syn", st);
            kern<<<grid, nt, smem, st>>>(asrc, astride, x, n, sg.lv0, dst, dstride, lc, pl, lift_prefetch(false, grid));
        }
        if (!check_launch("fused_lift1d_syn")) return WB200_ECUDA;
    }
    return WB200_OK;
}

} // namespace l1

template <typename T>
int32_t fused_lift1d(const PassOp<T> &op, T *y, const T *x, const ArrayGeom &g, int L, bool fw,
                     void *workspace, size_t ws_bytes, cudaStream_t st) {
    if (!op.lifting || op.generic_only || g.ndim != 1 || g.C != 1 || L < 1) return -1;
    if (l1::env_l("WB200_DISABLE_LIFT1D", 0)) return -1;
    const int64_t n = g.dim[0], B = g.batch;
    if (n > ((int64_t)1 << 30) || B > 65535 || B < 1) return -1;
    if (((uintptr_t)x | (uintptr_t)y) & 15) return -1;
    const int id = l1::shape_of<T>(op.sc, fw);
    const int HM = (id == 2) ? 0 : 2;
    l1::PlanL p = l1::plan_l<T>(n, L, HM, fw);
    if (id == 0 && p.nstages > 0) {          // a scheme the tile kernels are not specialised for: tail kernel only (any step table)
        p = l1::PlanL();
        p.m = n; p.lv = 0;
        p.tail = (size_t)n * sizeof(T) <= 96 * 1024;
        p.ok = p.tail && n >= 256;
    }
    if (!p.ok) return -1;
    const bool inplace = (y == x) && p.nstages > 0;          // the tail kernel owns its whole column: in place as it stands
    const l1::ScratchL<T> so = l1::scratch_l<T>(p, n, B, L);
    const size_t copy_bytes = inplace ? (((size_t)(n * B) * sizeof(T) + 255) & ~(size_t)255) : 0;
    const size_t need = copy_bytes + so.total * sizeof(T);
    char *base = nullptr;
    bool own = false;
    if (need) {
        if (workspace != nullptr) {
            if (ws_bytes < need) { set_error("workspace too small: %zu bytes given, %zu needed", ws_bytes, need); return WB200_EWORKSPACE; }
            base = (char *)workspace;
        } else {
            if (scratch_alloc((void **)&base, need, st) != cudaSuccess) { (void)cudaGetLastError(); set_error("cudaMallocAsync(lift1d scratch) failed"); return WB200_ECUDA; }
            own = true;
        }
    }
    const T *xin = x;
    if (inplace) {   // tiles of a column read x while others write its detail bands: stage a copy
        if (cudaMemcpyAsync(base, x, (size_t)(n * B) * sizeof(T), cudaMemcpyDeviceToDevice, st) != cudaSuccess) {
            (void)cudaGetLastError(); set_error("cudaMemcpyAsync failed");
            if (own) cudaFreeAsync(base, st);
            return WB200_ECUDA;
        }
        xin = (const T *)base;
    }
    T *sc = (T *)(base + copy_bytes);
    int32_t rc;
#define WB_L1(SF, SI_) rc = op.strict ? l1::run_l1<T, SF, SI_, true>(op, y, xin, n, B, L, fw, p, sc, so, st) \
                                      : l1::run_l1<T, SF, SI_, false>(op, y, xin, n, B, L, fw, p, sc, so, st)
    switch (id) {
    case 1: WB_L1(ShapeCdf97F, ShapeCdf97I); break;
    case 2: WB_L1(ShapeHaarF, ShapeHaarI); break;
    default: WB_L1(ShapeDb2F, ShapeDb2I); break;
    }
#undef WB_L1
    if (own) cudaFreeAsync(base, st);
    return rc;
}

template int32_t fused_lift1d<float>(const PassOp<float> &, float *, const float *, const ArrayGeom &, int, bool, void *, size_t, cudaStream_t);
template int32_t fused_lift1d<double>(const PassOp<double> &, double *, const double *, const ArrayGeom &, int, bool, void *, size_t, cudaStream_t);

} // namespace wb
