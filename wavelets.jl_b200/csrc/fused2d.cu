// fused2d.cu -- one-launch-per-level 2-D lifting transform (cdf 9/7, Haar, db2 schemes) for sm_100a.
//
// The reference's 2-D lifting level (src/Transforms/transforms_lifting.jl:158-191) makes, per line and per
// dimension, a strided split pass, one pass per lifting step and a normalise/scatter pass -- its own KA GPU
// extension launches 12 kernels per cdf97 level, with the dim-2 pass uncoalesced.  Here one CTA
// stages a TI x TJ tile (+ the scheme's halo in both dimensions, periodic wrap folded into the load) in shared
// memory, de-interleaved along dim 1 while loading, and runs
//      forward:  dim-2 pass (rows)    -> dim-1 pass (columns) -> four coalesced quadrant stores
//      inverse:  dim-1 pass (columns) -> dim-2 pass (rows)    -> one coalesced interleaved store
// Each pass keeps a short segment of one line in REGISTERS and applies every predict/update step there (halo
// pairs are recomputed per segment), so a sample costs one shared-memory read and one write per pass instead of
// one per lifting step.  The shared row pitch is odd: the dim-2 pass walks rows with threads along dim 1
// (stride 1), the dim-1 pass walks along a row with threads across rows (stride = odd pitch): both conflict-free.
// A level therefore reads its input once and writes its output once: 2*sizeof(T) bytes per sample of the level.
//
// Arithmetic order is the reference's (lift_inbounds!/lift_perboundary!, normalize!): in
// STRICT mode interior elements use x + ((c0*a + c1*b)), elements whose taps wrap use ((x + c0*a) + c1*b).
#include "fused.cuh"
#include "tile2d_shapes.cuh"

#include <cstdlib>

namespace wb {

// ---------------------------------------------------------------------------------------------------
// tile configuration
// ---------------------------------------------------------------------------------------------------
template <class S, int TI_, int TJ_, int SI_, int SJ_> struct Cfg2d {
    static constexpr int TI = TI_, TJ = TJ_, SI = SI_, SJ = SJ_;
    static constexpr int HL = Halo<S>::left(), HR = Halo<S>::right();
    static constexpr int TIp = TI / 2, TJp = TJ / 2;
    static constexpr int CP = TIp + HL + HR;          // staged pairs along dim 1
    static constexpr int RJ = TJ + 2 * (HL + HR);     // staged rows (dim 2 samples)
    static constexpr int P = CP | 1;                  // odd pitch
    static constexpr int NPI = SI + HL + HR, NPJ = SJ + HL + HR;
    static constexpr int ROW_TASKS_F = 2 * CP * (TJp / SJ);   // forward dim-2 pass: every staged column pair
    static constexpr int COL_TASKS_F = TJ * (TIp / SI);       // forward dim-1 pass: owned rows only
    static constexpr int COL_TASKS_I = RJ * (TIp / SI);       // inverse dim-1 pass: every staged row
    static constexpr int ROW_TASKS_I = 2 * TIp * (TJp / SJ);  // inverse dim-2 pass: owned column pairs only
    static constexpr int MAXT = ROW_TASKS_F > COL_TASKS_I ? (ROW_TASKS_F > COL_TASKS_F ? ROW_TASKS_F : COL_TASKS_F)
                                                          : (COL_TASKS_I > ROW_TASKS_I ? COL_TASKS_I : ROW_TASKS_I);
    static constexpr int NT = (MAXT + 31) / 32 * 32;
    static_assert(TIp % SI == 0 && TJp % SJ == 0, "segments must tile the tile");
};


// one element, global -> shared, asynchronously (LDGSTS)
template <typename T> __device__ __forceinline__ void cp_async(T *dst_smem, const T *src) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst_smem);
    if constexpr (sizeof(T) == 4) asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(src) : "memory");
    else                          asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

// ---------------------------------------------------------------------------------------------------
// forward level: src (n x n corner, leading dimension ld_s) -> LL to `ll`, the three detail quadrants to `yd`
// ---------------------------------------------------------------------------------------------------
template <typename T, class S, bool STRICT, class C>
__global__ void __launch_bounds__(C::NT)
k_lift2d_fwd(const T *__restrict__ src, int64_t ld_s, int64_t bs_s, T *__restrict__ ll, int64_t ld_ll, int64_t bs_ll,
             T *__restrict__ yd, int64_t ld_y, int64_t bs_y, int n, const __grid_constant__ LiftCoefs<T> lc) {
    using fp = FP<STRICT>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *Se = reinterpret_cast<T *>(smem_raw);
    T *So = Se + C::RJ * C::P;
    const int nh = n >> 1;
    const int ip0 = blockIdx.x * C::TIp, j0 = blockIdx.y * C::TJ;
    const int64_t b = blockIdx.z;
    const T *sb = src + b * bs_s;
    const int tid = threadIdx.x;

    // ---- stage the tile with asynchronous copies (LDGSTS: global -> shared without a register round trip, so every
    //      thread keeps all of its copies in flight).  Even and odd dim-1 samples go to separate arrays (the
    //      polyphase split of the lifting scheme, fused into the load).  A thread owns one column pair and walks
    //      down the rows with incremental addressing.
    {
        constexpr int RT = C::NT / C::CP;                   // rows covered per sweep
        const int c = tid % C::CP, rsub = tid / C::CP;
        if (rsub < RT) {
            const int ipg = wrapi(ip0 - C::HL + c, nh);
            int j = wrapi(j0 - 2 * C::HL + rsub, n);
            const T *p = sb + (int64_t)j * ld_s + 2 * ipg;
            T *se = Se + rsub * C::P + c, *so = So + rsub * C::P + c;
            const int64_t pstep = (int64_t)RT * ld_s, pwrap = (int64_t)n * ld_s;
#pragma unroll 6
            for (int r = rsub; r < C::RJ; r += RT) {
                cp_async<T>(se, p);
                cp_async<T>(so, p + 1);
                se += RT * C::P; so += RT * C::P;
                j += RT; p += pstep;
                if (j >= n) { j -= n; p -= pwrap; }
            }
        }
        cp_async_wait_all();
    }
    __syncthreads();

    // ---- dim-2 pass (the reference's "rows"): lines run across the staged rows, one thread per (column pair, parity, segment)
    {
        T s[C::NPJ], d[C::NPJ];
        const bool act = tid < C::ROW_TASKS_F;
        int c = 0, q = 0;
        T *A = Se;
        if (act) {
            c = tid % C::CP;
            const int rest = tid / C::CP;
            A = (rest & 1) ? So : Se;
            q = rest >> 1;
#pragma unroll
            for (int pp = 0; pp < C::NPJ; ++pp) {
                s[pp] = A[(2 * (q * C::SJ + pp)) * C::P + c];
                d[pp] = A[(2 * (q * C::SJ + pp) + 1) * C::P + c];
            }
        }
        __syncthreads();
        if (act) {
            const int jp0 = (j0 >> 1) - C::HL + q * C::SJ;
            const bool edge = STRICT && (blockIdx.y == 0 || blockIdx.y == gridDim.y - 1);
            lift_regs<T, S, STRICT, C::NPJ>(s, d, lc, wrapi(jp0, nh), nh, edge);
#pragma unroll
            for (int pp = C::HL; pp < C::HL + C::SJ; ++pp) {
                A[(2 * (q * C::SJ + pp)) * C::P + c] = fp::mul(s[pp], lc.n1);
                A[(2 * (q * C::SJ + pp) + 1) * C::P + c] = fp::mul(d[pp], lc.n2);
            }
        }
        __syncthreads();
    }
    // ---- dim-1 pass ("columns"): lines run along a staged row, one thread per (owned row, segment) ----
    {
        T s[C::NPI], d[C::NPI];
        const bool act = tid < C::COL_TASKS_F;
        int r = 0, q = 0;
        if (act) {
            r = 2 * C::HL + tid % C::TJ;
            q = tid / C::TJ;
#pragma unroll
            for (int pp = 0; pp < C::NPI; ++pp) {
                s[pp] = Se[r * C::P + q * C::SI + pp];
                d[pp] = So[r * C::P + q * C::SI + pp];
            }
        }
        __syncthreads();
        if (act) {
            const bool edge = STRICT && (blockIdx.x == 0 || blockIdx.x == gridDim.x - 1);
            lift_regs<T, S, STRICT, C::NPI>(s, d, lc, wrapi(ip0 - C::HL + q * C::SI, nh), nh, edge);
#pragma unroll
            for (int pp = C::HL; pp < C::HL + C::SI; ++pp) {
                Se[r * C::P + q * C::SI + pp] = fp::mul(s[pp], lc.n1);
                So[r * C::P + q * C::SI + pp] = fp::mul(d[pp], lc.n2);
            }
        }
        __syncthreads();
    }
    // ---- stores: (s_i, s_j) -> LL, the other three combinations -> their quadrants of y.  A thread owns one
    //      (dim-1 pair, dim-1 parity, dim-2 parity) and walks the tile's dim-2 pairs: 128-byte coalesced rows.
    {
        T *llb = ll + b * bs_ll;
        T *yb = yd + b * bs_y;
        const int ipl = tid % C::TIp;
        const int rest = tid / C::TIp;
        if (rest < 4) {
            const int pi = rest & 1, pj = rest >> 1;
            const int jq0 = j0 >> 1;
            const T *srcp = (pi ? So : Se) + (2 * C::HL + pj) * C::P + C::HL + ipl;
            T *dp; int64_t dstep;
            if ((pi | pj) == 0) { dp = llb + (int64_t)jq0 * ld_ll + ip0 + ipl; dstep = ld_ll; }
            else                { dp = yb + (int64_t)(pj * nh + jq0) * ld_y + pi * nh + ip0 + ipl; dstep = ld_y; }
#pragma unroll 8
            for (int k = 0; k < C::TJp; ++k) {
                *dp = *srcp;
                srcp += 2 * C::P;
                dp += dstep;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// inverse level: LL from `ll`, detail quadrants from `xd` -> dst (n x n, interleaved)
// ---------------------------------------------------------------------------------------------------
template <typename T, class S, bool STRICT, class C>
__global__ void __launch_bounds__(C::NT)
k_lift2d_inv(const T *__restrict__ ll, int64_t ld_ll, int64_t bs_ll, const T *__restrict__ xd, int64_t ld_x, int64_t bs_x,
             T *__restrict__ dst, int64_t ld_d, int64_t bs_d, int n, const __grid_constant__ LiftCoefs<T> lc) {
    using fp = FP<STRICT>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *Se = reinterpret_cast<T *>(smem_raw);
    T *So = Se + C::RJ * C::P;
    const int nh = n >> 1;
    const int ip0 = blockIdx.x * C::TIp, j0 = blockIdx.y * C::TJ;
    const int64_t b = blockIdx.z;
    const T *llb = ll + b * bs_ll;
    const T *xb = xd + b * bs_x;
    const int tid = threadIdx.x;

    // ---- stage (asynchronous copies): staged row r is output column j = j0 - 2HL + r, i.e. band pj = j & 1 at
    //      index jq = j >> 1.  A thread owns one dim-1 pair and one dim-2 parity and walks the dim-2 pairs.
    {
        constexpr int RT = (C::NT / C::CP) & ~1;            // even, so that a thread keeps its dim-2 parity
        const int c = tid % C::CP, rsub = tid / C::CP;
        if (rsub < RT) {
            const int ipg = wrapi(ip0 - C::HL + c, nh);
            const int j = wrapi(j0 - 2 * C::HL + rsub, n);
            int jq = j >> 1;
            const int pj = j & 1;
            const T *pe, *po;
            int64_t estep, ostep;
            if (pj == 0) { pe = llb + (int64_t)jq * ld_ll + ipg; estep = ld_ll; }
            else         { pe = xb + (int64_t)(nh + jq) * ld_x + ipg; estep = ld_x; }
            po = xb + (int64_t)(pj * nh + jq) * ld_x + nh + ipg; ostep = ld_x;
            T *se = Se + rsub * C::P + c, *so = So + rsub * C::P + c;
#pragma unroll 6
            for (int r = rsub; r < C::RJ; r += RT) {
                cp_async<T>(se, pe);
                cp_async<T>(so, po);
                se += RT * C::P; so += RT * C::P;
                jq += RT / 2; pe += (RT / 2) * estep; po += (RT / 2) * ostep;
                if (jq >= nh) { jq -= nh; pe -= (int64_t)nh * estep; po -= (int64_t)nh * ostep; }
            }
        }
        cp_async_wait_all();
    }
    __syncthreads();
    // ---- dim-1 pass first (inverse order), on every staged row ----
    {
        T s[C::NPI], d[C::NPI];
        const bool act = tid < C::COL_TASKS_I;
        int r = 0, q = 0;
        if (act) {
            r = tid % C::RJ;
            q = tid / C::RJ;
#pragma unroll
            for (int pp = 0; pp < C::NPI; ++pp) {
                s[pp] = fp::mul(Se[r * C::P + q * C::SI + pp], lc.n1);   // normalize! (reciprocal norms) precedes the steps
                d[pp] = fp::mul(So[r * C::P + q * C::SI + pp], lc.n2);
            }
        }
        __syncthreads();
        if (act) {
            const bool edge = STRICT && (blockIdx.x == 0 || blockIdx.x == gridDim.x - 1);
            lift_regs<T, S, STRICT, C::NPI>(s, d, lc, wrapi(ip0 - C::HL + q * C::SI, nh), nh, edge);
#pragma unroll
            for (int pp = C::HL; pp < C::HL + C::SI; ++pp) {
                Se[r * C::P + q * C::SI + pp] = s[pp];
                So[r * C::P + q * C::SI + pp] = d[pp];
            }
        }
        __syncthreads();
    }
    // ---- dim-2 pass on the owned column pairs ----
    {
        T s[C::NPJ], d[C::NPJ];
        const bool act = tid < C::ROW_TASKS_I;
        int c = 0, q = 0;
        T *A = Se;
        if (act) {
            c = C::HL + tid % C::TIp;
            const int rest = tid / C::TIp;
            A = (rest & 1) ? So : Se;
            q = rest >> 1;
#pragma unroll
            for (int pp = 0; pp < C::NPJ; ++pp) {
                s[pp] = fp::mul(A[(2 * (q * C::SJ + pp)) * C::P + c], lc.n1);
                d[pp] = fp::mul(A[(2 * (q * C::SJ + pp) + 1) * C::P + c], lc.n2);
            }
        }
        __syncthreads();
        if (act) {
            const int jp0 = (j0 >> 1) - C::HL + q * C::SJ;
            const bool edge = STRICT && (blockIdx.y == 0 || blockIdx.y == gridDim.y - 1);
            lift_regs<T, S, STRICT, C::NPJ>(s, d, lc, wrapi(jp0, nh), nh, edge);
#pragma unroll
            for (int pp = C::HL; pp < C::HL + C::SJ; ++pp) {
                A[(2 * (q * C::SJ + pp)) * C::P + c] = s[pp];
                A[(2 * (q * C::SJ + pp) + 1) * C::P + c] = d[pp];
            }
        }
        __syncthreads();
    }
    // ---- merged store: out[2ip + {0,1}, j] = (Se, So)[row j][ip]; a thread owns one dim-1 pair and walks the rows
    {
        T *db = dst + b * bs_d;
        constexpr int RS = C::NT / C::TIp;
        const int ipl = tid % C::TIp, rsub = tid / C::TIp;
        if (rsub < RS) {
            const T *pe = Se + (2 * C::HL + rsub) * C::P + C::HL + ipl;
            const T *po = So + (2 * C::HL + rsub) * C::P + C::HL + ipl;
            T *p = db + (int64_t)(j0 + rsub) * ld_d + 2 * (ip0 + ipl);
            const int64_t pstep = (int64_t)RS * ld_d;
#pragma unroll 4
            for (int rr = rsub; rr < C::TJ; rr += RS) {
                if constexpr (sizeof(T) == 4) *reinterpret_cast<float2 *>(p) = make_float2(*pe, *po);
                else                          *reinterpret_cast<double2 *>(p) = make_double2(*pe, *po);
                pe += RS * C::P; po += RS * C::P; p += pstep;
            }
        }
    }
}

} // namespace wb
#include "fused2d_tma.cuh"
namespace wb {

// ---------------------------------------------------------------------------------------------------
// pyramid tail: every remaining level of an n x n (n <= 128) approximation in ONE launch, one CTA per image.
// In-place lifting on the dyadic lattice in shared memory (level l works on the samples whose indices are
// multiples of 2^(l-1)); any lifting scheme (runtime step table).  The coefficients are scattered to / gathered
// from their Mallat-pyramid positions by index arithmetic at the store / load.
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void pyr_pos(int i, int j, int n, int levels, int &oi, int &oj) {
    const int tzi = __ffs(i | n) - 1, tzj = __ffs(j | n) - 1;   // trailing zeros, capped by those of n (>= levels)
    const int tz = tzi < tzj ? tzi : tzj;
    if (tz >= levels) { oi = i >> levels; oj = j >> levels; return; }
    const int ii = i >> tz, jj = j >> tz;                        // lattice coordinates at level tz+1 (one of them odd)
    const int hh = (n >> tz) >> 1;
    oi = (ii & 1) * hh + (ii >> 1);
    oj = (jj & 1) * hh + (jj >> 1);
}

// one 1-level lifting pass over the lattice lines: line l = base + l*ls, sample k of a line at + k*ps
template <typename T, bool STRICT, bool FW>
__device__ __forceinline__ void tail_pass(T *A, int nl, int ls, int ps, bool line_fast, const LiftScheme<T> &sc) {
    using fp = FP<STRICT>;
    const int half = nl >> 1;
    const int total = nl * half;
    if (!FW) { // normalize! precedes the steps on the inverse path
        for (int idx = threadIdx.x; idx < nl * nl; idx += blockDim.x) {
            int line, k;
            if (line_fast) { line = idx % nl; k = idx / nl; } else { k = idx % nl; line = idx / nl; }
            T *e = A + line * ls + k * ps;
            *e = fp::mul(*e, (k & 1) ? sc.norm2 : sc.norm1);
        }
        __syncthreads();
    }
    for (int st = 0; st < sc.nsteps; ++st) {
        const int sh = sc.shift[st], nc = sc.nc[st];
        const bool pred = sc.is_predict[st] != 0;
        const int left = sh > 0 ? sh : 0;
        for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
            int line, p;
            if (line_fast) { line = idx % nl; p = idx / nl; } else { p = idx % half; line = idx / half; }
            T *L0 = A + line * ls;
            T *tg = L0 + (2 * p + (pred ? 0 : 1)) * ps;
            const int opar = pred ? 1 : 0;
            T v = *tg;
            const bool interior = (p >= left) && (p <= half + sh - nc) && (nc <= 3);
            if (interior && nc > 1) {
                T acc = fp::mul(sc.coef[st][0], L0[(2 * (p - sh) + opar) * ps]);
                for (int k = 1; k < nc; ++k) acc = fp::mac(acc, sc.coef[st][k], L0[(2 * (p + k - sh) + opar) * ps]);
                v = fp::add(v, acc);
            } else {
                for (int k = 0; k < nc; ++k) {
                    int q = (p + k - sh) % half;
                    if (q < 0) q += half;
                    v = fp::mac(v, sc.coef[st][k], L0[(2 * q + opar) * ps]);
                }
            }
            *tg = v;
        }
        __syncthreads();
    }
    if (FW) {
        for (int idx = threadIdx.x; idx < nl * nl; idx += blockDim.x) {
            int line, k;
            if (line_fast) { line = idx % nl; k = idx / nl; } else { k = idx % nl; line = idx / nl; }
            T *e = A + line * ls + k * ps;
            *e = fp::mul(*e, (k & 1) ? sc.norm2 : sc.norm1);
        }
        __syncthreads();
    }
}

template <typename T, bool STRICT>
__global__ void __launch_bounds__(512)
k_lift2d_tail_fwd(const T *__restrict__ src, int64_t ld_s, int64_t bs_s, T *__restrict__ y, int64_t ld_y, int64_t bs_y,
                  int n, int levels, const __grid_constant__ LiftScheme<T> sc) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *A = reinterpret_cast<T *>(smem_raw);
    const T *sb = src + (int64_t)blockIdx.x * bs_s;
    T *yb = y + (int64_t)blockIdx.x * bs_y;
    for (int idx = threadIdx.x; idx < n * n; idx += blockDim.x) {
        const int i = idx % n, j = idx / n;
        A[j * n + i] = sb[(int64_t)j * ld_s + i];
    }
    __syncthreads();
    for (int l = 1; l <= levels; ++l) {
        const int st = 1 << (l - 1), nl = n >> (l - 1);
        tail_pass<T, STRICT, true>(A, nl, st, st * n, true, sc);       // dim-2 lines (one per lattice row index i)
        tail_pass<T, STRICT, true>(A, nl, st * n, st, false, sc);      // dim-1 lines (one per lattice column j)
    }
    for (int idx = threadIdx.x; idx < n * n; idx += blockDim.x) {
        const int i = idx % n, j = idx / n;
        int oi, oj;
        pyr_pos(i, j, n, levels, oi, oj);
        yb[(int64_t)oj * ld_y + oi] = A[j * n + i];
    }
}

template <typename T, bool STRICT>
__global__ void __launch_bounds__(512)
k_lift2d_tail_inv(const T *__restrict__ x, int64_t ld_x, int64_t bs_x, T *__restrict__ dst, int64_t ld_d, int64_t bs_d,
                  int n, int levels, const __grid_constant__ LiftScheme<T> sc) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *A = reinterpret_cast<T *>(smem_raw);
    const T *xb = x + (int64_t)blockIdx.x * bs_x;
    T *db = dst + (int64_t)blockIdx.x * bs_d;
    for (int idx = threadIdx.x; idx < n * n; idx += blockDim.x) {
        const int i = idx % n, j = idx / n;
        int oi, oj;
        pyr_pos(i, j, n, levels, oi, oj);
        A[j * n + i] = xb[(int64_t)oj * ld_x + oi];
    }
    __syncthreads();
    for (int l = levels; l >= 1; --l) {
        const int st = 1 << (l - 1), nl = n >> (l - 1);
        tail_pass<T, STRICT, false>(A, nl, st * n, st, false, sc);     // inverse order: dim 1 first
        tail_pass<T, STRICT, false>(A, nl, st, st * n, true, sc);
    }
    for (int idx = threadIdx.x; idx < n * n; idx += blockDim.x) {
        const int i = idx % n, j = idx / n;
        db[(int64_t)j * ld_d + i] = A[j * n + i];
    }
}

// ---------------------------------------------------------------------------------------------------
// pyramid tail, register edition (the fused shapes, power-of-two corners 4 .. 64): a thread owns one whole LINE of the
// level's corner, holds it in registers, runs every predict / update step there with the periodic wrap and the
// interior / boundary distinction resolved at compile time, and writes the line back de-interleaved -- two barriers per
// level instead of one per lifting step, no strided lattice.  A 64 x 64 corner with six levels took 29 us in the lattice
// kernel above (one CTA per image: pure latency, which is what a single 4096^2 image -- BASELINE configs[2] as written --
// pays twice per dwt + idwt pair).
// ---------------------------------------------------------------------------------------------------
template <typename T, class S, bool STRICT, int NPAIR>
__device__ __forceinline__ void lift_line(T (&s)[NPAIR], T (&d)[NPAIR], const LiftCoefs<T> &lc) {
    using fp = FP<STRICT>;
#pragma unroll
    for (int st = 0; st < S::N; ++st) {
        const int sh = S::sh(st), nc = S::nc(st);
        const bool pred = S::pred(st) != 0;
        const int left = sh > 0 ? sh : 0;
#pragma unroll
        for (int p = 0; p < NPAIR; ++p) {       // a step reads only the OTHER band: in place
            T v = pred ? s[p] : d[p];
            const int i0 = ((p - sh) % NPAIR + NPAIR) % NPAIR, i1 = ((p + 1 - sh) % NPAIR + NPAIR) % NPAIR;
            const T t0 = pred ? d[i0] : s[i0];
            if (nc == 1) {
                v = fp::mac(v, lc.c[st][0], t0);
            } else {
                const T t1 = pred ? d[i1] : s[i1];
                const bool interior = (p >= left) && (p <= NPAIR + sh - nc);
                if (STRICT && interior) v = fp::add(v, fp::mac(fp::mul(lc.c[st][0], t0), lc.c[st][1], t1));
                else                    v = fp::mac(fp::mac(v, lc.c[st][0], t0), lc.c[st][1], t1);
            }
            if (pred) s[p] = v; else d[p] = v;
        }
    }
}
// one pass over the NL lines of an NL x NL corner; element k of line `line` sits at base[k * es]
template <typename T, class S, bool STRICT, bool FW, int NL>
__device__ __forceinline__ void tailfast_pass(T *A, int P, bool along_j, const LiftCoefs<T> &lc) {
    using fp = FP<STRICT>;
    if ((int)threadIdx.x < NL) {
        constexpr int NPAIR = NL / 2;
        T s[NPAIR], d[NPAIR];
        const int es = along_j ? P : 1;
        T *base = A + (along_j ? (int)threadIdx.x : (int)threadIdx.x * P);
        if (FW) {
#pragma unroll
            for (int p = 0; p < NPAIR; ++p) { s[p] = base[(2 * p) * es]; d[p] = base[(2 * p + 1) * es]; }
            lift_line<T, S, STRICT, NPAIR>(s, d, lc);
#pragma unroll
            for (int p = 0; p < NPAIR; ++p) { base[p * es] = fp::mul(s[p], lc.n1); base[(NPAIR + p) * es] = fp::mul(d[p], lc.n2); }
        } else {
#pragma unroll
            for (int p = 0; p < NPAIR; ++p) { s[p] = fp::mul(base[p * es], lc.n1); d[p] = fp::mul(base[(NPAIR + p) * es], lc.n2); }
            lift_line<T, S, STRICT, NPAIR>(s, d, lc);
#pragma unroll
            for (int p = 0; p < NPAIR; ++p) { base[(2 * p) * es] = s[p]; base[(2 * p + 1) * es] = d[p]; }
        }
    }
    __syncthreads();
}
template <typename T, class S, bool STRICT, bool FW, int NL>
__device__ __forceinline__ void tailfast_level(T *A, int P, const LiftCoefs<T> &lc) {
    if (FW) { tailfast_pass<T, S, STRICT, true, NL>(A, P, true, lc);  tailfast_pass<T, S, STRICT, true, NL>(A, P, false, lc); }   // dim 2, then dim 1
    else    { tailfast_pass<T, S, STRICT, false, NL>(A, P, false, lc); tailfast_pass<T, S, STRICT, false, NL>(A, P, true, lc); }  // dim 1, then dim 2
}
template <typename T, class S, bool STRICT, bool FW>
__global__ void __launch_bounds__(256)
k_lift2d_tailfast(const T *__restrict__ src, int64_t ld_s, int64_t bs_s, T *__restrict__ dst, int64_t ld_d, int64_t bs_d,
                  int nt, int levels, const __grid_constant__ LiftCoefs<T> lc) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *A = reinterpret_cast<T *>(smem_raw);
    const int P = nt + 1;                                      // odd pitch: a lane per row walks along dim 1 without bank conflicts
    const T *sb = src + (int64_t)blockIdx.x * bs_s;
    T *db = dst + (int64_t)blockIdx.x * bs_d;
    for (int idx = threadIdx.x; idx < nt * nt; idx += blockDim.x) {
        const int i = idx % nt, j = idx / nt;
        A[j * P + i] = sb[(int64_t)j * ld_s + i];
    }
    __syncthreads();
    for (int q = 0; q < levels; ++q) {
        const int l = FW ? q + 1 : levels - q;
        switch (nt >> (l - 1)) {
        case 64: tailfast_level<T, S, STRICT, FW, 64>(A, P, lc); break;
        case 32: tailfast_level<T, S, STRICT, FW, 32>(A, P, lc); break;
        case 16: tailfast_level<T, S, STRICT, FW, 16>(A, P, lc); break;
        case 8:  tailfast_level<T, S, STRICT, FW, 8>(A, P, lc); break;
        case 4:  tailfast_level<T, S, STRICT, FW, 4>(A, P, lc); break;
        default: tailfast_level<T, S, STRICT, FW, 2>(A, P, lc); break;
        }
    }
    for (int idx = threadIdx.x; idx < nt * nt; idx += blockDim.x) {
        const int i = idx % nt, j = idx / nt;
        db[(int64_t)j * ld_d + i] = A[j * P + i];
    }
}

// ===================================================================================================
// host side
// ===================================================================================================
static int env_int2(const char *name, int dflt) {
    const char *v = getenv(name);
    return (v && *v) ? atoi(v) : dflt;
}

template <typename T> static void fill_coefs(LiftCoefs<T> &lc, const LiftScheme<T> &sc) {
    for (int i = 0; i < 4; ++i)
        for (int k = 0; k < 2; ++k) lc.c[i][k] = (i < sc.nsteps && k < sc.nc[i]) ? sc.coef[i][k] : T(0);
    lc.n1 = sc.norm1; lc.n2 = sc.norm2;
}

template <typename T> struct Tile2d { static constexpr int TI = 128, TJ = 64; };
constexpr int TAIL2D_MAX = 64;    // largest corner handed to the pyramid-tail kernel (latency-bound: one CTA per image)

template <class S, typename T> using CfgFor = Cfg2d<S, Tile2d<T>::TI, Tile2d<T>::TJ, 16, 16>;

// 0: not a fused shape; 1: cdf97; 2: haar; 3: db2 (direction is implied by the scheme order)
template <typename T> static int shape_id(const LiftScheme<T> &sc, bool fw) {
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

template <typename T>
int fused2d_levels(const PassOp<T> &op, const ArrayGeom &g, int L, bool fw) {
    if (g.ndim != 2 || g.C != 1 || g.dim[0] != g.dim[1]) return 0;
    if (env_int2("WB200_DISABLE_FUSED2D", 0)) return 0;
    if (g.batch > 65535) return 0;
    if (!op.lifting) {       // orthogonal filter bank: tensor-map tile kernels only (fir2d_impl.cuh); the remainder is generic
        if (op.generic_only || env_int2("WB200_DISABLE_FIR2D", 0)) return 0;
        const int te = fir2d_tile_edge<T>(op.fc.F);
        int64_t n = g.dim[0];
        if (te == 0 || n > (int64_t)1 << 30 || !fir2d_available<T>()) return 0;
        int Lf = 0;
        while (Lf < L && n >= te && n % te == 0) { ++Lf; n >>= 1; }
        return Lf;
    }
    if (shape_id<T>(op.sc, fw) == 0) return 0;
    const int tmax = Tile2d<T>::TI > Tile2d<T>::TJ ? Tile2d<T>::TI : Tile2d<T>::TJ;
    int Lf = 0;
    int64_t n = g.dim[0];
    if (n > (int64_t)1 << 30) return 0;
    // corners of 128 and below belong to the pyramid-tail kernel (one launch for all of them)
    while (Lf < L && n >= tmax && n > TAIL2D_MAX && n % Tile2d<T>::TI == 0 && n % Tile2d<T>::TJ == 0) { ++Lf; n >>= 1; }
    return Lf;
}

template <typename T>
bool fused2d_tail_ok(const PassOp<T> &op, const ArrayGeom &g, int64_t nt, int levels) {
    if (!op.lifting || g.ndim != 2 || g.C != 1 || g.dim[0] != g.dim[1]) return false;
    if (env_int2("WB200_DISABLE_FUSED2D", 0) || env_int2("WB200_DISABLE_TAIL2D", 0)) return false;
    return levels >= 1 && nt >= 2 && nt <= TAIL2D_MAX && g.batch <= 0x7fffffff;
}

template <typename T>
int32_t fused2d_tail(const PassOp<T> &op, const T *src, int64_t ld_s, int64_t bs_s, T *dst, int64_t ld_d, int64_t bs_d,
                     int nt, int levels, int64_t B, bool fw, cudaStream_t st) {
    // register edition for the fused shapes on power-of-two corners
    const int sid = op.lifting ? shape_id<T>(op.sc, fw) : 0;
    if (sid != 0 && nt >= 4 && nt <= 64 && (nt & (nt - 1)) == 0 && levels >= 1 && (nt >> levels) >= 1 && !env_int2("WB200_DISABLE_TAILFAST", 0)) {
        LiftCoefs<T> lc;
        fill_coefs<T>(lc, op.sc);
        const size_t sm = (size_t)nt * (nt + 1) * sizeof(T);
#define WB_TF(SF, SI_)                                                                                               \
        {                                                                                                            \
            LaunchScope scope(fw ? "fused_lift2d_tail_fwd" : "fused_lift2d_tail_inv", st);                           \
            if (fw) { if (op.strict) k_lift2d_tailfast<T, SF, true, true><<<(unsigned)B, 256, sm, st>>>(src, ld_s, bs_s, dst, ld_d, bs_d, nt, levels, lc);   \
                      else           k_lift2d_tailfast<T, SF, false, true><<<(unsigned)B, 256, sm, st>>>(src, ld_s, bs_s, dst, ld_d, bs_d, nt, levels, lc); } \
            else    { if (op.strict) k_lift2d_tailfast<T, SI_, true, false><<<(unsigned)B, 256, sm, st>>>(src, ld_s, bs_s, dst, ld_d, bs_d, nt, levels, lc);  \
                      else           k_lift2d_tailfast<T, SI_, false, false><<<(unsigned)B, 256, sm, st>>>(src, ld_s, bs_s, dst, ld_d, bs_d, nt, levels, lc); } \
        }
        switch (sid) {
        case 1: WB_TF(ShapeCdf97F, ShapeCdf97I) break;
        case 2: WB_TF(ShapeHaarF, ShapeHaarI) break;
        default: WB_TF(ShapeDb2F, ShapeDb2I) break;
        }
#undef WB_TF
        return check_launch("fused_lift2d_tail(register edition)") ? WB200_OK : WB200_ECUDA;
    }
    const size_t smem = (size_t)nt * nt * sizeof(T);
#define WB_TAIL(KERN, NAME)                                                                                        \
    {                                                                                                              \
        auto kern = KERN;                                                                                          \
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {   \
            (void)cudaGetLastError(); set_error("cudaFuncSetAttribute(" NAME ") failed"); return WB200_ECUDA; }     \
        LaunchScope scope(NAME, st);                                                                               \
        kern<<<(unsigned)B, 512, smem, st>>>(src, ld_s, bs_s, dst, ld_d, bs_d, nt, levels, op.sc);                  \
    }
    if (fw) { if (op.strict) WB_TAIL((k_lift2d_tail_fwd<T, true>), "fused_lift2d_tail_fwd") else WB_TAIL((k_lift2d_tail_fwd<T, false>), "fused_lift2d_tail_fwd") }
    else    { if (op.strict) WB_TAIL((k_lift2d_tail_inv<T, true>), "fused_lift2d_tail_inv") else WB_TAIL((k_lift2d_tail_inv<T, false>), "fused_lift2d_tail_inv") }
#undef WB_TAIL
    return check_launch("fused_lift2d_tail") ? WB200_OK : WB200_ECUDA;
}

template <typename T> size_t fused2d_scratch_bytes(const ArrayGeom &g, int Lf) {
    // approximation ping-pong: LL_1 (n/2)^2 in buffer 0, LL_2 (n/4)^2 in buffer 1, ... (the last fused level writes y)
    if (Lf < 1) return 0;
    const size_t n = (size_t)g.dim[0];
    size_t b0 = (n / 2) * (n / 2) * (size_t)g.batch * sizeof(T);
    size_t b1 = (Lf >= 2) ? (n / 4) * (n / 4) * (size_t)g.batch * sizeof(T) : 0;
    return ((b0 + 255) & ~(size_t)255) + ((b1 + 255) & ~(size_t)255);
}

template <class S, typename T> using Cfg3For = Cfg3<T, S, Tile2d<T>::TI, Tile2d<T>::TJ, 16, 16>;

template <typename T, class S, bool STRICT, bool FW>
static int32_t launch_level(const T *a, int64_t lda, int64_t bsa, const T *xd, int64_t ldx, int64_t bsx,
                            T *o1, int64_t ld1, int64_t bs1, T *o2, int64_t ld2, int64_t bs2,
                            int n, int64_t B, const LiftCoefs<T> &lc, cudaStream_t st) {
    // ---- preferred: tensor-map TMA tiles (persistent double-buffered CTAs when there are enough tiles) ----
    // the TMA kernels move 16-byte pieces to and from global memory: every base and leading dimension must keep that
    // alignment (always true for the library's scratch and for arrays from a CUDA allocator; a caller's odd sub-view
    // takes the cp.async kernels below, which use element-wise global accesses)
    auto al16 = [](const void *q, int64_t ld, int64_t bs) {
        return q == nullptr || ((reinterpret_cast<uintptr_t>(q) & 15) == 0 && (ld * (int64_t)sizeof(T)) % 16 == 0 && (bs * (int64_t)sizeof(T)) % 16 == 0);
    };
    const bool aligned = al16(a, lda, bsa) && al16(xd, ldx, bsx) && al16(o1, ld1, bs1) && al16(o2, ld2, bs2);
    if (aligned && env_int2("WB200_LIFT2D_TMA", 1)) {
        using C3 = Cfg3For<S, T>;
        const int nx = n / C3::TI, ny = n / C3::TJ;
        const int64_t ntiles64 = (int64_t)nx * ny * B;
        if (ntiles64 <= 0x7fffffffLL) {
            const int ntiles = (int)ntiles64;
            int dev = 0, nsm = 148;
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
            const size_t one_f = C3::SMEM_F - 128, one_i = C3::SMEM_I - 128;
            if constexpr (FW) {
                TensorMap tm;
                if (make_tensor_map<T>(tm, a, n, n, B, lda, bsa, C3::PI, C3::RJ)) {
                    const int per_sm = (int)((220 * 1024) / (2 * one_f + 128));
                    const bool persist = env_int2("WB200_LIFT2D_PERSIST", 0) && per_sm >= 1 && ntiles >= 4 * nsm * per_sm;
                    const size_t smem = 128 + (persist ? 2 : 1) * one_f;
                    const int grid = persist ? nsm * per_sm : ntiles;
                    if (persist) {
                        auto kern = k_lift2d_fwd_tma_p<T, S, STRICT, C3, 2>;
                        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) == cudaSuccess) {
                            {
                                LaunchScope scope("fused_lift2d_fwd", st);
                                kern<<<grid, C3::NT, smem, st>>>(tm, a, lda, bsa, o1, ld1, bs1, o2, ld2, bs2, n, nx, ny, ntiles, lc);
                            }
                            return check_launch("fused_lift2d_fwd(tma,persistent)") ? WB200_OK : WB200_ECUDA;
                        }
                    } else {
                        auto kern = k_lift2d_fwd_tma<T, S, STRICT, C3>;
                        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) == cudaSuccess) {
                            {
                                LaunchScope scope("fused_lift2d_fwd", st);
                                kern<<<dim3((unsigned)nx, (unsigned)ny, (unsigned)B), C3::NT, smem, st>>>(tm, a, lda, bsa, o1, ld1, bs1, o2, ld2, bs2, n, lc);
                            }
                            return check_launch("fused_lift2d_fwd(tma)") ? WB200_OK : WB200_ECUDA;
                        }
                    }
                    (void)cudaGetLastError();
                }
            } else {
                TensorMap tml, tmx;
                const int nh = n / 2;
                if (make_tensor_map<T>(tml, a, nh, nh, B, lda, bsa, C3::PC, C3::JQ) &&
                    make_tensor_map<T>(tmx, xd, n, n, B, ldx, bsx, C3::PC, C3::JQ)) {
                    const int per_sm = (int)((220 * 1024) / (2 * one_i + 128));
                    const bool persist = env_int2("WB200_LIFT2D_PERSIST", 0) && per_sm >= 1 && ntiles >= 4 * nsm * per_sm;
                    const size_t smem = 128 + (persist ? 2 : 1) * one_i;
                    const int grid = persist ? nsm * per_sm : ntiles;
                    if (persist) {
                        auto kern = k_lift2d_inv_tma_p<T, S, STRICT, C3, 2>;
                        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) == cudaSuccess) {
                            {
                                LaunchScope scope("fused_lift2d_inv", st);
                                kern<<<grid, C3::NT, smem, st>>>(tml, tmx, a, lda, bsa, xd, ldx, bsx, o1, ld1, bs1, n, nx, ny, ntiles, lc);
                            }
                            return check_launch("fused_lift2d_inv(tma,persistent)") ? WB200_OK : WB200_ECUDA;
                        }
                    } else {
                        auto kern = k_lift2d_inv_tma<T, S, STRICT, C3>;
                        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) == cudaSuccess) {
                            {
                                LaunchScope scope("fused_lift2d_inv", st);
                                kern<<<dim3((unsigned)nx, (unsigned)ny, (unsigned)B), C3::NT, smem, st>>>(tml, tmx, a, lda, bsa, xd, ldx, bsx, o1, ld1, bs1, n, lc);
                            }
                            return check_launch("fused_lift2d_inv(tma)") ? WB200_OK : WB200_ECUDA;
                        }
                    }
                    (void)cudaGetLastError();
                }
            }
        }
    }
    // ---- fallback: cp.async staging ----
    using C = CfgFor<S, T>;
    const size_t smem = (size_t)2 * C::RJ * C::P * sizeof(T);
    dim3 grid((unsigned)(n / C::TI), (unsigned)(n / C::TJ), (unsigned)B);
    if constexpr (FW) {
        auto kern = k_lift2d_fwd<T, S, STRICT, C>;
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { (void)cudaGetLastError(); set_error("cudaFuncSetAttribute(k_lift2d_fwd) failed"); return WB200_ECUDA; }
        LaunchScope scope("fused_lift2d_fwd", st);
        kern<<<grid, C::NT, smem, st>>>(a, lda, bsa, o1, ld1, bs1, o2, ld2, bs2, n, lc);
    } else {
        auto kern = k_lift2d_inv<T, S, STRICT, C>;
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { (void)cudaGetLastError(); set_error("cudaFuncSetAttribute(k_lift2d_inv) failed"); return WB200_ECUDA; }
        LaunchScope scope("fused_lift2d_inv", st);
        kern<<<grid, C::NT, smem, st>>>(a, lda, bsa, xd, ldx, bsx, o1, ld1, bs1, n, lc);
    }
    return check_launch(FW ? "fused_lift2d_fwd" : "fused_lift2d_inv") ? WB200_OK : WB200_ECUDA;
}

template <typename T, class SF, class SI_, bool STRICT>
static int32_t run2d(const PassOp<T> &op, T *y, const T *x, const T *ll_src, int64_t ll_ld, int64_t ll_bs,
                     const ArrayGeom &g, int Lf, bool fw, void *scratch, cudaStream_t st, bool ll_to_scratch) {
    const int64_t N = g.dim[0], B = g.batch;
    const int64_t bsN = N * N;
    LiftCoefs<T> lc;
    fill_coefs<T>(lc, op.sc);
    T *buf[2];
    buf[0] = (T *)scratch;
    const size_t b0 = (((size_t)(N / 2) * (N / 2) * B * sizeof(T)) + 255) & ~(size_t)255;
    buf[1] = (T *)((char *)scratch + b0);
    // approximation of level l (1 <= l < Lf) lives compactly in buf[(l-1)&1] with leading dimension N>>l
    if (fw) {
        for (int l = 1; l <= Lf; ++l) {
            const int n = (int)(N >> (l - 1));
            const T *src = (l == 1) ? x : buf[(l - 2) & 1];
            const int64_t lds = (l == 1) ? N : n, bss = (l == 1) ? bsN : (int64_t)n * n;
            T *llo; int64_t ldl, bsl;
            if (l == Lf && !ll_to_scratch) { llo = y; ldl = N; bsl = bsN; }     // final approximation: y's corner
            else         { llo = buf[(l - 1) & 1]; ldl = n / 2; bsl = (int64_t)(n / 2) * (n / 2); }
            int32_t rc = launch_level<T, SF, STRICT, true>(src, lds, bss, nullptr, 0, 0, llo, ldl, bsl, y, N, bsN, n, B, lc, st);
            if (rc != WB200_OK) return rc;
        }
    } else {
        for (int l = Lf; l >= 1; --l) {
            const int n = (int)(N >> (l - 1));
            const T *lls; int64_t ldl, bsl;
            if (l == Lf) { lls = ll_src; ldl = ll_ld; bsl = ll_bs; }            // x / y corner, or a parked compact copy
            else         { lls = buf[(l - 1) & 1]; ldl = n / 2; bsl = (int64_t)(n / 2) * (n / 2); }
            T *dst; int64_t ldd, bsd;
            if (l == 1) { dst = y; ldd = N; bsd = bsN; }
            else        { dst = buf[(l - 2) & 1]; ldd = n; bsd = (int64_t)n * n; }
            int32_t rc = launch_level<T, SI_, STRICT, false>(lls, ldl, bsl, x, N, bsN, dst, ldd, bsd, nullptr, 0, 0, n, B, lc, st);
            if (rc != WB200_OK) return rc;
        }
    }
    return WB200_OK;
}

template <typename T>
int32_t fused2d_run(const PassOp<T> &op, T *y, const T *x, const T *ll_src, int64_t ll_ld, int64_t ll_bs,
                    const ArrayGeom &g, int Lf, bool fw, void *scratch, cudaStream_t st, bool ll_to_scratch) {
    if (!op.lifting) return fir2d_run<T>(op, y, x, ll_src, ll_ld, ll_bs, g, Lf, fw, scratch, st, ll_to_scratch);
    const int id = shape_id<T>(op.sc, fw);
#define WB_RUN(SF, SI_)                                                                                            \
    return op.strict ? run2d<T, SF, SI_, true>(op, y, x, ll_src, ll_ld, ll_bs, g, Lf, fw, scratch, st, ll_to_scratch)   \
                     : run2d<T, SF, SI_, false>(op, y, x, ll_src, ll_ld, ll_bs, g, Lf, fw, scratch, st, ll_to_scratch)
    switch (id) {
    case 1: WB_RUN(ShapeCdf97F, ShapeCdf97I);
    case 2: WB_RUN(ShapeHaarF, ShapeHaarI);
    case 3: WB_RUN(ShapeDb2F, ShapeDb2I);
    default: return -1;
    }
#undef WB_RUN
}

template <typename T>
int32_t lift2d_level(const PassOp<T> &op, bool fw, const T *a, int64_t lda, int64_t bsa, const T *xd, int64_t ldx, int64_t bsx,
                     T *o1, int64_t ld1, int64_t bs1, T *o2, int64_t ld2, int64_t bs2, int n, int64_t B, cudaStream_t st) {
    LiftCoefs<T> lc;
    fill_coefs<T>(lc, op.sc);
    const int id = shape_id<T>(op.sc, fw);
#define WB_L2(SF, SI_)                                                                                                    \
    if (fw) return op.strict ? launch_level<T, SF, true, true>(a, lda, bsa, xd, ldx, bsx, o1, ld1, bs1, o2, ld2, bs2, n, B, lc, st)    \
                             : launch_level<T, SF, false, true>(a, lda, bsa, xd, ldx, bsx, o1, ld1, bs1, o2, ld2, bs2, n, B, lc, st);  \
    return op.strict ? launch_level<T, SI_, true, false>(a, lda, bsa, xd, ldx, bsx, o1, ld1, bs1, o2, ld2, bs2, n, B, lc, st)         \
                     : launch_level<T, SI_, false, false>(a, lda, bsa, xd, ldx, bsx, o1, ld1, bs1, o2, ld2, bs2, n, B, lc, st)
    switch (id) {
    case 1: WB_L2(ShapeCdf97F, ShapeCdf97I);
    case 2: WB_L2(ShapeHaarF, ShapeHaarI);
    case 3: WB_L2(ShapeDb2F, ShapeDb2I);
    default: break;
    }
#undef WB_L2
    set_error("internal: lift2d_level called for a scheme the fused kernels do not know");
    return WB200_EARG;
}

#define WB_INST(T)                                                                                                 \
    template int32_t lift2d_level<T>(const PassOp<T> &, bool, const T *, int64_t, int64_t, const T *, int64_t, int64_t, T *, int64_t, int64_t, T *, int64_t, int64_t, int, int64_t, cudaStream_t); \
    template int fused2d_levels<T>(const PassOp<T> &, const ArrayGeom &, int, bool);                                \
    template size_t fused2d_scratch_bytes<T>(const ArrayGeom &, int);                                               \
    template int32_t fused2d_run<T>(const PassOp<T> &, T *, const T *, const T *, int64_t, int64_t, const ArrayGeom &, int, bool, void *, cudaStream_t, bool); \
    template bool fused2d_tail_ok<T>(const PassOp<T> &, const ArrayGeom &, int64_t, int);                          \
    template int32_t fused2d_tail<T>(const PassOp<T> &, const T *, int64_t, int64_t, T *, int64_t, int64_t, int, int, int64_t, bool, cudaStream_t);
WB_INST(float)
WB_INST(double)
#undef WB_INST

} // namespace wb
