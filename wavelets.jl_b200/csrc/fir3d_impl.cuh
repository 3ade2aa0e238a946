// fir3d_impl.cuh -- ONE-PASS 3-D orthogonal filter-bank level (even F <= 20): a level reads its corner once and writes its
// eight octants once, instead of the reference's three strided line passes (src/Transforms/transforms_filter.jl:246-288;
// forward order dim 3 -> dim 2 -> dim 1, inverse dim 1 -> dim 2 -> dim 3, kept here bit for bit in STRICT mode).
//
// A 3-D tile with a halo in all three dimensions does not fit shared memory for long filters (db6: 10 samples per side and
// dimension), so a CTA MARCHES along one dimension and keeps the F most recent slabs of that dimension's input in a register
// ring (the decimating FIR along the marching dimension needs no halo at all: each slab is consumed exactly once):
//
//   forward  k_fir3d_fwd   march along dim 2.  Per step two (dim 1 x dim 3) slabs j, j+1 arrive in shared memory
//                          (16-byte cp.async, periodic wrap resolved per chunk).  Phase A: a thread owns one dim-1
//                          position and 2*PP dim-3 outputs; it runs the dim-3 analysis of both slabs straight from shared
//                          memory into its register ring (F slabs x 2*PP positions) and, from the ring, the dim-2 analysis
//                          (a2[u] and d2[u + F/2 - 1] read the same F slabs -- the window trick of fused1d.cu).
//                          Phase B: the dim-1 analysis of the two finished slabs on 16-byte shared-memory windows, results
//                          leave as 16-byte stores into the eight octants.
//   inverse  k_fir3d_inv   the mirror image, marching along dim 3: per step the (dim 1 x dim 2) planes A3[s], D3[s + F/2 - 1]
//                          arrive.  Phase A: dim-1 synthesis on 16-byte windows.  Phase B: a thread owns one dim-1 position
//                          and 2*PP dim-2 outputs: dim-2 synthesis from shared memory into the register rings (F/2 planes of
//                          each band), dim-3 synthesis from the rings, coalesced stores of the two finished output planes.
//
// Arithmetic = filtdown!/filtup! closed forms in the reference's summation order; STRICT keeps multiply
// and add separately rounded.  One launch per level covers the whole batch and every octant (LLL included).
#pragma once
#include "fused.cuh"
#include "tile2d_shapes.cuh"

#include <cstdlib>

namespace wb {
namespace f3 {

__device__ __forceinline__ void cp_async16(void *dst_smem, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void cp_wait_1() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }

template <typename T> struct Vt;
template <> struct Vt<float> { using type = float4; };
template <> struct Vt<double> { using type = double2; };

template <typename T, int N> __device__ __forceinline__ void ld16(T (&w)[N], const T *p) {
    constexpr int V = 16 / (int)sizeof(T);
    static_assert(N % V == 0, "window must be whole 16-byte vectors");
#pragma unroll
    for (int i = 0; i < N / V; ++i) {
        const typename Vt<T>::type v = *reinterpret_cast<const typename Vt<T>::type *>(p + i * V);
        if constexpr (sizeof(T) == 4) { w[4 * i] = v.x; w[4 * i + 1] = v.y; w[4 * i + 2] = v.z; w[4 * i + 3] = v.w; }
        else                          { w[2 * i] = v.x; w[2 * i + 1] = v.y; }
    }
}
template <typename T, int N> __device__ __forceinline__ void st16(T *p, const T (&w)[N]) {
    constexpr int V = 16 / (int)sizeof(T);
#pragma unroll
    for (int i = 0; i < N / V; ++i) {
        typename Vt<T>::type v;
        if constexpr (sizeof(T) == 4) { v.x = w[4 * i]; v.y = w[4 * i + 1]; v.z = w[4 * i + 2]; v.w = w[4 * i + 3]; }
        else                          { v.x = w[2 * i]; v.y = w[2 * i + 1]; }
        *reinterpret_cast<typename Vt<T>::type *>(p + i * V) = v;
    }
}
// streaming 16-byte global stores (detail octants are written once and never re-read by this library's next launch)
template <typename T, int N> __device__ __forceinline__ void st16_cs(T *p, const T (&w)[N]) {
    constexpr int V = 16 / (int)sizeof(T);
#pragma unroll
    for (int i = 0; i < N / V; ++i) {
        typename Vt<T>::type v;
        if constexpr (sizeof(T) == 4) { v.x = w[4 * i]; v.y = w[4 * i + 1]; v.z = w[4 * i + 2]; v.w = w[4 * i + 3]; }
        else                          { v.x = w[2 * i]; v.y = w[2 * i + 1]; }
        __stcs(reinterpret_cast<typename Vt<T>::type *>(p + i * V), v);
    }
}

constexpr int cmax(int a, int b) { return a > b ? a : b; }
constexpr int rup(int a, int m) { return (a + m - 1) / m * m; }

// ---------------------------------------------------------------------------------------------------
// geometry.  TI: owned dim-1 samples per CTA; PP: output PAIRS per thread along the strided in-slab dimension
// (dim 3 forward, dim 2 inverse); NG: thread groups along it -> TO = 2*PP*NG owned samples; SI: pairs per dim-1 task.
// ---------------------------------------------------------------------------------------------------
template <typename T, int F_, int TI_, int PP_, int NG_, int SI_> struct Cfg {
    static constexpr int F = F_, TI = TI_, PP = PP_, NG = NG_, SI = SI_;
    static constexpr int V = 16 / (int)sizeof(T);
    static constexpr int Q = F / 2, H = Q - 1;              // taps per polyphase branch; halo in pairs
    static constexpr int TIp = TI / 2;
    static constexpr int TO = 2 * PP * NG, TOp = TO / 2;
    static constexpr int KR = TO + F - 2;                    // staged rows per slab (both directions)
    static constexpr int NSEG = TIp / SI;                    // dim-1 tasks per row
    static constexpr int NP = SI + 2 * H;                    // pairs a dim-1 task holds in registers
    // ---- forward: rows of HS + TI + HS samples (halo rounded to whole vectors)
    static constexpr int HS = rup(F - 2, V);
    static constexpr int WIS = HS + TI + HS;
    static constexpr int CPR_F = WIS / V;                    // 16-byte chunks per staged row
    static constexpr int PIN_F = WIS;                        // input rows: scalar access along dim 1 only
    static constexpr int PO = ((WIS / V) | 1) * V;           // marching-output rows: odd number of vectors (16-byte windows, 8 rows apart)
    static constexpr int OFF_F = HS - (F - 2);               // samples between a window's aligned start and its first pair
    static constexpr int WIN_F = rup(OFF_F + 2 * NP, V);
    static constexpr int NT_F = rup(WIS * NG, 32);
    static constexpr int STAGE_F = 2 * KR * PIN_F;           // two slabs
    static constexpr int OBUF_F = 2 * TO * PO;               // two output slabs
    static constexpr size_t SMEM_F = (size_t)(2 * STAGE_F + 2 * OBUF_F) * sizeof(T);
    // ---- inverse: a row is [a1 piece | d1 piece], each HA + TIp pairs
    static constexpr int HA = rup(H, V);
    static constexpr int PW = HA + TIp;
    static constexpr int CPR_I = 2 * PW / V;
    static constexpr int PIN_I = ((2 * PW / V) | 1) * V;     // 16-byte windows, 8 rows apart
    static constexpr int PS1 = ((TI / V) | 1) * V;           // dim-1 synthesis output rows
    static constexpr int OFF_A = HA - H;
    static constexpr int WA = rup(OFF_A + SI + H, V);
    static constexpr int WD = rup(SI + H, V);
    static constexpr int NT_I = rup(TI * NG, 32);
    static constexpr int STAGE_I = 2 * KR * PIN_I;
    static constexpr int SBUF_I = 2 * KR * PS1;
    static constexpr size_t SMEM_I = (size_t)(2 * STAGE_I + 2 * SBUF_I) * sizeof(T);
    static constexpr int MINB_F = NT_F <= 224 ? 3 : (NT_F <= 352 ? 2 : 1);      // resident CTAs the register budget is sized for
    static constexpr int MINB_I = NT_I <= 224 ? 3 : (NT_I <= 352 ? 2 : 1);
    static_assert(F % 2 == 0 && F >= 2, "even filter length");
    static_assert(TIp % SI == 0 && SI % V == 0 && NSEG % 4 == 0, "dim-1 tasks: whole vectors, four segments per warp row group");
    static_assert(TO % 8 == 0, "dim-1 tasks take rows in groups of eight");
    static_assert(TI % V == 0, "tile must be whole vectors");
};

// ===================================================================================================
// forward level
// ===================================================================================================
template <typename T, int F, bool STRICT, class C>
__global__ void __launch_bounds__(C::NT_F, C::MINB_F)
k_fir3d_fwd(const T *__restrict__ src, int64_t ld_s, int64_t ps_s, int64_t bs_s,
            T *__restrict__ ll, int64_t ld_ll, int64_t ps_ll, int64_t bs_ll,
            T *__restrict__ yd, int64_t ld_y, int64_t ps_y, int64_t bs_y,
            int nI, int nJ, int nK, int tilesI, int tilesK, int UC, const __grid_constant__ FirCoefs<T, F> fc) {
    using fp = FP<STRICT>;
    constexpr int V = C::V, Q = C::Q, H = C::H, PP = C::PP, NG = C::NG, TO = C::TO, TOp = C::TOp, KR = C::KR, NT = C::NT_F;
    constexpr int NX = 2 * PP + F - 2;                       // dim-3 inputs a thread reads per slab
    extern __shared__ __align__(128) unsigned char smem_raw[];
    T *In = reinterpret_cast<T *>(smem_raw);                  // [2 stages][2 slabs][KR][PIN_F]
    T *O2 = In + 2 * C::STAGE_F;                              // [2 buffers][2 slabs (a2, d2)][TO rows][PO]
    const int tid = threadIdx.x;
    int bx = blockIdx.x;
    const int ti = bx % tilesI; bx /= tilesI;
    const int tk = bx % tilesK; bx /= tilesK;
    const int ch = bx;                                        // chunk of the marching dimension
    const int64_t b = blockIdx.y;
    const int nhI = nI >> 1, nhJ = nJ >> 1, nhK = nK >> 1;
    const int i0 = ti * C::TI, p0 = i0 >> 1;
    const int q0 = tk * TOp;                                  // first dim-3 output pair of this tile
    const int u0 = ch * UC;                                   // first dim-2 output pair of this chunk
    const int steps = UC + H;                                 // H warm-up steps fill the ring
    const T *sb = src + b * bs_s;

    // ---- loader: thread -> (16-byte chunk column, row lane); rows rl, rl + RL, ... of both slabs ----
    constexpr int RL = NT / C::CPR_F;
    const int lcc = tid % C::CPR_F, lrl = tid / C::CPR_F;
    int lig = i0 - C::HS + lcc * V;                           // global dim-1 index of this chunk (periodic)
    if (lig < 0) lig += nI; else if (lig >= nI) lig -= nI;
    auto issue = [&](int t) {
        if (lrl < RL) {
            int j = 2 * (u0 + t);
            if (j >= nJ) j -= nJ;                             // 2 (u0 + t) < nJ + F - 2 <= 2 nJ
            const T *pj = sb + (int64_t)j * ld_s + lig;       // slab j; slab j + 1 is never the wrap of an even j
            T *dst = In + (t & 1) * C::STAGE_F + lcc * V;
#pragma unroll
            for (int rr = lrl, m = 0; m < (KR + RL - 1) / RL; ++m, rr += RL) {
                if (rr < KR) {
                    int k = 2 * q0 + rr;
                    if (k >= nK) k -= nK;
                    const T *p = pj + (int64_t)k * ps_s;
                    cp_async16(dst + rr * C::PIN_F, p);
                    cp_async16(dst + (KR + rr) * C::PIN_F, p + ld_s);
                }
            }
        }
        cp_commit();
    };
    issue(0);
    if (steps > 1) { issue(1); cp_wait_1(); } else cp_wait_all();
    __syncthreads();

    // ---- phase-A identity: dim-1 position il, group g of dim-3 outputs ----
    const bool actA = tid < C::WIS * NG;
    const int il = tid % C::WIS, g = tid / C::WIS;
    T ring[2 * PP][F];                                        // [position][slab slot]
#pragma unroll
    for (int a = 0; a < 2 * PP; ++a)
#pragma unroll
        for (int m = 0; m < F; ++m) ring[a][m] = T(0);
    // ---- phase-B identity: warp tasks of 8 rows x 4 segments ----
    const int lane = tid & 31, warp = tid >> 5;
    constexpr int NWARP = NT / 32;
    constexpr int RG = TO / 8, SG = C::NSEG / 4;
    constexpr int NWT = 2 * RG * SG;
    T *llb = ll + b * bs_ll;
    T *yb = yd + b * bs_y;

    for (int tb = 0; tb < steps; tb += Q) {
#pragma unroll
        for (int ph = 0; ph < Q; ++ph) {
            const int t = tb + ph;
            if (t < steps) {
                // ================= phase A: dim-3 analysis of the two new slabs -> ring; dim-2 analysis from the ring =============
                if (actA) {
                    const T *inb = In + (t & 1) * C::STAGE_F + il;
#pragma unroll
                    for (int s = 0; s < 2; ++s) {
                        T xin[NX];
#pragma unroll
                        for (int r = 0; r < NX; ++r) xin[r] = inb[(s * KR + g * 2 * PP + r) * C::PIN_F];
#pragma unroll
                        for (int pr = 0; pr < PP; ++pr) {
                            T a = fp::mul(fc.h[0], xin[2 * pr]);
#pragma unroll
                            for (int m = 1; m < F; ++m) a = fp::mac(a, fc.h[m], xin[2 * pr + m]);
                            T d = fp::mul(fc.g[F - 1], xin[2 * pr]);
#pragma unroll
                            for (int m = 1; m < F; ++m) d = fp::mac(d, fc.g[F - 1 - m], xin[2 * pr + m]);
                            ring[pr][(2 * ph + s) % F] = a;
                            ring[PP + pr][(2 * ph + s) % F] = d;
                        }
                    }
                    if (t >= H) {
                        T *ob = O2 + (t & 1) * C::OBUF_F + il;
#pragma unroll
                        for (int pos = 0; pos < 2 * PP; ++pos) {
                            // window in slab order: slot (2 ph + 2 + m) % F, m = 0 .. F-1
                            T a = fp::mul(fc.h[0], ring[pos][(2 * ph + 2) % F]);
#pragma unroll
                            for (int m = 1; m < F; ++m) a = fp::mac(a, fc.h[m], ring[pos][(2 * ph + 2 + m) % F]);
                            T d = fp::mul(fc.g[F - 1], ring[pos][(2 * ph + 2) % F]);
#pragma unroll
                            for (int m = 1; m < F; ++m) d = fp::mac(d, fc.g[F - 1 - m], ring[pos][(2 * ph + 2 + m) % F]);
                            const int krow = (pos < PP) ? (g * PP + pos) : (TOp + g * PP + (pos - PP));
                            ob[krow * C::PO] = a;
                            ob[(TO + krow) * C::PO] = d;
                        }
                    }
                }
                cp_wait_all();                                // this thread's share of step t + 1 has landed
                __syncthreads();                              // O2[t & 1] complete; In[(t + 1) & 1] complete; In[t & 1] free
                if (t + 2 < steps) issue(t + 2);
                // ================= phase B: dim-1 analysis of the two finished slabs, 16-byte stores into the octants =============
                if (t >= H) {
                    const int u = u0 + t - H;                 // a2 index; the d2 slab is index u + H (periodic)
                    int ud = u + H;
                    if (ud >= nhJ) ud -= nhJ;
                    for (int wt = warp; wt < NWT; wt += NWARP) {
                        const int sj = wt / (RG * SG), rg = (wt / SG) % RG, sg = wt % SG;
                        const int row = rg * 8 + (lane & 7), seg = sg * 4 + (lane >> 3);
                        T w[C::WIN_F];
                        ld16<T, C::WIN_F>(w, O2 + (t & 1) * C::OBUF_F + (sj * TO + row) * C::PO + 2 * seg * C::SI);
                        T s[C::NP], d[C::NP];
#pragma unroll
                        for (int pp = 0; pp < C::NP; ++pp) { s[pp] = w[C::OFF_F + 2 * pp]; d[pp] = w[C::OFF_F + 2 * pp + 1]; }
                        fir_ana_regs<T, F, STRICT, C::NP>(s, d, fc);
                        T oa[C::SI], od[C::SI];
#pragma unroll
                        for (int pp = 0; pp < C::SI; ++pp) { oa[pp] = s[H + pp]; od[pp] = d[H + pp]; }
                        const bool hk = row >= TOp;           // detail band along dim 3
                        int kq = q0 + (hk ? row - TOp + H : row);
                        if (kq >= nhK) kq -= nhK;
                        const int kidx = hk ? nhK + kq : kq;
                        const int jidx = sj ? nhJ + ud : u;
                        const int ip = p0 + seg * C::SI;
                        T *py = yb + (int64_t)kidx * ps_y + (int64_t)jidx * ld_y + ip;
                        if (hk || sj) st16_cs<T, C::SI>(py, oa);
                        else          st16<T, C::SI>(llb + (int64_t)kidx * ps_ll + (int64_t)jidx * ld_ll + ip, oa);
                        st16_cs<T, C::SI>(py + nhI, od);
                    }
                }
            }
        }
    }
}

// ===================================================================================================
// inverse level
// ===================================================================================================
template <typename T, int F, bool STRICT, class C>
__global__ void __launch_bounds__(C::NT_I, C::MINB_I)
k_fir3d_inv(const T *__restrict__ ll, int64_t ld_ll, int64_t ps_ll, int64_t bs_ll,
            const T *__restrict__ xd, int64_t ld_x, int64_t ps_x, int64_t bs_x,
            T *__restrict__ dst, int64_t ld_d, int64_t ps_d, int64_t bs_d,
            int nI, int nJ, int nK, int tilesI, int tilesJ, int WC, const __grid_constant__ FirCoefs<T, F> fc) {
    using fp = FP<STRICT>;
    constexpr int V = C::V, Q = C::Q, H = C::H, PP = C::PP, NG = C::NG, TOp = C::TOp, KR = C::KR, NT = C::NT_I;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    T *In = reinterpret_cast<T *>(smem_raw);                  // [2 stages][2 planes (A3, D3)][KR rows][PIN_I]
    T *S1 = In + 2 * C::STAGE_I;                              // [2 buffers][2 planes][KR rows][PS1]
    const int tid = threadIdx.x;
    int bx = blockIdx.x;
    const int ti = bx % tilesI; bx /= tilesI;
    const int tj = bx % tilesJ; bx /= tilesJ;
    const int ch = bx;
    const int64_t b = blockIdx.y;
    const int nhI = nI >> 1, nhJ = nJ >> 1, nhK = nK >> 1;
    const int i0 = ti * C::TI, p0 = i0 >> 1;
    const int v0 = tj * TOp;                                  // first dim-2 output pair of this tile
    const int w0 = ch * WC;                                   // first dim-3 output pair of this chunk
    const int steps = WC + H;
    const T *llb = ll + b * bs_ll;
    const T *xb = xd + b * bs_x;

    // ---- loader: thread -> (chunk column, row lane).  Chunk columns [0, PW/V) are the a1 piece, the rest the d1 piece ----
    constexpr int RL = NT / C::CPR_I;
    const int lcc = tid % C::CPR_I, lrl = tid / C::CPR_I;
    const bool lhi = lcc >= C::PW / V;                        // d1 piece (detail band along dim 1)
    int lp = lhi ? p0 + (lcc - C::PW / V) * V : p0 - C::HA + lcc * V;
    if (lp < 0) lp += nhI; else if (lp >= nhI) lp -= nhI;
    const int lig = lhi ? nhI + lp : lp;
    auto issue = [&](int t) {
        if (lrl < RL) {
            int ka = w0 - H + t;                              // A3 plane index (periodic in [0, nhK))
            if (ka < 0) ka += nhK; else if (ka >= nhK) ka -= nhK;
            int kd = w0 + t;
            if (kd >= nhK) kd -= nhK;
            T *dstp = In + (t & 1) * C::STAGE_I + lcc * V;
#pragma unroll
            for (int rr = lrl, m = 0; m < (KR + RL - 1) / RL; ++m, rr += RL) {
                if (rr < KR) {
                    const bool hj = rr >= TOp + H;            // D2 rows (detail band along dim 2)
                    int jq = hj ? v0 + (rr - (TOp + H)) : v0 - H + rr;
                    if (jq < 0) jq += nhJ; else if (jq >= nhJ) jq -= nhJ;
                    const int j = hj ? nhJ + jq : jq;
                    // plane 0: A3[ka] -- its (a1, a2) corner is the LLL octant, which lives in `ll`
                    const T *pa = (lhi || hj) ? xb + (int64_t)ka * ps_x + (int64_t)j * ld_x + lig
                                              : llb + (int64_t)ka * ps_ll + (int64_t)j * ld_ll + lig;
                    cp_async16(dstp + rr * C::PIN_I, pa);
                    cp_async16(dstp + (KR + rr) * C::PIN_I, xb + (int64_t)(nhK + kd) * ps_x + (int64_t)j * ld_x + lig);
                }
            }
        }
        cp_commit();
    };
    issue(0);
    if (steps > 1) { issue(1); cp_wait_1(); } else cp_wait_all();
    __syncthreads();

    // ---- phase-A identity (dim-1 synthesis): warp tasks of 8 rows x 4 segments over both planes ----
    const int lane = tid & 31, warp = tid >> 5;
    constexpr int NWARP = NT / 32;
    constexpr int RG = (KR + 7) / 8, SG = C::NSEG / 4;
    constexpr int NWT = 2 * RG * SG;
    // ---- phase-B identity: dim-1 position il, group g of dim-2 outputs ----
    const bool actB = tid < C::TI * NG;
    const int il = tid % C::TI, g = tid / C::TI;
    T ringA[2 * PP][Q], ringD[2 * PP][Q];
#pragma unroll
    for (int a = 0; a < 2 * PP; ++a)
#pragma unroll
        for (int m = 0; m < Q; ++m) { ringA[a][m] = T(0); ringD[a][m] = T(0); }
    T *db = dst + b * bs_d + (int64_t)(2 * (v0 + g * PP)) * ld_d + i0 + il;

    for (int tb = 0; tb < steps; tb += Q) {
#pragma unroll
        for (int ph = 0; ph < Q; ++ph) {
            const int t = tb + ph;
            if (t < steps) {
                // ================= phase A: dim-1 synthesis of every staged row of both planes =============
                for (int wt = warp; wt < NWT; wt += NWARP) {
                    const int pl = wt / (RG * SG), rg = (wt / SG) % RG, sg = wt % SG;
                    const int row = rg * 8 + (lane & 7), seg = sg * 4 + (lane >> 3);
                    if (row < KR) {
                        const T *rp = In + (t & 1) * C::STAGE_I + (pl * KR + row) * C::PIN_I + seg * C::SI;
                        T wa[C::WA], wd[C::WD];
                        ld16<T, C::WA>(wa, rp);
                        ld16<T, C::WD>(wd, rp + C::PW);
                        T s[C::NP], d[C::NP];
#pragma unroll
                        for (int pp = 0; pp < C::NP; ++pp) {
                            s[pp] = (pp < C::SI + H) ? wa[C::OFF_A + (pp < C::SI + H ? pp : 0)] : T(0);
                            d[pp] = (pp >= H) ? wd[pp >= H ? pp - H : 0] : T(0);
                        }
                        fir_syn_regs<T, F, STRICT, C::NP>(s, d, fc);
                        T o[2 * C::SI];
#pragma unroll
                        for (int pp = 0; pp < C::SI; ++pp) { o[2 * pp] = s[H + pp]; o[2 * pp + 1] = d[H + pp]; }
                        st16<T, 2 * C::SI>(S1 + (t & 1) * C::SBUF_I + (pl * KR + row) * C::PS1 + 2 * seg * C::SI, o);
                    }
                }
                cp_wait_all();
                __syncthreads();                              // S1[t & 1] complete; In[(t + 1) & 1] landed; In[t & 1] free
                if (t + 2 < steps) issue(t + 2);
                // ================= phase B: dim-2 synthesis -> rings; dim-3 synthesis from the rings -> global =============
                if (actB) {
                    const T *sbp = S1 + (t & 1) * C::SBUF_I + il;
#pragma unroll
                    for (int pl = 0; pl < 2; ++pl) {
                        T a2[PP + H], d2[PP + H];
#pragma unroll
                        for (int r = 0; r < PP + H; ++r) {
                            a2[r] = sbp[(pl * KR + g * PP + r) * C::PS1];
                            d2[r] = sbp[(pl * KR + TOp + H + g * PP + r) * C::PS1];
                        }
#pragma unroll
                        for (int pr = 0; pr < PP; ++pr) {
                            // A2[v - k] = a2[pr + H - k], D2[v + k] = d2[pr + k]
                            T rae = fp::mul(fc.h[2 * (Q - 1)], a2[pr]);
                            T rao = fp::mul(fc.h[2 * (Q - 1) + 1], a2[pr]);
#pragma unroll
                            for (int k = Q - 2; k >= 0; --k) {
                                rae = fp::mac(rae, fc.h[2 * k], a2[pr + H - k]);
                                rao = fp::mac(rao, fc.h[2 * k + 1], a2[pr + H - k]);
                            }
                            T rde, rdo;      // fast mode: the detail terms continue the approximation's chain (no second FMUL, no final FADD)
                            if constexpr (STRICT) { rde = fp::mul(fc.g[1], d2[pr]); rdo = fp::mul(fc.g[0], d2[pr]); }
                            else { rde = fp::mac(rae, fc.g[1], d2[pr]); rdo = fp::mac(rao, fc.g[0], d2[pr]); }
#pragma unroll
                            for (int k = 1; k < Q; ++k) {
                                rde = fp::mac(rde, fc.g[2 * k + 1], d2[pr + k]);
                                rdo = fp::mac(rdo, fc.g[2 * k], d2[pr + k]);
                            }
                            const T x0 = (STRICT ? fp::add(rae, rde) : rde), x1 = (STRICT ? fp::add(rao, rdo) : rdo);
                            if (pl == 0) { ringA[2 * pr][ph % Q] = x0; ringA[2 * pr + 1][ph % Q] = x1; }
                            else         { ringD[2 * pr][ph % Q] = x0; ringD[2 * pr + 1][ph % Q] = x1; }
                        }
                    }
                    if (t >= H) {
                        const int w = w0 + t - H;             // output pair along dim 3
                        T *pw = db + (int64_t)(2 * w) * ps_d;
#pragma unroll
                        for (int pos = 0; pos < 2 * PP; ++pos) {
                            // A3[w - k] sits in slot (ph - k) mod Q, D3[w + k] in slot (ph + k + 1) mod Q
                            T rae = fp::mul(fc.h[2 * (Q - 1)], ringA[pos][(ph + 1) % Q]);
                            T rao = fp::mul(fc.h[2 * (Q - 1) + 1], ringA[pos][(ph + 1) % Q]);
#pragma unroll
                            for (int k = Q - 2; k >= 0; --k) {
                                rae = fp::mac(rae, fc.h[2 * k], ringA[pos][(ph - k + Q) % Q]);
                                rao = fp::mac(rao, fc.h[2 * k + 1], ringA[pos][(ph - k + Q) % Q]);
                            }
                            T rde, rdo;      // fast mode: the detail terms continue the approximation's chain (no second FMUL, no final FADD)
                            if constexpr (STRICT) { rde = fp::mul(fc.g[1], ringD[pos][(ph + 1) % Q]); rdo = fp::mul(fc.g[0], ringD[pos][(ph + 1) % Q]); }
                            else { rde = fp::mac(rae, fc.g[1], ringD[pos][(ph + 1) % Q]); rdo = fp::mac(rao, fc.g[0], ringD[pos][(ph + 1) % Q]); }
#pragma unroll
                            for (int k = 1; k < Q; ++k) {
                                rde = fp::mac(rde, fc.g[2 * k + 1], ringD[pos][(ph + k + 1) % Q]);
                                rdo = fp::mac(rdo, fc.g[2 * k], ringD[pos][(ph + k + 1) % Q]);
                            }
                            T *po = pw + (int64_t)pos * ld_d;
                            po[0] = (STRICT ? fp::add(rae, rde) : rde);
                            po[ps_d] = (STRICT ? fp::add(rao, rdo) : rdo);
                        }
                    }
                }
            }
        }
    }
}

// ===================================================================================================
// host side
// ===================================================================================================
// configuration per element type, filter length and direction.  Float32 with up to 12 taps keeps four positions per thread
// (ring of 4 F registers) on a 64 x 16 tile, two CTAs per SM; longer filters and Float64 keep two positions.  The dim-1
// synthesis of the inverse takes 8-pair tasks (fewer 16-byte window loads per output), the analysis 4-pair tasks.
// r02 sweep on 512^3 db6 (profiles/r02_fir3d_tile_sweep.md): 128 x 16 one CTA/SM 0.603 / 0.596 ms (forward / inverse, L = 3),
// 128 x 8 0.583 / 0.666, 64 x 16 0.577 / 0.551, 64 x 8 0.72 / 0.72, 64 x 16 with 8-pair tasks 0.607 (spills) / 0.535.
template <typename T, int F, bool FW> struct Pick {
    static constexpr bool WIDE = sizeof(T) == 4 && F <= 12;
    static constexpr int TI = WIDE ? 64 : (sizeof(T) == 4 ? 128 : 64);
    static constexpr int PP = WIDE ? 2 : 1;
    static constexpr int NG = 4;
    static constexpr int SI = (WIDE && !FW) ? 8 : 4;
    using type = Cfg<T, F, TI, PP, NG, SI>;
};

static inline int env3(const char *name, int dflt) {
    const char *v = std::getenv(name);
    return (v && *v) ? std::atoi(v) : dflt;
}

// chunks of the marching dimension: enough CTAs to fill the device in whole waves, few enough that the H warm-up steps of a
// chunk stay small against its UC output steps
static inline int pick_chunks(int64_t tiles, int half, int H, int nsm) {
    int best = 1;
    double best_cost = 1e300;
    for (int c = 1; c <= half; c *= 2) {
        if (half % c) break;
        const int uc = half / c;
        if (uc < 2 && c > 1) break;
        const int64_t ctas = tiles * c;
        const int64_t waves = (ctas + nsm - 1) / nsm;
        const double cost = (double)waves * (uc + 0.5 * H + 1.0);
        if (cost < best_cost) { best_cost = cost; best = c; }
    }
    return best;
}

template <typename T, int F, bool STRICT>
static int32_t launch_fwd(const T *src, int64_t ld_s, int64_t ps_s, int64_t bs_s, T *ll, int64_t ld_ll, int64_t ps_ll, int64_t bs_ll,
                          T *y, int64_t ld_y, int64_t ps_y, int64_t bs_y, int nI, int nJ, int nK, int64_t B,
                          const FirCoefs<T, F> &fc, cudaStream_t st) {
    using C = typename Pick<T, F, true>::type;
    const int tilesI = nI / C::TI, tilesK = nK / C::TO;
    int dev = 0, nsm = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    int nch = env3("WB200_FIR3D_CHUNKS", 0);
    if (nch <= 0 || (nJ / 2) % nch) nch = pick_chunks((int64_t)tilesI * tilesK * B, nJ / 2, C::H, nsm);
    auto kern = k_fir3d_fwd<T, F, STRICT, C>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM_F) != cudaSuccess) {
        (void)cudaGetLastError(); set_error("cudaFuncSetAttribute(k_fir3d_fwd) failed"); return WB200_ECUDA;
    }
    {
        LaunchScope scope("fused_fir3d_fwd", st);
        kern<<<dim3((unsigned)(tilesI * tilesK * nch), (unsigned)B), C::NT_F, C::SMEM_F, st>>>(
            src, ld_s, ps_s, bs_s, ll, ld_ll, ps_ll, bs_ll, y, ld_y, ps_y, bs_y, nI, nJ, nK, tilesI, tilesK, (nJ / 2) / nch, fc);
    }
    return check_launch("fused_fir3d_fwd") ? WB200_OK : WB200_ECUDA;
}
template <typename T, int F, bool STRICT>
static int32_t launch_inv(const T *ll, int64_t ld_ll, int64_t ps_ll, int64_t bs_ll, const T *x, int64_t ld_x, int64_t ps_x, int64_t bs_x,
                          T *dst, int64_t ld_d, int64_t ps_d, int64_t bs_d, int nI, int nJ, int nK, int64_t B,
                          const FirCoefs<T, F> &fc, cudaStream_t st) {
    using C = typename Pick<T, F, false>::type;
    const int tilesI = nI / C::TI, tilesJ = nJ / C::TO;
    int dev = 0, nsm = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    int nch = env3("WB200_FIR3D_CHUNKS", 0);
    if (nch <= 0 || (nK / 2) % nch) nch = pick_chunks((int64_t)tilesI * tilesJ * B, nK / 2, C::H, nsm);
    auto kern = k_fir3d_inv<T, F, STRICT, C>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM_I) != cudaSuccess) {
        (void)cudaGetLastError(); set_error("cudaFuncSetAttribute(k_fir3d_inv) failed"); return WB200_ECUDA;
    }
    {
        LaunchScope scope("fused_fir3d_inv", st);
        kern<<<dim3((unsigned)(tilesI * tilesJ * nch), (unsigned)B), C::NT_I, C::SMEM_I, st>>>(
            ll, ld_ll, ps_ll, bs_ll, x, ld_x, ps_x, bs_x, dst, ld_d, ps_d, bs_d, nI, nJ, nK, tilesI, tilesJ, (nK / 2) / nch, fc);
    }
    return check_launch("fused_fir3d_inv") ? WB200_OK : WB200_ECUDA;
}

// a corner (nI, nJ, nK) is served when every dimension is whole tiles and 16-byte granular
template <typename T, int F> static bool corner_ok(int64_t nI, int64_t nJ, int64_t nK) {
    using C = typename Pick<T, F, true>::type;
    static_assert(C::TI == Pick<T, F, false>::type::TI && C::TO == Pick<T, F, false>::type::TO, "both directions tile alike");
    constexpr int V = C::V;
    return nI % C::TI == 0 && nJ % C::TO == 0 && nK % C::TO == 0 && (nI / 2) % V == 0 && nI >= C::TI && nJ >= C::TO && nK >= C::TO &&
           nJ >= F && nK >= F && nI < (1 << 30) && nJ < (1 << 30) && nK < (1 << 30);
}

template <typename T, int F>
static int levels_f(const ArrayGeom &g, int L) {
    int Lf = 0;
    int64_t a = g.dim[0], b = g.dim[1], c = g.dim[2];
    while (Lf < L && corner_ok<T, F>(a, b, c)) { ++Lf; a >>= 1; b >>= 1; c >>= 1; }
    return Lf;
}

template <typename T, int F, bool STRICT>
static int32_t run_f(const PassOp<T> &op, T *y, const T *x, const T *ll_src, int64_t ll_ld, int64_t ll_ps, int64_t ll_bs,
                     const ArrayGeom &g, int Lf, bool fw, void *scratch, cudaStream_t st) {
    const int64_t N1 = g.dim[0], N2 = g.dim[1], N3 = g.dim[2], B = g.batch;
    const int64_t ldN = N1, psN = N1 * N2, bsN = N1 * N2 * N3;
    FirCoefs<T, F> fc;
    for (int m = 0; m < F; ++m) { fc.h[m] = op.fc.h[m]; fc.g[m] = op.fc.g[m]; }
    T *buf[2];
    buf[0] = (T *)scratch;
    const size_t b0 = (((size_t)(N1 / 2) * (N2 / 2) * (N3 / 2) * B * sizeof(T)) + 255) & ~(size_t)255;
    buf[1] = (T *)((char *)scratch + b0);
    // the level-l approximation (1 <= l < Lf) lives compactly in buf[(l-1) & 1]
    if (fw) {
        for (int l = 1; l <= Lf; ++l) {
            const int64_t a = N1 >> (l - 1), b = N2 >> (l - 1), c = N3 >> (l - 1);
            const T *src = (l == 1) ? x : buf[(l - 2) & 1];
            const int64_t lds = (l == 1) ? ldN : a, pss = (l == 1) ? psN : a * b, bss = (l == 1) ? bsN : a * b * c;
            T *llo; int64_t ldl, psl, bsl;
            if (l == Lf) { llo = y; ldl = ldN; psl = psN; bsl = bsN; }
            else         { llo = buf[(l - 1) & 1]; ldl = a / 2; psl = (a / 2) * (b / 2); bsl = psl * (c / 2); }
            const int32_t rc = launch_fwd<T, F, STRICT>(src, lds, pss, bss, llo, ldl, psl, bsl, y, ldN, psN, bsN, (int)a, (int)b, (int)c, B, fc, st);
            if (rc != WB200_OK) return rc;
        }
    } else {
        for (int l = Lf; l >= 1; --l) {
            const int64_t a = N1 >> (l - 1), b = N2 >> (l - 1), c = N3 >> (l - 1);
            const T *lls; int64_t ldl, psl, bsl;
            if (l == Lf) { lls = ll_src; ldl = ll_ld; psl = ll_ps; bsl = ll_bs; }
            else         { lls = buf[(l - 1) & 1]; ldl = a / 2; psl = (a / 2) * (b / 2); bsl = psl * (c / 2); }
            T *dst; int64_t ldd, psd, bsd;
            if (l == 1) { dst = y; ldd = ldN; psd = psN; bsd = bsN; }
            else        { dst = buf[(l - 2) & 1]; ldd = a; psd = a * b; bsd = a * b * c; }
            const int32_t rc = launch_inv<T, F, STRICT>(lls, ldl, psl, bsl, x, ldN, psN, bsN, dst, ldd, psd, bsd, (int)a, (int)b, (int)c, B, fc, st);
            if (rc != WB200_OK) return rc;
        }
    }
    return WB200_OK;
}

} // namespace f3

template <typename T>
int fir3d_levels(const PassOp<T> &op, const ArrayGeom &g, int L, bool fw) {
    (void)fw;
    if (op.lifting || op.generic_only || g.ndim != 3 || g.C != 1 || g.batch > 65535 || g.batch < 1) return 0;
    if (f3::env3("WB200_DISABLE_FIR3D", 0)) return 0;
    switch (op.fc.F) {
#define WB_F3(FF) case FF: return f3::levels_f<T, FF>(g, L);
        WB_F3(2) WB_F3(4) WB_F3(6) WB_F3(8) WB_F3(10) WB_F3(12) WB_F3(14) WB_F3(16) WB_F3(18) WB_F3(20)
#undef WB_F3
    default: return 0;
    }
}
template <typename T>
size_t fir3d_scratch_bytes(const ArrayGeom &g, int Lf) {
    if (Lf < 1) return 0;
    const size_t n1 = (size_t)g.dim[0], n2 = (size_t)g.dim[1], n3 = (size_t)g.dim[2];
    const size_t b0 = (n1 / 2) * (n2 / 2) * (n3 / 2) * (size_t)g.batch * sizeof(T);
    const size_t b1 = (Lf >= 2) ? (n1 / 4) * (n2 / 4) * (n3 / 4) * (size_t)g.batch * sizeof(T) : 0;
    return ((b0 + 255) & ~(size_t)255) + ((b1 + 255) & ~(size_t)255);
}
template <typename T>
int32_t fir3d_run(const PassOp<T> &op, T *y, const T *x, const T *ll_src, int64_t ll_ld, int64_t ll_ps, int64_t ll_bs,
                  const ArrayGeom &g, int Lf, bool fw, void *scratch, cudaStream_t st) {
#define WB_F3(FF)                                                                                                   \
    case FF: return op.strict ? f3::run_f<T, FF, true>(op, y, x, ll_src, ll_ld, ll_ps, ll_bs, g, Lf, fw, scratch, st) \
                              : f3::run_f<T, FF, false>(op, y, x, ll_src, ll_ld, ll_ps, ll_bs, g, Lf, fw, scratch, st);
    switch (op.fc.F) {
        WB_F3(2) WB_F3(4) WB_F3(6) WB_F3(8) WB_F3(10) WB_F3(12) WB_F3(14) WB_F3(16) WB_F3(18) WB_F3(20)
    default: break;
    }
#undef WB_F3
    set_error("internal: fir3d_run called for an unsupported filter length %d", op.fc.F);
    return WB200_EARG;
}

} // namespace wb
