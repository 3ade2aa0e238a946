// fused2d_tma.cuh -- 2-D lifting level kernels, TMA tensor-map edition (included by fused2d.cu).
//
// Same algorithm as k_lift2d_fwd/inv (one launch per level, register-resident lifting, reference operation order)
// with the shared-memory traffic cut to what the MIO pipe can sustain next to HBM speed:
//   * the tile (+ halo) arrives with ONE `cp.async.bulk.tensor` (TMA, 3-D tensor map: dim 1, dim 2, image) per
//     source box, completing on an mbarrier -- no per-element LDGSTS/LDG+STS instructions at all;
//   * forward keeps the tile INTERLEAVED along dim 1 (as in memory): the dim-2 pass walks rows (threads along dim 1,
//     stride 1), the dim-1 pass reads its segment with 128-bit LDS ((s,d,s,d) per load) and writes it back with
//     128-bit STS; the row pitch is 4*odd samples (2*odd for double) so that 8 consecutive rows hit 8 different
//     16-byte bank groups -> conflict-free; the store phase reads 128-bit and writes two 64-bit quadrant pieces;
//   * inverse stages the four quadrants as four dense arrays (they ARE the polyphase components of both
//     dimensions), so neither pass needs a split and the final store merges with 128-bit LDS + 128-bit STG.
// Periodic boundary: a tensor-map box that leaves the array is zero-filled by the hardware; tiles on the array
// border then patch the wrapped halo cells with ordinary loads (border tiles only).
#pragma once

namespace wb {

// ---------------------------------------------------------------------------------------------------
// host: tensor maps through the driver entry point (no link-time libcuda dependency)
// ---------------------------------------------------------------------------------------------------
struct alignas(64) TensorMap { unsigned char opaque[128]; };

using EncodeTiledFn = int (*)(void *tensorMap, int dtype, uint32_t rank, void *gaddr, const uint64_t *gdim,
                              const uint64_t *gstride, const uint32_t *box, const uint32_t *estride, int interleave,
                              int swizzle, int l2promo, int oobfill);
static EncodeTiledFn get_encode_tiled() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
        else
            (void)cudaGetLastError();
    }
    return fn;
}
// 3-D column-major view (dim1 = n1 contiguous, dim2 = n2 with stride ld, images with stride bs); box (b1, b2, 1)
template <typename T>
static bool make_tensor_map(TensorMap &tm, const T *base, int64_t n1, int64_t n2, int64_t nb, int64_t ld, int64_t bs,
                            int b1, int b2) {
    EncodeTiledFn enc = get_encode_tiled();
    if (!enc) return false;
    if (((uintptr_t)base & 15) || (ld * sizeof(T)) % 16 || (bs * sizeof(T)) % 16 || b1 > 256 || b2 > 256) return false;
    const uint64_t gdim[3] = {(uint64_t)n1, (uint64_t)n2, (uint64_t)nb};
    const uint64_t gstr[2] = {(uint64_t)ld * sizeof(T), (uint64_t)bs * sizeof(T)};
    const uint32_t box[3] = {(uint32_t)b1, (uint32_t)b2, 1u};
    const uint32_t estr[3] = {1u, 1u, 1u};
    const int dtype = sizeof(T) == 4 ? 7 /*CU_TENSOR_MAP_DATA_TYPE_FLOAT32*/ : 8 /*FLOAT64*/;
    const int rc = enc(&tm, dtype, 3, const_cast<T *>(base), gdim, gstr, box, estr, 0 /*INTERLEAVE_NONE*/,
                       0 /*SWIZZLE_NONE*/, 0 /*L2_PROMOTION_NONE*/, 0 /*OOB_FILL_NONE = zeros*/);
    return rc == 0;
}

// ---------------------------------------------------------------------------------------------------
// device: mbarrier + tensor TMA
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t s2u(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mb_init(uint64_t *bar) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s2u(bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mb_expect(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s2u(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mb_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "W2D_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra W2D_DONE;\n\t"
        "bra W2D_LOOP;\n\t"
        "W2D_DONE:\n\t}" ::"r"(s2u(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_box3(void *dst, const TensorMap *tm, int c0, int c1, int c2, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(s2u(dst)), "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(s2u(bar)) : "memory");
}

template <typename T> struct V16;   // 16-byte vector of T
template <> struct V16<float> { using type = float4; static constexpr int N = 4; };
template <> struct V16<double> { using type = double2; static constexpr int N = 2; };
template <typename T, int N> __device__ __forceinline__ void lds16(T (&w)[N], const T *p) {
    static_assert(N % V16<T>::N == 0, "window must be whole 16-byte vectors");
#pragma unroll
    for (int i = 0; i < N / V16<T>::N; ++i) {
        const typename V16<T>::type v = *reinterpret_cast<const typename V16<T>::type *>(p + i * V16<T>::N);
        if constexpr (sizeof(T) == 4) { w[4 * i] = v.x; w[4 * i + 1] = v.y; w[4 * i + 2] = v.z; w[4 * i + 3] = v.w; }
        else                          { w[2 * i] = v.x; w[2 * i + 1] = v.y; }
    }
}
template <typename T, int N> __device__ __forceinline__ void sts16(T *p, const T (&w)[N]) {
#pragma unroll
    for (int i = 0; i < N / V16<T>::N; ++i) {
        typename V16<T>::type v;
        if constexpr (sizeof(T) == 4) { v.x = w[4 * i]; v.y = w[4 * i + 1]; v.z = w[4 * i + 2]; v.w = w[4 * i + 3]; }
        else                          { v.x = w[2 * i]; v.y = w[2 * i + 1]; }
        *reinterpret_cast<typename V16<T>::type *>(p + i * V16<T>::N) = v;
    }
}

// ---------------------------------------------------------------------------------------------------
// configuration
// ---------------------------------------------------------------------------------------------------
template <typename T, class S, int TI_, int TJ_, int SI_, int SJ_> struct Cfg3 {
    static constexpr int TI = TI_, TJ = TJ_, SI = SI_, SJ = SJ_;
    static constexpr int V = 16 / (int)sizeof(T);                 // samples per 16-byte vector
    static constexpr int HL = Halo<S>::left(), HR = Halo<S>::right();   // halo in polyphase pairs
    static constexpr int TIp = TI / 2, TJp = TJ / 2;
    static constexpr int NPI = SI + HL + HR, NPJ = SJ + HL + HR;
    // ---- forward: interleaved tile; local sample 0 along dim 1 is i0 - HLS (HLS: left halo rounded up to a vector)
    static constexpr int HLS = (2 * HL + V - 1) / V * V;
    static constexpr int FW_I = HLS + TI + 2 * HR;                // staged samples along dim 1
    static constexpr int FPI = ((FW_I + V - 1) / V) | 1;          // pitch in vectors, odd  -> conflict-free 16-byte rows
    static constexpr int PI = FPI * V;                            // pitch in samples (= TMA box width)
    static constexpr int RJ = TJ + 2 * (HL + HR);                 // staged rows
    static constexpr int WINF = ((HLS - 2 * HL) + 2 * NPI + V - 1) / V * V;   // dim-1 window (samples) of a segment
    static constexpr int ROW_TASKS_F = (2 * HL + TI + 2 * HR) * (TJp / SJ);
    static constexpr int COL_TASKS_F = TJ * (TIp / SI);
    // ---- inverse: four dense quadrant arrays [pi][pj][JQ][PC]; local pair 0 along dim 1 is ip0 - CO
    static constexpr int CO = (HL + V - 1) / V * V;
    static constexpr int IW = CO + TIp + HR;                      // staged pairs along dim 1
    static constexpr int IPC = ((IW + V - 1) / V) | 1;
    static constexpr int PC = IPC * V;
    static constexpr int JQ = TJp + HL + HR;                      // staged dim-2 pairs
    static constexpr int WINI = ((CO - HL) + NPI + V - 1) / V * V;
    static constexpr int COL_TASKS_I = 2 * JQ * (TIp / SI);
    static constexpr int ROW_TASKS_I = 2 * TIp * (TJp / SJ);
    static constexpr int MAXT_ = ROW_TASKS_F > COL_TASKS_F ? ROW_TASKS_F : COL_TASKS_F;
    static constexpr int MAXI_ = COL_TASKS_I > ROW_TASKS_I ? COL_TASKS_I : ROW_TASKS_I;
    static constexpr int NT = ((MAXT_ > MAXI_ ? MAXT_ : MAXI_) + 31) / 32 * 32;
    static constexpr size_t SMEM_F = 128 + ((size_t)RJ * PI * sizeof(T) + 127) / 128 * 128;   // per buffer (+128 header once)
    static constexpr int QSZ = (JQ * PC * (int)sizeof(T) + 127) / 128 * 128 / (int)sizeof(T);   // one quadrant array, 128-byte multiple (TMA destination alignment)
    static constexpr size_t SMEM_I = 128 + (size_t)4 * QSZ * sizeof(T);
    static_assert(TIp % SI == 0 && TJp % SJ == 0, "segments must tile the tile");
    static_assert((HLS - 2 * HL) % 2 == 0, "pair alignment");
};

// ---------------------------------------------------------------------------------------------------
// forward level
// ---------------------------------------------------------------------------------------------------
// NBUF = 1: one tile per CTA (grid = all tiles).  NBUF = 2: PERSISTENT CTAs walk the tile list with two shared-memory
// buffers -- the TMA box of tile k+1 is in flight while tile k is transformed (the top stall of the one-shot kernel is
// the wait for its own tile, profiles/r01_lift2d_f32_tma.md).
template <typename T, class S, bool STRICT, class C, int NBUF>
__global__ void __launch_bounds__(C::NT, (NBUF == 1 ? 4 : 2) / (int)(sizeof(T) / 4))   // resident CTAs the register budget is sized for
k_lift2d_fwd_tma_p(const __grid_constant__ TensorMap tm_src, const T *__restrict__ src, int64_t ld_s, int64_t bs_s,
                 T *__restrict__ ll, int64_t ld_ll, int64_t bs_ll, T *__restrict__ yd, int64_t ld_y, int64_t bs_y,
                 int n, int nx, int ny, int ntiles, const __grid_constant__ LiftCoefs<T> lc) {
    using fp = FP<STRICT>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw);
    constexpr int BUFSZ = (C::RJ * C::PI * (int)sizeof(T) + 127) / 128 * 128 / (int)sizeof(T);
    T *Sm0 = reinterpret_cast<T *>(smem_raw + 128);
    const int nh = n >> 1;
    const int tid = threadIdx.x;
    auto issue = [&](int tile, int buf) {
        const int tx = tile % nx, r = tile / nx;
        mb_expect(bars + buf, (uint32_t)(C::RJ * C::PI * sizeof(T)));
        tma_box3(Sm0 + buf * BUFSZ, &tm_src, tx * C::TI - C::HLS, (r % ny) * C::TJ - 2 * C::HL, r / ny, bars + buf);
    };
    if (tid == 0) {
        for (int q = 0; q < NBUF; ++q) mb_init(bars + q);
        if ((int)blockIdx.x < ntiles) issue(blockIdx.x, 0);
    }
    __syncthreads();
    int it = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
    const int cur = (NBUF == 2) ? (it & 1) : 0;
    T *Sm = Sm0 + cur * BUFSZ;
    const int txi = tile % nx, tr = tile / nx, tyi = tr % ny;
    const int b = tr / ny;
    const int i0 = txi * C::TI, j0 = tyi * C::TJ;
    const int iorg = i0 - C::HLS, jorg = j0 - 2 * C::HL;          // global coordinates of local (0, 0)
    if (NBUF == 2 && tid == 0 && tile + (int)gridDim.x < ntiles) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy accesses of that buffer are done (barrier below)
        issue(tile + gridDim.x, cur ^ 1);
    }
    mb_wait(bars + cur, (NBUF == 2) ? ((it >> 1) & 1) : 0);
    // ---- periodic wrap: border tiles patch the halo cells the tensor map zero-filled ----
    const bool edge_i = (txi == 0) || (txi == nx - 1);
    const bool edge_j = (tyi == 0) || (tyi == ny - 1);
    if (edge_i || edge_j) {
        const T *sb = src + (int64_t)b * bs_s;
        constexpr int WI = C::HLS + C::TI + 2 * C::HR;
        for (int idx = tid; idx < C::RJ * WI; idx += C::NT) {
            const int r = idx / WI, il = idx - r * WI;
            const int gi = iorg + il, gj = jorg + r;
            if (gi < 0 || gi >= n || gj < 0 || gj >= n)
                Sm[r * C::PI + il] = sb[(int64_t)wrapi(gj, n) * ld_s + wrapi(gi, n)];
        }
        __syncthreads();
    }
    // ---- dim-2 pass: one thread per (dim-1 sample, segment of dim-2 pairs) ----
    {
        T s[C::NPJ], d[C::NPJ];
        const bool act = tid < C::ROW_TASKS_F;
        constexpr int WI = 2 * C::HL + C::TI + 2 * C::HR;
        int il = 0, q = 0;
        if (act) {
            il = (C::HLS - 2 * C::HL) + tid % WI;
            q = tid / WI;
#pragma unroll
            for (int pp = 0; pp < C::NPJ; ++pp) {
                s[pp] = Sm[(2 * (q * C::SJ + pp)) * C::PI + il];
                d[pp] = Sm[(2 * (q * C::SJ + pp) + 1) * C::PI + il];
            }
        }
        __syncthreads();
        if (act) {
            const int jp0 = (j0 >> 1) - C::HL + q * C::SJ;
            lift_regs<T, S, STRICT, C::NPJ>(s, d, lc, wrapi(jp0, nh), nh, STRICT && edge_j);
#pragma unroll
            for (int pp = C::HL; pp < C::HL + C::SJ; ++pp) {
                Sm[(2 * (q * C::SJ + pp)) * C::PI + il] = fp::mul(s[pp], lc.n1);
                Sm[(2 * (q * C::SJ + pp) + 1) * C::PI + il] = fp::mul(d[pp], lc.n2);
            }
        }
        __syncthreads();
    }
    // ---- dim-1 pass: one thread per (owned row, segment); 16-byte shared-memory accesses ----
    {
        T w[C::WINF];
        T s[C::NPI], d[C::NPI];
        const bool act = tid < C::COL_TASKS_F;
        constexpr int OFF = C::HLS - 2 * C::HL;           // samples between the vector-aligned window start and pair 0
        int r = 0, q = 0;
        if (act) {
            r = 2 * C::HL + tid % C::TJ;
            q = tid / C::TJ;
            lds16<T, C::WINF>(w, Sm + r * C::PI + 2 * q * C::SI);
#pragma unroll
            for (int pp = 0; pp < C::NPI; ++pp) { s[pp] = w[OFF + 2 * pp]; d[pp] = w[OFF + 2 * pp + 1]; }
        }
        __syncthreads();
        if (act) {
            lift_regs<T, S, STRICT, C::NPI>(s, d, lc, wrapi((i0 >> 1) - C::HL + q * C::SI, nh), nh, STRICT && edge_i);
            T o[2 * C::SI];
#pragma unroll
            for (int pp = 0; pp < C::SI; ++pp) {
                o[2 * pp] = fp::mul(s[C::HL + pp], lc.n1);
                o[2 * pp + 1] = fp::mul(d[C::HL + pp], lc.n2);
            }
            sts16<T, 2 * C::SI>(Sm + r * C::PI + C::HLS + 2 * q * C::SI, o);   // HLS is a vector multiple: aligned
        }
        __syncthreads();
    }
    // ---- stores: a thread reads one 16-byte piece of an owned row and writes its s-part and d-part ----
    {
        T *llb = ll + (int64_t)b * bs_ll;
        T *yb = yd + (int64_t)b * bs_y;
        constexpr int VPR = C::TI / C::V;                 // vectors per owned row
        for (int idx = tid; idx < C::TJ * VPR; idx += C::NT) {
            const int t = idx % VPR, rr = idx / VPR;
            const int j = j0 + rr, jq = j >> 1, pj = j & 1;
            T w[C::V];
            lds16<T, C::V>(w, Sm + (2 * C::HL + rr) * C::PI + C::HLS + t * C::V);
            const int ip = (i0 >> 1) + t * (C::V / 2);
            T *ps = (pj == 0) ? (llb + (int64_t)jq * ld_ll + ip) : (yb + (int64_t)(nh + jq) * ld_y + ip);
            T *pd = yb + (int64_t)(pj * nh + jq) * ld_y + nh + ip;
            if constexpr (sizeof(T) == 4) {
                *reinterpret_cast<float2 *>(ps) = make_float2(w[0], w[2]);
                *reinterpret_cast<float2 *>(pd) = make_float2(w[1], w[3]);
            } else {
                ps[0] = w[0];
                pd[0] = w[1];
            }
        }
    }
    __syncthreads();   // every thread is done with this buffer before it is refilled
    }
}

// ---------------------------------------------------------------------------------------------------
// inverse level
// ---------------------------------------------------------------------------------------------------
template <typename T, class S, bool STRICT, class C, int NBUF>
__global__ void __launch_bounds__(C::NT, (NBUF == 1 ? 4 : 2) / (int)(sizeof(T) / 4))
k_lift2d_inv_tma_p(const __grid_constant__ TensorMap tm_ll, const __grid_constant__ TensorMap tm_x,
                 const T *__restrict__ ll, int64_t ld_ll, int64_t bs_ll, const T *__restrict__ xd, int64_t ld_x, int64_t bs_x,
                 T *__restrict__ dst, int64_t ld_d, int64_t bs_d, int n, int nx, int ny, int ntiles,
                 const __grid_constant__ LiftCoefs<T> lc) {
    using fp = FP<STRICT>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw);
    T *Sm0 = reinterpret_cast<T *>(smem_raw + 128);
    constexpr int QSZ = C::QSZ;                           // one quadrant array; order [pi][pj]
    const int nh = n >> 1;
    const int tid = threadIdx.x;
    auto issue = [&](int tile, int buf) {
        const int tx = tile % nx, r = tile / nx;
        const int c = tx * C::TIp - C::CO, q = (r % ny) * C::TJp - C::HL, bb = r / ny;
        T *B0 = Sm0 + buf * 4 * QSZ;
        mb_expect(bars + buf, (uint32_t)(4 * C::JQ * C::PC * sizeof(T)));
        tma_box3(B0 + 0 * QSZ, &tm_ll, c, q, bb, bars + buf);                 // (pi, pj) = (0, 0): LL
        tma_box3(B0 + 1 * QSZ, &tm_x, c, nh + q, bb, bars + buf);             // (0, 1)
        tma_box3(B0 + 2 * QSZ, &tm_x, nh + c, q, bb, bars + buf);             // (1, 0)
        tma_box3(B0 + 3 * QSZ, &tm_x, nh + c, nh + q, bb, bars + buf);        // (1, 1)
    };
    if (tid == 0) {
        for (int q = 0; q < NBUF; ++q) mb_init(bars + q);
        if ((int)blockIdx.x < ntiles) issue(blockIdx.x, 0);
    }
    __syncthreads();
    int it = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
    const int cur = (NBUF == 2) ? (it & 1) : 0;
    T *Sm = Sm0 + cur * 4 * QSZ;
    const int txi = tile % nx, tr = tile / nx, tyi = tr % ny;
    const int b = tr / ny;
    const int ip0 = txi * C::TIp, jq0 = tyi * C::TJp;
    const int corg = ip0 - C::CO, qorg = jq0 - C::HL;     // quadrant coordinates of local (0, 0)
    if (NBUF == 2 && tid == 0 && tile + (int)gridDim.x < ntiles) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        issue(tile + gridDim.x, cur ^ 1);
    }
    mb_wait(bars + cur, (NBUF == 2) ? ((it >> 1) & 1) : 0);
    const bool edge_i = (txi == 0) || (txi == nx - 1);
    const bool edge_j = (tyi == 0) || (tyi == ny - 1);
    if (edge_i || edge_j) {   // quadrant-relative wrap of the halo cells
        const T *llb = ll + (int64_t)b * bs_ll;
        const T *xb = xd + (int64_t)b * bs_x;
        for (int idx = tid; idx < 4 * C::JQ * C::IW; idx += C::NT) {
            const int c = idx % C::IW;
            const int rest = idx / C::IW;
            const int ql = rest % C::JQ, quad = rest / C::JQ;
            const int gc = corg + c, gq = qorg + ql;
            if (gc < 0 || gc >= nh || gq < 0 || gq >= nh) {
                const int wc = wrapi(gc, nh), wq = wrapi(gq, nh);
                const int pi = quad >> 1, pj = quad & 1;
                Sm[quad * QSZ + ql * C::PC + c] = (quad == 0) ? llb[(int64_t)wq * ld_ll + wc]
                                                              : xb[(int64_t)(pj * nh + wq) * ld_x + pi * nh + wc];
            }
        }
        __syncthreads();
    }
    // ---- dim-1 pass first (inverse order): s_i = quadrant (0, pj), d_i = quadrant (1, pj), on every staged dim-2 pair
    {
        T ws[C::WINI], wd[C::WINI];
        T s[C::NPI], d[C::NPI];
        const bool act = tid < C::COL_TASKS_I;
        constexpr int OFF = C::CO - C::HL;
        T *As = Sm, *Ad = Sm;
        int q = 0;
        if (act) {
            const int ql = tid % C::JQ;
            const int rest = tid / C::JQ;
            const int pj = rest & 1;
            q = rest >> 1;
            As = Sm + (0 * 2 + pj) * QSZ + ql * C::PC + q * C::SI;
            Ad = Sm + (1 * 2 + pj) * QSZ + ql * C::PC + q * C::SI;
            lds16<T, C::WINI>(ws, As);
            lds16<T, C::WINI>(wd, Ad);
#pragma unroll
            for (int pp = 0; pp < C::NPI; ++pp) { s[pp] = fp::mul(ws[OFF + pp], lc.n1); d[pp] = fp::mul(wd[OFF + pp], lc.n2); }
        }
        __syncthreads();
        if (act) {
            lift_regs<T, S, STRICT, C::NPI>(s, d, lc, wrapi(ip0 - C::HL + q * C::SI, nh), nh, STRICT && edge_i);
            T os[C::SI], od[C::SI];
#pragma unroll
            for (int pp = 0; pp < C::SI; ++pp) { os[pp] = s[C::HL + pp]; od[pp] = d[C::HL + pp]; }
            sts16<T, C::SI>(As + C::CO, os);
            sts16<T, C::SI>(Ad + C::CO, od);
        }
        __syncthreads();
    }
    // ---- dim-2 pass on the owned dim-1 pairs: s_j = quadrant (pi, 0), d_j = quadrant (pi, 1) ----
    {
        T s[C::NPJ], d[C::NPJ];
        const bool act = tid < C::ROW_TASKS_I;
        T *A0 = Sm, *A1 = Sm;
        int q = 0;
        if (act) {
            const int c = C::CO + tid % C::TIp;
            const int rest = tid / C::TIp;
            const int pi = rest & 1;
            q = rest >> 1;
            A0 = Sm + (pi * 2 + 0) * QSZ + (q * C::SJ) * C::PC + c;
            A1 = Sm + (pi * 2 + 1) * QSZ + (q * C::SJ) * C::PC + c;
#pragma unroll
            for (int pp = 0; pp < C::NPJ; ++pp) { s[pp] = fp::mul(A0[pp * C::PC], lc.n1); d[pp] = fp::mul(A1[pp * C::PC], lc.n2); }
        }
        __syncthreads();
        if (act) {
            lift_regs<T, S, STRICT, C::NPJ>(s, d, lc, wrapi(jq0 - C::HL + q * C::SJ, nh), nh, STRICT && edge_j);
#pragma unroll
            for (int pp = C::HL; pp < C::HL + C::SJ; ++pp) { A0[pp * C::PC] = s[pp]; A1[pp * C::PC] = d[pp]; }
        }
        __syncthreads();
    }
    // ---- merged store: out[2ip + pi, 2jq + pj] ----
    {
        T *db = dst + (int64_t)b * bs_d;
        constexpr int VPR = C::TIp / C::V;                // 16-byte pieces per owned quadrant row
        for (int idx = tid; idx < 2 * C::TJp * VPR; idx += C::NT) {
            const int t = idx % VPR;
            const int rest = idx / VPR;
            const int pj = rest & 1, qq = rest >> 1;       // owned dim-2 pair qq
            const int ql = C::HL + qq;
            T ws[C::V], wd[C::V];
            lds16<T, C::V>(ws, Sm + (0 * 2 + pj) * QSZ + ql * C::PC + C::CO + t * C::V);
            lds16<T, C::V>(wd, Sm + (1 * 2 + pj) * QSZ + ql * C::PC + C::CO + t * C::V);
            T *p = db + (int64_t)(2 * (jq0 + qq) + pj) * ld_d + 2 * (ip0 + t * C::V);
            if constexpr (sizeof(T) == 4) {
                *reinterpret_cast<float4 *>(p) = make_float4(ws[0], wd[0], ws[1], wd[1]);
                *reinterpret_cast<float4 *>(p + 4) = make_float4(ws[2], wd[2], ws[3], wd[3]);
            } else {
                *reinterpret_cast<double2 *>(p) = make_double2(ws[0], wd[0]);
                *reinterpret_cast<double2 *>(p + 2) = make_double2(ws[1], wd[1]);
            }
        }
    }
    __syncthreads();   // every thread is done with this buffer before it is refilled
    }
}

// ===================================================================================================
// one-shot variants (one tile per CTA): used when there are too few tiles to keep persistent CTAs busy
// ===================================================================================================
// one tile per CTA
template <typename T, class S, bool STRICT, class C>
__global__ void __launch_bounds__(C::NT)
k_lift2d_fwd_tma(const __grid_constant__ TensorMap tm_src, const T *__restrict__ src, int64_t ld_s, int64_t bs_s,
                 T *__restrict__ ll, int64_t ld_ll, int64_t bs_ll, T *__restrict__ yd, int64_t ld_y, int64_t bs_y,
                 int n, const __grid_constant__ typename CoefsOf<S, T>::type lc) {
    using fp = FP<STRICT>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw);
    T *Sm = reinterpret_cast<T *>(smem_raw + 128);
    const int nh = n >> 1;
    const int i0 = blockIdx.x * C::TI, j0 = blockIdx.y * C::TJ;
    const int b = blockIdx.z;
    const int tid = threadIdx.x;
    const int iorg = i0 - C::HLS, jorg = j0 - 2 * C::HL;          // global coordinates of local (0, 0)

    if (tid == 0) {
        mb_init(bar);
        mb_expect(bar, (uint32_t)(C::RJ * C::PI * sizeof(T)));
        tma_box3(Sm, &tm_src, iorg, jorg, b, bar);
    }
    __syncthreads();
    mb_wait(bar, 0);
    // ---- periodic wrap: border tiles patch the halo cells the tensor map zero-filled ----
    const bool edge_i = (blockIdx.x == 0) || (blockIdx.x == gridDim.x - 1);
    const bool edge_j = (blockIdx.y == 0) || (blockIdx.y == gridDim.y - 1);
    if (edge_i || edge_j) {
        // only the four halo strips can hold zero-filled cells: walk those (a 512^2 plane has 62 % border tiles, and a
        // sweep over the whole tile doubled their instruction count -- profiles/r01d_fir3d_db6_f32.md)
        const T *sb = src + (int64_t)b * bs_s;
        constexpr int WI = C::HLS + C::TI + 2 * C::HR;
        auto patch = [&](int r, int il) {
            const int gi = iorg + il, gj = jorg + r;
            if (gi < 0 || gi >= n || gj < 0 || gj >= n)
                Sm[r * C::PI + il] = sb[(int64_t)wrapi(gj, n) * ld_s + wrapi(gi, n)];
        };
        if (blockIdx.x == 0) {
            if constexpr (C::HLS > 0)
                for (int idx = tid; idx < C::RJ * C::HLS; idx += C::NT) patch(idx / C::HLS, idx % C::HLS);
        }
        if (blockIdx.x == gridDim.x - 1) {
            if constexpr (C::HR > 0)
                for (int idx = tid; idx < C::RJ * 2 * C::HR; idx += C::NT) patch(idx / (2 * C::HR), C::HLS + C::TI + idx % (2 * C::HR));
        }
        if (blockIdx.y == 0) {
            if constexpr (C::HL > 0)
                for (int idx = tid; idx < 2 * C::HL * WI; idx += C::NT) patch(idx / WI, idx % WI);
        }
        if (blockIdx.y == gridDim.y - 1) {
            if constexpr (C::HR > 0)
                for (int idx = tid; idx < 2 * C::HR * WI; idx += C::NT) patch(2 * C::HL + C::TJ + idx / WI, idx % WI);
        }
        __syncthreads();
    }
    // ---- dim-2 pass: one thread per (dim-1 sample, segment of dim-2 pairs) ----
    {
        T s[C::NPJ], d[C::NPJ];
        const bool act = tid < C::ROW_TASKS_F;
        constexpr int WI = 2 * C::HL + C::TI + 2 * C::HR;
        int il = 0, q = 0;
        if (act) {
            il = (C::HLS - 2 * C::HL) + tid % WI;
            q = tid / WI;
#pragma unroll
            for (int pp = 0; pp < C::NPJ; ++pp) {
                s[pp] = Sm[(2 * (q * C::SJ + pp)) * C::PI + il];
                d[pp] = Sm[(2 * (q * C::SJ + pp) + 1) * C::PI + il];
            }
        }
        __syncthreads();
        if (act) {
            const int jp0 = (j0 >> 1) - C::HL + q * C::SJ;
            TileTransform<T, S, STRICT, C::NPJ>::run(s, d, lc, wrapi(jp0, nh), nh, STRICT && edge_j);
#pragma unroll
            for (int pp = C::HL; pp < C::HL + C::SJ; ++pp) {
                if constexpr (STRICT && HasNorm<S>::value) {
                    Sm[(2 * (q * C::SJ + pp)) * C::PI + il] = fp::mul(s[pp], lc.n1);
                    Sm[(2 * (q * C::SJ + pp) + 1) * C::PI + il] = fp::mul(d[pp], lc.n2);
                } else {                                  // fast mode: the row factor rides in the dim-1 pass below
                    Sm[(2 * (q * C::SJ + pp)) * C::PI + il] = s[pp];
                    Sm[(2 * (q * C::SJ + pp) + 1) * C::PI + il] = d[pp];
                }
            }
        }
        __syncthreads();
    }
    // ---- dim-1 pass: one thread per (owned row, segment); 16-byte shared-memory accesses ----
    {
        T w[C::WINF];
        T s[C::NPI], d[C::NPI];
        const bool act = tid < C::COL_TASKS_F;
        constexpr int OFF = C::HLS - 2 * C::HL;           // samples between the vector-aligned window start and pair 0
        int r = 0, q = 0;
        if (act) {
            r = 2 * C::HL + tid % C::TJ;
            q = tid / C::TJ;
            lds16<T, C::WINF>(w, Sm + r * C::PI + 2 * q * C::SI);
#pragma unroll
            for (int pp = 0; pp < C::NPI; ++pp) { s[pp] = w[OFF + 2 * pp]; d[pp] = w[OFF + 2 * pp + 1]; }
        }
        __syncthreads();
        if (act) {
            TileTransform<T, S, STRICT, C::NPI>::run(s, d, lc, wrapi((i0 >> 1) - C::HL + q * C::SI, nh), nh, STRICT && edge_i);
            T os[C::SI], od[C::SI];
            if constexpr (HasNorm<S>::value) {
                T f1 = band_n1<S, T>(lc), f2 = band_n2<S, T>(lc);
                if constexpr (!STRICT) {                  // fast mode: (row factor) x (column factor) in one multiply
                    const T rowf = (r & 1) ? f2 : f1;     // 2*HL is even: staged row parity = dim-2 parity
                    f1 = f1 * rowf;
                    f2 = f2 * rowf;
                }
#pragma unroll
                for (int pp = 0; pp < C::SI; ++pp) {
                    os[pp] = fp::mul(s[C::HL + pp], f1);
                    od[pp] = fp::mul(d[C::HL + pp], f2);
                }
            } else {
#pragma unroll
                for (int pp = 0; pp < C::SI; ++pp) { os[pp] = s[C::HL + pp]; od[pp] = d[C::HL + pp]; }
            }
            // de-interleaved: the row's s-part at [HLS, HLS + TIp), its d-part behind it (every window of this row was
            // read before the barrier above), so the store phase moves whole 16-byte pieces
            sts16<T, C::SI>(Sm + r * C::PI + C::HLS + q * C::SI, os);
            sts16<T, C::SI>(Sm + r * C::PI + C::HLS + C::TIp + q * C::SI, od);
        }
        __syncthreads();
    }
    // ---- stores: a group of VPR lanes keeps one dim-2 parity and one 16-byte piece column and walks the tile's dim-2
    //      pairs, so every address is "previous + constant" (the index arithmetic of a flat loop was 38 % of the
    //      kernel's instructions); the shared-memory reads of a batch are issued before its stores.
    {
        T *llb = ll + (int64_t)b * bs_ll;
        T *yb = yd + (int64_t)b * bs_y;
        constexpr int VPR = C::TIp / C::V;                // 16-byte pieces per half row
        constexpr int G = (C::NT / VPR) & ~1;             // lane groups (even: half of them per dim-2 parity)
        static_assert(G >= 2 && C::TIp % C::V == 0, "store phase needs two lane groups");
        constexpr int GH = G / 2;
        constexpr int NIT = (C::TJp + GH - 1) / GH;
        constexpr int UB = 2;
        const int t = tid % VPR, g = tid / VPR;
        if (g < G) {
            const int pj = g & 1, q0 = g >> 1;
            const int ip = (i0 >> 1) + t * C::V;
            const int jq = (j0 >> 1) + q0;
            T *ps = (pj == 0) ? (llb + (int64_t)jq * ld_ll + ip) : (yb + (int64_t)(nh + jq) * ld_y + ip);
            T *pd = yb + (int64_t)(pj * nh + jq) * ld_y + nh + ip;
            const int64_t step_s = (int64_t)GH * ((pj == 0) ? ld_ll : ld_y), step_d = (int64_t)GH * ld_y;
            const T *sp = Sm + (2 * C::HL + 2 * q0 + pj) * C::PI + C::HLS + t * C::V;
#pragma unroll
            for (int it0 = 0; it0 < NIT; it0 += UB) {
                T ws[UB][C::V], wd[UB][C::V];
#pragma unroll
                for (int u = 0; u < UB; ++u)
                    if (it0 + u < NIT && (C::TJp % GH == 0 || q0 + (it0 + u) * GH < C::TJp)) {
                        lds16<T, C::V>(ws[u], sp + (it0 + u) * 2 * GH * C::PI);
                        lds16<T, C::V>(wd[u], sp + (it0 + u) * 2 * GH * C::PI + C::TIp);
                    }
#pragma unroll
                for (int u = 0; u < UB; ++u)
                    if (it0 + u < NIT && (C::TJp % GH == 0 || q0 + (it0 + u) * GH < C::TJp)) {
                        sts16<T, C::V>(ps, ws[u]);        // (generic 16-byte store helper: global pointers are fine)
                        sts16<T, C::V>(pd, wd[u]);
                        ps += step_s;
                        pd += step_d;
                    }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// inverse level
// ---------------------------------------------------------------------------------------------------
template <typename T, class S, bool STRICT, class C>
__global__ void __launch_bounds__(C::NT)
k_lift2d_inv_tma(const __grid_constant__ TensorMap tm_ll, const __grid_constant__ TensorMap tm_x,
                 const T *__restrict__ ll, int64_t ld_ll, int64_t bs_ll, const T *__restrict__ xd, int64_t ld_x, int64_t bs_x,
                 T *__restrict__ dst, int64_t ld_d, int64_t bs_d, int n, const __grid_constant__ typename CoefsOf<S, T>::type lc) {
    using fp = FP<STRICT>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw);
    T *Sm = reinterpret_cast<T *>(smem_raw + 128);
    constexpr int QSZ = C::QSZ;                           // one quadrant array; order [pi][pj]
    const int nh = n >> 1;
    const int ip0 = blockIdx.x * C::TIp, jq0 = blockIdx.y * C::TJp;
    const int b = blockIdx.z;
    const int tid = threadIdx.x;
    const int corg = ip0 - C::CO, qorg = jq0 - C::HL;     // quadrant coordinates of local (0, 0)

    if (tid == 0) {
        mb_init(bar);
        mb_expect(bar, (uint32_t)(4 * C::JQ * C::PC * sizeof(T)));
        tma_box3(Sm + 0 * QSZ, &tm_ll, corg, qorg, b, bar);                 // (pi, pj) = (0, 0): LL
        tma_box3(Sm + 1 * QSZ, &tm_x, corg, nh + qorg, b, bar);             // (0, 1)
        tma_box3(Sm + 2 * QSZ, &tm_x, nh + corg, qorg, b, bar);             // (1, 0)
        tma_box3(Sm + 3 * QSZ, &tm_x, nh + corg, nh + qorg, b, bar);        // (1, 1)
    }
    __syncthreads();
    mb_wait(bar, 0);
    const bool edge_i = (blockIdx.x == 0) || (blockIdx.x == gridDim.x - 1);
    const bool edge_j = (blockIdx.y == 0) || (blockIdx.y == gridDim.y - 1);
    if (edge_i || edge_j) {   // quadrant-relative wrap of the halo cells: the four halo strips of each quadrant array
        const T *llb = ll + (int64_t)b * bs_ll;
        const T *xb = xd + (int64_t)b * bs_x;
        auto patch = [&](int quad, int ql, int c) {
            const int gc = corg + c, gq = qorg + ql;
            if (gc < 0 || gc >= nh || gq < 0 || gq >= nh) {
                const int wc = wrapi(gc, nh), wq = wrapi(gq, nh);
                const int pi = quad >> 1, pj = quad & 1;
                Sm[quad * QSZ + ql * C::PC + c] = (quad == 0) ? llb[(int64_t)wq * ld_ll + wc]
                                                              : xb[(int64_t)(pj * nh + wq) * ld_x + pi * nh + wc];
            }
        };
        if (blockIdx.x == 0) {
            if constexpr (C::CO > 0)
                for (int idx = tid; idx < 4 * C::JQ * C::CO; idx += C::NT) patch(idx / (C::JQ * C::CO), (idx / C::CO) % C::JQ, idx % C::CO);
        }
        if (blockIdx.x == gridDim.x - 1) {
            if constexpr (C::HR > 0)
                for (int idx = tid; idx < 4 * C::JQ * C::HR; idx += C::NT) patch(idx / (C::JQ * C::HR), (idx / C::HR) % C::JQ, C::CO + C::TIp + idx % C::HR);
        }
        if (blockIdx.y == 0) {
            if constexpr (C::HL > 0)
                for (int idx = tid; idx < 4 * C::HL * C::IW; idx += C::NT) patch(idx / (C::HL * C::IW), (idx / C::IW) % C::HL, idx % C::IW);
        }
        if (blockIdx.y == gridDim.y - 1) {
            if constexpr (C::HR > 0)
                for (int idx = tid; idx < 4 * C::HR * C::IW; idx += C::NT) patch(idx / (C::HR * C::IW), C::HL + C::TJp + (idx / C::IW) % C::HR, idx % C::IW);
        }
        __syncthreads();
    }
    // ---- dim-1 pass first (inverse order): s_i = quadrant (0, pj), d_i = quadrant (1, pj), on every staged dim-2 pair
    {
        T ws[C::WINI], wd[C::WINI];
        T s[C::NPI], d[C::NPI];
        const bool act = tid < C::COL_TASKS_I;
        constexpr int OFF = C::CO - C::HL;
        T *Ao = Sm;
        int q = 0;
        if (act) {
            const int ql = tid % C::JQ;
            const int rest = tid / C::JQ;
            const int pj = rest & 1;
            q = rest >> 1;
            constexpr int SPH = C::TIp / (2 * C::SI);     // merged segments per quadrant-array row
            static_assert(C::TIp % (2 * C::SI) == 0, "merged segments must tile a quadrant row");
            const T *As = Sm + (0 * 2 + pj) * QSZ + ql * C::PC + q * C::SI;
            const T *Ad = Sm + (1 * 2 + pj) * QSZ + ql * C::PC + q * C::SI;
            Ao = Sm + ((q / SPH) * 2 + pj) * QSZ + ql * C::PC + C::CO + (q % SPH) * 2 * C::SI;
            lds16<T, C::WINI>(ws, As);
            lds16<T, C::WINI>(wd, Ad);
            if constexpr (HasNorm<S>::value) {
                T f1 = band_n1<S, T>(lc), f2 = band_n2<S, T>(lc);
                if constexpr (!STRICT) {                  // fast mode: both (reciprocal) factors of a quadrant in one multiply
                    const T rowf = pj ? f2 : f1;
                    f1 = f1 * rowf;
                    f2 = f2 * rowf;
                }
#pragma unroll
                for (int pp = 0; pp < C::NPI; ++pp) { s[pp] = fp::mul(ws[OFF + pp], f1); d[pp] = fp::mul(wd[OFF + pp], f2); }
            } else {
#pragma unroll
                for (int pp = 0; pp < C::NPI; ++pp) { s[pp] = ws[OFF + pp]; d[pp] = wd[OFF + pp]; }
            }
        }
        __syncthreads();
        if (act) {
            TileTransform<T, S, STRICT, C::NPI>::run(s, d, lc, wrapi(ip0 - C::HL + q * C::SI, nh), nh, STRICT && edge_i);
            T o[2 * C::SI];
#pragma unroll
            for (int pp = 0; pp < C::SI; ++pp) { o[2 * pp] = s[C::HL + pp]; o[2 * pp + 1] = d[C::HL + pp]; }
            // merged (interleaved) along dim 1 already here: output sample i of this row lives in array (i / TIp, pj) at
            // column CO + i % TIp -- the dim-2 pass below only needs the two pj arrays to agree on the column labels
            sts16<T, 2 * C::SI>(Ao, o);
        }
        __syncthreads();
    }
    // ---- dim-2 pass on the owned dim-1 pairs: s_j = quadrant (pi, 0), d_j = quadrant (pi, 1) ----
    {
        T s[C::NPJ], d[C::NPJ];
        const bool act = tid < C::ROW_TASKS_I;
        T *A0 = Sm, *A1 = Sm;
        int q = 0;
        if (act) {
            const int c = C::CO + tid % C::TIp;
            const int rest = tid / C::TIp;
            const int pi = rest & 1;
            q = rest >> 1;
            A0 = Sm + (pi * 2 + 0) * QSZ + (q * C::SJ) * C::PC + c;
            A1 = Sm + (pi * 2 + 1) * QSZ + (q * C::SJ) * C::PC + c;
#pragma unroll
            for (int pp = 0; pp < C::NPJ; ++pp) {
                if constexpr (STRICT && HasNorm<S>::value) { s[pp] = fp::mul(A0[pp * C::PC], lc.n1); d[pp] = fp::mul(A1[pp * C::PC], lc.n2); }
                else { s[pp] = A0[pp * C::PC]; d[pp] = A1[pp * C::PC]; }      // lifting, fast mode: scaled in the dim-1 pass
            }
        }
        __syncthreads();
        if (act) {
            TileTransform<T, S, STRICT, C::NPJ>::run(s, d, lc, wrapi(jq0 - C::HL + q * C::SJ, nh), nh, STRICT && edge_j);
#pragma unroll
            for (int pp = C::HL; pp < C::HL + C::SJ; ++pp) { A0[pp * C::PC] = s[pp]; A1[pp * C::PC] = d[pp]; }
        }
        __syncthreads();
    }
    // ---- store: the rows are already merged along dim 1; a group of VPR lanes keeps one dim-2 parity and one 16-byte
    //      piece column and walks the tile's dim-2 pairs (every address is "previous + constant").
    {
        T *db = dst + (int64_t)b * bs_d;
        constexpr int VPR = C::TI / C::V;                 // pieces per output row
        constexpr int VPH = C::TIp / C::V;                // pieces per quadrant-array row
        constexpr int G = (C::NT / VPR) & ~1;
        static_assert(G >= 2, "store phase needs two lane groups");
        constexpr int GH = G / 2;
        constexpr int NIT = (C::TJp + GH - 1) / GH;
        constexpr int UB = 4;
        const int t = tid % VPR, g = tid / VPR;
        if (g < G) {
            const int pj = g & 1, q0 = g >> 1;
            const T *sp = Sm + ((t / VPH) * 2 + pj) * QSZ + (C::HL + q0) * C::PC + C::CO + (t % VPH) * C::V;
            T *p = db + (int64_t)(2 * (jq0 + q0) + pj) * ld_d + 2 * ip0 + t * C::V;
            const int64_t step = (int64_t)2 * GH * ld_d;
#pragma unroll
            for (int it0 = 0; it0 < NIT; it0 += UB) {
                T w[UB][C::V];
#pragma unroll
                for (int u = 0; u < UB; ++u)
                    if (it0 + u < NIT && (C::TJp % GH == 0 || q0 + (it0 + u) * GH < C::TJp))
                        lds16<T, C::V>(w[u], sp + (it0 + u) * GH * C::PC);
#pragma unroll
                for (int u = 0; u < UB; ++u)
                    if (it0 + u < NIT && (C::TJp % GH == 0 || q0 + (it0 + u) * GH < C::TJp)) {
                        sts16<T, C::V>(p, w[u]);
                        p += step;
                    }
            }
        }
    }
}

} // namespace wb
