// generic_kernels.cu -- one-level analysis / synthesis passes over arbitrary strided line sets.
//
// These kernels implement the reference's 1-level primitives for ANY line length, filter length (odd
// lengths, n < flen multi-wrap), lifting scheme and memory layout:
//   filter analysis  = filtdown! x2 (approx + detail)        src/Transforms/transforms_filter.jl:387-433
//   filter synthesis = filtup!  x2 (approx, then detail add) src/Transforms/transforms_filter.jl:467-541
//   lifting forward  = split! -> lift! per step -> normalize! src/Transforms/transforms_lifting.jl:91-104
//   lifting inverse  = normalize! -> lift! per step -> merge! src/Transforms/transforms_lifting.jl:105-119
// in closed form with the reference's summation order.  They are the coverage path:
// every N-D / WPT / odd-size configuration is correct through them; the fused sm_100a kernels
// (fused1d.cu, fused2d.cu) take over for the bandwidth-critical shapes.
#include "common.cuh"

namespace wb {

// ---------------------------------------------------------------------------------------------------
// index helpers
// ---------------------------------------------------------------------------------------------------
struct Coords {
    int64_t c1, c2, c3;
};
__device__ __forceinline__ Coords split_y(int64_t yy, const Extent &e) {
    Coords c;
    c.c1 = yy % e.n[1];
    int64_t r = yy / e.n[1];
    c.c2 = r % e.n[2];
    c.c3 = r / e.n[2];
    return c;
}
__device__ __forceinline__ int64_t wrap_mod(int64_t a, int64_t n) {
    int64_t r = a % n;
    return r < 0 ? r + n : r;
}

// ---------------------------------------------------------------------------------------------------
// filter analysis: one thread per output pair (a[k], d[k])
//   a[k] = sum_{m=0..F-1}      h[m]     * x[(2k+m)   mod n]   (increasing m)
//   d[k] = sum_{p=0..F-1}      g[F-1-p] * x[(2k+2-F+p) mod n] (increasing x index)
// ---------------------------------------------------------------------------------------------------
template <typename T, bool STRICT, bool KFAST>
__global__ void __launch_bounds__(256)
k_filter_analysis(View<const T> src, View<T> dlo, View<T> dhi, Extent e,
                  const __grid_constant__ FilterCoefs<T> fc, const uint8_t *__restrict__ active) {
    using fp = FP<STRICT>;
    const int64_t n = e.len, nh = n >> 1;
    const int64_t fastN = nh * e.n[0];
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= fastN) return;
    int64_t k, c0;
    if (KFAST) { k = r % nh; c0 = r / nh; } else { c0 = r % e.n[0]; k = r / e.n[0]; }
    const int64_t NY = e.n[1] * e.n[2] * e.n[3];
    const int F = fc.F;
    for (int64_t yy = blockIdx.y; yy < NY; yy += gridDim.y) {
        const Coords c = split_y(yy, e);
        const T *x = src.line(c0, c.c1, c.c2, c.c3);
        T *lo = dlo.line(c0, c.c1, c.c2, c.c3);
        T *hi = dhi.line(c0, c.c1, c.c2, c.c3);
        if (active != nullptr && !active[c.c1]) { // WPT leaf: carry the samples through unchanged
            const int64_t j0 = 2 * k, j1 = 2 * k + 1;
            const T v0 = x[j0 * src.ls], v1 = x[j1 * src.ls];
            if (j0 < nh) lo[j0 * dlo.ls] = v0; else hi[(j0 - nh) * dhi.ls] = v0;
            if (j1 < nh) lo[j1 * dlo.ls] = v1; else hi[(j1 - nh) * dhi.ls] = v1;
            continue;
        }
        T a, d;
        const int64_t b0 = 2 * k;
        if (b0 + F <= n && b0 + 2 - F >= 0) { // no wrap
            const T *xa = x + b0 * src.ls;
            a = fp::mul(fc.h[0], xa[0]);
            for (int m = 1; m < F; ++m) a = fp::mac(a, fc.h[m], xa[m * src.ls]);
            const T *xd = x + (b0 + 2 - F) * src.ls;
            d = fp::mul(fc.g[F - 1], xd[0]);
            for (int p = 1; p < F; ++p) d = fp::mac(d, fc.g[F - 1 - p], xd[p * src.ls]);
        } else {
            int64_t ia = b0 % n; // b0 < n always
            a = fp::mul(fc.h[0], x[ia * src.ls]);
            for (int m = 1; m < F; ++m) {
                if (++ia == n) ia = 0;
                a = fp::mac(a, fc.h[m], x[ia * src.ls]);
            }
            int64_t id = wrap_mod(b0 + 2 - F, n);
            d = fp::mul(fc.g[F - 1], x[id * src.ls]);
            for (int p = 1; p < F; ++p) {
                if (++id == n) id = 0;
                d = fp::mac(d, fc.g[F - 1 - p], x[id * src.ls]);
            }
        }
        lo[k * dlo.ls] = a;
        hi[k * dhi.ls] = d;
    }
}

// ---------------------------------------------------------------------------------------------------
// filter synthesis: one thread per output pair (x[2u], x[2u+1])
//   x[t] = ( sum_{m == t (2), decreasing m} h[m] a[((t-m)/2) mod nh] )
//        + ( sum_{m != t (2), increasing m} g[m] d[((t+m-1)/2) mod nh] )
// both partial sums run over increasing band index; the two bands are summed separately, then added.
// ---------------------------------------------------------------------------------------------------
template <typename T, bool STRICT, bool KFAST>
__global__ void __launch_bounds__(256)
k_filter_synthesis(View<const T> slo, View<const T> shi, View<const T> salt, Extent e,
                   int64_t thr0, int64_t thr1, int64_t thr2, int64_t thr3, int has_alt, View<T> dst,
                   const __grid_constant__ FilterCoefs<T> fc, const uint8_t *__restrict__ active) {
    using fp = FP<STRICT>;
    const int64_t n = e.len, nh = n >> 1;
    const int64_t fastN = nh * e.n[0];
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= fastN) return;
    int64_t u, c0;
    if (KFAST) { u = r % nh; c0 = r / nh; } else { c0 = r % e.n[0]; u = r / e.n[0]; }
    const int64_t NY = e.n[1] * e.n[2] * e.n[3];
    const int F = fc.F;
    const int me = (F - 1) & ~1;      // largest even tap index
    const int mo = (F - 1) - ((F - 1) & 1 ? 0 : 1); // largest odd tap index (>= 1 since F >= 2)
    for (int64_t yy = blockIdx.y; yy < NY; yy += gridDim.y) {
        const Coords c = split_y(yy, e);
        const bool alt = has_alt && c0 < thr0 && c.c1 < thr1 && c.c2 < thr2 && c.c3 < thr3;
        const T *a;
        int64_t als;
        if (alt) { a = salt.line(c0, c.c1, c.c2, c.c3); als = salt.ls; }
        else     { a = slo.line(c0, c.c1, c.c2, c.c3);  als = slo.ls; }
        const T *d = shi.line(c0, c.c1, c.c2, c.c3);
        T *o = dst.line(c0, c.c1, c.c2, c.c3);
        if (active != nullptr && !active[c.c1]) {
            o[u * dst.ls] = a[u * als];
            o[(nh + u) * dst.ls] = d[u * shi.ls];
            continue;
        }
        // even output t = 2u: approx taps m = me, me-2, ..., 0 -> a[u - m/2]; detail taps m = 1,3,.. -> d[u + (m-1)/2]
        T ra, rd, x0, x1;
        {
            int64_t ia = wrap_mod(u - me / 2, nh);
            ra = fp::mul(fc.h[me], a[ia * als]);
            for (int m = me - 2; m >= 0; m -= 2) {
                if (++ia == nh) ia = 0;
                ra = fp::mac(ra, fc.h[m], a[ia * als]);
            }
            int64_t id = u;
            rd = fp::mul(fc.g[1], d[id * shi.ls]);
            for (int m = 3; m < F; m += 2) {
                if (++id == nh) id = 0;
                rd = fp::mac(rd, fc.g[m], d[id * shi.ls]);
            }
            x0 = fp::add(ra, rd);
        }
        // odd output t = 2u+1: approx taps m = mo, mo-2, ..., 1 -> a[u - (m-1)/2]; detail taps m = 0,2,.. -> d[u + m/2]
        {
            int64_t ia = wrap_mod(u - (mo - 1) / 2, nh);
            ra = fp::mul(fc.h[mo], a[ia * als]);
            for (int m = mo - 2; m >= 1; m -= 2) {
                if (++ia == nh) ia = 0;
                ra = fp::mac(ra, fc.h[m], a[ia * als]);
            }
            int64_t id = u;
            rd = fp::mul(fc.g[0], d[id * shi.ls]);
            for (int m = 2; m < F; m += 2) {
                if (++id == nh) id = 0;
                rd = fp::mac(rd, fc.g[m], d[id * shi.ls]);
            }
            x1 = fp::add(ra, rd);
        }
        o[(2 * u) * dst.ls] = x0;
        o[(2 * u + 1) * dst.ls] = x1;
    }
}

// ---------------------------------------------------------------------------------------------------
// lifting: shared-memory tile kernel.  A CTA owns `tp` polyphase pairs of `tln` lines; it stages the pairs
// plus the scheme's halo (or the whole line when it fits: `whole`), runs every predict/update step in
// shared memory and writes the scaled halves (forward) or the merged line (inverse).
//   Predict: s[i] += sum_k c[k] d[(i+k-shift) mod half]   (writes the first half)
//   Update : d[i] += sum_k c[k] s[(i+k-shift) mod half]   (writes the second half)
// FP order (lift_inbounds!/lift_perboundary!, transforms_lifting.jl:437-483): elements whose taps do not
// wrap use x + ((c0*p0 + c1*p1) + c2*p2) for nc <= 3; wrapped elements and nc > 3 accumulate sequentially.
// ---------------------------------------------------------------------------------------------------
struct LiftTile {
    int tp;     // owned pairs per tile
    int tln;    // lines per CTA
    int W;      // staged pairs per line (tp + halo_l + halo_r, or half when whole)
    int whole;  // 1: the whole line is resident, periodic indexing inside the tile
    int64_t ntiles;  // tiles along a line
    int64_t NL;      // total number of lines
};

template <typename T, bool STRICT>
__device__ __forceinline__ T lift_value(const LiftScheme<T> &sc, int st, T v, const T *__restrict__ other,
                                        int64_t pos, int64_t gi, int64_t half, bool whole, int stride) {
    using fp = FP<STRICT>;
    const int nc = sc.nc[st];
    const int sh = sc.shift[st];
    const bool interior = (gi >= (sh > 0 ? sh : 0)) && (gi <= half + sh - nc) && (gi <= half - 1);
    auto tap = [&](int k) -> T {
        int64_t q = pos + k - sh;
        if (whole) q = wrap_mod(q, half);
        return other[q * stride];
    };
    if (interior && nc <= 3) {
        T acc = fp::mul(sc.coef[st][0], tap(0));
        if (nc >= 2) acc = fp::mac(acc, sc.coef[st][1], tap(1));
        if (nc >= 3) acc = fp::mac(acc, sc.coef[st][2], tap(2));
        return fp::add(v, acc);
    }
    for (int k = 0; k < nc; ++k) v = fp::mac(v, sc.coef[st][k], tap(k));
    return v;
}

template <typename T, bool STRICT, bool FW, bool KFAST>
__global__ void __launch_bounds__(256)
k_lifting(View<const T> s0 /*FW: src ; INV: approx half*/, View<const T> s1 /*INV: detail half*/,
          View<const T> salt, int64_t thr0, int64_t thr1, int64_t thr2, int64_t thr3, int has_alt,
          View<T> d0 /*FW: approx half ; INV: dst*/, View<T> d1 /*FW: detail half*/,
          Extent e, const __grid_constant__ LiftScheme<T> sc, LiftTile tile,
          const uint8_t *__restrict__ active) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int W = tile.W, TLN = tile.tln;
    T *S = reinterpret_cast<T *>(smem_raw);
    T *D = S + (size_t)W * TLN;
    const size_t data_bytes = (((size_t)2 * W * TLN * sizeof(T)) + 15) & ~(size_t)15;
    int64_t *offs = reinterpret_cast<int64_t *>(smem_raw + data_bytes); // [4][TLN]: in0, in1, out0, out1
    int *lact = reinterpret_cast<int *>(offs + 4 * TLN);             // [TLN]  : 1 active, 0 leaf copy, -1 absent
    const int64_t half = e.len >> 1;
    const int HL = tile.whole ? 0 : sc.halo_l;
    const int nthreads = blockDim.x;
    const int tid = threadIdx.x;
    const int64_t ngroups = (tile.NL + TLN - 1) / TLN;

    for (int64_t grp = blockIdx.y; grp < ngroups; grp += gridDim.y) {
        const int64_t p0 = (int64_t)blockIdx.x * tile.tp; // first owned pair
        const int64_t l0 = grp * TLN;
        __syncthreads(); // previous iteration's smem fully consumed
        // ---- per-line base offsets ----
        for (int l = tid; l < TLN; l += nthreads) {
            int64_t ln = l0 + l;
            if (ln >= tile.NL) { lact[l] = -1; continue; }
            int64_t c0 = ln % e.n[0]; int64_t r = ln / e.n[0];
            int64_t c1 = r % e.n[1]; r /= e.n[1];
            int64_t c2 = r % e.n[2]; int64_t c3 = r / e.n[2];
            const bool alt = has_alt && c0 < thr0 && c1 < thr1 && c2 < thr2 && c3 < thr3;
            if (FW) {
                offs[0 * TLN + l] = s0.line(c0, c1, c2, c3) - s0.p;
                offs[1 * TLN + l] = 0;
                offs[2 * TLN + l] = d0.line(c0, c1, c2, c3) - d0.p;
                offs[3 * TLN + l] = d1.line(c0, c1, c2, c3) - d1.p;
            } else {
                offs[0 * TLN + l] = alt ? (salt.line(c0, c1, c2, c3) - salt.p) : (s0.line(c0, c1, c2, c3) - s0.p);
                offs[1 * TLN + l] = s1.line(c0, c1, c2, c3) - s1.p;
                offs[2 * TLN + l] = d0.line(c0, c1, c2, c3) - d0.p;
                offs[3 * TLN + l] = 0;
            }
            int a = 1;
            if (active != nullptr && !active[c1]) a = 0;
            lact[l] = a | (alt ? 2 : 0);
        }
        __syncthreads();
        // ---- load (split fused into the load; inverse: normalise while loading) ----
        const int64_t nload = (int64_t)2 * W * TLN;
        for (int64_t it = tid; it < nload; it += nthreads) {
            int l, q;
            if (KFAST) { q = (int)(it % (2 * W)); l = (int)(it / (2 * W)); }
            else       { l = (int)(it % TLN); q = (int)(it / TLN); }
            const int la = lact[l];
            if (la < 0) continue;
            int pos, par;
            if (FW) { pos = q >> 1; par = q & 1; }           // interleaved source: coalesced along the line
            else    { par = q >= W; pos = par ? q - W : q; } // two dense halves
            int64_t gi = tile.whole ? pos : p0 - HL + pos;
            if (!tile.whole) { if (gi < 0) gi += half; else if (gi >= half) gi -= half; }
            if (tile.whole && pos >= half) continue;
            T v;
            if (FW) {
                v = s0.p[offs[l] + (2 * gi + par) * s0.ls];
            } else {
                const bool alt = (la & 2) != 0;
                if (par == 0) v = alt ? salt.p[offs[l] + gi * salt.ls] : s0.p[offs[l] + gi * s0.ls];
                else          v = s1.p[offs[TLN + l] + gi * s1.ls];
                if (la & 1) v = FP<STRICT>::mul(v, par ? sc.norm2 : sc.norm1); // normalize! before the steps
            }
            const size_t si = KFAST ? (size_t)l * W + pos : (size_t)pos * TLN + l;
            (par ? D : S)[si] = v;
        }
        __syncthreads();
        // ---- lifting steps ----
        int lo_s = 0, hi_s = tile.whole ? (int)half : W, lo_d = lo_s, hi_d = hi_s;
        const int stride = KFAST ? 1 : TLN;
        for (int st = 0; st < sc.nsteps; ++st) {
            const int sh = sc.shift[st], nc = sc.nc[st];
            const int left = sh > 0 ? sh : 0, right = (nc - 1 - sh) > 0 ? (nc - 1 - sh) : 0;
            int lo, hi;
            const bool pred = sc.is_predict[st] != 0;
            if (tile.whole) { lo = 0; hi = (int)half; }
            else if (pred) { lo = max(lo_s, lo_d + left); hi = min(hi_s, hi_d - right); lo_s = lo; hi_s = hi; }
            else           { lo = max(lo_d, lo_s + left); hi = min(hi_d, hi_s - right); lo_d = lo; hi_d = hi; }
            const int cnt = hi - lo;
            T *tgt = pred ? S : D;
            const T *oth = pred ? D : S;
            for (int64_t it = tid; it < (int64_t)cnt * TLN; it += nthreads) {
                int l, pos;
                if (KFAST) { pos = lo + (int)(it % cnt); l = (int)(it / cnt); }
                else       { l = (int)(it % TLN); pos = lo + (int)(it / TLN); }
                if (!(lact[l] > 0 && (lact[l] & 1))) continue;
                int64_t gi = tile.whole ? pos : p0 - HL + pos;
                if (!tile.whole) { if (gi < 0) gi += half; else if (gi >= half) gi -= half; }
                const size_t base = KFAST ? (size_t)l * W : (size_t)l;
                const size_t si = KFAST ? base + pos : (size_t)pos * TLN + l;
                tgt[si] = lift_value<T, STRICT>(sc, st, tgt[si], oth + base, pos, gi, half, tile.whole != 0, stride);
            }
            __syncthreads();
        }
        // ---- store owned pairs ----
        const int own0 = HL;
        int64_t ownn = tile.whole ? half : min((int64_t)tile.tp, half - p0);
        if (ownn < 0) ownn = 0;
        const int64_t nstore = (int64_t)2 * ownn * TLN;
        for (int64_t it = tid; it < nstore; it += nthreads) {
            int l; int64_t q;
            if (KFAST) { q = it % (2 * ownn); l = (int)(it / (2 * ownn)); }
            else       { l = (int)(it % TLN); q = it / TLN; }
            const int la = lact[l];
            if (la < 0) continue;
            int64_t j; int par;
            if (FW) { par = q >= ownn; j = par ? q - ownn : q; } // two dense halves out
            else    { j = q >> 1; par = (int)(q & 1); }          // merged line out
            const int pos = own0 + (int)j;
            const int64_t gi = tile.whole ? j : p0 + j;
            const size_t si = KFAST ? (size_t)l * W + pos : (size_t)pos * TLN + l;
            T v = (par ? D : S)[si];
            if (FW) {
                if (la & 1) {
                    v = FP<STRICT>::mul(v, par ? sc.norm2 : sc.norm1);
                    if (par) d1.p[offs[3 * TLN + l] + gi * d1.ls] = v;
                    else     d0.p[offs[2 * TLN + l] + gi * d0.ls] = v;
                } else { // leaf: sample 2gi+par goes back to position 2gi+par of the [lo|hi] line
                    const int64_t jj = 2 * gi + par;
                    if (jj < half) d0.p[offs[2 * TLN + l] + jj * d0.ls] = v;
                    else           d1.p[offs[3 * TLN + l] + (jj - half) * d1.ls] = v;
                }
            } else {
                if (la & 1) d0.p[offs[2 * TLN + l] + (2 * gi + par) * d0.ls] = v;
                else        d0.p[offs[2 * TLN + l] + (par ? half + gi : gi) * d0.ls] = v;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// strided line copy
// ---------------------------------------------------------------------------------------------------
template <typename T, bool KFAST>
__global__ void __launch_bounds__(256)
k_copy_lines(View<const T> src, View<T> dst, Extent e) {
    const int64_t fastN = e.len * e.n[0];
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= fastN) return;
    int64_t k, c0;
    if (KFAST) { k = r % e.len; c0 = r / e.len; } else { c0 = r % e.n[0]; k = r / e.n[0]; }
    const int64_t NY = e.n[1] * e.n[2] * e.n[3];
    for (int64_t yy = blockIdx.y; yy < NY; yy += gridDim.y) {
        const Coords c = split_y(yy, e);
        dst.line(c0, c.c1, c.c2, c.c3)[k * dst.ls] = src.line(c0, c.c1, c.c2, c.c3)[k * src.ls];
    }
}

// ---------------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------------
static inline bool grid_for(const Extent &e, int64_t per_line, dim3 &grid, dim3 &block) {
    const int64_t fastN = per_line * e.n[0];
    const int64_t NY = e.n[1] * e.n[2] * e.n[3];
    if (fastN <= 0 || NY <= 0) return false;
    const int64_t gx = (fastN + 255) / 256;
    if (gx > 0x7fffffffLL) { set_error("line set too large for one launch"); return false; }
    block = dim3(256);
    grid = dim3((unsigned)gx, (unsigned)(NY < 65535 ? NY : 65535));
    return true;
}

// Short lines with a trivial coordinate 0 (packet nodes, late levels): the x dimension of the grid only spans
// (position, coordinate 0), so fold coordinate 1 into coordinate 0 to keep whole CTAs busy.
template <typename VT> static inline void swap_slots01(VT &v) { const int64_t t = v.s[0]; v.s[0] = v.s[1]; v.s[1] = t; }
static inline bool want_swap(const Extent &e, const uint8_t *active) {
    return active == nullptr && e.n[0] == 1 && e.n[1] > 1 && e.len < 1024;
}

template <typename T>
bool launch_filter_analysis(const View<const T> &src_, const View<T> &dlo_, const View<T> &dhi_,
                            const Extent &e_, const FilterCoefs<T> &fc, bool strict, cudaStream_t st,
                            const uint8_t *active) {
    View<const T> src = src_; View<T> dlo = dlo_, dhi = dhi_; Extent e = e_;
    if (want_swap(e, active)) { swap_slots01(src); swap_slots01(dlo); swap_slots01(dhi); e.n[0] = e.n[1]; e.n[1] = 1; }
    dim3 grid, block;
    if (!grid_for(e, e.len / 2, grid, block)) return true; // nothing to do
    const bool kfast = (src.ls == 1) || e.n[0] == 1;
    LaunchScope scope("generic_filter_analysis", st);
#define WB_LAUNCH(S, K) k_filter_analysis<T, S, K><<<grid, block, 0, st>>>(src, dlo, dhi, e, fc, active)
    if (strict) { if (kfast) WB_LAUNCH(true, true); else WB_LAUNCH(true, false); }
    else        { if (kfast) WB_LAUNCH(false, true); else WB_LAUNCH(false, false); }
#undef WB_LAUNCH
    return check_launch("filter_analysis");
}

template <typename T>
bool launch_filter_synthesis(const View<const T> &slo_, const View<const T> &shi_, const View<const T> &salt_,
                             const int64_t thr_[4], bool has_alt, const View<T> &dst_,
                             const Extent &e_, const FilterCoefs<T> &fc, bool strict, cudaStream_t st,
                             const uint8_t *active) {
    View<const T> slo = slo_, shi = shi_, salt = salt_; View<T> dst = dst_; Extent e = e_;
    int64_t thr[4] = {thr_[0], thr_[1], thr_[2], thr_[3]};
    if (want_swap(e, active)) {
        swap_slots01(slo); swap_slots01(shi); swap_slots01(salt); swap_slots01(dst);
        e.n[0] = e.n[1]; e.n[1] = 1;
        const int64_t t = thr[0]; thr[0] = thr[1]; thr[1] = t;
    }
    dim3 grid, block;
    if (!grid_for(e, e.len / 2, grid, block)) return true;
    const bool kfast = (dst.ls == 1) || e.n[0] == 1;
    LaunchScope scope("generic_filter_synthesis", st);
#define WB_LAUNCH(S, K) k_filter_synthesis<T, S, K><<<grid, block, 0, st>>>(slo, shi, salt, e, thr[0], thr[1], thr[2], thr[3], has_alt ? 1 : 0, dst, fc, active)
    if (strict) { if (kfast) WB_LAUNCH(true, true); else WB_LAUNCH(true, false); }
    else        { if (kfast) WB_LAUNCH(false, true); else WB_LAUNCH(false, false); }
#undef WB_LAUNCH
    return check_launch("filter_synthesis");
}

template <typename T>
static bool plan_lift_tile(const Extent &e, const LiftScheme<T> &sc, bool kfast, LiftTile &t, size_t &smem) {
    const int64_t half = e.len / 2;
    const int64_t NL = e.n[0] * e.n[1] * e.n[2] * e.n[3];
    const int cap = (int)(40 * 1024 / (2 * sizeof(T))); // staged pairs per CTA (both arrays within ~40 KB)
    const int halo = sc.halo_l + sc.halo_r;
    t.NL = NL;
    if (kfast) {
        if (half <= 1024 && half <= cap) {
            t.whole = 1; t.W = (int)half; t.tp = (int)half;
            int64_t tl = 1024 / half; if (tl < 1) tl = 1; if (tl > NL) tl = NL; if (tl * half > cap) tl = cap / half;
            // very short lines: the per-line bookkeeping (36 bytes) outweighs the samples -- keep the CTA within the default
            // 48 KB of dynamic shared memory (1024 lines of 2 samples asked for 53 KB: launch failure)
            const int64_t per_line_bytes = 2 * half * (int64_t)sizeof(T) + 4 * (int64_t)sizeof(int64_t) + (int64_t)sizeof(int);
            if (tl * per_line_bytes > 40 * 1024) tl = (40 * 1024) / per_line_bytes;
            t.tln = (int)(tl < 1 ? 1 : tl);
        } else {
            t.whole = 0; t.tln = 1; t.tp = 1024 - halo; if (t.tp > cap - halo) t.tp = cap - halo;
            if (t.tp < 16) { set_error("lifting halo %d too large for the generic tile", halo); return false; }
            t.W = t.tp + halo;
        }
    } else {
        t.tln = (int)(e.n[0] < 32 ? e.n[0] : 32);
        if (NL < t.tln) t.tln = (int)NL;
        const int per_line = cap / t.tln;
        if (half <= per_line) { t.whole = 1; t.W = (int)half; t.tp = (int)half; }
        else {
            t.whole = 0; t.tp = per_line - halo; if (t.tp > 64) t.tp = 64;
            if (t.tp < 4) { set_error("lifting halo %d too large for the generic tile", halo); return false; }
            t.W = t.tp + halo;
        }
    }
    if (!t.whole && half < t.W) { // cannot happen by construction (half > cap >= W); guard anyway
        set_error("internal: windowed lifting tile wider than the line"); return false;
    }
    t.ntiles = t.whole ? 1 : (half + t.tp - 1) / t.tp;
    smem = ((((size_t)2 * t.W * t.tln * sizeof(T)) + 15) & ~(size_t)15) + (size_t)t.tln * (4 * sizeof(int64_t) + sizeof(int));
    smem = (smem + 15) & ~(size_t)15;
    return true;
}

template <typename T, bool FW>
static bool launch_lifting(const View<const T> &s0, const View<const T> &s1, const View<const T> &salt,
                           const int64_t thr[4], bool has_alt, const View<T> &d0, const View<T> &d1,
                           const Extent &e, const LiftScheme<T> &sc, bool strict, cudaStream_t st,
                           const uint8_t *active) {
    if (e.len < 2) return true;
    const bool kfast = FW ? ((s0.ls == 1) || e.n[0] == 1) : ((d0.ls == 1) || e.n[0] == 1);
    LiftTile t; size_t smem;
    if (!plan_lift_tile(e, sc, kfast, t, smem)) return false;
    if (t.NL <= 0) return true;
    const int64_t ngroups = (t.NL + t.tln - 1) / t.tln;
    dim3 grid((unsigned)t.ntiles, (unsigned)(ngroups < 65535 ? ngroups : 65535)), block(256);
    LaunchScope scope(FW ? "generic_lifting_forward" : "generic_lifting_inverse", st);
#define WB_LAUNCH(S, K) k_lifting<T, S, FW, K><<<grid, block, smem, st>>>(s0, s1, salt, thr[0], thr[1], thr[2], thr[3], has_alt ? 1 : 0, d0, d1, e, sc, t, active)
    if (strict) { if (kfast) WB_LAUNCH(true, true); else WB_LAUNCH(true, false); }
    else        { if (kfast) WB_LAUNCH(false, true); else WB_LAUNCH(false, false); }
#undef WB_LAUNCH
    return check_launch("lifting");
}

template <typename T>
bool launch_lifting_analysis(const View<const T> &src, const View<T> &dlo, const View<T> &dhi,
                             const Extent &e, const LiftScheme<T> &sc, bool strict, cudaStream_t st,
                             const uint8_t *active) {
    const int64_t thr[4] = {0, 0, 0, 0};
    View<const T> none{nullptr, 0, {0, 0, 0, 0}};
    return launch_lifting<T, true>(src, none, none, thr, false, dlo, dhi, e, sc, strict, st, active);
}
template <typename T>
bool launch_lifting_synthesis(const View<const T> &slo, const View<const T> &shi, const View<const T> &salt,
                              const int64_t thr[4], bool has_alt, const View<T> &dst,
                              const Extent &e, const LiftScheme<T> &sc, bool strict, cudaStream_t st,
                              const uint8_t *active) {
    View<T> none{nullptr, 0, {0, 0, 0, 0}};
    return launch_lifting<T, false>(slo, shi, salt, thr, has_alt, dst, none, e, sc, strict, st, active);
}

template <typename T>
bool launch_copy_lines(const View<const T> &src, const View<T> &dst, const Extent &e, cudaStream_t st) {
    dim3 grid, block;
    if (!grid_for(e, e.len, grid, block)) return true;
    const bool kfast = (src.ls == 1 && dst.ls == 1) || e.n[0] == 1;
    LaunchScope scope("copy_lines", st);
    if (kfast) k_copy_lines<T, true><<<grid, block, 0, st>>>(src, dst, e);
    else       k_copy_lines<T, false><<<grid, block, 0, st>>>(src, dst, e);
    return check_launch("copy_lines");
}

#define WB_INST(T)                                                                                                      \
    template bool launch_filter_analysis<T>(const View<const T> &, const View<T> &, const View<T> &, const Extent &,   \
                                            const FilterCoefs<T> &, bool, cudaStream_t, const uint8_t *);              \
    template bool launch_filter_synthesis<T>(const View<const T> &, const View<const T> &, const View<const T> &,      \
                                             const int64_t[4], bool, const View<T> &, const Extent &,                  \
                                             const FilterCoefs<T> &, bool, cudaStream_t, const uint8_t *);             \
    template bool launch_lifting_analysis<T>(const View<const T> &, const View<T> &, const View<T> &, const Extent &,  \
                                             const LiftScheme<T> &, bool, cudaStream_t, const uint8_t *);              \
    template bool launch_lifting_synthesis<T>(const View<const T> &, const View<const T> &, const View<const T> &,     \
                                              const int64_t[4], bool, const View<T> &, const Extent &,                 \
                                              const LiftScheme<T> &, bool, cudaStream_t, const uint8_t *);             \
    template bool launch_copy_lines<T>(const View<const T> &, const View<T> &, const Extent &, cudaStream_t);
WB_INST(float)
WB_INST(double)
#undef WB_INST

} // namespace wb
