// lift3d.cu -- 3-D LIFTING levels (cube, fused shapes cdf97 / Haar / db2) in two passes per level instead of the reference's
// 3 x (split + one sweep per step + normalize) line sweeps (src/Transforms/transforms_lifting.jl:200-278):
//   dim 3   k_walk_lift_fwd / _inv : lines along dim 3 are strided by a whole plane, so threads run along dim 1 (every load
//                                    and store is a coalesced row), each thread holds RK pairs (+ the scheme's halo, periodic
//                                    wrap by index) of ITS line in registers, runs all predict / update steps there and
//                                    writes the two de-interleaved halves (forward) or the merged line (inverse);
//   dims 2, 1                      : the tensor-map TMA 2-D lifting level kernel of fused2d_tma.cuh on the planes of the
//                                    volume (batch = planes).
// Forward order dim 3 -> (dim 2, dim 1), inverse (dim 1, dim 2) -> dim 3, as upstream; in place like upstream (every pass
// is out of place between the array and a compact scratch volume).  Before this file a 512^3 cdf97 transform (L = 3) took
// 15.8 ms per dwt + idwt pair on the generic one-launch-per-level-and-dimension kernels.
#include "fused.cuh"
#include "tile2d_shapes.cuh"

#include <cstdlib>

namespace wb {
namespace l3 {

constexpr int RK = 16;          // output pairs per thread

// halo in pairs, symmetric and even (as lift1d.cu)
template <class S> struct HaloE {
    static constexpr int L_ = Halo<S>::left(), R_ = Halo<S>::right();
    static constexpr int value = ((L_ > R_ ? L_ : R_) + 1) & ~1;
};

// Lines along a strided dimension: element k of the line of (i, j, b) sits at base + i + j*sj + b*sb + k*ls.
struct Lines { int64_t ls, sj, sb; int ni, nj, nb; };

template <typename T, class S, bool STRICT>
__global__ void __launch_bounds__(128)
k_walk_lift_fwd(const T *__restrict__ src, Lines gs, T *__restrict__ lo, T *__restrict__ hi, Lines gd, int n, int nseg,
                const __grid_constant__ LiftCoefs<T> lc) {
    using fp = FP<STRICT>;
    constexpr int HM = HaloE<S>::value, NP = RK + 2 * HM;
    const int i = (int)(blockIdx.x * blockDim.x + threadIdx.x);
    if (i >= gs.ni) return;
    int64_t r = blockIdx.y;
    const int seg = (int)(r % nseg); r /= nseg;
    const int j = (int)(r % gs.nj);
    const int b = (int)(r / gs.nj);
    const int nh = n >> 1;
    const T *x = src + i + (int64_t)j * gs.sj + (int64_t)b * gs.sb;
    const int k0 = seg * RK;
    T s[NP], d[NP];
    {
        int p = k0 - HM;                       // first pair of the window (periodic)
        p %= nh; if (p < 0) p += nh;
#pragma unroll
        for (int pp = 0; pp < NP; ++pp) {
            s[pp] = x[(int64_t)(2 * p) * gs.ls];
            d[pp] = x[(int64_t)(2 * p + 1) * gs.ls];
            if (++p == nh) p = 0;
        }
    }
    int g0 = (k0 - HM) % nh; if (g0 < 0) g0 += nh;
    lift_regs<T, S, STRICT, NP>(s, d, lc, g0, nh, STRICT);
    T *plo = lo + i + (int64_t)j * gd.sj + (int64_t)b * gd.sb;
    T *phi = hi + i + (int64_t)j * gd.sj + (int64_t)b * gd.sb;
#pragma unroll
    for (int k = 0; k < RK; ++k) {
        if (k0 + k < nh) {
            plo[(int64_t)(k0 + k) * gd.ls] = fp::mul(s[HM + k], lc.n1);
            phi[(int64_t)(k0 + k) * gd.ls] = fp::mul(d[HM + k], lc.n2);
        }
    }
}

template <typename T, class S, bool STRICT>
__global__ void __launch_bounds__(128)
k_walk_lift_inv(const T *__restrict__ lo, const T *__restrict__ hi, Lines gs, T *__restrict__ dst, Lines gd, int n, int nseg,
                const __grid_constant__ LiftCoefs<T> lc) {
    using fp = FP<STRICT>;
    constexpr int HM = HaloE<S>::value, NP = RK + 2 * HM;
    const int i = (int)(blockIdx.x * blockDim.x + threadIdx.x);
    if (i >= gs.ni) return;
    int64_t r = blockIdx.y;
    const int seg = (int)(r % nseg); r /= nseg;
    const int j = (int)(r % gs.nj);
    const int b = (int)(r / gs.nj);
    const int nh = n >> 1;
    const T *pl = lo + i + (int64_t)j * gs.sj + (int64_t)b * gs.sb;
    const T *ph = hi + i + (int64_t)j * gs.sj + (int64_t)b * gs.sb;
    const int k0 = seg * RK;
    T s[NP], d[NP];
    {
        int p = (k0 - HM) % nh; if (p < 0) p += nh;
#pragma unroll
        for (int pp = 0; pp < NP; ++pp) {
            s[pp] = fp::mul(pl[(int64_t)p * gs.ls], lc.n1);     // normalize! (reciprocal norms) precedes the steps
            d[pp] = fp::mul(ph[(int64_t)p * gs.ls], lc.n2);
            if (++p == nh) p = 0;
        }
    }
    int g0 = (k0 - HM) % nh; if (g0 < 0) g0 += nh;
    lift_regs<T, S, STRICT, NP>(s, d, lc, g0, nh, STRICT);
    T *o = dst + i + (int64_t)j * gd.sj + (int64_t)b * gd.sb;
#pragma unroll
    for (int k = 0; k < RK; ++k) {
        if (k0 + k < nh) {
            o[(int64_t)(2 * (k0 + k)) * gd.ls] = s[HM + k];
            o[(int64_t)(2 * (k0 + k) + 1) * gd.ls] = d[HM + k];
        }
    }
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

template <typename T, class SF, class SI_, bool STRICT>
static bool walk_pass(bool fw, const T *a, const T *a2, const Lines &gs, T *o, T *o2, const Lines &gd, int n,
                      const LiftCoefs<T> &lc, cudaStream_t st) {
    const int nh = n / 2, nseg = (nh + RK - 1) / RK;
    const int64_t gy = (int64_t)nseg * gs.nj * gs.nb;
    if (gy > 65535LL * 32768LL || gy > 0x7fffffffLL) { set_error("lift3d: too many lines for one launch"); return false; }
    dim3 grid((unsigned)((gs.ni + 127) / 128), (unsigned)gy);
    if (gy > 65535) { set_error("lift3d: grid.y limit"); return false; }
    if (fw) {
        LaunchScope scope("walk_lift_fwd", st);
        k_walk_lift_fwd<T, SF, STRICT><<<grid, 128, 0, st>>>(a, gs, o, o2, gd, n, nseg, lc);
    } else {
        LaunchScope scope("walk_lift_inv", st);
        k_walk_lift_inv<T, SI_, STRICT><<<grid, 128, 0, st>>>(a, a2, gs, o, gd, n, nseg, lc);
    }
    return check_launch("walk_lift");
}

} // namespace l3

// number of leading levels of a cube the two-pass 3-D lifting route takes (its planes must be tiles of the 2-D level kernel)
template <typename T>
int lift3d_levels(const PassOp<T> &op, const ArrayGeom &g, int L, bool fw) {
    if (!op.lifting || op.generic_only || g.ndim != 3 || g.C != 1 || g.dim[0] != g.dim[1] || g.dim[0] != g.dim[2]) return 0;
    const char *e = std::getenv("WB200_DISABLE_LIFT3D");
    if (e && *e && std::atoi(e)) return 0;
    if (l3::shape_of<T>(op.sc, fw) == 0 || g.batch < 1 || g.batch > 64) return 0;
    int Lf = 0;
    int64_t c = g.dim[0];
    // the per-line grid packs (segment, j, volume) into gridDim.y <= 65535
    while (Lf < L && c >= 128 && c % 128 == 0 && c < 32768 && ((c / 2 + l3::RK - 1) / l3::RK) * c * g.batch <= 65535) { ++Lf; c >>= 1; }
    return Lf;
}

// levels 1..Lf (forward) / Lf..1 (inverse).  W: compact scratch of N^3 * batch elements.  Forward leaves the level-Lf
// approximation in y's corner; inverse expects it there (l < L) or in x (l == L).
template <typename T>
int32_t lift3d_run(const PassOp<T> &op, T *y, const T *x, const ArrayGeom &g, int L, int Lf, bool fw, T *W, cudaStream_t st) {
    const int64_t N = g.dim[0], B = g.batch, NN = N * N, NNN = NN * N;
    LiftCoefs<T> lc;
    l3::fill_lc<T>(lc, op.sc);
    const int id = l3::shape_of<T>(op.sc, fw);
    auto walk = [&](const T *a, const T *a2, const l3::Lines &gs, T *o, T *o2, const l3::Lines &gd, int n) -> bool {
#define WB_W(SF, SI_) return op.strict ? l3::walk_pass<T, SF, SI_, true>(fw, a, a2, gs, o, o2, gd, n, lc, st) \
                                       : l3::walk_pass<T, SF, SI_, false>(fw, a, a2, gs, o, o2, gd, n, lc, st)
        switch (id) {
        case 1: WB_W(ShapeCdf97F, ShapeCdf97I);
        case 2: WB_W(ShapeHaarF, ShapeHaarI);
        default: WB_W(ShapeDb2F, ShapeDb2I);
        }
#undef WB_W
    };
    for (int it = 0; it < Lf; ++it) {
        const int l = fw ? it + 1 : Lf - it;
        const int64_t c = N >> (l - 1), h = c / 2;
        l3::Lines ga{NN, N, NNN, (int)c, (int)c, (int)B};           // the corner inside the full array
        l3::Lines gw{c * c, c, c * c * c, (int)c, (int)c, (int)B};  // the compact scratch volume
        if (fw) {
            const T *src = (l == 1) ? x : y;
            if (!walk(src, nullptr, ga, W, W + h * c * c, gw, (int)c)) return WB200_ECUDA;
            for (int64_t b = 0; b < B; ++b) {
                T *yb = y + b * NNN;
                const int32_t rc = lift2d_level<T>(op, true, W + b * c * c * c, c, c * c, nullptr, 0, 0, yb, N, NN, yb, N, NN, (int)c, c, st);
                if (rc != WB200_OK) return rc;
            }
        } else {
            for (int64_t b = 0; b < B; ++b) {
                const T *xb = x + b * NNN;
                const T *llb = (l < L) ? (const T *)(y + b * NNN) : xb;      // LLL octant: the previous inverse level left it in y's corner
                T *wb_ = W + b * c * c * c;
                int32_t rc = lift2d_level<T>(op, false, llb, N, NN, xb, N, NN, wb_, c, c * c, nullptr, 0, 0, (int)c, h, st);
                if (rc != WB200_OK) return rc;
                rc = lift2d_level<T>(op, false, xb + h * NN, N, NN, xb + h * NN, N, NN, wb_ + h * c * c, c, c * c, nullptr, 0, 0, (int)c, c - h, st);
                if (rc != WB200_OK) return rc;
            }
            if (!walk(W, W + h * c * c, gw, y, nullptr, ga, (int)c)) return WB200_ECUDA;
        }
    }
    return WB200_OK;
}

#define WB_INST(T)                                                                                       \
    template int lift3d_levels<T>(const PassOp<T> &, const ArrayGeom &, int, bool);                      \
    template int32_t lift3d_run<T>(const PassOp<T> &, T *, const T *, const ArrayGeom &, int, int, bool, T *, cudaStream_t);
WB_INST(float)
WB_INST(double)
#undef WB_INST

} // namespace wb
