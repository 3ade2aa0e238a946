// api.cu -- C ABI entry points and the level drivers of libwavelets_b200.
//
// The drivers restate the level / dimension ORDER of the reference drivers
//   _dwt! filter 1-D/2-D/3-D   src/Transforms/transforms_filter.jl:20-62, 123-188, 202-294
//   _dwt! lifting 1-D/2-D/3-D  src/Transforms/transforms_lifting.jl:30-76, 128-194, 200-278
//   _wpt! filter / lifting     src/Transforms/transforms_filter.jl:309-359, transforms_lifting.jl:283-319
// on top of one-level GPU passes.  Every pass is out of place between distinct buffers (x, y and the
// workspace), so no pass ever reads what another CTA of the same launch writes.
#include "common.cuh"
#include "fused.cuh"

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <algorithm>
#include <vector>

namespace wb {

// ---------------------------------------------------------------------------------------------------
// diagnostics
// ---------------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
static thread_local int64_t g_launches = 0;

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
// ---- per-kernel device timing (opt-in) ----
struct ProfRec { const char *name; cudaEvent_t e0, e1; };
static thread_local bool g_prof_on = false;
static thread_local std::vector<ProfRec> g_prof;

LaunchScope::LaunchScope(const char *n, cudaStream_t s) : name(n), st(s) {
    ++g_launches;
    if (g_prof_on) {
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        cudaEventRecord(e0, st);
    }
}
LaunchScope::~LaunchScope() {
    if (e0) {
        cudaEventRecord(e1, st);
        g_prof.push_back({name, e0, e1});
    }
}
bool check_launch(const char *what) {
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) {
        set_error("CUDA launch of %s failed: %s", what, cudaGetErrorString(err));
        return false;
    }
    return true;
}
static bool cuda_ok(cudaError_t err, const char *what) {
    if (err != cudaSuccess) {
        set_error("%s failed: %s", what, cudaGetErrorString(err));
        return false;
    }
    return true;
}

// ---------------------------------------------------------------------------------------------------
// Util mirrors (src/Util/non_dyadic.jl:14-22, util_main.jl:21-27, 301-313)
// ---------------------------------------------------------------------------------------------------
static inline bool suffpow2(int64_t n, int L) { return L < 62 && (n % ((int64_t)1 << L)) == 0; }
static int maxtransformlevels(int64_t n) {
    if (n <= 1) return 0;
    int tl = 0;
    while (suffpow2(n, tl)) ++tl;
    return tl - 1;
}
static bool isvalidtree(int64_t n, const uint8_t *b, int64_t nb) {
    const int ns = maxtransformlevels(n);
    if (nb != (((int64_t)1 << ns) - 1) || nb == 0 || b == nullptr) return false;
    for (int64_t i = 1; i <= (((int64_t)1 << (ns - 1)) - 1); ++i)
        if (!b[i - 1] && (b[2 * i - 1] || b[2 * i])) return false;
    return true;
}

// ---------------------------------------------------------------------------------------------------
// workspace: caller-provided or taken from the stream-ordered pool
// ---------------------------------------------------------------------------------------------------
static inline size_t align_up(size_t b) { return (b + 255) & ~(size_t)255; }

// Scratch comes from a LIBRARY-PRIVATE stream-ordered pool (one per device, created on first use), not from the device's
// default pool: the default pool hands freed memory back to the OS at the next synchronisation, so a transform called
// in a loop would re-map its scratch (milliseconds per GB) on every call; keeping it needs a raised release threshold,
// and raising it on the DEFAULT pool would change the behaviour of every other cudaMallocAsync user in the process.
// The private pool keeps what it has been given until wb200_trim_pool() (or process exit) releases it.
static std::mutex g_pool_mu;
static cudaMemPool_t g_pools[64] = {nullptr};
static cudaMemPool_t scratch_pool() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) { (void)cudaGetLastError(); return nullptr; }
    std::lock_guard<std::mutex> lk(g_pool_mu);
    if (g_pools[dev]) return g_pools[dev];
    cudaMemPoolProps props;
    memset(&props, 0, sizeof(props));
    props.allocType = cudaMemAllocationTypePinned;
    props.handleTypes = cudaMemHandleTypeNone;
    props.location.type = cudaMemLocationTypeDevice;
    props.location.id = dev;
    cudaMemPool_t pool = nullptr;
    if (cudaMemPoolCreate(&pool, &props) != cudaSuccess) { (void)cudaGetLastError(); return nullptr; }
    uint64_t thr = ~0ull;
    if (cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr) != cudaSuccess) (void)cudaGetLastError();
    g_pools[dev] = pool;
    return pool;
}
cudaError_t scratch_alloc(void **p, size_t bytes, cudaStream_t st) {
    cudaMemPool_t pool = scratch_pool();
    if (pool) return cudaMallocFromPoolAsync(p, bytes, pool, st);
    return cudaMallocAsync(p, bytes, st);          // pool creation refused: the default pool (its own threshold) still works
}
int64_t trim_pool(int64_t keep_bytes) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) { (void)cudaGetLastError(); return -1; }
    std::lock_guard<std::mutex> lk(g_pool_mu);
    if (!g_pools[dev]) return 0;
    if (cudaMemPoolTrimTo(g_pools[dev], keep_bytes > 0 ? (size_t)keep_bytes : 0) != cudaSuccess) { (void)cudaGetLastError(); return -1; }
    uint64_t reserved = 0;
    if (cudaMemPoolGetAttribute(g_pools[dev], cudaMemPoolAttrReservedMemCurrent, &reserved) != cudaSuccess) { (void)cudaGetLastError(); return -1; }
    return (int64_t)reserved;
}

struct Workspace {
    char *base = nullptr;
    size_t size = 0, used = 0;
    bool owned = false;
    cudaStream_t st = nullptr;
    int32_t init(void *user, size_t user_bytes, size_t need, cudaStream_t stream) {
        st = stream;
        if (need == 0) return WB200_OK;
        if (user != nullptr) {
            if (user_bytes < need) {
                set_error("workspace too small: %zu bytes given, %zu needed", user_bytes, need);
                return WB200_EWORKSPACE;
            }
            base = (char *)user; size = user_bytes;
            return WB200_OK;
        }
        if (!cuda_ok(scratch_alloc((void **)&base, need, stream), "cudaMallocAsync(workspace)")) return WB200_ECUDA;
        size = need; owned = true;
        return WB200_OK;
    }
    void *take(size_t bytes) {
        bytes = align_up(bytes);
        if (used + bytes > size) return nullptr;
        void *p = base + used;
        used += bytes;
        return p;
    }
    ~Workspace() {
        if (owned && base) cudaFreeAsync(base, st);
    }
};

// ---------------------------------------------------------------------------------------------------
// coefficient preparation
// ---------------------------------------------------------------------------------------------------
// WT.makereverseqmfpair(f, fw, T): the Float64 qmf is rounded to T first (src/WT/wt_main.jl:172-183);
// g = mirror(h) = h .* (-1)^(0:len-1) (src/Util/util_main.jl:30).
template <typename T> static void make_filter(FilterCoefs<T> &fc, const double *qmf, int flen) {
    fc.F = flen;
    for (int m = 0; m < MAXF; ++m) { fc.h[m] = 0; fc.g[m] = 0; }
    for (int m = 0; m < flen; ++m) {
        const T h = (T)qmf[m];
        fc.h[m] = h;
        fc.g[m] = (m % 2 == 0) ? h : -h;
    }
}
// makescheme(T, scheme, fw), src/Transforms/transforms_lifting.jl:13-25
template <typename T>
static void make_scheme(LiftScheme<T> &sc, const wb200_lift_step *steps, int nsteps, double norm1, double norm2, bool fw) {
    memset(&sc, 0, sizeof(sc));
    sc.nsteps = nsteps;
    sc.halo_l = sc.halo_r = 0;
    for (int i = 0; i < nsteps; ++i) {
        const int j = fw ? i : nsteps - 1 - i;
        sc.is_predict[i] = steps[j].is_predict ? 1 : 0;
        sc.shift[i] = steps[j].shift;
        sc.nc[i] = steps[j].nc;
        for (int k = 0; k < steps[j].nc; ++k) sc.coef[i][k] = (T)(steps[j].coef[k] * (fw ? -1.0 : 1.0));
        const int left = steps[j].shift > 0 ? steps[j].shift : 0;
        const int right = (steps[j].nc - 1 - steps[j].shift) > 0 ? (steps[j].nc - 1 - steps[j].shift) : 0;
        sc.halo_l += left;
        sc.halo_r += right;
    }
    sc.norm1 = (T)(fw ? norm1 : 1.0 / norm1);
    sc.norm2 = (T)(fw ? norm2 : 1.0 / norm2);
}

// Build the line set for a pass along array axis `ax` (1..3) over the corner `cor` (extent per dim).
// Outer coordinates are ordered by memory stride with extent-1 axes dropped (so that slot 0 is the
// contiguous coordinate when the lines are strided); the batch always sits in slot 3.
// thr_half: for each slot, the "low half" threshold used to select LL... lines on the inverse path.
template <typename T>
static void make_lines(const ArrayGeom &g, const int64_t cor[3], int ax, T *base, View<T> &v, Extent &e,
                       int64_t thr_half[4]) {
    e.len = cor[ax - 1];
    v.p = base;
    v.ls = g.stride(ax);
    int slot = 0;
    for (int q = 0; q < 4; ++q) { e.n[q] = 1; v.s[q] = 0; thr_half[q] = 1; }
    if (g.C > 1) { e.n[slot] = g.C; v.s[slot] = 1; thr_half[slot] = g.C; ++slot; }
    for (int a = 1; a <= 3; ++a) {
        if (a == ax) continue;
        if (cor[a - 1] == 1) continue;
        e.n[slot] = cor[a - 1];
        v.s[slot] = g.stride(a);
        thr_half[slot] = (a <= g.ndim) ? cor[a - 1] / 2 : cor[a - 1];
        ++slot;
    }
    // with C == 2 and three other axes present slot can reach 3 only if ndim == 3 and ax excluded one: max 1+2 = 3
    e.n[3] = g.batch;
    v.s[3] = g.slice();
    thr_half[3] = g.batch;
}

template <typename T> static View<const T> cview(const View<T> &v) {
    View<const T> c;
    c.p = v.p; c.ls = v.ls;
    for (int q = 0; q < 4; ++q) c.s[q] = v.s[q];
    return c;
}
template <typename T> static View<T> offset_view(const View<T> &v, int64_t elems) {
    View<T> o = v;
    o.p = v.p + elems;
    return o;
}

// ---------------------------------------------------------------------------------------------------
// 1-D driver (filter or lifting): y = [a_L | d_L | ... | d_1]
//   forward : level l reads a_{l-1} (x, or the ping-pong scratch) and writes d_l straight into y and a_l into
//             the other scratch buffer (into y at the last level)            transforms_filter.jl:45-60
//   inverse : level l reads a_l (x at l = L, else scratch) and d_l from x, writes a_{l-1} (y at l = 1)
// ---------------------------------------------------------------------------------------------------
template <typename T>
static int32_t run_1d(const PassOp<T> &op, T *y, const T *x, const ArrayGeom &g, int L, bool fw, Workspace &ws) {
    const int64_t n = g.dim[0], C = g.C, B = g.batch;
    T *buf[2] = {nullptr, nullptr}; // buf[1]: a_1, a_3, ... (n/2 per line) ; buf[0]: a_2, a_4, ... (n/4 per line)
    if (L >= 2) {
        buf[1] = (T *)ws.take(sizeof(T) * (size_t)((n / 2) * C * B));
        if (L >= 3) buf[0] = (T *)ws.take(sizeof(T) * (size_t)((n / 4) * C * B));
        if (!buf[1] || (L >= 3 && !buf[0])) { set_error("internal: 1-D workspace plan mismatch"); return WB200_EWORKSPACE; }
    }
    auto lines = [&](T *p, int64_t len, int64_t bstride, View<T> &v, Extent &e) {
        e.len = len;
        e.n[0] = C; e.n[1] = 1; e.n[2] = 1; e.n[3] = B;
        v.p = p; v.ls = C;
        v.s[0] = 1; v.s[1] = 0; v.s[2] = 0; v.s[3] = bstride;
    };
    const int64_t thr[4] = {0, 0, 0, 0};
    if (fw) {
        for (int l = 1; l <= L; ++l) {
            const int64_t nin = n >> (l - 1), nout = n >> l;
            View<T> src, dlo, dhi; Extent e, e2;
            if (l == 1) lines(const_cast<T *>(x), nin, n * C, src, e);
            else        lines(buf[(l - 1) & 1], nin, nin * C, src, e);
            lines(y + nout * C, nout, n * C, dhi, e2);
            if (l == L) lines(y, nout, n * C, dlo, e2);
            else        lines(buf[l & 1], nout, nout * C, dlo, e2);
            if (!op.analysis(cview(src), dlo, dhi, e)) return WB200_ECUDA;
        }
    } else {
        for (int l = L; l >= 1; --l) {
            const int64_t nout = n >> (l - 1), nin = n >> l;
            View<T> slo, shi, dst; Extent e, e2;
            if (l == L) lines(const_cast<T *>(x), nin, n * C, slo, e2);
            else        lines(buf[l & 1], nin, nin * C, slo, e2);
            lines(const_cast<T *>(x) + nin * C, nin, n * C, shi, e2);
            if (l == 1) lines(y, nout, n * C, dst, e);
            else        lines(buf[(l - 1) & 1], nout, nout * C, dst, e);
            if (!op.synthesis(cview(slo), cview(shi), cview(slo), thr, false, dst, e)) return WB200_ECUDA;
        }
    }
    return WB200_OK;
}

// ---------------------------------------------------------------------------------------------------
// 2-D / 3-D driver (filter or lifting).  Per level, forward: dim 3 -> dim 2 -> dim 1; inverse: 1 -> 2 -> 3
// (transforms_filter.jl:165-183, 246-288; transforms_lifting.jl:160-188, 232-272), on the leading corner.
// Buffers rotate src -> W1 (-> W2) -> y so that every pass is out of place; W1/W2 share the array layout.
// ---------------------------------------------------------------------------------------------------
template <typename T>
static int32_t run_nd(const PassOp<T> &op, T *y, const T *x, const ArrayGeom &g, int L, bool fw, Workspace &ws,
                      int l_lo = 1, int l_hi = -1) {
    // Processes levels l_lo..l_hi (forward: ascending, inverse: descending; default: all L levels).  A forward call
    // with l_lo > 1 expects the level-(l_lo-1) approximation in y's leading corner; an inverse call with l_lo > 1
    // leaves the level-(l_lo-1) approximation there (the fused kernels take over from it).
    if (l_hi < 0) l_hi = L;
    const int nd = g.ndim;
    // scratch arrays are compact: they only ever hold the corner of the largest processed level
    ArrayGeom gw = g;
    for (int a = 0; a < nd; ++a) gw.dim[a] = g.dim[a] >> (l_lo - 1);
    T *W[2] = {nullptr, nullptr};
    for (int q = 0; q < nd - 1; ++q) {
        W[q] = (T *)ws.take(sizeof(T) * (size_t)gw.total());
        if (!W[q]) { set_error("internal: N-D workspace plan mismatch"); return WB200_EWORKSPACE; }
    }
    const int nlev = l_hi - l_lo + 1;
    // 3-D filter bank on a volume whose (dim 1, dim 2) faces are squares the tile kernels take: the dim-2 and dim-1
    // passes of a level run as ONE fused 2-D launch per volume over the corner's planes (fir2d_impl.cuh), so a level is
    // two passes over the corner instead of three
    const bool fuse3_base = nd == 3 && !op.lifting && !op.generic_only && g.C == 1 && g.dim[0] == g.dim[1] && g.batch <= 16 &&
                            fir2d_tile_edge<T>(op.fc.F) > 0 && fir2d_available<T>() && std::getenv("WB200_DISABLE_FIR2D") == nullptr &&
                            ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(W[0])) & 15) == 0;
    for (int it = 0; it < nlev; ++it) {
        const int l = fw ? l_lo + it : l_hi - it;
        int64_t cor[3];
        for (int a = 0; a < 3; ++a) cor[a] = (a < nd) ? (g.dim[a] >> (l - 1)) : 1;
        const int te = fuse3_base ? fir2d_tile_edge<T>(op.fc.F) : 0;
        const bool fuse3 = fuse3_base && cor[0] >= te && cor[0] % te == 0 && cor[2] >= 2 && cor[2] <= 65535;
        const int64_t plane = g.dim[0] * g.dim[1], wplane = gw.dim[0] * gw.dim[1];
        for (int p = 0; p < nd; ++p) {
            const int ax = fw ? nd - p : p + 1;
            if (fuse3 && fw && p == 1) {          // planes of W[0] -> quadrants of y's planes
                for (int64_t b = 0; b < g.batch; ++b) {
                    T *yb = y + b * g.slice();
                    const int32_t rc = fir2d_level<T>(op, true, W[0] + b * gw.slice(), gw.dim[0], wplane, nullptr, 0, 0,
                                                      yb, g.dim[0], plane, yb, g.dim[0], plane, (int)cor[0], cor[2], op.st);
                    if (rc != WB200_OK) return rc;
                }
                break;
            }
            if (fuse3 && !fw && p == 0) {         // quadrants of x's planes (LLL corner from y below level L) -> planes of W[0]
                const int64_t dh = cor[2] / 2;
                for (int64_t b = 0; b < g.batch; ++b) {
                    const T *xb = x + b * g.slice();
                    const T *yb = y + b * g.slice();
                    T *wb_ = W[0] + b * gw.slice();
                    int32_t rc;
                    if (l < L) {
                        rc = fir2d_level<T>(op, false, yb, g.dim[0], plane, xb, g.dim[0], plane, wb_, gw.dim[0], wplane, nullptr, 0, 0, (int)cor[0], dh, op.st);
                        if (rc != WB200_OK) return rc;
                        rc = fir2d_level<T>(op, false, xb + dh * plane, g.dim[0], plane, xb + dh * plane, g.dim[0], plane,
                                            wb_ + dh * wplane, gw.dim[0], wplane, nullptr, 0, 0, (int)cor[0], cor[2] - dh, op.st);
                    } else {
                        rc = fir2d_level<T>(op, false, xb, g.dim[0], plane, xb, g.dim[0], plane, wb_, gw.dim[0], wplane, nullptr, 0, 0, (int)cor[0], cor[2], op.st);
                    }
                    if (rc != WB200_OK) return rc;
                }
                p = 1;                            // the dim-2 pass is done too; the dim-3 pass below reads W[0]
                continue;
            }
            View<T> vs, vd; Extent e, e2; int64_t thr[4], thr2[4];
            // source of this pass
            if (p == 0) make_lines<T>(g, cor, ax, const_cast<T *>((fw && l > 1) ? y : x), vs, e, thr);
            else        make_lines<T>(gw, cor, ax, (fuse3 && !fw) ? W[0] : W[p - 1], vs, e, thr);
            // destination of this pass
            if (p == nd - 1) make_lines<T>(g, cor, ax, y, vd, e2, thr2);
            else             make_lines<T>(gw, cor, ax, W[p], vd, e2, thr2);
            const int64_t nh = cor[ax - 1] / 2;
            if (fw) {
                if (!op.analysis(cview(vs), vd, offset_view(vd, nh * vd.ls), e)) return WB200_ECUDA;
            } else {
                // first pass of an inverse level: details (and, at level L, the approximation) come from x; below
                // level L the LL.. corner was produced by the previous level into y
                const bool has_alt = (p == 0) && (l < L);
                View<T> valt; Extent e3; int64_t thr3[4];
                make_lines<T>(g, cor, ax, y, valt, e3, thr3);
                if (!op.synthesis(cview(vs), cview(offset_view(vs, nh * vs.ls)), cview(valt), thr, has_alt, vd, e)) return WB200_ECUDA;
            }
        }
    }
    return WB200_OK;
}

// ---------------------------------------------------------------------------------------------------
// WPT driver: one sweep per tree level; active nodes get a one-level transform, leaves are carried along.
// Sweeps ping-pong between y and a full-size scratch so that each is out of place.
// ---------------------------------------------------------------------------------------------------
template <typename T>
static int32_t run_wpt(const PassOp<T> &op, T *y, const T *x, int64_t n, int64_t C, int64_t B,
                       const uint8_t *tree, int64_t ntree, bool fw, Workspace &ws) {
    const size_t tot = (size_t)(n * C * B);
    if (!tree[0]) {
        if (y != x && !cuda_ok(cudaMemcpyAsync(y, x, tot * sizeof(T), cudaMemcpyDeviceToDevice, op.st), "cudaMemcpyAsync")) return WB200_ECUDA;
        return WB200_OK;
    }
    const int Lmax = maxtransformlevels(n);
    int deep = 0; // number of tree levels holding at least one active node
    for (int lv = 0; lv < Lmax; ++lv) {
        bool any = false;
        for (int64_t k = 0; k < ((int64_t)1 << lv); ++k) any = any || tree[(((int64_t)1 << lv) - 1) + k];
        if (any) deep = lv + 1;
    }
    T *W = (T *)ws.take(sizeof(T) * tot);
    uint8_t *dtree = (uint8_t *)ws.take((size_t)ntree);
    if (!W || !dtree) { set_error("internal: WPT workspace plan mismatch"); return WB200_EWORKSPACE; }
    if (!cuda_ok(cudaMemcpyAsync(dtree, tree, (size_t)ntree, cudaMemcpyHostToDevice, op.st), "cudaMemcpyAsync(tree)")) return WB200_ECUDA;

    // Sweeps: one per tree level, except that a run of FULL levels whose nodes fit shared memory is taken by one
    // subtree launch (filter path).  lv_sub = first level of that run (deep: none).
    int lv_sub = deep;
    if (!op.lifting && !op.generic_only && C == 1) {
        int lv = deep;
        while (lv > 0) {                       // extend the run upwards while the level is fully active
            const int64_t nodes = (int64_t)1 << (lv - 1);
            bool all = true;
            for (int64_t k = 0; k < nodes; ++k) all = all && tree[(nodes - 1) + k];
            if (!all || (n >> (lv - 1)) > wpt_subtree_max_samples((int)sizeof(T))) break;
            --lv;
        }
        if (deep - lv >= 2) lv_sub = lv;       // worth a dedicated launch only for two or more levels
    }
    struct Sweep { int lv, count; };
    std::vector<Sweep> sweeps;
    // levels above the subtree run: runs of FULL levels are fused K at a time (wptfused.cu, filter path), the rest go one
    // sweep per level
    const bool fuse_ok = !op.lifting && !op.generic_only && C == 1 && ((n * (int64_t)sizeof(T)) % 16) == 0 &&
                         ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(W)) & 15) == 0;
    const int kfuse = fuse_ok ? wpt_fused_max_levels(op.fc.F) : 0;
    auto level_full = [&](int lv) {
        const int64_t nodes = (int64_t)1 << lv;
        for (int64_t k = 0; k < nodes; ++k) if (!tree[(nodes - 1) + k]) return false;
        return true;
    };
    for (int lv = 0; lv < lv_sub;) {
        int run = 0;
        if (kfuse >= 2) while (lv + run < lv_sub && run < kfuse && level_full(lv + run)) ++run;
        while (run >= 2 && !wpt_fused_ok((int)sizeof(T), op.fc.F, n >> lv, run)) --run;   // the kernels' own acceptance test
        if (run >= 2) { sweeps.push_back({lv, run}); lv += run; }
        else { sweeps.push_back({lv, 1}); ++lv; }
    }
    if (lv_sub < deep) sweeps.push_back({lv_sub, deep - lv_sub});
    if (!fw) std::reverse(sweeps.begin(), sweeps.end());
    // sweep i writes D_i; the last sweep must land in y unless that would make sweep 0 in place (x == y)
    const int ns = (int)sweeps.size();
    auto dst_of = [&](int i) -> T * { return ((ns - 1 - i) % 2 == 0) ? y : W; };
    bool flip = (x == y) && (dst_of(0) == y);
    auto dst2 = [&](int i) -> T * { T *d = dst_of(i); return flip ? (d == y ? W : y) : d; };
    const int64_t thr[4] = {0, 0, 0, 0};
    for (int i = 0; i < ns; ++i) {
        const int lv = sweeps[i].lv;
        const int64_t nj = n >> lv, nodes = (int64_t)1 << lv;
        const T *S = (i == 0) ? x : dst2(i - 1);
        T *D = dst2(i);
        if (sweeps[i].count > 1 && lv >= lv_sub) {
            const int r = fast_wpt_subtree<T>(S, D, n, nj, sweeps[i].count, nodes, B, op.fc, op.strict, fw, op.st);
            if (r < 0) return WB200_ECUDA;
            if (r > 0) continue;
            set_error("internal: packet subtree kernel rejected a planned sweep");
            return WB200_ECUDA;
        }
        if (sweeps[i].count > 1) {
            const int r = fast_wpt_fused_levels<T>(S, D, n, nj, sweeps[i].count, nodes, B, op.fc, op.strict, fw, op.st);
            if (r < 0) return WB200_ECUDA;
            if (r > 0) continue;
            set_error("internal: fused packet-level kernel rejected a planned sweep");
            return WB200_ECUDA;
        }
        Extent e; e.len = nj; e.n[0] = C; e.n[1] = nodes; e.n[2] = 1; e.n[3] = B;
        View<T> vs, vd;
        vs.p = const_cast<T *>(S); vs.ls = C; vs.s[0] = 1; vs.s[1] = nj * C; vs.s[2] = 0; vs.s[3] = n * C;
        vd = vs; vd.p = D;
        bool all_active = true;
        for (int64_t k = 0; k < nodes; ++k) all_active = all_active && tree[(nodes - 1) + k];
        const uint8_t *act = all_active ? nullptr : dtree + (nodes - 1);   // full levels take the fast line kernels
        const int64_t half = (nj / 2) * C;
        if (fw) {
            if (!op.analysis(cview(vs), vd, offset_view(vd, half), e, act)) return WB200_ECUDA;
        } else {
            if (!op.synthesis(cview(vs), cview(offset_view(vs, half)), cview(vs), thr, false, vd, e, act)) return WB200_ECUDA;
        }
    }
    if (dst2(ns - 1) != y)
        if (!cuda_ok(cudaMemcpyAsync(y, W, tot * sizeof(T), cudaMemcpyDeviceToDevice, op.st), "cudaMemcpyAsync")) return WB200_ECUDA;
    return WB200_OK;
}

// ---------------------------------------------------------------------------------------------------
// argument checking shared by the entry points
// ---------------------------------------------------------------------------------------------------
struct Call {
    ArrayGeom g;
    int esize;      // bytes per real element
    bool is_f64;
};
static int32_t parse_geom(Call &c, int32_t ndim, const int64_t *dims, int64_t batch, int32_t dtype) {
    if (ndim < 1 || ndim > 3 || dims == nullptr) { set_error("ndim must be 1, 2 or 3"); return WB200_EDIMS; }
    if (dtype < WB200_F32 || dtype > WB200_C128) { set_error("unsupported dtype %d", dtype); return WB200_EDTYPE; }
    c.is_f64 = (dtype == WB200_F64 || dtype == WB200_C128);
    c.esize = c.is_f64 ? 8 : 4;
    c.g.C = (dtype == WB200_C64 || dtype == WB200_C128) ? 2 : 1;
    c.g.ndim = ndim;
    for (int a = 0; a < 3; ++a) c.g.dim[a] = (a < ndim) ? dims[a] : 1;
    c.g.batch = batch;
    for (int a = 0; a < ndim; ++a)
        if (dims[a] < 1) { set_error("dims[%d] = %lld", a, (long long)dims[a]); return WB200_EDIMS; }
    if (batch < 0) { set_error("batch = %lld", (long long)batch); return WB200_EDIMS; }
    return WB200_OK;
}
// workspace plan (bytes); must stay in lock-step with run_1d / run_nd / run_wpt and the fused paths
static size_t plan_dwt(const Call &c, int L, bool lifting, bool inplace, uint32_t flags) {
    (void)lifting; (void)flags;
    const ArrayGeom &g = c.g;
    if (L <= 0 || g.batch == 0) return 0;
    size_t need = 0;
    if (g.ndim == 1) {
        if (inplace) need += align_up((size_t)g.total() * c.esize);
        if (L >= 2) need += align_up((size_t)((g.dim[0] / 2) * g.C * g.batch) * c.esize);
        if (L >= 3) need += align_up((size_t)((g.dim[0] / 4) * g.C * g.batch) * c.esize);
    } else {
        need += (size_t)(g.ndim - 1) * align_up((size_t)g.total() * c.esize);
    }
    return need;
}
// scratch of the generic N-D driver when it starts at level l_lo (compact corner buffers)
static size_t plan_nd_from(const Call &c, int l_lo) {
    const ArrayGeom &g = c.g;
    size_t corner = (size_t)g.C * (size_t)g.batch;
    for (int a = 0; a < g.ndim; ++a) corner *= (size_t)(g.dim[a] >> (l_lo - 1));
    return (size_t)(g.ndim - 1) * align_up(corner * c.esize);
}
template <typename T>
static size_t plan_fused2d(const PassOp<T> &op, const Call &c, int L, bool fw, bool inplace, int &Lf, bool &tail) {
    Lf = fused2d_levels<T>(op, c.g, L, fw);
    tail = (L > Lf) && fused2d_tail_ok<T>(op, c.g, c.g.dim[0] >> Lf, L - Lf);
    if (Lf == 0 && !tail) return 0;
    size_t need = align_up(fused2d_scratch_bytes<T>(c.g, Lf > 0 ? Lf : 1)) + 512;
    if (inplace && Lf > 0) need += align_up((size_t)c.g.total() * c.esize);
    if (L > Lf && !tail) need += plan_nd_from(c, Lf + 1);
    return need;
}

template <typename T>
static int32_t dispatch_dwt(PassOp<T> &op, void *y, const void *x, const Call &c, int L, bool fw, bool lifting,
                            void *workspace, size_t ws_bytes, cudaStream_t st, uint32_t flags) {
    const ArrayGeom &g = c.g;
    const size_t bytes = (size_t)g.total() * sizeof(T);
    if (g.batch == 0 || bytes == 0) return WB200_OK;
    if (L == 0) { // identity: copyto!(y, x) (filter) / return y (lifting)
        if (y != x && !cuda_ok(cudaMemcpyAsync(y, x, bytes, cudaMemcpyDeviceToDevice, st), "cudaMemcpyAsync")) return WB200_ECUDA;
        return WB200_OK;
    }
    const bool inplace = (y == x);
    // fused sm_100a kernels first; they report WB200_OK when they took the call, or -1 when the shape is not theirs
    if (!(flags & WB200_FLAG_FORCE_GENERIC) && g.C == 1) {
        int32_t rc = fused_dwt<T>(op, (T *)y, (const T *)x, g, L, fw, workspace, ws_bytes, st, flags);
        if (rc >= 0) return rc;
        if (lifting && g.ndim == 1) {
            rc = fused_lift1d<T>(op, (T *)y, (const T *)x, g, L, fw, workspace, ws_bytes, st);
            if (rc >= 0) return rc;
        }
    }
    // 2-D lifting: fused level kernels for the large levels, then the pyramid-tail kernel (or, for shapes it does not
    // take, the generic passes) for the small remainder
    // (orthogonal filters: tensor-map tile kernels, which move 16-byte pieces -- x and y must keep that alignment)
    if (!(flags & WB200_FLAG_FORCE_GENERIC) && g.C == 1 && g.ndim == 2 &&
        (lifting || (((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0))) {
        int Lf = 0;
        bool tail = false;
        const size_t need = plan_fused2d<T>(op, c, L, fw, inplace, Lf, tail);
        if (Lf > 0 || tail) {
            Workspace ws;
            int32_t rc = ws.init(workspace, ws_bytes, need, st);
            if (rc != WB200_OK) return rc;
            const int64_t N = g.dim[0], nr = N >> Lf;          // nr: corner left to the remainder
            T *scratch = (T *)ws.take(fused2d_scratch_bytes<T>(g, Lf > 0 ? Lf : 1) + 256);
            const T *xin = (const T *)x;
            if (inplace && Lf > 0) { // the level kernels read x while other CTAs write y: stage a copy
                T *x0 = (T *)ws.take(bytes);
                if (!x0 || !scratch) { set_error("internal: fused 2-D workspace plan mismatch"); return WB200_EWORKSPACE; }
                if (!cuda_ok(cudaMemcpyAsync(x0, x, bytes, cudaMemcpyDeviceToDevice, st), "cudaMemcpyAsync")) return WB200_ECUDA;
                xin = x0;
            }
            // compact buffer that carries the level-Lf approximation between the tile levels and the remainder
            T *mid = (Lf > 0) ? (T *)((char *)scratch + (((Lf - 1) & 1) ? align_up((size_t)(N / 2) * (N / 2) * g.batch * sizeof(T)) : 0)) : nullptr;
            if (fw) {
                if (Lf > 0) {
                    rc = fused2d_run<T>(op, (T *)y, xin, nullptr, 0, 0, g, Lf, true, scratch, st, tail && L > Lf);
                    if (rc != WB200_OK) return rc;
                }
                if (L == Lf) return WB200_OK;
                if (tail) return fused2d_tail<T>(op, Lf > 0 ? mid : xin, Lf > 0 ? nr : N, Lf > 0 ? nr * nr : N * N,
                                                 (T *)y, N, N * N, (int)nr, L - Lf, g.batch, true, st);
                return run_nd<T>(op, (T *)y, (const T *)y, g, L, true, ws, Lf + 1, L);
            }
            if (L == Lf) return fused2d_run<T>(op, (T *)y, xin, xin, N, N * N, g, Lf, false, scratch, st);
            if (tail) {
                if (Lf == 0) return fused2d_tail<T>(op, xin, N, N * N, (T *)y, N, N * N, (int)N, L, g.batch, false, st);
                rc = fused2d_tail<T>(op, xin, N, N * N, mid, nr, nr * nr, (int)nr, L - Lf, g.batch, false, st);
                if (rc != WB200_OK) return rc;
                return fused2d_run<T>(op, (T *)y, xin, mid, nr, nr * nr, g, Lf, false, scratch, st);
            }
            rc = run_nd<T>(op, (T *)y, xin, g, L, false, ws, Lf + 1, L);
            if (rc != WB200_OK) return rc;
            if (Lf >= 2) // level Lf reads y's corner and writes scratch; y itself is only written by level 1, later
                return fused2d_run<T>(op, (T *)y, xin, (const T *)y, N, N * N, g, Lf, false, scratch, st);
            // Lf == 1: the single fused level would read y's corner while other CTAs overwrite y -> park the corner
            View<const T> vs; View<T> vd; Extent e;
            e.len = nr; e.n[0] = 1; e.n[1] = nr; e.n[2] = 1; e.n[3] = g.batch;
            vs.p = (const T *)y; vs.ls = 1; vs.s[0] = 0; vs.s[1] = N; vs.s[2] = 0; vs.s[3] = N * N;
            vd.p = scratch; vd.ls = 1; vd.s[0] = 0; vd.s[1] = nr; vd.s[2] = 0; vd.s[3] = nr * nr;
            if (!launch_copy_lines<T>(vs, vd, e, st)) return WB200_ECUDA;
            return fused2d_run<T>(op, (T *)y, xin, (const T *)scratch, nr, nr * nr, g, Lf, false, scratch, st);
        }
    }
    // 3-D orthogonal filter bank: one-pass marching level kernels (fir3d_impl.cuh) for the levels whose corner is whole
    // tiles, the generic / line passes for the small remainder
    if (!(flags & WB200_FLAG_FORCE_GENERIC) && g.C == 1 && g.ndim == 3 && !lifting &&
        ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0) {
        const int Lf = fir3d_levels<T>(op, g, L, fw);
        if (Lf > 0) {
            const int64_t N1 = g.dim[0], N2 = g.dim[1], N3 = g.dim[2];
            const int64_t r1 = N1 >> Lf, r2 = N2 >> Lf, r3 = N3 >> Lf;          // corner left to the remainder
            const bool park = !fw && L > Lf && Lf == 1;                        // the single fused level would read y's corner while y is written
            const size_t park_bytes = park ? align_up((size_t)(r1 * r2 * r3 * g.batch) * sizeof(T)) : 0;
            const size_t need = align_up(fir3d_scratch_bytes<T>(g, Lf)) + 512 + park_bytes + (L > Lf ? plan_nd_from(c, Lf + 1) : 0);
            Workspace ws;
            int32_t rc = ws.init(workspace, ws_bytes, need, st);
            if (rc != WB200_OK) return rc;
            T *scratch = (T *)ws.take(fir3d_scratch_bytes<T>(g, Lf) + 256);
            T *parked = park ? (T *)ws.take(park_bytes) : nullptr;
            if (!scratch || (park && !parked)) { set_error("internal: fused 3-D workspace plan mismatch"); return WB200_EWORKSPACE; }
            const T *xin = (const T *)x;
            if (fw) {
                rc = fir3d_run<T>(op, (T *)y, xin, nullptr, 0, 0, 0, g, Lf, true, scratch, st);
                if (rc != WB200_OK || L == Lf) return rc;
                return run_nd<T>(op, (T *)y, (const T *)y, g, L, true, ws, Lf + 1, L);
            }
            if (L == Lf) return fir3d_run<T>(op, (T *)y, xin, xin, N1, N1 * N2, g.slice(), g, Lf, false, scratch, st);
            rc = run_nd<T>(op, (T *)y, xin, g, L, false, ws, Lf + 1, L);        // leaves the level-Lf approximation in y's corner
            if (rc != WB200_OK) return rc;
            if (!park) return fir3d_run<T>(op, (T *)y, xin, (const T *)y, N1, N1 * N2, g.slice(), g, Lf, false, scratch, st);
            View<const T> vs; View<T> vd; Extent e;
            e.len = r1; e.n[0] = 1; e.n[1] = r2; e.n[2] = r3; e.n[3] = g.batch;
            vs.p = (const T *)y; vs.ls = 1; vs.s[0] = 0; vs.s[1] = N1; vs.s[2] = N1 * N2; vs.s[3] = g.slice();
            vd.p = parked; vd.ls = 1; vd.s[0] = 0; vd.s[1] = r1; vd.s[2] = r1 * r2; vd.s[3] = r1 * r2 * r3;
            if (!launch_copy_lines<T>(vs, vd, e, st)) return WB200_ECUDA;
            return fir3d_run<T>(op, (T *)y, xin, parked, r1, r1 * r2, r1 * r2 * r3, g, Lf, false, scratch, st);
        }
    }
    // 3-D lifting on a cube: register walk along dim 3 + the 2-D lifting level kernel on the planes (lift3d.cu), generic
    // passes for the levels whose corner is smaller than a 2-D tile
    if (!(flags & WB200_FLAG_FORCE_GENERIC) && g.C == 1 && g.ndim == 3 && lifting &&
        ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0) {
        const int Lf = lift3d_levels<T>(op, g, L, fw);
        if (Lf > 0) {
            const size_t wbytes = align_up((size_t)g.total() * sizeof(T));
            const size_t need = wbytes + (L > Lf ? plan_nd_from(c, Lf + 1) : 0);
            Workspace ws;
            int32_t rc = ws.init(workspace, ws_bytes, need, st);
            if (rc != WB200_OK) return rc;
            T *W = (T *)ws.take(wbytes);
            if (!W) { set_error("internal: 3-D lifting workspace plan mismatch"); return WB200_EWORKSPACE; }
            if (fw) {
                rc = lift3d_run<T>(op, (T *)y, (const T *)x, g, L, Lf, true, W, st);
                if (rc != WB200_OK || L == Lf) return rc;
                return run_nd<T>(op, (T *)y, (const T *)y, g, L, true, ws, Lf + 1, L);
            }
            if (L > Lf) {
                rc = run_nd<T>(op, (T *)y, (const T *)x, g, L, false, ws, Lf + 1, L);    // leaves the level-Lf approximation in y's corner
                if (rc != WB200_OK) return rc;
            }
            return lift3d_run<T>(op, (T *)y, (const T *)x, g, L, Lf, false, W, st);
        }
    }
    Workspace ws;
    int32_t rc = ws.init(workspace, ws_bytes, plan_dwt(c, L, lifting, inplace, flags | WB200_FLAG_FORCE_GENERIC), st);
    if (rc != WB200_OK) return rc;
    const T *xin = (const T *)x;
    if (g.ndim == 1 && inplace) { // 1-D in place: level 1 would read what it writes, so stage a copy of x
        T *x0 = (T *)ws.take(bytes);
        if (!x0) { set_error("internal: in-place workspace plan mismatch"); return WB200_EWORKSPACE; }
        if (!cuda_ok(cudaMemcpyAsync(x0, x, bytes, cudaMemcpyDeviceToDevice, st), "cudaMemcpyAsync")) return WB200_ECUDA;
        xin = x0;
    }
    if (g.ndim == 1) return run_1d<T>(op, (T *)y, xin, g, L, fw, ws);
    return run_nd<T>(op, (T *)y, xin, g, L, fw, ws);
}

template <typename T>
static int32_t idwt_thr_T(void *y, const void *x, int64_t n, int64_t batch, const double *qmf, int32_t flen, int32_t L,
                          const ThreshEpi &epi, cudaStream_t st, uint32_t flags) {
    PassOp<T> op;
    op.lifting = false; op.strict = (flags & WB200_FLAG_STRICT_FP) != 0; op.st = st; op.generic_only = false;
    op.epi = epi;
    make_filter<T>(op.fc, qmf, flen);
    ArrayGeom g;
    g.C = 1; g.ndim = 1; g.dim[0] = n; g.dim[1] = 1; g.dim[2] = 1; g.batch = batch;
    return fused_dwt<T>(op, (T *)y, (const T *)x, g, L, false, nullptr, 0, st, flags);     // -1: not a fused shape
}
int32_t idwt_filter_thresholded(void *y, const void *x, int64_t n, int64_t batch, const double *qmf, int32_t flen, int32_t L,
                                int32_t dtype, const ThreshEpi &epi, cudaStream_t st, uint32_t flags) {
    if ((flags & WB200_FLAG_FORCE_GENERIC) || L < 1 || n < 2 || batch < 1 || y == x || qmf == nullptr || !suffpow2(n, L)) return -1;
    if (flen < 2 || flen > WB200_MAX_FILTER_LEN || std::getenv("WB200_DISABLE_THRESH_EPILOGUE") != nullptr) return -1;
    if (dtype == WB200_F64) return idwt_thr_T<double>(y, x, n, batch, qmf, flen, L, epi, st, flags);
    if (dtype == WB200_F32) return idwt_thr_T<float>(y, x, n, batch, qmf, flen, L, epi, st, flags);
    return -1;
}

} // namespace wb

using namespace wb;

// ===================================================================================================
// C ABI
// ===================================================================================================
extern "C" int32_t wb200_version(void) { return WB200_VERSION; }
extern "C" const char *wb200_last_error_string(void) { return g_err; }
extern "C" int64_t wb200_launch_count(int32_t reset) {
    int64_t v = g_launches;
    if (reset) g_launches = 0;
    return v;
}
extern "C" int64_t wb200_trim_pool(int64_t keep_bytes) { return trim_pool(keep_bytes); }
extern "C" void wb200_profile_enable(int32_t on) { g_prof_on = on != 0; }
// Waits for the recorded launches, then writes one line per kernel name: "<name> <launches> <total_ms>\n".
// Returns the number of bytes written (0 when nothing was recorded).  Clears the record.
extern "C" int64_t wb200_profile_collect(char *buf, int64_t buflen) {
    struct Agg { const char *name; int64_t n; double ms; };
    std::vector<Agg> agg;
    for (auto &r : g_prof) {
        float ms = 0.f;
        cudaEventSynchronize(r.e1);
        cudaEventElapsedTime(&ms, r.e0, r.e1);
        cudaEventDestroy(r.e0);
        cudaEventDestroy(r.e1);
        bool found = false;
        for (auto &a : agg) if (strcmp(a.name, r.name) == 0) { a.n++; a.ms += ms; found = true; break; }
        if (!found) agg.push_back({r.name, 1, (double)ms});
    }
    g_prof.clear();
    int64_t off = 0;
    for (auto &a : agg) {
        int w = snprintf(buf ? buf + off : nullptr, buf && buflen > off ? (size_t)(buflen - off) : 0, "%s %lld %.6f\n",
                         a.name, (long long)a.n, a.ms);
        if (w < 0) break;
        off += w;
        if (buf && off >= buflen) { off = buflen; break; }
    }
    return off;
}
extern "C" const char *wb200_status_string(int32_t s) {
    switch (s) {
    case WB200_OK: return "ok";
    case WB200_EDIMS: return "in and out array size must match";
    case WB200_ELEVEL: return "L must be positive";
    case WB200_EPOW2: return "size must have a sufficient power of 2 factor";
    case WB200_EALIAS: return "in array is out array";
    case WB200_ENOTCUBE: return "array must be square/cube";
    case WB200_ETREE: return "invalid tree";
    case WB200_EDTYPE: return "unsupported element type";
    case WB200_EARG: return "invalid argument";
    case WB200_EWORKSPACE: return "workspace too small";
    case WB200_ECUDA: return "CUDA error";
    default: return "unknown status";
    }
}
extern "C" int32_t wb200_maxtransformlevels(int64_t n) { return maxtransformlevels(n); }
extern "C" int32_t wb200_isvalidtree(int64_t n, const uint8_t *tree, int64_t ntree) { return isvalidtree(n, tree, ntree) ? 1 : 0; }

static int32_t check_filter_args(const double *qmf, int32_t flen) {
    if (qmf == nullptr || flen < 2 || flen > WB200_MAX_FILTER_LEN) {
        set_error("filter length %d outside [2, %d] or qmf is NULL", flen, WB200_MAX_FILTER_LEN);
        return WB200_EARG;
    }
    return WB200_OK;
}
static int32_t check_steps(const wb200_lift_step *steps, int32_t nsteps, double norm1, double norm2) {
    if (steps == nullptr || nsteps < 0 || nsteps > WB200_MAX_LIFT_STEPS) {
        set_error("nsteps %d outside [0, %d] or steps is NULL", nsteps, WB200_MAX_LIFT_STEPS);
        return WB200_EARG;
    }
    for (int i = 0; i < nsteps; ++i)
        if (steps[i].nc < 1 || steps[i].nc > WB200_MAX_LIFT_COEF) {
            set_error("step %d has %d coefficients (max %d)", i, steps[i].nc, WB200_MAX_LIFT_COEF);
            return WB200_EARG;
        }
    if (norm1 == 0.0 || norm2 == 0.0) { set_error("zero normalisation"); return WB200_EARG; }
    return WB200_OK;
}

extern "C" int32_t wb200_dwt_filter(void *y, const void *x, int32_t ndim, const int64_t *dims, int64_t batch,
                         const double *qmf, int32_t flen, int32_t L, int32_t fw, int32_t dtype,
                         void *workspace, size_t workspace_bytes, void *stream, uint32_t flags) {
    g_err[0] = 0;
    Call c;
    int32_t rc = parse_geom(c, ndim, dims, batch, dtype);
    if (rc != WB200_OK) return rc;
    if ((rc = check_filter_args(qmf, flen)) != WB200_OK) return rc;
    // order of checks = transforms_filter.jl:25-34
    if (L < 0) { set_error("L must be positive"); return WB200_ELEVEL; }
    for (int a = 0; a < ndim; ++a)
        if (!suffpow2(dims[a], L)) { set_error("size must have a sufficient power of 2 factor"); return WB200_EPOW2; }
    if (y == nullptr || x == nullptr) { set_error("null array pointer"); return WB200_EARG; }
    if (y == x) { set_error("in array is out array"); return WB200_EALIAS; }
    cudaStream_t st = (cudaStream_t)stream;
    if (c.is_f64) {
        PassOp<double> op; op.lifting = false; op.strict = (flags & WB200_FLAG_STRICT_FP) != 0; op.st = st; op.generic_only = (flags & WB200_FLAG_FORCE_GENERIC) != 0;
        make_filter<double>(op.fc, qmf, flen);
        return dispatch_dwt<double>(op, y, x, c, L, fw != 0, false, workspace, workspace_bytes, st, flags);
    } else {
        PassOp<float> op; op.lifting = false; op.strict = (flags & WB200_FLAG_STRICT_FP) != 0; op.st = st; op.generic_only = (flags & WB200_FLAG_FORCE_GENERIC) != 0;
        make_filter<float>(op.fc, qmf, flen);
        return dispatch_dwt<float>(op, y, x, c, L, fw != 0, false, workspace, workspace_bytes, st, flags);
    }
}

extern "C" int32_t wb200_dwt_lifting(void *y, const void *x, int32_t ndim, const int64_t *dims, int64_t batch,
                          const wb200_lift_step *steps, int32_t nsteps, double norm1, double norm2,
                          int32_t L, int32_t fw, int32_t dtype,
                          void *workspace, size_t workspace_bytes, void *stream, uint32_t flags) {
    g_err[0] = 0;
    Call c;
    int32_t rc = parse_geom(c, ndim, dims, batch, dtype);
    if (rc != WB200_OK) return rc;
    if ((rc = check_steps(steps, nsteps, norm1, norm2)) != WB200_OK) return rc;
    // order of checks = transforms_lifting.jl:131-136
    for (int a = 1; a < ndim; ++a)
        if (dims[a] != dims[0]) { set_error("array must be square/cube"); return WB200_ENOTCUBE; }
    if (L < 0) { set_error("L must be positive"); return WB200_ELEVEL; }
    if (!suffpow2(dims[0], L)) { set_error("size must have a sufficient power of 2 factor"); return WB200_EPOW2; }
    if (y == nullptr || x == nullptr) { set_error("null array pointer"); return WB200_EARG; }
    cudaStream_t st = (cudaStream_t)stream;
    if (c.is_f64) {
        PassOp<double> op; op.lifting = true; op.strict = (flags & WB200_FLAG_STRICT_FP) != 0; op.st = st; op.generic_only = (flags & WB200_FLAG_FORCE_GENERIC) != 0;
        make_scheme<double>(op.sc, steps, nsteps, norm1, norm2, fw != 0);
        return dispatch_dwt<double>(op, y, x, c, L, fw != 0, true, workspace, workspace_bytes, st, flags);
    } else {
        PassOp<float> op; op.lifting = true; op.strict = (flags & WB200_FLAG_STRICT_FP) != 0; op.st = st; op.generic_only = (flags & WB200_FLAG_FORCE_GENERIC) != 0;
        make_scheme<float>(op.sc, steps, nsteps, norm1, norm2, fw != 0);
        return dispatch_dwt<float>(op, y, x, c, L, fw != 0, true, workspace, workspace_bytes, st, flags);
    }
}

static size_t plan_wpt(int64_t n, int64_t C, int64_t batch, int64_t ntree, int esize) {
    return align_up((size_t)(n * C * batch) * esize) + align_up((size_t)ntree);
}

template <typename T>
static int32_t wpt_common(PassOp<T> &op, void *y, const void *x, int64_t n, int64_t C, int64_t batch,
                          const uint8_t *tree, int64_t ntree, bool fw, void *workspace, size_t ws_bytes, cudaStream_t st) {
    if (batch == 0 || n == 0) return WB200_OK;
    Workspace ws;
    int32_t rc = ws.init(workspace, ws_bytes, tree[0] ? plan_wpt(n, C, batch, ntree, sizeof(T)) : 0, st);
    if (rc != WB200_OK) return rc;
    return run_wpt<T>(op, (T *)y, (const T *)x, n, C, batch, tree, ntree, fw, ws);
}

extern "C" int32_t wb200_wpt_filter(void *y, const void *x, int64_t n, int64_t batch,
                         const double *qmf, int32_t flen, const uint8_t *tree, int64_t ntree,
                         int32_t fw, int32_t dtype,
                         void *workspace, size_t workspace_bytes, void *stream, uint32_t flags) {
    g_err[0] = 0;
    Call c;
    int64_t dims[1] = {n};
    int32_t rc = parse_geom(c, 1, dims, batch, dtype);
    if (rc != WB200_OK) return rc;
    if ((rc = check_filter_args(qmf, flen)) != WB200_OK) return rc;
    if (y == nullptr || x == nullptr) { set_error("null array pointer"); return WB200_EARG; }
    if (y == x) { set_error("in array is out array"); return WB200_EALIAS; }      // transforms_filter.jl:313
    if (!isvalidtree(n, tree, ntree)) { set_error("invalid tree"); return WB200_ETREE; }
    cudaStream_t st = (cudaStream_t)stream;
    if (c.is_f64) {
        PassOp<double> op; op.lifting = false; op.strict = (flags & WB200_FLAG_STRICT_FP) != 0; op.st = st; op.generic_only = (flags & WB200_FLAG_FORCE_GENERIC) != 0;
        make_filter<double>(op.fc, qmf, flen);
        return wpt_common<double>(op, y, x, n, c.g.C, batch, tree, ntree, fw != 0, workspace, workspace_bytes, st);
    } else {
        PassOp<float> op; op.lifting = false; op.strict = (flags & WB200_FLAG_STRICT_FP) != 0; op.st = st; op.generic_only = (flags & WB200_FLAG_FORCE_GENERIC) != 0;
        make_filter<float>(op.fc, qmf, flen);
        return wpt_common<float>(op, y, x, n, c.g.C, batch, tree, ntree, fw != 0, workspace, workspace_bytes, st);
    }
}

extern "C" int32_t wb200_wpt_lifting(void *y, const void *x, int64_t n, int64_t batch,
                          const wb200_lift_step *steps, int32_t nsteps, double norm1, double norm2,
                          const uint8_t *tree, int64_t ntree, int32_t fw, int32_t dtype,
                          void *workspace, size_t workspace_bytes, void *stream, uint32_t flags) {
    g_err[0] = 0;
    Call c;
    int64_t dims[1] = {n};
    int32_t rc = parse_geom(c, 1, dims, batch, dtype);
    if (rc != WB200_OK) return rc;
    if ((rc = check_steps(steps, nsteps, norm1, norm2)) != WB200_OK) return rc;
    if (y == nullptr || x == nullptr) { set_error("null array pointer"); return WB200_EARG; }
    if (!isvalidtree(n, tree, ntree)) { set_error("invalid tree"); return WB200_ETREE; }
    cudaStream_t st = (cudaStream_t)stream;
    if (c.is_f64) {
        PassOp<double> op; op.lifting = true; op.strict = (flags & WB200_FLAG_STRICT_FP) != 0; op.st = st; op.generic_only = (flags & WB200_FLAG_FORCE_GENERIC) != 0;
        make_scheme<double>(op.sc, steps, nsteps, norm1, norm2, fw != 0);
        return wpt_common<double>(op, y, x, n, c.g.C, batch, tree, ntree, fw != 0, workspace, workspace_bytes, st);
    } else {
        PassOp<float> op; op.lifting = true; op.strict = (flags & WB200_FLAG_STRICT_FP) != 0; op.st = st; op.generic_only = (flags & WB200_FLAG_FORCE_GENERIC) != 0;
        make_scheme<float>(op.sc, steps, nsteps, norm1, norm2, fw != 0);
        return wpt_common<float>(op, y, x, n, c.g.C, batch, tree, ntree, fw != 0, workspace, workspace_bytes, st);
    }
}

extern "C" size_t wb200_workspace_bytes(int32_t kind, int32_t ndim, const int64_t *dims, int64_t batch,
                             int32_t L, int32_t dtype, uint32_t flags) {
    Call c;
    if (parse_geom(c, ndim, dims, batch, dtype) != WB200_OK) return 0;
    switch (kind) {
    case 0: case 1: case 2: {
        size_t generic = plan_dwt(c, L, kind != 0, kind == 2, flags);
        size_t fused = fused_workspace_bytes(c.g, c.esize, L, kind != 0, kind == 2, flags);
        // (the fused 2-D lifting path needs at most scratch (n^2/4 + n^2/16) + an in-place copy + the generic
        //  remainder, which never exceeds the generic plan's full-size buffer plus the copy)
        if (kind == 2 && c.g.ndim == 2) generic += align_up((size_t)c.g.total() * c.esize) + 1024;
        return generic > fused ? generic : fused;
    }
    case 3: case 4: {
        const int Lmax = maxtransformlevels(dims[0]);
        return plan_wpt(dims[0], c.g.C, batch, (((int64_t)1) << Lmax) - 1, c.esize);
    }
    default: return 0;
    }
}

// ---------------------------------------------------------------------------------------------------
// host-buffer entry points: chunked H2D -> transform -> D2H pipeline on three streams
// ---------------------------------------------------------------------------------------------------
namespace {
struct HostPipe {
    int device = -1;
    size_t buf_bytes = 0, ws_bytes = 0;
    void *dx[3] = {nullptr, nullptr, nullptr}, *dy[3] = {nullptr, nullptr, nullptr}, *dw[3] = {nullptr, nullptr, nullptr};
    cudaStream_t st[3] = {nullptr, nullptr, nullptr};
    void release() {
        for (int i = 0; i < 3; ++i) {
            if (dx[i]) cudaFree(dx[i]);
            if (dy[i]) cudaFree(dy[i]);
            if (dw[i]) cudaFree(dw[i]);
            dx[i] = dy[i] = dw[i] = nullptr;
        }
        buf_bytes = ws_bytes = 0;
    }
    bool ensure(int dev, size_t bb, size_t wb) {
        if (dev != device) {
            if (device >= 0) { cudaSetDevice(device); release(); for (auto &s : st) if (s) { cudaStreamDestroy(s); s = nullptr; } }
            device = dev;
        }
        if (!cuda_ok(cudaSetDevice(dev), "cudaSetDevice")) return false;
        for (auto &s : st)
            if (!s && !cuda_ok(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking), "cudaStreamCreate")) return false;
        if (bb > buf_bytes || wb > ws_bytes) {
            release();
            for (int i = 0; i < 3; ++i) {
                if (!cuda_ok(cudaMalloc(&dx[i], bb), "cudaMalloc(chunk x)")) return false;
                if (!cuda_ok(cudaMalloc(&dy[i], bb), "cudaMalloc(chunk y)")) return false;
                if (wb && !cuda_ok(cudaMalloc(&dw[i], wb), "cudaMalloc(chunk workspace)")) return false;
            }
            buf_bytes = bb; ws_bytes = wb;
        }
        return true;
    }
};
HostPipe g_pipe;
std::mutex g_pipe_mu;

template <typename F>
int32_t run_host(void *y_host, const void *x_host, int32_t ndim, const int64_t *dims, int64_t batch, int32_t dtype,
                 int32_t L, int32_t kind, int32_t device, uint32_t flags, F &&call) {
    g_err[0] = 0;
    Call c;
    int32_t rc = parse_geom(c, ndim, dims, batch, dtype);
    if (rc != WB200_OK) return rc;
    if (y_host == nullptr || x_host == nullptr) { set_error("null host pointer"); return WB200_EARG; }
    if (batch == 0) return WB200_OK;
    const size_t slice_bytes = (size_t)c.g.slice() * c.esize;
    // chunk of the batch: ~64 MiB per stage (WB200_HOST_CHUNK_MB).  The first chunk's H2D and the last chunk's D2H cannot
    // overlap anything, so a call pays about two chunk copies of pipeline fill / drain: 256 MiB chunks cost ~10 ms of a
    // 48 ms call at the 55 GB/s a PCIe 5 x16 link delivers, 64 MiB chunks ~2.5 ms
    const char *cm = std::getenv("WB200_HOST_CHUNK_MB");
    size_t chunk_mb = (cm && *cm) ? (size_t)std::atoll(cm) : 64;
    if (chunk_mb < 1) chunk_mb = 1;
    if (chunk_mb > 2048) chunk_mb = 2048;
    int64_t cb = (int64_t)((chunk_mb << 20) / (slice_bytes ? slice_bytes : 1));
    if (cb < 1) cb = 1;
    if (cb > batch) cb = batch;
    if (batch >= 3 && cb > (batch + 2) / 3) cb = (batch + 2) / 3; // at least three chunks when possible
    const size_t ws_need = wb200_workspace_bytes(kind, ndim, dims, cb, L, dtype, flags);
    std::lock_guard<std::mutex> lock(g_pipe_mu);
    if (!g_pipe.ensure(device, slice_bytes * (size_t)cb, ws_need)) return WB200_ECUDA;
    int32_t result = WB200_OK;
    int slot = 0;
    for (int64_t b0 = 0; b0 < batch; b0 += cb, slot = (slot + 1) % 3) {
        const int64_t nb = (batch - b0 < cb) ? batch - b0 : cb;
        cudaStream_t st = g_pipe.st[slot];
        const char *xs = (const char *)x_host + (size_t)b0 * slice_bytes;
        char *ys = (char *)y_host + (size_t)b0 * slice_bytes;
        if (!cuda_ok(cudaMemcpyAsync(g_pipe.dx[slot], xs, slice_bytes * nb, cudaMemcpyHostToDevice, st), "H2D copy")) { result = WB200_ECUDA; break; }
        rc = call(g_pipe.dy[slot], g_pipe.dx[slot], nb, g_pipe.dw[slot], g_pipe.ws_bytes, (void *)st);
        if (rc != WB200_OK) { result = rc; break; }
        if (!cuda_ok(cudaMemcpyAsync(ys, g_pipe.dy[slot], slice_bytes * nb, cudaMemcpyDeviceToHost, st), "D2H copy")) { result = WB200_ECUDA; break; }
    }
    for (auto &s : g_pipe.st)
        if (s && !cuda_ok(cudaStreamSynchronize(s), "cudaStreamSynchronize") && result == WB200_OK) result = WB200_ECUDA;
    return result;
}
} // namespace

extern "C" int32_t wb200_dwt_filter_host(void *y_host, const void *x_host, int32_t ndim, const int64_t *dims,
                              int64_t batch, const double *qmf, int32_t flen, int32_t L, int32_t fw,
                              int32_t dtype, int32_t device, uint32_t flags) {
    return run_host(y_host, x_host, ndim, dims, batch, dtype, L, 0, device, flags,
                    [&](void *dy, const void *dx, int64_t nb, void *w, size_t wb, void *st) {
                        return wb200_dwt_filter(dy, dx, ndim, dims, nb, qmf, flen, L, fw, dtype, w, wb, st, flags);
                    });
}
extern "C" int32_t wb200_dwt_lifting_host(void *y_host, const void *x_host, int32_t ndim, const int64_t *dims,
                               int64_t batch, const wb200_lift_step *steps, int32_t nsteps,
                               double norm1, double norm2, int32_t L, int32_t fw,
                               int32_t dtype, int32_t device, uint32_t flags) {
    return run_host(y_host, x_host, ndim, dims, batch, dtype, L, 1, device, flags,
                    [&](void *dy, const void *dx, int64_t nb, void *w, size_t wb, void *st) {
                        return wb200_dwt_lifting(dy, dx, ndim, dims, nb, steps, nsteps, norm1, norm2, L, fw, dtype, w, wb, st, flags);
                    });
}

