/* wavelets_oracle.c -- CPU ORACLE.  TEST INFRASTRUCTURE, NOT PRODUCT.
 *
 * A plain-C restatement of the forward/inverse DWT / WPT hot path of
 * JuliaDSP/Wavelets.jl @ a5ad928 (src/Transforms/transforms_filter.jl,
 * src/Transforms/transforms_lifting.jl, src/Util/util_main.jl,
 * src/Util/non_dyadic.jl, src/WT/wt_main.jl:172-183).  The reference is pure
 * Julia and cannot run in this image (no julia binary), so this file follows the
 * reference loop-for-loop -- including the transposed-direct-form shift register
 * of filtdown!/filtup! and the interior/boundary split of lift! -- so that the
 * floating-point operation order is the reference's.  Build with
 * -ffp-contract=off (Julia never fuses a*b+c on this path).
 *
 * Pinned by: tests/golden/wavelab_golden.json (the reference's own WaveLab /
 * PyWavelets known-answer vectors, test/transforms.jl:2-55) and the reference's
 * relational tests (lifting == filter for db1/db2, wpt == dwt tree, round trips).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference
 * arm may load this library.  The product (libwavelets_b200.so) never does.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_OK 0
#define ORC_EDIMS 1    /* DimensionMismatch("in and out array size must match") / bad ndim */
#define ORC_ELEVEL 2   /* ArgumentError("L must be positive") */
#define ORC_EPOW2 3    /* ArgumentError("size must have a sufficient power of 2 factor") */
#define ORC_EALIAS 4   /* ArgumentError("in array is out array") */
#define ORC_ENOTCUBE 5 /* ArgumentError("array must be square/cube") */
#define ORC_ETREE 6    /* ArgumentError("invalid tree") */

#define ORC_MAX_STEPS 16
#define ORC_MAX_COEF 8

/* one lifting step as stored in WT.SCHEMES (src/WT/wt_main.jl:195-209, 451-480) */
typedef struct {
    int32_t is_predict; /* 1: PredictStep (writes the first half), 0: UpdateStep */
    int32_t shift;
    int32_t nc;
    double coef[ORC_MAX_COEF];
} orc_step;

typedef struct { int64_t a, b; } range_t; /* inclusive a:b, empty when b < a */

/* Julia mod (result has the sign of the divisor) and mod1 */
static inline int64_t jmod(int64_t a, int64_t n) { int64_t r = a % n; return r < 0 ? r + n : r; }
static inline int64_t jmod1(int64_t a, int64_t n) { return jmod(a - 1, n) + 1; }
/* arithmetic shift right by one (Julia >> on negative Int) */
static inline int64_t asr1(int64_t a) { return (a >= 0) ? (a >> 1) : -((-a + 1) >> 1); }

/* src/Util/non_dyadic.jl:5-22, src/Util/util_main.jl:21-27 (exact for n % 2^l == 0) */
static inline int64_t detailn(int64_t n, int l) { return n >> l; }
static inline int64_t detailindex(int64_t n, int l, int64_t i) { return (n >> l) + i; }
static inline int suffpow2(int64_t n, int L) { return L < 62 && (n % ((int64_t)1 << L)) == 0; }
static int maxtransformlevels(int64_t n)
{
    if (n <= 1) return 0;
    int tl = 0;
    while (suffpow2(n, tl)) tl += 1;
    return tl - 1;
}
int orc_maxtransformlevels(int64_t n) { return maxtransformlevels(n); }

/* isvalidtree, src/Util/util_main.jl:301-313 (tree is 1-based heap order, one byte per node) */
static int isvalidtree(int64_t n, const uint8_t *b, int64_t nb)
{
    const int ns = maxtransformlevels(n);
    if (nb != (((int64_t)1 << ns) - 1)) return 0;
    if (nb == 0) return 0;
    for (int64_t i = 1; i <= (((int64_t)1 << (ns - 1)) - 1); ++i)
        if (!b[i - 1] && (b[(i << 1) - 1] || b[(i << 1)])) return 0;
    return 1;
}
int orc_isvalidtree(int64_t n, const uint8_t *b, int64_t nb) { return isvalidtree(n, b, nb); }

/* splitdownrangeper, transforms_filter.jl:436-456 */
static void splitdownrangeper(int64_t istart, int64_t ix, int64_t nx, int64_t shift,
                              range_t *r1, range_t *rin, range_t *r2)
{
    const int64_t ixsh = -1 + shift + ix;
    if (jmod(shift, nx) + ix == 1 + ixsh) {
        int64_t iend = nx - 1;
        while (jmod(iend - 1 + shift, nx) + ix != iend + ixsh) iend -= 1;
        *r1 = (range_t){0, -1}; *rin = (range_t){1, iend}; *r2 = (range_t){iend + 1, nx - 1 + istart};
    } else if (jmod(istart - 1 + shift, nx) + ix == istart + ixsh) {
        int64_t iend = nx - 1 + istart;
        while (jmod(iend - 1 + shift, nx) + ix != iend + ixsh) iend -= 1;
        *r1 = (range_t){1, istart - 1}; *rin = (range_t){istart, iend}; *r2 = (range_t){iend + 1, nx - 1 + istart};
    } else {
        *r1 = (range_t){0, -1}; *rin = (range_t){0, -1}; *r2 = (range_t){1, nx - 1 + istart};
    }
}
/* splituprangeper, transforms_filter.jl:544-564 */
static void splituprangeper(int64_t istart, int64_t ix, int64_t nx, int64_t nout, int64_t shift,
                            range_t *r1, range_t *rin, range_t *r2)
{
    const int64_t sh = asr1(shift);
    const int64_t ixsh = sh + ix;
    if (jmod(sh, nx) + ix == ixsh) {
        int64_t iend = nout - 1;
        while (jmod(((iend - 1) >> 1) + sh, nx) + ix != ((iend - 1) >> 1) + ixsh) iend -= 1;
        *r1 = (range_t){0, -1}; *rin = (range_t){1, iend}; *r2 = (range_t){iend + 1, nout - 1 + istart};
    } else if (jmod(((istart - 1) >> 1) + sh, nx) + ix == ((istart - 1) >> 1) + ixsh) {
        int64_t iend = nout - 1 + istart;
        while (jmod(((iend - 1) >> 1) + sh, nx) + ix != ((iend - 1) >> 1) + ixsh) iend -= 1;
        *r1 = (range_t){1, istart - 1}; *rin = (range_t){istart, iend}; *r2 = (range_t){iend + 1, nout - 1 + istart};
    } else {
        *r1 = (range_t){0, -1}; *rin = (range_t){0, -1}; *r2 = (range_t){1, nout - 1 + istart};
    }
}

/* irlimits + getliftranges, transforms_lifting.jl:383-434 */
static void getliftranges(int64_t half, int nc, int64_t shift, int is_predict,
                          range_t *lhsr, range_t *irange, range_t *rhsr, int64_t *rhsis)
{
    int64_t irmin = (shift + 1 > 1 - nc + shift) ? shift + 1 : 1 - nc + shift;
    int64_t irmax = (half + 1 + shift - nc < half + shift) ? half + 1 + shift - nc : half + shift;
    const int64_t off = is_predict ? 0 : half;
    *rhsis = is_predict ? (-shift + half) : (-shift - half);
    int empty;
    if (irmin > half || irmax < 1) {
        empty = 1;
    } else {
        if (irmin < 1) irmin = 1;
        if (irmax > half) irmax = half;
        *irange = (range_t){irmin + off, irmax + off};
        empty = (irmax < irmin);
    }
    if (empty) {
        *irange = (range_t){1, 0};
        *lhsr = (range_t){1 + off, half + off};
        *rhsr = (range_t){1 + off, 0 + off};
    } else {
        *lhsr = (range_t){1 + off, irmin - 1 + off};
        *rhsr = (range_t){irmax + 1 + off, half + off};
    }
}

#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)

#define T double
#define FN(name) CAT(name, _f64)
#include "oracle_impl.inc"
#undef T
#undef FN

#define T float
#define ORC_IS_F32 1
#define FN(name) CAT(name, _f32)
#include "oracle_impl.inc"
#undef T
#undef FN
#undef ORC_IS_F32
