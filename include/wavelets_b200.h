/* wavelets_b200.h -- C ABI of libwavelets_b200.so
 *
 * B200-native (sm_100a) forward/inverse discrete wavelet transform hot path: the drop-in
 * replacement for the per-level kernels behind JuliaDSP/Wavelets.jl's dispatch seam
 *
 *     Transforms._dwt!(y, x, filter::OrthoFilter, L, fw)      src/Transforms/transforms_filter.jl:13,113,192
 *     Transforms._dwt!(y, scheme::GLS, L, fw)                 src/Transforms/transforms_lifting.jl:30,128,200
 *     Transforms._wpt!(y, x, filter, tree::BitVector, fw)     src/Transforms/transforms_filter.jl:301
 *     Transforms._wpt!(y, scheme, tree, fw)                   src/Transforms/transforms_lifting.jl:283
 *
 * (the same four method families the reference's own KernelAbstractions extension overrides,
 * ext/WaveletsGPUExt/WaveletsGPUExt.jl:11).  A Julia shim adds these methods for CUDA device
 * arrays and `ccall`s the entry points below; see INTEGRATION.md and julia/WaveletsB200.jl.
 *
 * Conventions
 *   - plain C types only; `x`, `y`, `workspace` are DEVICE pointers owned by the caller; `stream` is a
 *     cudaStream_t passed as void* (NULL = legacy default stream).  Every entry point only enqueues
 *     work on `stream`; it never synchronises the device.
 *   - arrays are column-major (Julia layout): dims[0] is the contiguous dimension.  `batch` independent
 *     arrays are laid out back to back (slice stride = prod(dims)); batch = 1 reproduces the reference
 *     call, batch > 1 is the column-wise `dwtc`/`idwtc` the reference advertises but never implemented
 *     (src/Transforms/transforms_main.jl:179-181).
 *   - wavelet coefficients always arrive as Float64 HOST arrays and are rounded to `dtype` inside,
 *     mirroring WT.makereverseqmfpair(f, fw, T) (src/WT/wt_main.jl:172-183) and
 *     makescheme(T, scheme, fw) (src/Transforms/transforms_lifting.jl:13-25).  Reversal / mirroring of
 *     the qmf, the -1 factor and step reversal of lifting schemes and the reciprocal norms are derived
 *     inside the library exactly as those functions do.
 *   - every function returns a wb200_status; nothing throws or aborts across the ABI.
 */
#ifndef WAVELETS_B200_H
#define WAVELETS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WB200_VERSION 100 /* 0.1.0 */

/* status codes; the messages are the reference's exception texts
 * (transforms_filter.jl:25-34,129-138,311-319; transforms_lifting.jl:34-39,131-140,286-289) */
typedef enum {
    WB200_OK = 0,
    WB200_EDIMS = 1,    /* DimensionMismatch("in and out array size must match") / unsupported ndim */
    WB200_ELEVEL = 2,   /* ArgumentError("L must be positive") */
    WB200_EPOW2 = 3,    /* ArgumentError("size must have a sufficient power of 2 factor") */
    WB200_EALIAS = 4,   /* ArgumentError("in array is out array") */
    WB200_ENOTCUBE = 5, /* ArgumentError("array must be square/cube") */
    WB200_ETREE = 6,    /* ArgumentError("invalid tree") */
    WB200_EDTYPE = 7,   /* unsupported element type */
    WB200_EARG = 8,     /* bad filter length / step table / null pointer */
    WB200_EWORKSPACE = 9, /* caller workspace too small (see wb200_workspace_bytes) */
    WB200_ECUDA = 10    /* a CUDA runtime call failed; see wb200_last_error_string() */
} wb200_status;

/* element types (ValueType = Union{AbstractFloat, Complex}, transforms_main.jl:7) */
typedef enum {
    WB200_F32 = 0,
    WB200_F64 = 1,
    WB200_C64 = 2,  /* ComplexF32: interleaved (re, im) Float32 pairs */
    WB200_C128 = 3  /* ComplexF64 */
} wb200_dtype;

/* flags */
#define WB200_FLAG_STRICT_FP 1u     /* no FMA contraction: results are bit-identical to the reference CPU path
                                       (same operation order, separately rounded multiply and add).  Without the flag:
                                       the same order with FMAs, one accumulation chain per synthesis output, and the
                                       last four levels of a full-depth packet tree as one 16 x 16 map per node --
                                       within the reference's own Float32 GPU tolerance (1e-5, test/gpu.jl:24) */
#define WB200_FLAG_FORCE_GENERIC 2u /* bypass the fused sm_100a kernels (testing / A-B comparison) */

#define WB200_MAX_FILTER_LEN 64
#define WB200_MAX_LIFT_STEPS 16
#define WB200_MAX_LIFT_COEF 8

/* one lifting step, as stored in WT.SCHEMES (LSStep/LSStepParam, src/WT/wt_main.jl:195-209) */
typedef struct wb200_lift_step {
    int32_t is_predict; /* 1: PredictStep (updates the first half from the second), 0: UpdateStep */
    int32_t shift;      /* LSStepParam.shift */
    int32_t nc;         /* number of coefficients, <= WB200_MAX_LIFT_COEF */
    double coef[WB200_MAX_LIFT_COEF];
} wb200_lift_step;

/* ---- filter-bank DWT: replaces _dwt!(y, x, filter::OrthoFilter, L, fw) for 1-D/2-D/3-D arrays ----------
 * (transforms_filter.jl:13-62, 113-188, 192-294).  Out of place: y must not alias x.
 * ndim in {1,2,3}; every dims[i] % 2^L == 0; L == 0 copies x to y. fw != 0: dwt, fw == 0: idwt.
 * Output layout is the reference's: [a_L | d_L | ... | d_1] in 1-D, the Mallat square pyramid in N-D. */
int32_t wb200_dwt_filter(void *y, const void *x, int32_t ndim, const int64_t *dims, int64_t batch,
                         const double *qmf, int32_t flen, int32_t L, int32_t fw, int32_t dtype,
                         void *workspace, size_t workspace_bytes, void *stream, uint32_t flags);

/* ---- lifting DWT: replaces _dwt!(y, scheme::GLS, L, fw) (transforms_lifting.jl:30-76, 128-194, 200-278).
 * The reference transforms y in place; here x == y selects the in-place form (`dwt!(y, scheme, L)`),
 * x != y the allocating form (`dwt(x, scheme, L)` = similar + copyto! + in-place, transforms_main.jl:119-124)
 * without the extra copy.  2-D/3-D arrays must be square/cube (Util.iscube). */
int32_t wb200_dwt_lifting(void *y, const void *x, int32_t ndim, const int64_t *dims, int64_t batch,
                          const wb200_lift_step *steps, int32_t nsteps, double norm1, double norm2,
                          int32_t L, int32_t fw, int32_t dtype,
                          void *workspace, size_t workspace_bytes, void *stream, uint32_t flags);

/* ---- wavelet packet transform: replaces _wpt!(y, x, filter, tree, fw) (transforms_filter.jl:301-359) and
 * _wpt!(y, scheme, tree, fw) (transforms_lifting.jl:283-319).  1-D signals of length n, `batch` of them.
 * `tree` is the reference's BitVector expanded to one byte per node (1-based heap order: node i has
 * children 2i, 2i+1), ntree == 2^maxtransformlevels(n) - 1; validity is Util.isvalidtree
 * (src/Util/util_main.jl:301-313).  HOST pointer.  Output is in natural (Paley) order. */
int32_t wb200_wpt_filter(void *y, const void *x, int64_t n, int64_t batch,
                         const double *qmf, int32_t flen, const uint8_t *tree, int64_t ntree,
                         int32_t fw, int32_t dtype,
                         void *workspace, size_t workspace_bytes, void *stream, uint32_t flags);
int32_t wb200_wpt_lifting(void *y, const void *x, int64_t n, int64_t batch,
                          const wb200_lift_step *steps, int32_t nsteps, double norm1, double norm2,
                          const uint8_t *tree, int64_t ntree, int32_t fw, int32_t dtype,
                          void *workspace, size_t workspace_bytes, void *stream, uint32_t flags);

/* ---- maximal-overlap (undecimated) DWT: replaces modwt(x, wt, L) / imodwt(xw, wt)
 * (src/Transforms/transforms_maximal_overlap.jl:44-62, 98-108; SURVEY 8f "next" row 1).  x: n samples x `batch`
 * signals; y / xw: n x (L+1) column-major per signal (W_1 .. W_L, V_L), signals back to back.  Float32/Float64.
 * 1 <= L <= floor(log2(n)) (WB200_ELEVEL otherwise: "Too many transform levels (length(x) < 2^L)" / "L must be >= 1").
 * Scratch: n * batch elements (workspace = NULL: stream-ordered pool). */
int32_t wb200_modwt(void *y, const void *x, int64_t n, int64_t batch, const double *qmf, int32_t flen, int32_t L,
                    int32_t dtype, void *workspace, size_t workspace_bytes, void *stream, uint32_t flags);
int32_t wb200_imodwt(void *x, const void *xw, int64_t n, int64_t batch, const double *qmf, int32_t flen, int32_t ncols,
                     int32_t dtype, void *workspace, size_t workspace_bytes, void *stream, uint32_t flags);
int32_t wb200_maxmodwttransformlevels(int64_t n);                             /* non_dyadic.jl:24-25 */

/* ---- thresholding and denoising: the main caller of the transforms, device end to end ----
 * threshold!(x, TH, t)            src/Threshold/threshold_main.jl:21-117
 * noisest(x, wt, L = 1)           src/Threshold/denoising.jl:88-106: y = dwt(x, wt, L), MAD of y[detailrange(y, L)] =
 *                                 y[round(n1/2^L + 1) : round(n1/2^(L-1))] (LINEAR indices: a piece of the FIRST column,
 *                                 whatever ndim) over 0.6745.  Returns the number, so it waits for the stream.
 * denoise(x, wt; L, dnt, TI, nspin)  denoising.jl:22-82, dnt = VisuShrink(th_kind, tfac): t = sigma * tfac.  Pure enqueue:
 *                                 sigma = NaN estimates the noise level on the device (noisest) and the threshold kernel
 *                                 reads it from device memory; a finite sigma is the caller's `estnoise` result.
 * wkind 0: wt = nothing, 1: OrthoFilter (qmf, flen), 2: GLS (steps, nsteps, norm1, norm2).  One array per call. */
#define WB200_TH_HARD 0
#define WB200_TH_SOFT 1
#define WB200_TH_SEMISOFT 2
#define WB200_TH_STEIN 3
#define WB200_TH_NEG 4
#define WB200_TH_POS 5
int32_t wb200_threshold(void *x, int64_t count, int32_t kind, double t, int32_t dtype, void *stream);
/* threshold!(x, BiggestTH(), m): keep the m entries of largest magnitude (threshold_main.jl:21-33); ties at the cut are
 * dropped in index order (the reference's unstable QuickSort leaves that order unspecified). */
int32_t wb200_threshold_biggest(void *x, int64_t count, int64_t m, int32_t dtype, void *stream);
int32_t wb200_noisest(double *sigma_out, const void *x, int32_t ndim, const int64_t *dims, int32_t wkind,
                      const double *qmf, int32_t flen, const wb200_lift_step *steps, int32_t nsteps, double norm1,
                      double norm2, int32_t L, int32_t dtype, void *stream, uint32_t flags);
int32_t wb200_denoise(void *y, const void *x, int32_t ndim, const int64_t *dims, int32_t wkind, const double *qmf,
                      int32_t flen, const wb200_lift_step *steps, int32_t nsteps, double norm1, double norm2, int32_t L,
                      int32_t th_kind, double tfac, double sigma, int32_t TI, const int32_t *nspin, int32_t dtype,
                      void *stream, uint32_t flags);

/* ---- best basis: coefentropy / bestbasistree, src/Threshold/entropy.jl:16-129 ----
 * et 0: ShannonEntropy, 1: LogEnergyEntropy.  coefentropy: nrm = NaN means norm(x).  bestbasistree: y on the device (n samples),
 * tree / besttree host byte arrays (2^Lmax - 1 nodes, heap order); entr_bf (ntree) / entr_af (2^(Lmax-1)) optional host outputs
 * of the entropy tables.  Both return host values and therefore wait for the stream.
 * Ties: upstream keeps a node unsplit when `entr_bf[i] <= bestsubtree_entropy` (entropy.jl:97) on sums accumulated
 * sequentially in T; the device sums are reduced in double in a different order, so the comparison carries a relative slack
 * (1e-13 Float64, 1e-6 Float32) and a tie -- exact upstream, last-bit different here -- resolves as upstream: not split. */
int32_t wb200_coefentropy(double *out, const void *x, int64_t count, int32_t et, double nrm, int32_t dtype, void *stream);
int32_t wb200_bestbasistree(uint8_t *besttree, double *entr_bf, double *entr_af, const void *y, int64_t n, int32_t wkind,
                            const double *qmf, int32_t flen, const wb200_lift_step *steps, int32_t nsteps, double norm1,
                            double norm2, const uint8_t *tree, int64_t ntree, int32_t et, int32_t dtype, void *stream,
                            uint32_t flags);

/* ---- host-buffer forms (end-to-end path): x_host / y_host are HOST pointers (pinned memory gives full
 * PCIe bandwidth).  The batch is cut into chunks that are copied in, transformed and copied out on
 * alternating streams so the three stages overlap; the call returns after the last chunk has landed
 * in y_host.  `device` is the CUDA device ordinal. */
int32_t wb200_dwt_filter_host(void *y_host, const void *x_host, int32_t ndim, const int64_t *dims,
                              int64_t batch, const double *qmf, int32_t flen, int32_t L, int32_t fw,
                              int32_t dtype, int32_t device, uint32_t flags);
int32_t wb200_dwt_lifting_host(void *y_host, const void *x_host, int32_t ndim, const int64_t *dims,
                               int64_t batch, const wb200_lift_step *steps, int32_t nsteps,
                               double norm1, double norm2, int32_t L, int32_t fw,
                               int32_t dtype, int32_t device, uint32_t flags);

/* ---- workspace: bytes of device scratch a call needs (the reference allocates its own `si`, `snew`,
 * `tmpbuffer` per call, transforms_filter.jl:16-23,117-118).  Pass workspace = NULL to let the library take it
 * from the stream-ordered CUDA memory pool (cudaMallocAsync on `stream`).
 * kind: 0 = dwt_filter, 1 = dwt_lifting (out of place), 2 = dwt_lifting in place, 3 = wpt_filter, 4 = wpt_lifting */
size_t wb200_workspace_bytes(int32_t kind, int32_t ndim, const int64_t *dims, int64_t batch,
                             int32_t L, int32_t dtype, uint32_t flags);

/* ---- helpers mirrored from src/Util (host side, no GPU work) */
int32_t wb200_maxtransformlevels(int64_t n);                                  /* non_dyadic.jl:14-22 */
int32_t wb200_isvalidtree(int64_t n, const uint8_t *tree, int64_t ntree);     /* util_main.jl:301-313 */

/* ---- diagnostics */
const char *wb200_status_string(int32_t status); /* the reference's exception text for a status code */
const char *wb200_last_error_string(void);       /* thread-local detail of the last failure on this thread */
int32_t wb200_version(void);
/* number of kernel launches issued by this thread since the last reset (bench.py's gpu_launches) */
int64_t wb200_launch_count(int32_t reset);
/* opt-in per-kernel device timing: while enabled, every kernel launch of this thread is bracketed by CUDA
 * events on its stream.  wb200_profile_collect waits for them and writes "<kernel> <launches> <total_ms>\n"
 * lines into buf (returns bytes written) and clears the record. */
void wb200_profile_enable(int32_t on);
int64_t wb200_profile_collect(char *buf, int64_t buflen);


/* Scratch the library allocates itself (calls made with workspace == NULL) comes from a library-private
 * stream-ordered memory pool of the current device, which keeps freed scratch for the next call (the reference
 * re-allocates `si`, `snew`, `tmpbuffer` per call: src/Transforms/transforms_filter.jl:20-23; here that would be
 * a re-map of up to gigabytes per call).  wb200_trim_pool releases everything above keep_bytes back to the
 * device and returns the bytes the pool still reserves (-1 on error).  Callers that pass their own workspace
 * never touch the pool. */
int64_t wb200_trim_pool(int64_t keep_bytes);

#ifdef __cplusplus
}
#endif
#endif /* WAVELETS_B200_H */
