/*
 * b200itp.h — C ABI of libb200itp: the B200 (sm_100a) implementation of the
 * uniform-grid B-spline hot path of JuliaMath/Interpolations.jl v0.16.3.
 *
 * Plain C, `extern "C"`, plain pointers and sizes.  Every entry point names the
 * reference interface it replaces (paths relative to the reference checkout).
 *
 * Conventions
 *  - Arrays are dense, column-major (first index contiguous): exactly
 *    `parent(coefs)` of the reference (src/b-splines/prefiltering.jl:126-127).
 *  - All device buffers are owned by the caller.  The library allocates nothing
 *    that outlives a call except a small per-device scratch cache.
 *  - Return value: 0 success; <0 CUDA error (text via b200itp_last_error);
 *    >0 unsupported/invalid argument (B200ITP_E_*).
 *  - Out-of-bounds queries under `Throw` are not an ABI error: the smallest
 *    offending query index is returned through `first_oob` and the output there
 *    is NaN; the host glue raises BoundsError (src/b-splines/indexing.jl:6,
 *    src/extrapolation/extrapolation.jl:115-129).
 *  - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 *  - Indices in the reference's index space are 1-based, as in Julia.
 */
#ifndef B200ITP_H
#define B200ITP_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define B200ITP_API __attribute__((visibility("default")))
#else
#define B200ITP_API
#endif

#define B200ITP_VERSION 100
#define B200ITP_MAX_DIMS 4
#define B200ITP_MAX_WOODBURY 8

/* element types of arrays */
enum { B200ITP_F32 = 0, B200ITP_F64 = 1 };

/* Julia-style promotion tag of a bound / range element: an Int or Rational bound
 * does not widen a Float32 coordinate, a Float64 bound does
 * (src/extrapolation/extrapolation.jl:153-156 `maxp/minp`, Base.clamp). */
enum { B200ITP_TAG_EXACT = 0, B200ITP_TAG_F32 = 1, B200ITP_TAG_F64 = 2 };

/* interpolation degree per axis (src/b-splines/{linear,quadratic,cubic}.jl) */
enum { B200ITP_LINEAR = 1, B200ITP_QUADRATIC = 2, B200ITP_CUBIC = 3 };
/* degree 0, `Constant{Nearest|Previous|Next}` (src/b-splines/constant.jl:69-108): one tap per axis, weight 1 (an Int:
 * it does not widen the result type), gradient and hessian weights 0; bc is THROW (OnGrid) or PERIODIC (OnCell).
 * Evaluated by the run-time-degree kernels; never prefiltered. */
enum { B200ITP_CONSTANT_NEAREST = 4, B200ITP_CONSTANT_PREVIOUS = 5, B200ITP_CONSTANT_NEXT = 6 };
#define B200ITP_IS_CONSTANT(deg) ((deg) >= B200ITP_CONSTANT_NEAREST && (deg) <= B200ITP_CONSTANT_NEXT)
/* `NoInterp()` axis (src/nointerp/nointerp.jl:1-25, indexing.jl:74,108,137-140): the coordinate is an index — it must
 * hold an integer value (`Int(x)`; anything else is reported through first_oob like a BoundsError, the reference
 * raises InexactError) — weight 1; the axis has NO gradient component: out_grad holds G = (number of axes that are not
 * NOINTERP) components per query.  bc/gridtype of such an axis are THROW/ONGRID; bare or scaled interpolants only
 * (the range of a NoInterp axis is the axis itself, scaling.jl:46,108). */
enum { B200ITP_NOINTERP = 7 };
#define B200ITP_IS_INDEXLIKE(deg) (B200ITP_IS_CONSTANT(deg) || (deg) == B200ITP_NOINTERP)

/* boundary conditions (src/Interpolations.jl:89-118); also the extrapolation flags
 * (src/extrapolation/extrapolation.jl:106-151) */
enum {
  B200ITP_BC_THROW = 0,
  B200ITP_BC_FLAT = 1,
  B200ITP_BC_LINE = 2, /* == Natural */
  B200ITP_BC_FREE = 3,
  B200ITP_BC_PERIODIC = 4,
  B200ITP_BC_REFLECT = 5,
  /* Quadratic(InPlace(OnCell())) / Quadratic(InPlaceQ(OnCell())) (src/Interpolations.jl:104-109,
   * quadratic.jl:61-65,91-110): unpadded coefficients, the outer taps are clamped to the axis; not extrapolation flags */
  B200ITP_BC_INPLACE = 6,
  B200ITP_BC_INPLACEQ = 7
};
#define B200ITP_IS_INPLACE_BC(bc) ((bc) == B200ITP_BC_INPLACE || (bc) == B200ITP_BC_INPLACEQ)

/* grid type (src/Interpolations.jl:53-56) */
enum { B200ITP_ONGRID = 0, B200ITP_ONCELL = 1 };

/* extrapolation wrapper kind */
enum { B200ITP_ETP_NONE = 0, B200ITP_ETP_FLAGS = 1, B200ITP_ETP_FILL = 2 };

/* error codes (>0) */
enum {
  B200ITP_E_BADARG = 1,
  B200ITP_E_UNSUPPORTED = 2, /* descriptor outside the hot path: host glue throws ArgumentError */
  B200ITP_E_DTYPE = 3,
  B200ITP_E_SIZE = 4
};

/* bits of the per-query branch record written by b200itp_eval when dbg_branch != NULL.
 * One nibble per axis d (bits 4d..4d+3) plus global flags.  These are the
 * "boundary / extrapolation branch choices" that must match the reference bit for bit. */
#define B200ITP_BR_BELOW 0x1u  /* x < lower bound of the outermost wrapper            */
#define B200ITP_BR_ABOVE 0x2u  /* x > upper bound                                     */
#define B200ITP_BR_MOVED 0x4u  /* extrapolation mapped x to a different xs (xs != x)  */
#define B200ITP_BR_EDGE 0x8u   /* the degree's edge rule fired (cubic.jl:42-43, quadratic
                                  ceil branch indexing.jl:176, linear.jl:47)           */
#define B200ITP_BR_SLOWPATH (1u << 16) /* extrapolation.jl:55-57 gradient path taken   */
#define B200ITP_BR_FILLED (1u << 17)   /* filled.jl:38 fill value returned             */
#define B200ITP_BR_THROW (1u << 18)    /* BoundsError would be thrown                  */

/*
 * POD description of `extrapolate(scale(interpolate(A, BSpline(deg(bc(gt)))), ranges...), et)`
 * after `Adapt.adapt` (src/gpu_support.jl:4-8,32-48).  Supported wrapper order is
 * BSplineInterpolation [-> ScaledInterpolation] [-> Extrapolation | FilledExtrapolation].
 * All numbers the host can compute with the reference's own functions
 * (`lbounds/ubounds`, `first/step` of ranges) are passed down as doubles together
 * with the Julia type they had, so the kernel reproduces the promotion rules.
 */
typedef struct b200itp_desc {
  int32_t ndims;      /* 1..4                                                         */
  int32_t coef_dtype; /* B200ITP_F32 / F64 : eltype(itp.coefs)                         */
  int32_t coord_dtype;/* eltype of the broadcast arguments xs...                       */
  int32_t etp_pair_mask; /* bit d set: axis d was given an explicit (lo, hi) flag pair   */
  int64_t size[B200ITP_MAX_DIMS];       /* size(parent(itp.coefs)) (padded)           */
  int64_t parent_len[B200ITP_MAX_DIMS]; /* length.(itp.parentaxes); parent axes 1:n   */
  int32_t coef_first[B200ITP_MAX_DIMS]; /* first index of coefs' axis: 0 padded, 1 not
                                           (cubic.jl:86-87, quadratic.jl:64-65, linear.jl:60) */
  int32_t degree[B200ITP_MAX_DIMS];
  int32_t bc[B200ITP_MAX_DIMS];         /* only PERIODIC changes evaluation (modrange taps) */
  int32_t gridtype[B200ITP_MAX_DIMS];
  /* bounds(itp::BSplineInterpolation), b-splines.jl:141-158 */
  double inner_lb[B200ITP_MAX_DIMS], inner_ub[B200ITP_MAX_DIMS];
  int32_t inner_tag[B200ITP_MAX_DIMS];
  /* ScaledInterpolation, scaling.jl:5-8,60-75,102-115 */
  int32_t has_scale;
  int32_t range_tag[B200ITP_MAX_DIMS]; /* eltype(ranges[d])                            */
  double range_first[B200ITP_MAX_DIMS], range_step[B200ITP_MAX_DIMS];
  double outer_lb[B200ITP_MAX_DIMS], outer_ub[B200ITP_MAX_DIMS];
  int32_t outer_tag[B200ITP_MAX_DIMS];
  /* Extrapolation / FilledExtrapolation */
  int32_t etp_kind;
  int32_t etp_lo[B200ITP_MAX_DIMS], etp_hi[B200ITP_MAX_DIMS]; /* per-axis, per-side flags */
  int32_t fill_tag; /* Julia type of the fill value                                    */
  double fill_value;
  /* first.(itp.parentaxes): 0 or 1 = the axes 1:n of `interpolate`; 2 = the inset axes 2:n-1 that `interpolate!`
   * leaves on an axis whose degree pads (b-splines.jl:236-241 `padinset`): the coefficients then occupy 1:n
   * (coef_first = 1, size = parent_len + 2) and the bounds are those of 2:n-1.  Not combined with scale. */
  int32_t parent_first[B200ITP_MAX_DIMS];
} b200itp_desc;

/*
 * One axis of the prefilter: the factorisation `prefiltering_system` returns
 * (cubic.jl:114-264, quadratic.jl:84-189): `M = Woodbury(lut!(dl,d,du), U, C, V)`
 * or a bare LU.  All arrays are HOST pointers of element type `dtype`, read during
 * the call.  U has unit columns e_{urow[j]}, V unit rows e_{vcol[j]}', Cp is W.Cp.
 */
typedef struct b200itp_axis_system {
  int32_t dtype;
  int32_t k;       /* Woodbury rank, 0 for a bare LU (quadratic.jl:84-89)               */
  int64_t n;       /* size(M,1) == size(coefs, axis)                                   */
  const void *dl;  /* n-1 : F.factors.dl (LU multipliers)                              */
  const void *d;   /* n   : F.factors.d  (LU pivots)                                   */
  const void *du;  /* n-1 : F.factors.du (unchanged super-diagonal)                    */
  const int32_t *urow; /* k, 1-based                                                   */
  const int32_t *vcol; /* k, 1-based                                                   */
  const void *Cp;      /* k*k column-major                                             */
} b200itp_axis_system;

/* ---- library ---------------------------------------------------------------- */
B200ITP_API int b200itp_version(void);
/* copies the text of the last error of the calling thread; returns its length */
B200ITP_API size_t b200itp_last_error(char *buf, size_t buflen);
/* number of kernel launches issued by this process so far (bench.py's gpu_launches) */
B200ITP_API uint64_t b200itp_launch_count(void);

/* Per-stage timing for bench.py's roofline: when enabled, every kernel launch of the library is
 * bracketed by CUDA events on the caller's stream.  enable(1) clears the record list and starts,
 * enable(0) stops (records stay readable).  get(i) synchronises on record i and returns its name
 * (kernel / stage) and duration in milliseconds. */
B200ITP_API int b200itp_profile_enable(int on);
B200ITP_API int b200itp_profile_count(void);
B200ITP_API int b200itp_profile_get(int i, char *name, size_t namelen, double *ms);

/* ---- device memory helpers for hosts without their own CUDA binding ---------- */
B200ITP_API int b200itp_malloc(int dev, size_t bytes, void **out);
B200ITP_API int b200itp_free(int dev, void *p);
B200ITP_API int b200itp_memcpy_h2d(int dev, void *dst, const void *src, size_t bytes, void *stream);
B200ITP_API int b200itp_memcpy_d2h(int dev, void *dst, const void *src, size_t bytes, void *stream);
B200ITP_API int b200itp_sync(int dev, void *stream);

/* ---- construction: interpolate(A, it) ---------------------------------------- */

/* Replaces `copy_with_padding(TC, A, it)` (prefiltering.jl:17-28): `dst` has size
 * src_size[d] + 2*pad[d]; borders are zero, the interior is `src`. Device pointers. */
B200ITP_API int b200itp_copy_with_padding(int dev, const void *src, void *dst, int dtype, int ndims,
                              const int64_t *src_size, const int32_t *pad, void *stream);

/* Host-side restatement of `prefiltering_system(T, TC, n, degree)` + `lut!`
 * (cubic.jl:101-264, quadratic.jl:71-189, b-splines.jl:257) and of the
 * `Woodbury(A,U,C,V)` constructor's `Cp = inv(inv(C) + V*(A\U))`.  Output buffers:
 * dl[n-1], d[n], du[n-1], urow/vcol[B200ITP_MAX_WOODBURY], Cp[64]; *k receives the
 * rank (0: bare LU).  Returns B200ITP_E_UNSUPPORTED when the reference has no
 * method (e.g. Cubic(Reflect)), in which case the axis is left unfiltered
 * (prefiltering.jl:124). */
B200ITP_API int b200itp_prefiltering_system(int dtype, int64_t n, int degree, int bc, int gridtype,
                                void *dl, void *d, void *du, int32_t *k, int32_t *urow,
                                int32_t *vcol, void *Cp);

/* Replaces `A_ldiv_B_md!(dest, M, src, dim, b)` with dest === src and b == 0
 * (filter1d.jl:8-85; called from prefiltering.jl:49-59,108-122): solves M x = v in
 * place along `axis` (0-based) for every line of the device array `coefs`.
 * One HBM read and one write of the array. */
B200ITP_API int b200itp_prefilter_axis(int dev, void *coefs, int dtype, int ndims, const int64_t *size,
                           int axis, const b200itp_axis_system *sys, void *stream);

/* Replaces `prefilter!(TW, ret, it)` (prefiltering.jl:49-59,108-122): all axes in
 * order, systems assembled internally by b200itp_prefiltering_system. `size` is the
 * padded size. Axes whose degree is LINEAR are skipped (prefiltering.jl:30). */
B200ITP_API int b200itp_prefilter(int dev, void *coefs, int dtype, int ndims, const int64_t *size,
                      const int32_t *degree, const int32_t *bc, const int32_t *gridtype,
                      void *stream);
/* `prefilter(TW, TC, A, it)` for an array that needs no padding (prefiltering.jl:32-38 with padded_axes == axes):
 * the first axis pass reads the samples from `src` and writes `coefs` (same size and layout, must not overlap unless
 * equal), the other passes run in place on `coefs`: the copy the reference makes first is folded into the first pass
 * and the caller's array is left untouched. */
B200ITP_API int b200itp_prefilter_from(int dev, const void *src, void *coefs, int dtype, int ndims, const int64_t *size,
                           const int32_t *degree, const int32_t *bc, const int32_t *gridtype, void *stream);
/* `prefilter(TW, TC, A, it)` for an array that DOES need padding: `src` is the unpadded sample array (src_size), `coefs`
 * receives the padded (src_size + 2*pad per axis), prefiltered coefficient array.  `copy_with_padding`
 * (prefiltering.jl:17-28) is folded into the first axis pass where possible (the contiguous-axis streaming kernel reads
 * the samples and treats the line ends and the border lines as zeros), otherwise the two steps run one after the other;
 * same results either way. */
B200ITP_API int b200itp_prefilter_padded_from(int dev, const void *src, void *coefs, int dtype, int ndims,
                                  const int64_t *src_size, const int32_t *pad, const int32_t *degree,
                                  const int32_t *bc, const int32_t *gridtype, void *stream);
/* Long lines of arrays with few lines are cut into segments (one warm-up and one look-ahead block each) to fill the
 * machine; a segmented pass agrees with the unsegmented one to round-off (<= 1 ulp in ~1e-5 of the elements) but not
 * bit for bit.  Callers that need results independent of how an array is split into slabs (the slab-sharded prefilter:
 * same bits at every GPU count) switch segmentation off around their calls.  Returns the previous setting. */
B200ITP_API int b200itp_prefilter_set_segmentation(int enable);
/* Verification mode: every prefilter call runs `A_ldiv_B_md!` exactly as the reference does (filter1d.jl:21-85) —
 * forward / backward LU sweeps with the zero offset, the explicit six-step Woodbury correction with a full-size
 * temporary and a second sweep, no fused multiply-add — one thread per line.  Given the same factors it reproduces the
 * reference's CPU arithmetic bit for bit, so |fast - strict| isolates the fast kernels' truncation and re-association
 * error.  Slow (3 passes + temporary); never the default.  Returns the previous setting. */
B200ITP_API int b200itp_prefilter_set_strict(int enable);

/* ---- evaluation: itp.(xs...) and Interpolations.gradient.(Ref(itp), xs...) ---- */

/* Replaces `Base.copy(::Broadcasted)` of `itp.(xs...)` / `gradient.(Ref(itp), xs...)`
 * reached through src/gpu_support.jl:52-57,75-80; per element it computes what
 * indexing.jl:5-9,25-31, scaling.jl:78-83,123-128, extrapolation.jl:50-58,71-82 and
 * filled.jl:31-40,58-67 compute.
 *   coefs      device, dtype d->coef_dtype, size d->size
 *   coords     host array of d->ndims DEVICE pointers (SoA), dtype d->coord_dtype, nq each
 *   out_value  device or NULL, nq elements of b200itp_out_dtype(d)
 *   out_grad   device or NULL, nq*G elements, G contiguous per query (SVector); G = ndims minus the NOINTERP axes
 *   first_oob  host out (may be NULL): smallest query index that throws, -1 if none.
 *              When non-NULL the call synchronises `stream`.
 *   dbg_index  device or NULL: nq*ndims int32, the base tap index per axis (bit-exact contract)
 *   dbg_branch device or NULL: nq uint32 branch records (B200ITP_BR_*)
 */
B200ITP_API int b200itp_eval(int dev, const b200itp_desc *d, const void *coefs, const void *const *coords,
                 int64_t nq, void *out_value, void *out_grad, int64_t *first_oob,
                 int32_t *dbg_index, uint32_t *dbg_branch, void *stream);

/* cudaLimitMaxL2FetchGranularity of the device's context (32, 64 or 128 bytes; 0 = only query).  Returns the previous
 * value (negative: CUDA error). */
B200ITP_API int b200itp_set_l2_fetch_granularity(int dev, int bytes);

/* Pipelining hook for multi-GPU hosts: `cuda_event` (a cudaEvent_t, or NULL to clear) marks the moment the coefficient
 * array of the NEXT evaluation call of the CALLING THREAD (b200itp_eval / _eval_sorted / _gradient_sorted / _hessian /
 * _eval_grid, any stream, taken once) becomes
 * valid — e.g. the end of an NCCL broadcast or of the prefilter on another stream.  The call then waits for it only
 * right before its first kernel that reads coefficients: the ordered pipeline's count / scan / split kernels, which
 * only read the query coordinates, run while the coefficients are still on their way. */
B200ITP_API int b200itp_eval_set_coefs_event(void *cuda_event);

/* Tensor-product (grid) evaluation `itp(xvec, yvec, ...)` (src/b-splines/indexing.jl:16-23, src/scaling/scaling.jl:98-100,
 * src/extrapolation/extrapolation.jl:59-70): `axis_coords[d]` is a DEVICE vector of `axis_len[d]` coordinates of
 * d->coord_dtype; `out_value` receives prod(axis_len) elements of b200itp_out_dtype(d), first axis fastest (the Julia
 * result array).  Every element is what b200itp_eval returns for that coordinate tuple.  The tuples are expanded on the
 * device into `scratch` (at least b200itp_eval_grid_scratch_bytes bytes) and evaluated by the ordered pipeline where it
 * applies, by the direct kernel otherwise.  `first_oob` is the linear index (first axis fastest) of the first tuple that
 * throws. */
B200ITP_API size_t b200itp_eval_grid_scratch_bytes(const b200itp_desc *d, const int64_t *axis_len);
/* Where the per-axis structure allows (every case but Line extrapolation, fill values and NoInterp axes) the grid is
 * evaluated SEPARABLY: the reference's nested reduction (Interpolations.jl:304-316) run as ndims one-dimensional
 * resampling passes, last axis first — the same operations in the same order, so the same bits as the pointwise
 * evaluation, but coalesced streams instead of prod(axis_len) scattered stencil gathers (`eachvalue(sitp)`,
 * scaling.jl:174-256, is this call on the knots of the ranges).  set_separable(0) forces the tuple-expansion path
 * (verification); separable_calls() counts the calls the separable kernels served. */
B200ITP_API int b200itp_eval_grid_set_separable(int enable);
B200ITP_API uint64_t b200itp_eval_grid_separable_calls(void);
B200ITP_API int b200itp_eval_grid(int dev, const b200itp_desc *d, const void *coefs, const void *const *axis_coords,
                      const int64_t *axis_len, void *out_value, int64_t *first_oob, void *scratch,
                      size_t scratch_bytes, void *stream);

/* Replaces `Base.copy(::Broadcasted)` of `Interpolations.hessian.(Ref(itp), xs...)` (src/b-splines/indexing.jl:37-41,
 * 114-144,201-229; src/scaling/scaling.jl:150-160; src/extrapolation/extrapolation.jl:84-90): `out_hess` receives
 * nq*ndims*ndims elements of b200itp_out_dtype(d), the full symmetric matrix per query (SMatrix layout).  Extrapolation
 * wrappers evaluate in bounds only: a query outside is reported through `first_oob` like a BoundsError (the reference
 * raises "extrapolation of hessian not yet implemented"); FilledExtrapolation has no hessian method (E_UNSUPPORTED). */
B200ITP_API int b200itp_hessian(int dev, const b200itp_desc *d, const void *coefs, const void *const *coords,
                    int64_t nq, void *out_hess, int64_t *first_oob, void *stream);

/* Element type of the outputs of b200itp_eval for this descriptor:
 * promote(coordinate type after all wrappers, eltype(coefs)). <0 on invalid descriptor. */
B200ITP_API int b200itp_out_dtype(const b200itp_desc *d);

/* Same as b200itp_eval (values only, no debug outputs), but the queries are first
 * ordered on the device by the brick of coefficients they touch (two streaming
 * multisplit passes), evaluated brick by brick from shared memory, and the results
 * are streamed back into the caller's query order (two more multisplit passes):
 * every pass moves whole cache lines, no scattered DRAM access.  Results are
 * bit-identical to b200itp_eval.  `scratch` is a device buffer of at least
 * b200itp_eval_sorted_scratch_bytes(d, nq) bytes (about 32 bytes per query).
 * Calls with fewer than b200itp_eval_sorted_min_queries() queries, or a descriptor
 * outside the ordered path's scope, are served by the direct kernel. */
B200ITP_API size_t b200itp_eval_sorted_scratch_bytes(const b200itp_desc *d, int64_t nq);
B200ITP_API int b200itp_eval_sorted(int dev, const b200itp_desc *d, const void *coefs,
                        const void *const *coords, int64_t nq, void *out_value,
                        int64_t *first_oob, void *scratch, size_t scratch_bytes, void *stream);
/* `Interpolations.gradient.(Ref(itp), xs...)` through the same ordered pipeline: `out_grad` receives nq*ndims elements,
 * ndims contiguous per query (SVector layout), bit-identical to b200itp_eval's gradients.  Same scope, scratch size and
 * fallback rule as b200itp_eval_sorted. */
B200ITP_API int b200itp_gradient_sorted(int dev, const b200itp_desc *d, const void *coefs,
                            const void *const *coords, int64_t nq, void *out_grad, int64_t *first_oob,
                            void *scratch, size_t scratch_bytes, void *stream);
/* threshold (queries per call) from which the ordered pipeline is used; default 2^20, and for 2-D arrays only
 * from 128 MiB of coefficients (smaller images stay L2-resident, where the direct kernel is faster; 3-D arrays of any
 * size take the ordered path).  Setting the
 * threshold to 1 (tests) also lifts the array-size condition. */
B200ITP_API int64_t b200itp_eval_sorted_min_queries(void);
B200ITP_API void b200itp_eval_sorted_set_min_queries(int64_t n);
/* 1 (default): the evaluation kernel of the ordered pipeline fetches a brick + halo tile with one tensor-map bulk copy
 * (cp.async.bulk.tensor) whenever no axis of the tile wraps around a periodic boundary and the array qualifies
 * (16-byte aligned base and rows); 0: every tile is staged with cp.async.  Identical results; returns the previous
 * setting. */
B200ITP_API int b200itp_eval_sorted_set_tma(int on);
/* number of ordered-pipeline calls of this process whose evaluation kernel used the tensor-map tile copy */
B200ITP_API uint64_t b200itp_eval_sorted_tma_calls(void);
/* number of b200itp_eval_sorted calls of this process that were served by the ordered pipeline (the others were
 * forwarded to the direct kernel): lets tests and bench.py assert which path ran */
B200ITP_API uint64_t b200itp_eval_sorted_ordered_calls(void);

/* ---- multi-GPU on one box: one host process drives the devices, NCCL over NVLink / NVSwitch ------------
 * libnccl.so.2 is loaded at run time by b200itp_comm_init (no link-time dependency).  Evaluation needs no collective of
 * its own: queries are independent, so each device evaluates a contiguous chunk of the query arrays against its replica
 * of the coefficient array (b200itp_eval* with that device's pointers). */
typedef struct b200itp_comm b200itp_comm;
/* ncclCommInitAll over `devs[0..ngpu)` plus one stream per device. */
B200ITP_API int b200itp_comm_init(int ngpu, const int *devs, b200itp_comm **comm);
B200ITP_API int b200itp_comm_size(const b200itp_comm *comm);
B200ITP_API int b200itp_comm_destroy(b200itp_comm *comm);
/* Replicate the coefficient array held by device `root` on every device: coefs_per_dev[i] is the buffer (device
 * pointer, `bytes` bytes) on devs[i] (one ncclBroadcast; returns when the copies have landed). */
B200ITP_API int b200itp_bcast_coefs(b200itp_comm *comm, int root, void *const *coefs_per_dev, size_t bytes);
/* `interpolate(A, it)` for a large 3-D array, slab-sharded: samples_slab[i] is devs[i]'s z-slab of the samples
 * (size[0] x size[1] x nz_i column-major, the z-planes split as evenly as possible in device order, sizes differing by at
 * most one); coefs_full[i] receives the FULL coefficient array (size[0] x size[1] x size[2]) on every device, ready for
 * query-sharded evaluation.  Passes x and y are local to the slabs, one all-to-all (ncclSend / ncclRecv) re-shards along
 * y, the z pass is local, one all-gather replicates.  Same pass order and systems as the reference
 * (prefiltering.jl:49-59), same bits as b200itp_prefilter on one device.  Unpadded specs (Periodic, InPlace, Linear
 * axes); for padding boundary conditions pass slabs of the padded array.  ms_out (nullable, 2 doubles): device time up
 * to the end of the z pass, and including the replication. */
B200ITP_API int b200itp_prefilter_sharded(b200itp_comm *comm, const void *const *samples_slab, void *const *coefs_full,
                              int dtype, const int64_t *size, const int32_t *degree, const int32_t *bc,
                              const int32_t *gridtype, double *ms_out);

#ifdef __cplusplus
}
#endif
#endif /* B200ITP_H */
