// Host side of b200itp_eval / b200itp_out_dtype / b200itp_eval_sorted (include/b200itp.h):
// validates the descriptor, derives the Julia promotion chain, fills EvalParams and launches the
// kernels of eval.cuh.
#include "eval.cuh"

#include <climits>
#include <cmath>

namespace b200itp {

template <typename TC, typename TX, typename TW>
int launch_eval_combo(const EvalParams &P, int deg, int mode, int blocks, cudaStream_t st);

namespace {

struct Chain {
  int t0;          // coordinate tag
  int t1[4];       // after the extrapolation stage
  int t2[4];       // after coordlookup
  int tw[4];       // weights type
  int out_f64;
  bool homogeneous;
  int e_wide, s_wide, w_wide;
  int s_wide_x, w_wide_x;  // the in-bounds gradient path of an Extrapolation uses x itself
};

inline int tmax(int a, int b) { return a > b ? a : b; }

// Julia's promotion through maxp/minp (extrapolation.jl:153-156), coordlookup (scaling.jl:102-115) and
// clamp (Interpolations.jl:450-452)
int derive_chain(const b200itp_desc *D, Chain &c) {
  if (!D) return B200ITP_E_BADARG;
  if (D->ndims < 1 || D->ndims > B200ITP_MAX_DIMS) return B200ITP_E_BADARG;
  if ((D->coef_dtype != B200ITP_F32 && D->coef_dtype != B200ITP_F64) ||
      (D->coord_dtype != B200ITP_F32 && D->coord_dtype != B200ITP_F64))
    return B200ITP_E_DTYPE;
  c.t0 = D->coord_dtype == B200ITP_F32 ? B200ITP_TAG_F32 : B200ITP_TAG_F64;
  int tacc = D->coef_dtype == B200ITP_F32 ? B200ITP_TAG_F32 : B200ITP_TAG_F64;
  c.homogeneous = true;
  for (int d = 0; d < D->ndims; ++d) {
    int t = c.t0;
    if (D->etp_kind == B200ITP_ETP_FLAGS &&
        !(D->etp_lo[d] == B200ITP_BC_THROW && D->etp_hi[d] == B200ITP_BC_THROW))
      t = tmax(t, D->has_scale ? D->outer_tag[d] : D->inner_tag[d]);
    c.t1[d] = t;
    if (D->has_scale) t = tmax(t, D->range_tag[d]);
    c.t2[d] = t;
    if (D->has_scale) t = tmax(t, D->inner_tag[d]);
    c.tw[d] = t;
    if (!B200ITP_IS_INDEXLIKE(D->degree[d])) tacc = tmax(tacc, t);  // Constant / NoInterp weights are Ints (constant.jl:106)
  }
  {
    // every axis that computes weights runs in one type; a NoInterp axis holds an integer, which every type carries
    int ref = -1;
    for (int d = 0; d < D->ndims; ++d)
      if (D->degree[d] != B200ITP_NOINTERP) {
        if (ref < 0) ref = d;
        if (c.t1[d] != c.t1[ref] || c.t2[d] != c.t2[ref] || c.tw[d] != c.tw[ref]) c.homogeneous = false;
      }
    if (ref > 0) {  // the kernel reads the types of axis 0
      c.t1[0] = c.t1[ref];
      c.t2[0] = c.t2[ref];
      c.tw[0] = c.tw[ref];
    }
  }
  if (D->etp_kind == B200ITP_ETP_FILL) tacc = tmax(tacc, D->fill_tag);
  c.out_f64 = tacc == B200ITP_TAG_F64;
  c.e_wide = c.t1[0] == B200ITP_TAG_F64;
  c.s_wide = c.t2[0] == B200ITP_TAG_F64;
  c.w_wide = c.tw[0] == B200ITP_TAG_F64;
  {
    int t = c.t0;
    if (D->has_scale) t = tmax(t, D->range_tag[0]);
    c.s_wide_x = t == B200ITP_TAG_F64;
    if (D->has_scale) t = tmax(t, D->inner_tag[0]);
    c.w_wide_x = t == B200ITP_TAG_F64;
  }
  return 0;
}

inline double round_tag(int tag, double v) { return tag == B200ITP_TAG_F32 ? (double)(float)v : v; }

int fill_params(const b200itp_desc *D, const Chain &c, const void *coefs, const void *const *coords, int64_t nq,
                void *out_value, void *out_grad, int32_t *dbg_index, uint32_t *dbg_branch,
                unsigned long long *first_oob_dev, EvalParams &P) {
  memset(&P, 0, sizeof P);
  P.ndims = D->ndims;
  long long stride = 1;
  for (int d = 0; d < D->ndims; ++d) {
    if (D->size[d] < 1 || D->parent_len[d] < 1) B200_FAIL(B200ITP_E_SIZE, "eval: empty axis %d", d);
    if (D->parent_len[d] > INT_MAX / 2) B200_FAIL(B200ITP_E_SIZE, "eval: axis %d too long", d);
    const int deg = D->degree[d];
    if ((deg < B200ITP_LINEAR || deg > B200ITP_CUBIC) && !B200ITP_IS_INDEXLIKE(deg))
      B200_FAIL(B200ITP_E_UNSUPPORTED, "eval: degree %d on axis %d is not on the hot path", deg, d);
    if (deg == B200ITP_NOINTERP) {
      if (D->etp_kind != B200ITP_ETP_NONE)
        B200_FAIL(B200ITP_E_UNSUPPORTED, "eval: extrapolation of NoInterp axes is not on the hot path");
      if (D->bc[d] == B200ITP_BC_PERIODIC) B200_FAIL(B200ITP_E_BADARG, "eval: NoInterp axes have no boundary condition");
      if (D->has_scale && !(D->range_first[d] == 1.0 && D->range_step[d] == 1.0))
        B200_FAIL(B200ITP_E_BADARG, "eval: the range of a NoInterp axis must equal the axis (scaling.jl:46)");
    }
    if (B200ITP_IS_CONSTANT(deg)) {
      if (D->bc[d] != B200ITP_BC_THROW && D->bc[d] != B200ITP_BC_PERIODIC)
        B200_FAIL(B200ITP_E_BADARG, "eval: Constant takes Throw(OnGrid()) or Periodic(OnCell()) (constant.jl:33)");
      if (D->etp_kind == B200ITP_ETP_FLAGS && (D->etp_lo[d] == B200ITP_BC_LINE || D->etp_hi[d] == B200ITP_BC_LINE))
        B200_FAIL(B200ITP_E_UNSUPPORTED, "eval: Line extrapolation of a Constant axis is not on the hot path");
      if (D->has_scale && out_grad)
        B200_FAIL(B200ITP_E_UNSUPPORTED, "eval: gradients of scaled Constant axes are not on the hot path");
    }
    int cf = D->coef_first[d];
    if (cf != 0 && cf != 1) B200_FAIL(B200ITP_E_BADARG, "eval: coef_first must be 0 or 1");
    const int pf = D->parent_first[d] == 0 ? 1 : D->parent_first[d];
    if (pf != 1 && pf != 2) B200_FAIL(B200ITP_E_BADARG, "eval: parent_first must be 1 or 2");
    if (B200ITP_IS_INPLACE_BC(D->bc[d])) {
      if (deg != B200ITP_QUADRATIC || D->gridtype[d] != B200ITP_ONCELL || cf != 1)
        B200_FAIL(B200ITP_E_BADARG, "eval: InPlace/InPlaceQ go with unpadded Quadratic OnCell axes (quadratic.jl:61-65,91-110)");
      P.clamp_taps[d] = 1;
    }
    if (pf == 2) {  // interpolate!: coefficients 1:n, parent axes 2:n-1 -> the padded layout seen from x - 1
      if (cf != 1) B200_FAIL(B200ITP_E_BADARG, "eval: inset parent axes go with unpadded coefficients (coef_first = 1)");
      if (D->has_scale) B200_FAIL(B200ITP_E_UNSUPPORTED, "eval: scale of an interpolate! object with inset axes is not supported");
      cf = 0;
      P.xshift[d] = 1;
    }
    const int64_t need = D->parent_len[d] + (cf == 0 ? 2 : 0);
    if (D->size[d] != need) B200_FAIL(B200ITP_E_SIZE, "eval: size[%d]=%lld inconsistent with parent length %lld", d, (long long)D->size[d], (long long)D->parent_len[d]);
    if (D->bc[d] == B200ITP_BC_PERIODIC && cf != 1) B200_FAIL(B200ITP_E_BADARG, "eval: periodic axes are unpadded");
    P.n[d] = (int)D->parent_len[d];
    P.coef_first[d] = cf;
    P.degree[d] = deg;
    P.periodic[d] = D->bc[d] == B200ITP_BC_PERIODIC;
    P.stride[d] = stride;
    stride *= D->size[d];
    const double lo = D->has_scale ? D->outer_lb[d] : D->inner_lb[d];
    const double hi = D->has_scale ? D->outer_ub[d] : D->inner_ub[d];
    const int bt = D->has_scale ? D->outer_tag[d] : D->inner_tag[d];
    if (!(lo < hi)) B200_FAIL(B200ITP_E_BADARG, "eval: empty bounds on axis %d", d);
    P.lb[d] = lo;
    P.ub[d] = hi;
    P.span[d] = round_tag(bt, hi - lo);
    P.span2[d] = round_tag(bt, 2.0 * P.span[d]);
    P.ub2[d] = round_tag(bt, 2.0 * hi);
    P.inner_lb[d] = D->inner_lb[d];
    P.inner_ub[d] = D->inner_ub[d];
    if (D->has_scale) {
      if (!(D->range_step[d] > 0)) B200_FAIL(B200ITP_E_BADARG, "eval: ranges must have positive steps (scaling.jl:48)");
      P.rfirst[d] = D->range_first[d];
      P.rstep[d] = D->range_step[d];
      P.range_f32[d] = D->range_tag[d] == B200ITP_TAG_F32;
    }
    P.etp_lo[d] = D->etp_lo[d];
    P.etp_hi[d] = D->etp_hi[d];
    if (D->etp_kind == B200ITP_ETP_FLAGS) {
      for (int f : {D->etp_lo[d], D->etp_hi[d]})
        if (f != B200ITP_BC_THROW && f != B200ITP_BC_FLAT && f != B200ITP_BC_LINE && f != B200ITP_BC_PERIODIC &&
            f != B200ITP_BC_REFLECT)
          B200_FAIL(B200ITP_E_UNSUPPORTED, "eval: extrapolation flag %d is not supported", f);
      if (D->etp_lo[d] == B200ITP_BC_LINE || D->etp_hi[d] == B200ITP_BC_LINE) P.any_line = 1;
      if (!((D->etp_pair_mask >> d) & 1) && D->etp_lo[d] != D->etp_hi[d])
        B200_FAIL(B200ITP_E_BADARG, "eval: axis %d has two different flags but is not marked as a pair", d);
    }
  }
  P.ngrad = 0;
  for (int d = 0; d < D->ndims; ++d) P.grad_slot[d] = D->degree[d] == B200ITP_NOINTERP ? -1 : P.ngrad++;
  P.total = stride;
  P.vec_ok = ((uintptr_t)coefs % 16) == 0;
  P.has_scale = D->has_scale != 0;
  P.e_wide = c.e_wide;
  P.s_wide = c.s_wide;
  P.s_wide_x = c.s_wide_x;
  P.w_wide_x = c.w_wide_x;
  P.etp_kind = D->etp_kind;
  P.pair_mask = D->etp_kind == B200ITP_ETP_FLAGS ? D->etp_pair_mask : 0;
  P.fill = D->fill_value;
  P.out_f64 = c.out_f64;
  for (int d = 0; d < D->ndims; ++d) P.coords[d] = coords[d];
  P.coefs = coefs;
  P.out_value = out_value;
  P.out_grad = out_grad;
  P.dbg_index = dbg_index;
  P.dbg_branch = dbg_branch;
  P.first_oob = first_oob_dev;
  P.nq = nq;
  P.perm = nullptr;
  return 0;
}

int dispatch(const b200itp_desc *D, const Chain &c, const EvalParams &P, int mode, int blocks, cudaStream_t st) {
  int deg = D->degree[0];
  for (int d = 1; d < D->ndims; ++d)
    if (D->degree[d] != deg) deg = 0;  // mixed per-axis degrees: run-time degree variant
  if (B200ITP_IS_INDEXLIKE(deg)) deg = 0;  // Constant / NoInterp axes are served by the run-time degree variant only
  for (int d = 0; d < D->ndims; ++d)
    if (P.xshift[d] || P.clamp_taps[d]) deg = 0;  // ... and so are the inset axes of interpolate! and InPlace axes
  const bool cf32 = D->coef_dtype == B200ITP_F32, xf32 = D->coord_dtype == B200ITP_F32, wf32 = !c.w_wide;
  StageScope sc(mode == 0 ? "eval_direct_value" : (mode == 1 ? "eval_direct_gradient" : "eval_direct_hessian"), st);
  int rc;
  if (cf32 && xf32 && wf32)
    rc = launch_eval_combo<float, float, float>(P, deg, mode, blocks, st);
  else if (cf32 && xf32 && !wf32)
    rc = launch_eval_combo<float, float, double>(P, deg, mode, blocks, st);
  else if (cf32 && !xf32)
    rc = launch_eval_combo<float, double, double>(P, deg, mode, blocks, st);
  else if (!cf32 && !xf32)
    rc = launch_eval_combo<double, double, double>(P, deg, mode, blocks, st);
  else if (!cf32 && xf32 && wf32)
    rc = launch_eval_combo<double, float, float>(P, deg, mode, blocks, st);
  else
    rc = launch_eval_combo<double, float, double>(P, deg, mode, blocks, st);
  if (rc) B200_FAIL(rc, "eval: no kernel for ndims=%d degree=%d", D->ndims, deg);
  count_launch();
  B200_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace

}  // namespace b200itp

using namespace b200itp;

extern "C" int b200itp_out_dtype(const b200itp_desc *d) {
  Chain c;
  if (derive_chain(d, c)) return -1;
  return c.out_f64 ? B200ITP_F64 : B200ITP_F32;
}

static int eval_common(int dev, const b200itp_desc *D, const void *coefs, const void *const *coords, int64_t nq,
                       void *out_value, void *out_grad, int64_t *first_oob, int32_t *dbg_index, uint32_t *dbg_branch,
                       const int *perm, cudaStream_t st) {
  if (!D || !coefs || !coords) B200_FAIL(B200ITP_E_BADARG, "eval: null pointer");
  if (nq < 0) B200_FAIL(B200ITP_E_BADARG, "eval: negative query count");
  Chain c;
  int rc = derive_chain(D, c);
  if (rc) B200_FAIL(rc, "eval: invalid descriptor");
  if (!c.homogeneous)
    B200_FAIL(B200ITP_E_UNSUPPORTED,
              "eval: Float32 coordinates that are widened to Float64 on some axes only (mixed OnGrid/OnCell or mixed "
              "range element types) are not on the hot path; pass Float64 coordinates");
  if (D->etp_kind != B200ITP_ETP_NONE && D->etp_kind != B200ITP_ETP_FLAGS && D->etp_kind != B200ITP_ETP_FILL)
    B200_FAIL(B200ITP_E_BADARG, "eval: bad etp_kind");
  if (out_grad && D->etp_kind == B200ITP_ETP_FLAGS && D->etp_pair_mask)
    B200_FAIL(B200ITP_E_UNSUPPORTED,
              "eval: gradient with per-side extrapolation pairs has no method in the reference (extrapolation.jl:184-185)");
  for (int d = 0; d < D->ndims; ++d)
    if (!coords[d]) B200_FAIL(B200ITP_E_BADARG, "eval: null coordinate array %d", d);
  DeviceGuard g(dev);
  if (!g.ok) B200_FAIL(-(int)cudaErrorInvalidDevice, "eval: cannot select device %d", dev);
  if (first_oob) *first_oob = -1;
  if (cudaEvent_t ev = take_coefs_event()) B200_CUDA_OK(cudaStreamWaitEvent(st, ev, 0));  // coefficients produced elsewhere
  if (nq == 0 || (!out_value && !out_grad)) return 0;

  void *flag = nullptr;
  rc = scratch_get(dev, 2, st, 256, &flag);
  if (rc) return rc;
  B200_CUDA_OK(cudaMemsetAsync(flag, 0xFF, sizeof(unsigned long long), st));

  EvalParams P;
  const int blocks = (int)std::min<int64_t>((nq + 255) / 256, (int64_t)sm_count(dev) * 32);
  if (out_value) {
    rc = fill_params(D, c, coefs, coords, nq, out_value, nullptr, dbg_index, dbg_branch, (unsigned long long *)flag, P);
    if (rc) return rc;
    P.perm = perm;
    rc = dispatch(D, c, P, 0, blocks, st);
    if (rc) return rc;
  }
  if (out_grad) {
    // when both outputs are requested the debug record describes the value evaluation
    rc = fill_params(D, c, coefs, coords, nq, nullptr, out_grad, out_value ? nullptr : dbg_index,
                     out_value ? nullptr : dbg_branch, (unsigned long long *)flag, P);
    if (rc) return rc;
    P.perm = perm;
    rc = dispatch(D, c, P, 1, blocks, st);
    if (rc) return rc;
  }
  if (first_oob) {
    unsigned long long h = ~0ULL;
    B200_CUDA_OK(cudaMemcpyAsync(&h, flag, sizeof h, cudaMemcpyDeviceToHost, st));
    B200_CUDA_OK(cudaStreamSynchronize(st));
    *first_oob = h == ~0ULL ? -1 : (int64_t)h;
  }
  return 0;
}

extern "C" int b200itp_eval(int dev, const b200itp_desc *d, const void *coefs, const void *const *coords, int64_t nq,
                            void *out_value, void *out_grad, int64_t *first_oob, int32_t *dbg_index,
                            uint32_t *dbg_branch, void *stream) {
  return eval_common(dev, d, coefs, coords, nq, out_value, out_grad, first_oob, dbg_index, dbg_branch, nullptr,
                     (cudaStream_t)stream);
}

extern "C" int b200itp_hessian(int dev, const b200itp_desc *Din, const void *coefs, const void *const *coords, int64_t nq,
                               void *out_hess, int64_t *first_oob, void *stream) {
  if (!Din || !coefs || !coords || !out_hess) B200_FAIL(B200ITP_E_BADARG, "hessian: null pointer");
  if (nq < 0) B200_FAIL(B200ITP_E_BADARG, "hessian: negative query count");
  if (Din->etp_kind == B200ITP_ETP_FILL)
    B200_FAIL(B200ITP_E_UNSUPPORTED, "hessian: FilledExtrapolation has no hessian method in the reference (filled.jl:75)");
  if (Din->etp_kind != B200ITP_ETP_NONE && Din->etp_kind != B200ITP_ETP_FLAGS) B200_FAIL(B200ITP_E_BADARG, "hessian: bad etp_kind");
  // hessian(etp, x) is hessian(parent(etp), x) in bounds and an error outside (extrapolation.jl:84-90): exactly the
  // unwrapped interpolant, including its types (the extrapolation bounds never meet the coordinates)
  b200itp_desc Dcopy = *Din;
  Dcopy.etp_kind = B200ITP_ETP_NONE;
  const b200itp_desc *D = &Dcopy;
  Chain c;
  int rc = derive_chain(D, c);
  if (rc) B200_FAIL(rc, "hessian: invalid descriptor");
  if (!c.homogeneous) B200_FAIL(B200ITP_E_UNSUPPORTED, "hessian: mixed per-axis promotion is not on the hot path");
  for (int d = 0; d < D->ndims; ++d) {
    if (!coords[d]) B200_FAIL(B200ITP_E_BADARG, "hessian: null coordinate array %d", d);
    if (D->has_scale && B200ITP_IS_CONSTANT(D->degree[d]))
      B200_FAIL(B200ITP_E_UNSUPPORTED, "hessian: scaled Constant axes are not on the hot path");
    if (D->degree[d] == B200ITP_NOINTERP)
      B200_FAIL(B200ITP_E_UNSUPPORTED, "hessian: NoInterp axes are not on the hot path");
  }
  DeviceGuard g(dev);
  if (!g.ok) B200_FAIL(-(int)cudaErrorInvalidDevice, "hessian: cannot select device %d", dev);
  if (first_oob) *first_oob = -1;
  if (nq == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (cudaEvent_t ev = take_coefs_event()) B200_CUDA_OK(cudaStreamWaitEvent(st, ev, 0));  // coefficients produced elsewhere
  void *flag = nullptr;
  rc = scratch_get(dev, 2, st, 256, &flag);
  if (rc) return rc;
  B200_CUDA_OK(cudaMemsetAsync(flag, 0xFF, sizeof(unsigned long long), st));
  // the hessian of an Extrapolation is the parent's, in bounds only: the chain of the in-bounds gradient path
  // (no extrapolation-stage promotion) decides the output type
  EvalParams P;
  const int blocks = (int)std::min<int64_t>((nq + 255) / 256, (int64_t)sm_count(dev) * 32);
  rc = fill_params(D, c, coefs, coords, nq, nullptr, nullptr, nullptr, nullptr, (unsigned long long *)flag, P);
  if (rc) return rc;
  P.out_hess = out_hess;
  rc = dispatch(D, c, P, 2, blocks, st);
  if (rc) return rc;
  if (first_oob) {
    unsigned long long h = ~0ULL;
    B200_CUDA_OK(cudaMemcpyAsync(&h, flag, sizeof h, cudaMemcpyDeviceToHost, st));
    B200_CUDA_OK(cudaStreamSynchronize(st));
    *first_oob = h == ~0ULL ? -1 : (int64_t)h;
  }
  return 0;
}

namespace b200itp {
// used by eval_sorted.cu: validated parameters for a values-only evaluation, plus whether the descriptor is
// a bare Float32 B-spline interpolant (the ordered path's scope) and its uniform degree (0: mixed)
// grid evaluation (eval_grid.cu): the per-axis parameters of the direct kernel with the coordinate VECTORS as `coords`
int fill_eval_params_grid(const b200itp_desc *D, const void *coefs, const void *const *coords, EvalParams &P,
                          bool *w_wide, bool *homogeneous) {
  Chain c;
  int rc = derive_chain(D, c);
  if (rc) B200_FAIL(rc, "eval_grid: invalid descriptor");
  rc = fill_params(D, c, coefs, coords, 0, nullptr, nullptr, nullptr, nullptr, nullptr, P);
  if (rc) return rc;
  *w_wide = c.w_wide != 0;
  *homogeneous = c.homogeneous;
  return 0;
}

int fill_eval_params_public(const b200itp_desc *D, const void *coefs, const void *const *coords, int64_t nq,
                            void *out_value, unsigned long long *first_oob_dev, EvalParams &P, int *uniform_degree,
                            bool *plain_f32) {
  Chain c;
  int rc = derive_chain(D, c);
  if (rc) B200_FAIL(rc, "eval_sorted: invalid descriptor");
  if (nq < 0) B200_FAIL(B200ITP_E_BADARG, "eval_sorted: negative query count");
  for (int d = 0; d < D->ndims; ++d)
    if (!coords[d]) B200_FAIL(B200ITP_E_BADARG, "eval_sorted: null coordinate array %d", d);
  rc = fill_params(D, c, coefs, coords, nq, out_value, nullptr, nullptr, nullptr, first_oob_dev, P);
  if (rc) return rc;
  int deg = D->degree[0];
  for (int d = 1; d < D->ndims; ++d)
    if (D->degree[d] != deg) deg = 0;
  *uniform_degree = deg;
  // Float32 throughout: coefficients, coordinates, and — for scaled / extrapolated objects — every stage of the
  // coordinate chain (no Float64 bound or range widens it: c.w_wide covers the extrapolation and coordlookup stages, which
  // only ever widen).  Line extrapolation (value + gradient * dx) and fill values stay with the direct kernel.
  bool etp_ok = D->etp_kind == B200ITP_ETP_NONE;
  if (D->etp_kind == B200ITP_ETP_FLAGS) {
    etp_ok = !P.any_line;
  }
  *plain_f32 = c.homogeneous && etp_ok && D->coef_dtype == B200ITP_F32 && D->coord_dtype == B200ITP_F32 && !c.w_wide &&
               !(P.xshift[0] | P.xshift[1] | P.xshift[2] | P.xshift[3] | P.clamp_taps[0] | P.clamp_taps[1] |
                 P.clamp_taps[2] | P.clamp_taps[3]);
  return 0;
}
}  // namespace b200itp
