/*
 * TEST INFRASTRUCTURE — NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this library.
 *
 * CPU oracle: a literal restatement of the B-spline hot path of
 * JuliaMath/Interpolations.jl v0.16.3 (the reference cannot run here: no Julia).
 * Each function cites the reference file:line it follows (paths relative to the
 * reference checkout).  Built with -ffp-contract=off: Julia never fuses a*b+c.
 *
 * Pinning: the reference ships no golden vectors for this path; the oracle is
 * pinned by the reference's known-answer doc examples (docs/src/devdocs.md:96-190:
 * value 8.7, weights, gradient (1.0, 3.000000000000001, 9.0)) and by every
 * property its tests assert (tests/test_oracle_*.py), plus independent SciPy
 * splines.  PARITY UNPINNED at the last ulp of the third-party Woodbury solve.
 *
 * Floating-point types.  A coordinate may be Float32 or Float64 and Julia widens
 * it when it meets a Float64 bound/range (`maxp/minp`, `clamp`, arithmetic
 * promotion).  The general evaluator below carries every scalar as a double plus
 * a type tag and rounds to Float32 after each operation when the tag says so.
 * This is exact: for +,-,*,/ computing in binary64 and rounding to binary32 gives
 * the correctly rounded binary32 result (53 >= 2*24+2), floor/ceil/fmod are exact.
 */
#include "../include/b200itp.h"

#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------ prefilter */
#define REAL float
#define NAME(x) x##_f32
#include "prefilter_impl.h"
#undef REAL
#undef NAME
#define REAL double
#define NAME(x) x##_f64
#include "prefilter_impl.h"
#undef REAL
#undef NAME

int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* prefiltering_system + lut! + Woodbury Cp; see b200itp_prefiltering_system for the layout */
int orc_prefiltering_system(int dtype, int64_t n, int degree, int bc, int grid, void *dl, void *d,
                            void *du, int32_t *k, int32_t *urow, int32_t *vcol, void *Cp) {
  int kk = 0, ur[B200ITP_MAX_WOODBURY], vc[B200ITP_MAX_WOODBURY], rc;
  if (dtype == B200ITP_F32) {
    float cv[B200ITP_MAX_WOODBURY];
    rc = orc_system_f32(n, degree, bc, grid, (float *)dl, (float *)d, (float *)du, &kk, ur, vc, cv);
    if (!rc) rc = orc_woodbury_cp_f32(n, (float *)dl, (float *)d, (float *)du, kk, ur, vc, cv, (float *)Cp);
  } else {
    double cv[B200ITP_MAX_WOODBURY];
    rc = orc_system_f64(n, degree, bc, grid, (double *)dl, (double *)d, (double *)du, &kk, ur, vc, cv);
    if (!rc) rc = orc_woodbury_cp_f64(n, (double *)dl, (double *)d, (double *)du, kk, ur, vc, cv, (double *)Cp);
  }
  *k = kk;
  for (int j = 0; j < kk; j++) {
    urow[j] = ur[j];
    vcol[j] = vc[j];
  }
  return rc;
}

/* interpolate(A, it)'s prefilter (prefiltering.jl:32-38): pad + filter. `A` has size
 * `size_in`; `coefs` has size_in[d] + 2*pad[d] where pad = padded_axis rule
 * (cubic.jl:86-87, quadratic.jl:64-65, linear.jl:60). */
int orc_prefilter(int dtype, int ndims, const int64_t *size_in, const void *A, const int32_t *degree,
                  const int32_t *bc, const int32_t *grid, void *coefs, int nthreads) {
  int pad[B200ITP_MAX_DIMS], deg[B200ITP_MAX_DIMS], b[B200ITP_MAX_DIMS], g[B200ITP_MAX_DIMS];
  int64_t psz[B200ITP_MAX_DIMS];
  for (int i = 0; i < ndims; i++) {
    deg[i] = degree[i];
    b[i] = bc[i];
    g[i] = grid[i];
    pad[i] = ((deg[i] == B200ITP_CUBIC || deg[i] == B200ITP_QUADRATIC) && b[i] != B200ITP_BC_PERIODIC &&
              !B200ITP_IS_INPLACE_BC(b[i])) ? 1 : 0; /* cubic.jl:86-87, quadratic.jl:64-65 */
    psz[i] = size_in[i] + 2 * pad[i];
  }
  if (dtype == B200ITP_F32) {
    orc_pad_f32((const float *)A, (float *)coefs, ndims, size_in, pad);
    return orc_prefilter_inplace_f32((float *)coefs, ndims, psz, deg, b, g, nthreads);
  }
  orc_pad_f64((const double *)A, (double *)coefs, ndims, size_in, pad);
  return orc_prefilter_inplace_f64((double *)coefs, ndims, psz, deg, b, g, nthreads);
}

/* prefilter!(TW, ret, it) on an already padded array */
int orc_prefilter_inplace(int dtype, int ndims, const int64_t *size, void *coefs, const int32_t *degree,
                          const int32_t *bc, const int32_t *grid, int nthreads) {
  int deg[B200ITP_MAX_DIMS], b[B200ITP_MAX_DIMS], g[B200ITP_MAX_DIMS];
  for (int i = 0; i < ndims; i++) {
    deg[i] = degree[i];
    b[i] = bc[i];
    g[i] = grid[i];
  }
  if (dtype == B200ITP_F32) return orc_prefilter_inplace_f32((float *)coefs, ndims, size, deg, b, g, nthreads);
  return orc_prefilter_inplace_f64((double *)coefs, ndims, size, deg, b, g, nthreads);
}

/* --------------------------------------------------------- tagged arithmetic */
typedef struct {
  double v;
  int t; /* B200ITP_TAG_EXACT / F32 / F64 */
} tv;

static inline double rnd(int t, double v) { return t == B200ITP_TAG_F32 ? (double)(float)v : v; }
static inline tv mk(double v, int t) {
  tv r;
  r.v = v;
  r.t = t;
  return r;
}
static inline int tmax(int a, int b) { return a > b ? a : b; }
/* convert(T, x): rounds when an Int/Float64 lands in Float32 */
static inline tv conv(tv a, int t) { return mk(t > a.t ? rnd(t, a.v) : a.v, tmax(a.t, t)); }
static inline tv tv_add(tv a, tv b) {
  int t = tmax(a.t, b.t);
  return mk(rnd(t, conv(a, t).v + conv(b, t).v), t);
}
static inline tv tv_sub(tv a, tv b) {
  int t = tmax(a.t, b.t);
  return mk(rnd(t, conv(a, t).v - conv(b, t).v), t);
}
static inline tv tv_mul(tv a, tv b) {
  int t = tmax(a.t, b.t);
  return mk(rnd(t, conv(a, t).v * conv(b, t).v), t);
}
static inline tv tv_div(tv a, tv b) {
  int t = tmax(a.t, b.t);
  if (t == B200ITP_TAG_EXACT) t = B200ITP_TAG_F64; /* Int/Int -> Float64 */
  return mk(rnd(t, conv(a, t).v / conv(b, t).v), t);
}
/* Julia mod(x::T, y::T) for floats: r = rem(x,y); r==0 ? copysign(r,y) : (r>0) xor (y>0) ? r+y : r */
static inline tv tv_mod(tv a, tv b) {
  int t = tmax(a.t, b.t);
  double x = conv(a, t).v, y = conv(b, t).v;
  double r = fmod(x, y);
  if (r == 0)
    r = copysign(r, y);
  else if ((r > 0) != (y > 0))
    r = rnd(t, r + y);
  return mk(r, t);
}

/* ------------------------------------------------------------------ evaluation */
typedef struct {
  int L;          /* number of taps */
  int64_t pos[4]; /* 0-based position of each tap along this axis of the coefficient array */
  int base;       /* first tap in the reference's (1-based parent) index space */
  int edge;       /* the degree's edge rule fired */
  tv wv[4], wg[4], wh[4]; /* value / gradient / hessian weights */
} axis_taps;

/* modrange(x, 1:n) = mod(x - 1, n) + 1   (utils.jl:7) */
static inline int64_t modrange(int64_t x, int64_t n) {
  int64_t m = (x - 1) % n;
  if (m < 0) m += n;
  return m + 1;
}

/* positions + value_weights + gradient_weights for one axis.
 * x: coordinate in the interpolant's index space (tag F32/F64). */
static void axis_setup(const b200itp_desc *D, int d, tv x, axis_taps *A) {
  const int t = x.t;
  const int64_t n = D->parent_len[d];
  const int periodic = D->bc[d] == B200ITP_BC_PERIODIC;
  const int deg = D->degree[d];
  const tv one = mk(1.0, t);
  tv dx;
  int64_t xi;
  A->edge = 0;
  if (B200ITP_IS_INDEXLIKE(deg)) {
    /* constant.jl:69-108 with floorbounds / ceilbounds / roundbounds (indexing.jl:168-197), ax = 1:n;
     * NoInterp: weightedindex_parts = Int(x) (nointerp.jl:17) */
    double xm;
    if (deg == B200ITP_NOINTERP) {
      xm = x.v;
    } else if (deg == B200ITP_CONSTANT_PREVIOUS) {
      if (periodic)
        xm = floor(x.v); /* constant.jl:91-98 */
      else
        xm = x.v < 1.0 ? floor(rnd(t, x.v + 0.5)) : floor(rnd(t, x.v + 0.0));
    } else if (deg == B200ITP_CONSTANT_NEXT) {
      xm = x.v > (double)n ? ceil(rnd(t, x.v + 0.5)) : ceil(rnd(t, x.v + 0.0));
    } else {
      double xh = rnd(t, x.v + 0.5);
      xm = x.v < (double)n + 0.5 ? floor(xh) : rnd(t, ceil(xh) - 1.0);
    }
    xi = isnan(x.v) ? 1 : (int64_t)xm;
    if (periodic && (deg == B200ITP_CONSTANT_PREVIOUS || deg == B200ITP_CONSTANT_NEXT)) xi = modrange(xi, n); /* constant.jl:97,103 */
    A->L = 1;
    A->base = (int)xi;
    A->pos[0] = xi - D->coef_first[d];
    /* value_weights (1,), gradient_weights (0,), hessian_weights (0,): Ints (constant.jl:106-108) */
    A->wv[0] = mk(1.0, B200ITP_TAG_EXACT);
    A->wg[0] = mk(0.0, B200ITP_TAG_EXACT);
    A->wh[0] = mk(0.0, B200ITP_TAG_EXACT);
    return;
  }
  if (deg == B200ITP_CUBIC) {
    double xf;
    if (!periodic) {
      /* cubic.jl:40-46 ; floorbounds indexing.jl:183-187 : l = first(ax) = 1, h = half(x) */
      if (x.v < 1.0) {
        xf = floor(rnd(t, x.v + 0.5));
        A->edge = 1;
      } else {
        xf = floor(rnd(t, x.v + 0.0));
      }
      if (xf > (double)(n - 1)) { /* xf -= ifelse(xf > last(ax)-1, 1, 0) */
        xf = rnd(t, xf - 1.0);
        A->edge = 1;
      }
    } else {
      xf = floor(x.v); /* cubic.jl:50-57 */
    }
    dx = mk(rnd(t, x.v - xf), t);
    xi = isnan(x.v) ? 1 : (int64_t)xf; /* fast_trunc, utils.jl:40 (NaN is undefined there: node 1, result NaN) */
    A->L = 4;
    for (int j = 0; j < 4; j++) { /* expand_index cubic.jl:59-61 */
      int64_t idx = xi - 1 + j;
      if (periodic) idx = modrange(idx, n);
      if (j == 0) A->base = (int)idx;
      A->pos[j] = idx - D->coef_first[d];
    }
    /* value_weights cubic.jl:63-69 */
    tv omd = tv_sub(one, dx);
    tv x3 = tv_mul(tv_mul(dx, dx), dx);
    tv xc3 = tv_mul(tv_mul(omd, omd), omd);
    tv r16 = mk(rnd(t, rnd(t, 1.0) / rnd(t, 6.0)), t), r23 = mk(rnd(t, 2.0 / 3.0), t);
    tv r12 = mk(0.5, t), r32 = mk(1.5, t);
    A->wv[0] = tv_mul(r16, xc3);
    A->wv[1] = tv_add(tv_sub(r23, tv_mul(dx, dx)), tv_mul(r12, x3));
    A->wv[2] = tv_add(tv_sub(r23, tv_mul(omd, omd)), tv_mul(r12, xc3));
    A->wv[3] = tv_mul(r16, x3);
    /* gradient_weights cubic.jl:71-77 */
    tv x2 = tv_mul(dx, dx), xc2 = tv_mul(omd, omd);
    A->wg[0] = tv_mul(mk(-0.5, t), xc2);
    A->wg[1] = tv_add(tv_mul(mk(-2.0, t), dx), tv_mul(r32, x2));
    A->wg[2] = tv_sub(tv_mul(mk(2.0, t), omd), tv_mul(r32, xc2));
    A->wg[3] = tv_mul(r12, x2);
    /* hessian_weights cubic.jl:79: (1-dx, 3*dx-2, 3*(1-dx)-2, dx) */
    A->wh[0] = omd;
    A->wh[1] = tv_sub(tv_mul(mk(3.0, t), dx), mk(2.0, t));
    A->wh[2] = tv_sub(tv_mul(mk(3.0, t), omd), mk(2.0, t));
    A->wh[3] = dx;
  } else if (deg == B200ITP_QUADRATIC) {
    /* quadratic.jl:39-43 ; roundbounds indexing.jl:172-177: ifelse(x < u + half(u), floor(x+h), ceil(x+h)-1) */
    double xh = rnd(t, x.v + 0.5), xm;
    if (x.v < (double)n + 0.5) {
      xm = floor(xh);
    } else {
      xm = rnd(t, ceil(xh) - 1.0);
      A->edge = 1;
    }
    dx = mk(rnd(t, x.v - xm), t);
    xi = isnan(x.v) ? 1 : (int64_t)xm;
    A->L = 3;
    for (int j = 0; j < 3; j++) { /* quadratic.jl:57-62 */
      int64_t idx = xi - 1 + j;
      if (periodic) idx = modrange(idx, n);
      if (B200ITP_IS_INPLACE_BC(D->bc[d])) idx = idx < 1 ? 1 : (idx > n ? n : idx); /* (max(xi-1, first), xi, min(xi+1, last)) */
      if (j == 0) A->base = (int)idx;
      A->pos[j] = idx - D->coef_first[d];
    }
    /* quadratic.jl:45-53 */
    tv half = mk(0.5, t), r34 = mk(0.75, t), two = mk(2.0, t);
    tv a = tv_sub(dx, half), b = tv_add(dx, half);
    A->wv[0] = tv_div(tv_mul(a, a), two);
    A->wv[1] = tv_sub(r34, tv_mul(dx, dx));
    A->wv[2] = tv_div(tv_mul(b, b), two);
    A->wg[0] = a;
    A->wg[1] = tv_mul(mk(-2.0, t), dx);
    A->wg[2] = b;
    /* hessian_weights quadratic.jl:55: (1, -2, 1) of dx's type */
    A->wh[0] = mk(1.0, t);
    A->wh[1] = mk(-2.0, t);
    A->wh[2] = mk(1.0, t);
  } else { /* linear.jl:43-58 */
    double f = floor(x.v);
    if (x.v == (double)n) {
      f = rnd(t, f - 1.0);
      A->edge = 1;
    }
    xi = isnan(x.v) ? 1 : (int64_t)f;
    dx = mk(rnd(t, x.v - f), t);
    A->L = 2;
    for (int j = 0; j < 2; j++) {
      int64_t idx = xi + j;
      if (periodic) idx = modrange(idx, n);
      if (j == 0) A->base = (int)idx;
      A->pos[j] = idx - D->coef_first[d];
    }
    A->wv[0] = tv_sub(one, dx);
    A->wv[1] = dx;
    A->wg[0] = mk(-1.0, t);
    A->wg[1] = mk(1.0, t);
    A->wh[0] = mk(0.0, t); /* linear.jl:58 */
    A->wh[1] = mk(0.0, t);
  }
}

typedef struct {
  const b200itp_desc *D;
  const void *coefs;
  int64_t stride[4];
  const axis_taps *ax;
  const tv *w[4]; /* weights to use on each axis */
  int lt[5];      /* result type of each nesting level */
} reduce_ctx;

/* interp_getindex (Interpolations.jl:304-316): axis 1 outermost, taps accumulated as
 * w_k*v_k + (previous), first tap is a bare product (`_weight_itp(w,i,::Zero) = w*i`). */
static double reduce_level(const reduce_ctx *C, int d, int64_t off) {
  const int N = C->D->ndims;
  const axis_taps *A = &C->ax[d];
  const int t = C->lt[d];
  double acc = 0;
  for (int k = 0; k < A->L; k++) {
    int64_t o = off + A->pos[k] * C->stride[d];
    double v;
    if (d == N - 1)
      v = C->D->coef_dtype == B200ITP_F32 ? (double)((const float *)C->coefs)[o] : ((const double *)C->coefs)[o];
    else
      v = reduce_level(C, d + 1, o);
    double term = rnd(t, C->w[d][k].v * v);
    acc = (k == 0) ? term : rnd(t, term + acc);
  }
  return acc;
}

/* value and/or gradient of the bare BSplineInterpolation at x (already in bounds).
 * indexing.jl:5-9 and :25-31 (slot_substitute :93-112). */
static void bspline_eval(const b200itp_desc *D, const void *coefs, const tv *x, int want_val, int want_grad,
                         tv *val, tv *grad, int32_t *base, uint32_t *edge_bits) {
  axis_taps ax[4];
  reduce_ctx C;
  const int N = D->ndims;
  C.D = D;
  C.coefs = coefs;
  C.ax = ax;
  int64_t s = 1;
  /* interpolate! (b-splines.jl:236-241): parent axes 2:n-1 over coefficients 1:n are the padded layout one index
   * further; x - 1 is exact, so positions and weights are those of the reference */
  b200itp_desc DD = *D;
  for (int d = 0; d < N; d++) {
    const int sh = D->parent_first[d] == 2;
    DD.coef_first[d] -= sh;
    axis_setup(&DD, d, mk(x[d].v - sh, x[d].t), &ax[d]);
    ax[d].base += sh;
    C.stride[d] = s;
    s *= D->size[d];
    if (base) base[d] = ax[d].base;
    if (edge_bits && ax[d].edge) *edge_bits |= (B200ITP_BR_EDGE << (4 * d));
  }
  C.lt[N] = D->coef_dtype == B200ITP_F32 ? B200ITP_TAG_F32 : B200ITP_TAG_F64;
  for (int d = N - 1; d >= 0; d--) /* Constant axes: Int weights do not widen the partial sums */
    C.lt[d] = tmax(B200ITP_IS_INDEXLIKE(D->degree[d]) ? B200ITP_TAG_EXACT : x[d].t, C.lt[d + 1]);
  if (want_val) {
    for (int d = 0; d < N; d++) C.w[d] = ax[d].wv;
    *val = mk(reduce_level(&C, 0, 0), C.lt[0]);
  }
  if (want_grad) {
    for (int g = 0; g < N; g++) {
      for (int d = 0; d < N; d++) C.w[d] = (d == g) ? ax[d].wg : ax[d].wv;
      grad[g] = mk(reduce_level(&C, 0, 0), C.lt[0]);
    }
  }
}

/* itp below the extrapolation wrapper: bare BSpline, or ScaledInterpolation called with
 * @inbounds (scaling.jl:78-83 / :123-128: coordlookup, maybe_clamp, parent call, rescale). */
static void inner_eval(const b200itp_desc *D, const void *coefs, const tv *x, int want_val, int want_grad,
                       tv *val, tv *grad, int32_t *base, uint32_t *edge_bits) {
  const int N = D->ndims;
  tv xl[4];
  for (int d = 0; d < N; d++) {
    if (D->has_scale) {
      /* coordlookup(r, x) = (x - first(r))/step(r) + one(eltype(r))   scaling.jl:102-115 */
      tv f = mk(D->range_first[d], D->range_tag[d]), st = mk(D->range_step[d], D->range_tag[d]);
      tv one = mk(1.0, D->range_tag[d]);
      tv q = tv_add(tv_div(tv_sub(x[d], f), st), one);
      /* maybe_clamp -> clamp(x, lo, hi) with promotion (Interpolations.jl:450-452) */
      int t = tmax(q.t, D->inner_tag[d]);
      double lo = D->inner_lb[d], hi = D->inner_ub[d];
      double c = q.v > hi ? rnd(t, hi) : (q.v < lo ? rnd(t, lo) : q.v);
      xl[d] = mk(c, t);
    } else {
      xl[d] = x[d];
    }
  }
  bspline_eval(D, coefs, xl, want_val, want_grad, val, grad, base, edge_bits);
  if (want_grad && D->has_scale)
    for (int d = 0; d < N; d++) /* rescale_gradient scaling.jl:116,137 */
      grad[d] = tv_div(grad[d], mk(D->range_step[d], D->range_tag[d]));
}

/* periodic / reflect (extrapolation.jl:158-163) with l,u of type btag */
static tv etp_periodic(tv y, double l, double u, int bt) {
  tv L = mk(l, bt), span = mk(rnd(bt, u - l), bt);
  return tv_add(tv_mod(tv_sub(y, L), span), L);
}
static tv etp_reflect(tv y, double l, double u, int bt) {
  tv L = mk(l, bt), span2 = mk(rnd(bt, 2.0 * rnd(bt, u - l)), bt), U2 = mk(rnd(bt, 2.0 * u), bt);
  tv yr = tv_add(tv_mod(tv_sub(y, L), span2), L);
  if (yr.v > u) return tv_sub(U2, yr);
  return yr;
}

/* inbounds_index (extrapolation.jl:106-151); returns 1 if BoundsError */
static int etp_inbounds_index(int flo, int fhi, int pair, double l, double u, int bt, tv x, tv *out) {
  if (!pair) { /* a single flag: one-sided specialisations :110-116 */
    if (flo == B200ITP_BC_THROW) {
      *out = x;
      return !(l <= x.v && x.v <= u);
    }
    if (flo == B200ITP_BC_PERIODIC) {
      *out = etp_periodic(x, l, u, bt);
      return 0;
    }
    if (flo == B200ITP_BC_REFLECT) {
      *out = etp_reflect(x, l, u, bt);
      return 0;
    }
  }
  /* left then right :118-151 */
  switch (flo) {
    case B200ITP_BC_THROW:
      if (!(l <= x.v)) {
        *out = x;
        return 1;
      }
      break;
    case B200ITP_BC_FLAT:
    case B200ITP_BC_LINE: { /* maxp(l, x) */
      int t = tmax(x.t, bt);
      double a = rnd(t, l), b = conv(x, t).v;
      x = mk(b < a ? a : b, t); /* max(l, x): NaN x propagates */
      break;
    }
    case B200ITP_BC_PERIODIC: x = etp_periodic(x, l, u, bt); break;
    case B200ITP_BC_REFLECT: x = etp_reflect(x, l, u, bt); break;
    default: break;
  }
  switch (fhi) {
    case B200ITP_BC_THROW:
      if (!(x.v <= u)) {
        *out = x;
        return 1;
      }
      break;
    case B200ITP_BC_FLAT:
    case B200ITP_BC_LINE: { /* minp(x, u) */
      int t = tmax(x.t, bt);
      double a = conv(x, t).v, b = rnd(t, u);
      x = mk(a > b ? b : a, t); /* min(x, u): NaN x propagates */
      break;
    }
    case B200ITP_BC_PERIODIC: x = etp_periodic(x, l, u, bt); break;
    case B200ITP_BC_REFLECT: x = etp_reflect(x, l, u, bt); break;
    default: break;
  }
  *out = x;
  return 0;
}

/* One query.  Outputs are doubles already rounded to their Julia type.
 * returns 1 when the reference would throw BoundsError. */
static int eval_one(const b200itp_desc *D, const void *coefs, const double *xin, int want_val, int want_grad,
                    double *oval, double *ograd, int32_t *base, uint32_t *branch) {
  const int N = D->ndims;
  const int xt = D->coord_dtype == B200ITP_F32 ? B200ITP_TAG_F32 : B200ITP_TAG_F64;
  const double *lb = D->has_scale ? D->outer_lb : D->inner_lb;
  const double *ub = D->has_scale ? D->outer_ub : D->inner_ub;
  const int32_t *bt = D->has_scale ? D->outer_tag : D->inner_tag;
  tv x[4], xs[4], val = mk(0, xt), g[4];
  uint32_t br = 0;
  int inb = 1;
  for (int d = 0; d < N; d++) {
    x[d] = mk(xin[d], xt);
    if (!(lb[d] <= xin[d])) {
      br |= B200ITP_BR_BELOW << (4 * d);
      inb = 0;
    }
    if (!(xin[d] <= ub[d])) {
      br |= B200ITP_BR_ABOVE << (4 * d);
      inb = 0;
    }
    if (D->degree[d] == B200ITP_NOINTERP && !(xin[d] == floor(xin[d]))) inb = 0; /* Int(x): InexactError */
    g[d] = mk(0, xt);
  }
  const double nan = NAN;
  int threw = 0;
  if (D->etp_kind == B200ITP_ETP_NONE) {
    /* indexing.jl:5-9,25-31 / scaling.jl:78-83,123-128 with @boundscheck */
    if (!inb) {
      threw = 1;
    } else {
      inner_eval(D, coefs, x, want_val, want_grad, &val, g, base, &br);
    }
  } else if (D->etp_kind == B200ITP_ETP_FILL) {
    /* filled.jl:31-40,58-67 */
    if (inb) {
      inner_eval(D, coefs, x, want_val, want_grad, &val, g, base, &br);
    } else {
      br |= B200ITP_BR_FILLED;
      /* Tret = promote_type(eltype(coefs), typeof.(x)..., T) */
      int t = tmax(tmax(xt, D->coef_dtype == B200ITP_F32 ? B200ITP_TAG_F32 : B200ITP_TAG_F64), D->fill_tag);
      val = mk(rnd(t, D->fill_value), t);
      for (int d = 0; d < N; d++) g[d] = mk(0.0, t);
      if (base)
        for (int d = 0; d < N; d++) base[d] = INT32_MIN;
    }
  } else {
    /* extrapolation.jl:50-58 (value) and :71-82 (gradient) */
    int moved = 0;
    for (int d = 0; d < N; d++) {
      int pair = (D->etp_pair_mask >> d) & 1;
      if (etp_inbounds_index(D->etp_lo[d], D->etp_hi[d], pair, lb[d], ub[d], bt[d], x[d], &xs[d])) threw = 1;
      if (xs[d].v != x[d].v) {
        moved = 1;
        br |= B200ITP_BR_MOVED << (4 * d);
      }
    }
    if (!threw) {
      if (want_val) {
        tv gg[4];
        if (!moved) {
          inner_eval(D, coefs, xs, 1, 0, &val, NULL, base, &br);
        } else {
          br |= B200ITP_BR_SLOWPATH;
          inner_eval(D, coefs, xs, 1, 1, &val, gg, base, &br);
          /* extrapolate_value / extrapolate_axis :166-182 */
          for (int d = 0; d < N; d++) {
            int pair = (D->etp_pair_mask >> d) & 1;
            tv delta = tv_add(val, tv_mul(tv_sub(x[d], xs[d]), gg[d]));
            if (!pair) {
              if (D->etp_lo[d] == B200ITP_BC_LINE) val = delta;
            } else {
              if (D->etp_lo[d] == B200ITP_BC_LINE && x[d].v < xs[d].v) val = delta;
              if (D->etp_hi[d] == B200ITP_BC_LINE) {
                tv delta2 = tv_add(val, tv_mul(tv_sub(x[d], xs[d]), gg[d]));
                if (x[d].v > xs[d].v) val = delta2;
              }
            }
          }
        }
      }
      if (want_grad) {
        if (inb) {
          inner_eval(D, coefs, x, 0, 1, NULL, g, want_val ? NULL : base, want_val ? NULL : &br);
        } else {
          inner_eval(D, coefs, xs, 0, 1, NULL, g, want_val ? NULL : base, want_val ? NULL : &br);
          for (int d = 0; d < N; d++) /* extrapolate_gradient :184-185 */
            if (D->etp_lo[d] == B200ITP_BC_FLAT && xs[d].v != x[d].v) g[d] = mk(0.0, g[d].t);
        }
      }
    }
  }
  if (threw) {
    br |= B200ITP_BR_THROW;
    if (want_val) *oval = nan;
    if (want_grad)
      for (int d = 0; d < N; d++) ograd[d] = nan;
    if (base)
      for (int d = 0; d < N; d++) base[d] = INT32_MIN;
  } else {
    if (want_val) *oval = val.v;
    if (want_grad)
      for (int d = 0; d < N; d++) ograd[d] = g[d].v;
  }
  if (branch) *branch = br;
  return threw;
}

/*
 * itp.(xs...) and gradient.(Ref(itp), xs...) : the serial broadcast loop of
 * src/gpu_support.jl:52-57 over `nq` queries (OpenMP over queries when nthreads > 1).
 * coords: ndims host pointers of coord_dtype.  out_value[nq], out_grad[nq*ndims] are doubles
 * holding values already rounded to their Julia type.  Per-side extrapolation pairs are
 * flagged in D->etp_pair_mask bit d.
 */
int orc_eval(const b200itp_desc *D, const void *coefs, const void *const *coords, int64_t nq,
             double *out_value, double *out_grad, int64_t *first_oob, int32_t *dbg_index,
             uint32_t *dbg_branch, int nthreads) {
  const int N = D->ndims;
  int64_t foob = -1;
  if (N < 1 || N > 4) return B200ITP_E_BADARG;
  if (out_grad && D->etp_kind == B200ITP_ETP_FLAGS && D->etp_pair_mask) return B200ITP_E_UNSUPPORTED; /* MethodError in the reference */
#pragma omp parallel for schedule(static) num_threads(nthreads > 0 ? nthreads : 1)
  for (int64_t q = 0; q < nq; q++) {
    double x[4], v = 0, g[4] = {0, 0, 0, 0};
    for (int d = 0; d < N; d++)
      x[d] = D->coord_dtype == B200ITP_F32 ? (double)((const float *)coords[d])[q] : ((const double *)coords[d])[q];
    int32_t base[4];
    uint32_t br;
    int threw = eval_one(D, coefs, x, out_value != NULL, out_grad != NULL, &v, g, dbg_index ? base : NULL, &br);
    if (out_value) out_value[q] = v;
    if (out_grad) { /* NoInterp axes have no gradient component (indexing.jl:108) */
      int G = 0, k = 0;
      for (int d = 0; d < N; d++) G += D->degree[d] != B200ITP_NOINTERP;
      for (int d = 0; d < N; d++)
        if (D->degree[d] != B200ITP_NOINTERP) out_grad[q * G + k++] = g[d];
    }
    if (dbg_index)
      for (int d = 0; d < N; d++) dbg_index[q * N + d] = base[d];
    if (dbg_branch) dbg_branch[q] = br;
    if (threw) {
#pragma omp critical
      if (foob < 0 || q < foob) foob = q;
    }
  }
  if (first_oob) *first_oob = foob;
  return 0;
}

/* hessian(itp, x...) (indexing.jl:37-41): components in the order of `weightedindexes` (indexing.jl:114-144):
 * (1,1), (1,2), ..., (1,N), (2,2), ..., (N,N); axis i and j carry gradient weights (hessian weights when i == j),
 * the others value weights; each component is its own interp_getindex reduction.  Scaled: scaling.jl:150-159,
 * h ./ (steps .* steps').  Output: the full symmetric N x N matrix (symmatrix, indexing.jl:201-229). */
static void inner_hess(const b200itp_desc *D, const void *coefs, const tv *x, tv *h /* N*N */) {
  const int N = D->ndims;
  tv xl[4];
  for (int d = 0; d < N; d++) {
    if (D->has_scale) {
      tv f = mk(D->range_first[d], D->range_tag[d]), st = mk(D->range_step[d], D->range_tag[d]);
      tv one = mk(1.0, D->range_tag[d]);
      tv q = tv_add(tv_div(tv_sub(x[d], f), st), one);
      int t = tmax(q.t, D->inner_tag[d]);
      double lo = D->inner_lb[d], hi = D->inner_ub[d];
      double c = q.v > hi ? rnd(t, hi) : (q.v < lo ? rnd(t, lo) : q.v);
      xl[d] = mk(c, t);
    } else {
      xl[d] = x[d];
    }
  }
  axis_taps ax[4];
  reduce_ctx C;
  C.D = D;
  C.coefs = coefs;
  C.ax = ax;
  int64_t s = 1;
  b200itp_desc DD = *D;
  for (int d = 0; d < N; d++) {
    const int sh = D->parent_first[d] == 2; /* interpolate!: see bspline_eval */
    DD.coef_first[d] -= sh;
    axis_setup(&DD, d, mk(xl[d].v - sh, xl[d].t), &ax[d]);
    C.stride[d] = s;
    s *= D->size[d];
  }
  C.lt[N] = D->coef_dtype == B200ITP_F32 ? B200ITP_TAG_F32 : B200ITP_TAG_F64;
  for (int d = N - 1; d >= 0; d--)
    C.lt[d] = tmax(B200ITP_IS_INDEXLIKE(D->degree[d]) ? B200ITP_TAG_EXACT : xl[d].t, C.lt[d + 1]);
  for (int i = 0; i < N; i++)
    for (int j = i; j < N; j++) {
      for (int d = 0; d < N; d++)
        C.w[d] = (d == i && d == j) ? ax[d].wh : ((d == i || d == j) ? ax[d].wg : ax[d].wv);
      tv v = mk(reduce_level(&C, 0, 0), C.lt[0]);
      if (D->has_scale) {
        tv si = mk(D->range_step[i], D->range_tag[i]), sj = mk(D->range_step[j], D->range_tag[j]);
        v = tv_div(v, tv_mul(si, sj));
      }
      h[i * N + j] = v;
      h[j * N + i] = v;
    }
}

/* hessian.(Ref(itp), xs...): out_hess[nq * N * N] doubles already rounded to their Julia type.  Extrapolation
 * wrappers only evaluate in bounds (extrapolation.jl:84-90: "extrapolation of hessian not yet implemented" is
 * reported like a BoundsError); FilledExtrapolation has no hessian method. */
int orc_hessian(const b200itp_desc *D, const void *coefs, const void *const *coords, int64_t nq, double *out_hess,
                int64_t *first_oob, int nthreads) {
  const int N = D->ndims;
  int64_t foob = -1;
  if (N < 1 || N > 4) return B200ITP_E_BADARG;
  if (D->etp_kind == B200ITP_ETP_FILL) return B200ITP_E_UNSUPPORTED;
  const int xt = D->coord_dtype == B200ITP_F32 ? B200ITP_TAG_F32 : B200ITP_TAG_F64;
  const double *lb = D->has_scale ? D->outer_lb : D->inner_lb;
  const double *ub = D->has_scale ? D->outer_ub : D->inner_ub;
#pragma omp parallel for schedule(static) num_threads(nthreads > 0 ? nthreads : 1)
  for (int64_t q = 0; q < nq; q++) {
    tv x[4], h[16];
    int inb = 1;
    for (int d = 0; d < N; d++) {
      const double xv = D->coord_dtype == B200ITP_F32 ? (double)((const float *)coords[d])[q] : ((const double *)coords[d])[q];
      x[d] = mk(xv, xt);
      if (!(lb[d] <= xv) || !(xv <= ub[d])) inb = 0;
    }
    if (inb) {
      inner_hess(D, coefs, x, h);
      for (int k = 0; k < N * N; k++) out_hess[q * N * N + k] = h[k].v;
    } else {
      for (int k = 0; k < N * N; k++) out_hess[q * N * N + k] = NAN;
#pragma omp critical
      if (foob < 0 || q < foob) foob = q;
    }
  }
  if (first_oob) *first_oob = foob;
  return 0;
}

/* value_weights / gradient_weights for one axis at offset dx (cubic.jl:63-77, quadratic.jl:45-53,
 * linear.jl:56-57); used by the known-answer tests (Interpolations.jl:326-329). */
int orc_weights(int degree, int tag, double x, int64_t n, double *wv, double *wg, int32_t *base) {
  b200itp_desc D;
  memset(&D, 0, sizeof D);
  D.ndims = 1;
  D.degree[0] = degree;
  D.parent_len[0] = n;
  D.coef_first[0] = (degree == B200ITP_LINEAR || B200ITP_IS_INDEXLIKE(degree)) ? 1 : 0;
  axis_taps A;
  axis_setup(&D, 0, mk(rnd(tag, x), tag), &A);
  for (int j = 0; j < A.L; j++) {
    wv[j] = A.wv[j].v;
    wg[j] = A.wg[j].v;
  }
  *base = A.base;
  return A.L;
}

/* ------------------------------------------------------------------ type-specialised evaluator (timing legs)
 * What Julia compiles for `itp.(xs, ys, zs)` with itp = interpolate(A::Array{Float32,3}, BSpline(Cubic(Periodic(OnGrid())))):
 * every type is known, so the loop body is straight-line Float32 code — no tags, no recursion.  Same operations in the
 * same order as the interpreter above (cubic.jl:50-69 positions/weights, Interpolations.jl:304-316 reduction);
 * tests/test_oracle_pinning.py checks it against the interpreter bit for bit.  Used by bench.py's CPU legs only: the
 * interpreter (1.5 Mevals/s/thread) would understate what the reference does on a CPU. */
static inline void cubic_periodic_axis_f32(float x, int64_t n, int64_t *idx, float *w) {
  const float xf = floorf(x);                 /* cubic.jl:50-57 */
  const float dx = x - xf;
  const int64_t xi = (int64_t)xf;
  for (int j = 0; j < 4; j++) idx[j] = modrange(xi - 1 + j, n) - 1; /* expand_index cubic.jl:59-61, 0-based */
  const float omd = 1.0f - dx;                /* value_weights cubic.jl:63-69 */
  const float x3 = (dx * dx) * dx, xc3 = (omd * omd) * omd;
  const float r16 = 1.0f / 6.0f, r23 = (float)(2.0 / 3.0);
  w[0] = r16 * xc3;
  w[1] = (r23 - dx * dx) + 0.5f * x3;
  w[2] = (r23 - omd * omd) + 0.5f * xc3;
  w[3] = r16 * x3;
}

int orc_eval_tricubic_periodic_f32(const float *coefs, int64_t n0, int64_t n1, int64_t n2, const float *xs,
                                   const float *ys, const float *zs, int64_t nq, float *out, int nthreads) {
#pragma omp parallel for schedule(static) num_threads(nthreads > 0 ? nthreads : 1)
  for (int64_t q = 0; q < nq; q++) {
    int64_t i0[4], i1[4], i2[4];
    float w0[4], w1[4], w2[4];
    cubic_periodic_axis_f32(xs[q], n0, i0, w0);
    cubic_periodic_axis_f32(ys[q], n1, i1, w1);
    cubic_periodic_axis_f32(zs[q], n2, i2, w2);
    float val = 0.f;
    for (int a = 0; a < 4; a++) {             /* axis 1 outermost; first tap a bare product, then w*v + acc */
      float v0 = 0.f;
      for (int b = 0; b < 4; b++) {
        float v1 = 0.f;
        for (int c = 0; c < 4; c++) {
          const float t = w2[c] * coefs[(i2[c] * n1 + i1[b]) * n0 + i0[a]];
          v1 = c == 0 ? t : t + v1;
        }
        const float t = w1[b] * v1;
        v0 = b == 0 ? t : t + v0;
      }
      const float t = w0[a] * v0;
      val = a == 0 ? t : t + val;
    }
    out[q] = val;
  }
  return 0;
}
