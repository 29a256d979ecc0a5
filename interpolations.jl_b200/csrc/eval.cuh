// Evaluation of `extrapolate(scale(interpolate(A, BSpline(...)), ranges...), et)` at scattered
// query points: `itp.(xs...)` and `Interpolations.gradient.(Ref(itp), xs...)`.
//
// One thread per query computes, entirely in registers, what the reference computes per element of
// the broadcast (paths relative to the Interpolations.jl v0.16.3 checkout):
//   extrapolation  src/extrapolation/extrapolation.jl:50-58,71-82,106-185, filled.jl:31-40,58-67
//   scaling        src/scaling/scaling.jl:78-83,102-141 ; clamp src/Interpolations.jl:450-452
//   positions      src/b-splines/cubic.jl:40-61, quadratic.jl:39-43,57-60, linear.jl:43-54,
//                  src/b-splines/indexing.jl:168-199, src/utils.jl:7,40
//   weights        cubic.jl:63-77, quadratic.jl:45-53, linear.jl:56-57
//   reduction      src/Interpolations.jl:304-316 (axis 1 outermost, taps accumulated w*v + acc)
// The arithmetic follows the reference operation by operation in the reference's types (this file is
// compiled with -fmad=false: Julia never contracts a*b+c), so stencil indices, branch choices AND
// values are bit-identical to the CPU path given the same coefficients.
#pragma once

#include "common.cuh"

namespace b200itp {

struct EvalParams {
  int ndims;
  int n[4];           // parent length per axis
  int coef_first[4];  // 0 padded, 1 unpadded
  int degree[4];
  int periodic[4];
  long long stride[4];  // element stride of each axis in the coefficient array
  long long total;      // number of coefficients
  int vec_ok;           // coefficient base pointer is 16-byte aligned
  // bounds seen by the outermost wrapper (scaled bounds if has_scale, else the B-spline's)
  double lb[4], ub[4];
  double span[4], span2[4], ub2[4];  // u-l, 2(u-l), 2u rounded in the bounds' own Julia type
  // scale
  int has_scale;
  double rfirst[4], rstep[4];
  double inner_lb[4], inner_ub[4];
  // where a Float32 coordinate becomes Float64 (only meaningful when TX=float, TW=double)
  int e_wide;  // extrapolation stage runs in TW
  int s_wide;  // coordlookup runs in TW
  int s_wide_x;  // same, for the in-bounds gradient path that bypasses the extrapolation mapping
  int w_wide_x;  // weights of that path are computed in TW (else in TX)
  // extrapolation
  int etp_kind;
  int etp_lo[4], etp_hi[4];
  int pair_mask;
  int any_line;
  double fill;
  int out_f64;  // output element type
  // buffers
  const void *coords[4];
  const void *coefs;
  void *out_value;
  void *out_grad;
  void *out_hess;     // nq * N * N, the full symmetric matrix per query
  int range_f32[4];   // eltype(ranges[d]) == Float32: steps[i] * steps[j] is rounded to Float32 (scaling.jl:157-160)
  int *dbg_index;
  unsigned *dbg_branch;
  unsigned long long *first_oob;
  long long nq;
  const int *perm;  // optional: query q reads coordinates / writes results at perm[q]
  // (new fields go last: the offsets of everything above are part of what the tuned kernels were compiled against)
  int grad_slot[4];  // position of the axis' component in the gradient (-1: NoInterp axis, no component)
  int ngrad;         // gradient components per query
  int xshift[4];     // first(parentaxes) - 1: subtracted from the coordinate before the B-spline stage (interpolate!)
  int clamp_taps[4]; // Quadratic(InPlace/InPlaceQ): outer taps clamped to the axis (quadratic.jl:61-62)
};

// Scattered coefficient gathers.  B200ITP_GATHER_VOLATILE=1 issues them as `ld.volatile.global` (no L1 allocation):
// measured on B200 (tools/ubench_gather.cu, profiles/r2_ubench_gather.txt), random 16-byte gathers over 2 GiB run at
// 32 G/s through `ld.global.nc` and 39 G/s volatile — the HBM-missing gather is bound by the ACCESS RATE (every miss
// fills a 128-byte line: 124 B of DRAM reads per 16-byte gather), not by bytes.
#ifndef B200ITP_GATHER_VOLATILE
#define B200ITP_GATHER_VOLATILE 0
#endif
template <typename T>
__device__ __forceinline__ T gather_ld(const T *p) {
#if B200ITP_GATHER_VOLATILE
  return *reinterpret_cast<const volatile T *>(p);
#else
  return __ldg(p);
#endif
}
__device__ __forceinline__ float4 gather_ld4(const float4 *p) {
#if B200ITP_GATHER_VOLATILE
  float4 r;
  asm volatile("ld.volatile.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
#else
  return __ldg(p);
#endif
}

template <typename A, typename B>
struct Promote {
  using type = double;
};
template <>
struct Promote<float, float> {
  using type = float;
};

template <typename S>
__device__ __forceinline__ S s_floor(S x) {
  return floor(x);
}
template <>
__device__ __forceinline__ float s_floor<float>(float x) {
  return floorf(x);
}
template <typename S>
__device__ __forceinline__ S s_ceil(S x) {
  return ceil(x);
}
template <>
__device__ __forceinline__ float s_ceil<float>(float x) {
  return ceilf(x);
}
template <typename S>
__device__ __forceinline__ S s_fmod(S a, S b) {
  return fmod(a, b);
}
template <>
__device__ __forceinline__ float s_fmod<float>(float a, float b) {
  return fmodf(a, b);
}

// Julia `mod(a, b)` for b > 0: rem, then shifted into [0, b)   (Base.mod for floats)
template <typename S>
__device__ __forceinline__ S julia_mod_pos(S a, S b) {
  S r = s_fmod<S>(a, b);
  if (r < S(0)) r = r + b;
  return r;
}

// periodic / reflect (extrapolation.jl:158-163)
template <typename S>
__device__ __forceinline__ S etp_periodic(const EvalParams &P, int d, S y) {
  const S l = (S)P.lb[d];
  return julia_mod_pos<S>(y - l, (S)P.span[d]) + l;
}
template <typename S>
__device__ __forceinline__ S etp_reflect(const EvalParams &P, int d, S y) {
  const S l = (S)P.lb[d], u = (S)P.ub[d];
  const S yr = julia_mod_pos<S>(y - l, (S)P.span2[d]) + l;
  return yr > u ? (S)P.ub2[d] - yr : yr;
}

// inbounds_index (extrapolation.jl:106-151) in type S; returns true when BoundsError
template <typename S>
__device__ __forceinline__ bool etp_inbounds(const EvalParams &P, int d, S x, S &xs) {
  const S l = (S)P.lb[d], u = (S)P.ub[d];
  const int flo = P.etp_lo[d], fhi = P.etp_hi[d];
  const bool pair = (P.pair_mask >> d) & 1;
  xs = x;
  if (!pair) {
    if (flo == B200ITP_BC_THROW) return !(l <= x && x <= u);
    if (flo == B200ITP_BC_PERIODIC) {
      xs = etp_periodic<S>(P, d, x);
      return false;
    }
    if (flo == B200ITP_BC_REFLECT) {
      xs = etp_reflect<S>(P, d, x);
      return false;
    }
  }
  if (flo == B200ITP_BC_THROW) {
    if (!(l <= x)) return true;
  } else if (flo == B200ITP_BC_FLAT || flo == B200ITP_BC_LINE) {
    x = x < l ? l : x;  // max(l, x); NaN propagates like Julia's max
  } else if (flo == B200ITP_BC_PERIODIC) {
    x = etp_periodic<S>(P, d, x);
  } else if (flo == B200ITP_BC_REFLECT) {
    x = etp_reflect<S>(P, d, x);
  }
  if (fhi == B200ITP_BC_THROW) {
    if (!(x <= u)) {
      xs = x;
      return true;
    }
  } else if (fhi == B200ITP_BC_FLAT || fhi == B200ITP_BC_LINE) {
    x = x > u ? u : x;  // min(x, u)
  } else if (fhi == B200ITP_BC_PERIODIC) {
    x = etp_periodic<S>(P, d, x);
  } else if (fhi == B200ITP_BC_REFLECT) {
    x = etp_reflect<S>(P, d, x);
  }
  xs = x;
  return false;
}

__device__ __forceinline__ int modrange1(int i, int n) {  // utils.jl:7 with r = 1:n
  int m = (i - 1) % n;
  if (m < 0) m += n;
  return m + 1;
}

// positions + weights of one axis.  LMAX taps; L = number actually used (LMAX unless DEG == 0).
template <int DEG, typename TW>
struct AxisTaps {
  static constexpr int LMAX = DEG == 0 ? 4 : DEG + 1;
  int pos[LMAX];  // 0-based position along the axis of the coefficient array
  TW wv[LMAX], wg[LMAX], wh[LMAX];  // value / gradient / hessian weights
  int L;
  int first;  // parent index (1-based, before any periodic wrap) of the first tap
  bool edge;
  bool contiguous;  // pos[k] == pos[0] + k
};

template <int DEG, typename TW, bool GRAD>
__device__ __forceinline__ void axis_setup(const EvalParams &P, int d, TW x, AxisTaps<DEG, TW> &A) {
  const int deg = DEG == 0 ? P.degree[d] : DEG;
  const int n = P.n[d];
  const bool per = P.periodic[d] != 0;
  const int cf = P.coef_first[d];
  A.edge = false;
  int first;  // index of the first tap in the parent index space
  // NaN coordinates (undefined in the reference: unsafe_trunc(Int, NaN)) must not fault: the taps are
  // taken at node 1 and the NaN weights make the result NaN.
  const bool isnan_x = !(x == x);
  if (DEG == 0 && B200ITP_IS_INDEXLIKE(deg)) {
    // constant.jl:69-108: floorbounds / ceilbounds / roundbounds (indexing.jl:168-197) with ax = 1:n;
    // NoInterp: Int(x) (nointerp.jl:17), x holds an integer (checked by the kernel)
    TW xm;
    if (deg == B200ITP_NOINTERP) {
      xm = x;
    } else if (deg == B200ITP_CONSTANT_PREVIOUS) {
      xm = per ? s_floor<TW>(x) : (x < TW(1) ? s_floor<TW>(x + TW(0.5)) : s_floor<TW>(x + TW(0)));
    } else if (deg == B200ITP_CONSTANT_NEXT) {
      xm = x > TW(n) ? s_ceil<TW>(x + TW(0.5)) : s_ceil<TW>(x + TW(0));
    } else {
      const TW xh = x + TW(0.5);
      xm = ((double)x < (double)n + 0.5) ? s_floor<TW>(xh) : s_ceil<TW>(xh) - TW(1);
    }
    first = isnan_x ? 1 : (int)xm;
    // Previous/Next with Periodic(OnCell) wrap (constant.jl:91-104); Nearest does not (its in-bounds index is in 1:n)
    if (per && (deg == B200ITP_CONSTANT_PREVIOUS || deg == B200ITP_CONSTANT_NEXT)) first = modrange1(first, n);
    A.L = 1;
    A.wv[0] = TW(1);
    A.wg[0] = TW(0);
    A.wh[0] = TW(0);
#pragma unroll
    for (int k = 1; k < 4; ++k) A.wv[k] = A.wg[k] = A.wh[k] = TW(0);
    A.first = first;
    A.contiguous = true;
#pragma unroll
    for (int k = 0; k < AxisTaps<DEG, TW>::LMAX; ++k) A.pos[k] = first - cf;
    return;
  }
  if (deg == B200ITP_CUBIC) {
    TW xf;
    if (!per) {  // cubic.jl:40-46, floorbounds indexing.jl:183-187
      if (x < TW(1)) {
        xf = s_floor<TW>(x + TW(0.5));
        A.edge = true;
      } else {
        xf = s_floor<TW>(x + TW(0));
      }
      if (xf > TW(n - 1)) {
        xf = xf - TW(1);
        A.edge = true;
      }
    } else {  // cubic.jl:50-57
      xf = s_floor<TW>(x);
    }
    const TW dx = x - xf;
    first = isnan_x ? 0 : (int)xf - 1;
    A.L = 4;
    // value_weights cubic.jl:63-69, constants convert(T, SimpleRatio) = T(p)/T(q)
    const TW omd = TW(1) - dx;
    const TW x3 = (dx * dx) * dx, xc3 = (omd * omd) * omd;
    const TW r16 = TW(1) / TW(6), r23 = TW(2) / TW(3);
    A.wv[0] = r16 * xc3;
    A.wv[1] = (r23 - dx * dx) + TW(0.5) * x3;
    A.wv[2] = (r23 - omd * omd) + TW(0.5) * xc3;
    A.wv[3] = r16 * x3;
    if (GRAD) {  // cubic.jl:71-77
      const TW x2 = dx * dx, xc2 = omd * omd;
      A.wg[0] = TW(-0.5) * xc2;
      A.wg[1] = TW(-2) * dx + TW(1.5) * x2;
      A.wg[2] = TW(2) * omd - TW(1.5) * xc2;
      A.wg[3] = TW(0.5) * x2;
      // hessian_weights cubic.jl:79
      A.wh[0] = omd;
      A.wh[1] = TW(3) * dx - TW(2);
      A.wh[2] = TW(3) * omd - TW(2);
      A.wh[3] = dx;
    }
  } else if (deg == B200ITP_QUADRATIC) {
    // quadratic.jl:39-43, roundbounds indexing.jl:172-177
    const TW xh = x + TW(0.5);
    TW xm;
    if ((double)x < (double)n + 0.5) {
      xm = s_floor<TW>(xh);
    } else {
      xm = s_ceil<TW>(xh) - TW(1);
      A.edge = true;
    }
    const TW dx = x - xm;
    first = isnan_x ? 0 : (int)xm - 1;
    A.L = 3;
    const TW a = dx - TW(0.5), b = dx + TW(0.5);  // quadratic.jl:45-53
    A.wv[0] = (a * a) / TW(2);
    A.wv[1] = TW(0.75) - dx * dx;
    A.wv[2] = (b * b) / TW(2);
    if (GRAD) {
      A.wg[0] = a;
      A.wg[1] = TW(-2) * dx;
      A.wg[2] = b;
      A.wh[0] = TW(1);  // quadratic.jl:55
      A.wh[1] = TW(-2);
      A.wh[2] = TW(1);
    }
    if (DEG == 0) {
      A.wv[3] = TW(0);
      A.wg[3] = TW(0);
      A.wh[3] = TW(0);
    }
  } else {  // linear.jl:43-58
    TW f = s_floor<TW>(x);
    if (x == TW(n)) {
      f = f - TW(1);
      A.edge = true;
    }
    const TW dx = x - f;
    first = isnan_x ? 1 : (int)f;
    A.L = 2;
    A.wv[0] = TW(1) - dx;
    A.wv[1] = dx;
    if (GRAD) {
      A.wg[0] = TW(-1);
      A.wg[1] = TW(1);
      A.wh[0] = TW(0);  // linear.jl:58
      A.wh[1] = TW(0);
    }
    if (DEG == 0) {
      A.wv[2] = A.wv[3] = TW(0);
      A.wg[2] = A.wg[3] = TW(0);
      A.wh[2] = A.wh[3] = TW(0);
    }
  }
  A.first = first;
  A.contiguous = true;
#pragma unroll
  for (int k = 0; k < AxisTaps<DEG, TW>::LMAX; ++k) {
    int idx = first + k;
    if (per) {
      idx = modrange1(idx, n);
      if (k > 0 && idx - cf != A.pos[0] + k) A.contiguous = false;
    }
    if (DEG == 0 && P.clamp_taps[d]) {  // expand_index of InPlace/InPlaceQ: (max(xi-1, 1), xi, min(xi+1, n))
      idx = idx < 1 ? 1 : (idx > n ? n : idx);
      A.contiguous = false;
    }
    A.pos[k] = idx - cf;
  }
  if (DEG == 0)
    for (int k = A.L; k < 4; ++k) A.pos[k] = A.pos[0];
}

// Nested tensor-product reduction (Interpolations.jl:304-316).  Returns the value of the sub-product over
// axes D..N-1 and, if GRAD, the gradient components of those axes (sharing the value partial sums is
// exact: component j of the gradient uses value weights on every axis but j).
template <int N, int DEG, typename TC, typename TW, bool GRAD, int D>
struct Reduce {
  using TA = typename Promote<TW, TC>::type;
  static constexpr int LMAX = AxisTaps<DEG, TW>::LMAX;
  __device__ __forceinline__ static void run(const EvalParams &P, const AxisTaps<DEG, TW> (&ax)[N], const TC *base,
                                             TA &val, TA (&g)[N]) {
    const AxisTaps<DEG, TW> &A = ax[D];
    TA gv = TA(0);
#pragma unroll
    for (int k = 0; k < LMAX; ++k) {
      if (DEG != 0 || k < A.L) {
        TA vk;
        TA gk[N];
        const TC *b = base + (long long)A.pos[k] * P.stride[D];
        if constexpr (D == N - 1) {
          vk = (TA)gather_ld(b);
        } else {
          Reduce<N, DEG, TC, TW, GRAD, D + 1>::run(P, ax, b, vk, gk);
        }
        const TA w = (TA)A.wv[k];
        if (k == 0) {
          val = w * vk;
        } else {
          val = w * vk + val;
        }
        if (GRAD) {
          const TA wgk = (TA)A.wg[k];
          if (k == 0)
            gv = wgk * vk;
          else
            gv = wgk * vk + gv;
          if constexpr (D < N - 1) {
#pragma unroll
            for (int j = D + 1; j < N; ++j) {
              if (k == 0)
                g[j] = w * gk[j];
              else
                g[j] = w * gk[j] + g[j];
            }
          }
        }
      }
    }
    if (GRAD) g[D] = gv;
  }
};

// One hessian component: the reference's nested reduction (Interpolations.jl:304-316) with, per axis, value (0),
// gradient (1) or hessian (2) weights (indexing.jl:114-144).
template <int N, int DEG, typename TC, typename TW, int D>
struct ReduceSel {
  using TA = typename Promote<TW, TC>::type;
  static constexpr int LMAX = AxisTaps<DEG, TW>::LMAX;
  __device__ __forceinline__ static TA run(const EvalParams &P, const AxisTaps<DEG, TW> (&ax)[N], const int (&sel)[N],
                                           const TC *base) {
    const AxisTaps<DEG, TW> &A = ax[D];
    TA val = TA(0);
#pragma unroll
    for (int k = 0; k < LMAX; ++k) {
      if (DEG != 0 || k < A.L) {
        const TC *b = base + (long long)A.pos[k] * P.stride[D];
        TA vk;
        if constexpr (D == N - 1) {
          vk = (TA)gather_ld(b);
        } else {
          vk = ReduceSel<N, DEG, TC, TW, D + 1>::run(P, ax, sel, b);
        }
        const TA w = (TA)(sel[D] == 0 ? A.wv[k] : (sel[D] == 1 ? A.wg[k] : A.wh[k]));
        val = k == 0 ? w * vk : w * vk + val;
      }
    }
    return val;
  }
};

// Value-only fast path for Float32 coefficients whose axis-1 taps are adjacent in memory: every
// (axis 2..N) row is fetched with two aligned 128-bit loads instead of 3-4 scalar loads, then the
// reduction runs in the reference order from registers.
template <int N, int DEG, typename TW>
__device__ __forceinline__ void load_row4(const float *coefs, long long lin, long long total, float (&c)[4]) {
  const long long a = lin & ~3LL;
  const int sft = (int)(lin - a);
  float4 q0;
  if (a + 4 <= total) {
    q0 = gather_ld4(reinterpret_cast<const float4 *>(coefs + a));
  } else {  // the aligned quad would extend past the allocation (total % 4 != 0, Quadratic: 3 taps end at total - 1)
    q0 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (a < total) q0.x = __ldg(coefs + a);
    if (a + 1 < total) q0.y = __ldg(coefs + a + 1);
    if (a + 2 < total) q0.z = __ldg(coefs + a + 2);
  }
  float4 q1 = make_float4(0.f, 0.f, 0.f, 0.f);
  if (sft != 0) {
    if (a + 8 <= total) {
      q1 = gather_ld4(reinterpret_cast<const float4 *>(coefs + a + 4));
    } else {  // last few elements of the array: stay inside the allocation
      if (a + 4 < total) q1.x = __ldg(coefs + a + 4);
      if (a + 5 < total) q1.y = __ldg(coefs + a + 5);
      if (a + 6 < total) q1.z = __ldg(coefs + a + 6);
    }
  }
  const float w[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    float v = w[k];
    v = sft == 1 ? w[k + 1] : v;
    v = sft == 2 ? w[k + 2] : v;
    v = sft == 3 ? w[k + 3] : v;
    c[k] = v;
  }
}

template <typename TO>
__device__ __forceinline__ void store_out(void *p, long long i, TO v, int out_f64) {
  if (out_f64)
    reinterpret_cast<double *>(p)[i] = (double)v;
  else
    reinterpret_cast<float *>(p)[i] = (float)v;
}

// B-spline stage of one query: positions, weights (type TWW) and the nested reduction; results are
// widened (exactly) to TAO.
// hessian of the bare B-spline at xl: N*N entries (symmetric), components evaluated in the reference's order
template <int N, int DEG, typename TC, typename TWW, typename TAO>
__device__ __forceinline__ void bspline_hessian(const EvalParams &P, const TC *coefs, const TWW (&xl)[N],
                                                TAO (&h)[N * N], int (&first_idx)[N], unsigned &br) {
  AxisTaps<DEG, TWW> ax[N];
#pragma unroll
  for (int d = 0; d < N; ++d) {
    axis_setup<DEG, TWW, true>(P, d, xl[d], ax[d]);
    first_idx[d] = ax[d].pos[0] + P.coef_first[d] + (DEG == 0 ? P.xshift[d] : 0);
    if (ax[d].edge) br |= B200ITP_BR_EDGE << (4 * d);
  }
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int j = i; j < N; ++j) {
      int sel[N];
#pragma unroll
      for (int d = 0; d < N; ++d) sel[d] = (d == i && d == j) ? 2 : ((d == i || d == j) ? 1 : 0);
      const TAO v = (TAO)ReduceSel<N, DEG, TC, TWW, 0>::run(P, ax, sel, coefs);
      h[i * N + j] = v;
      h[j * N + i] = v;
    }
}

template <int N, int DEG, typename TC, typename TWW, typename TAO>
__device__ __forceinline__ void bspline_stage(const EvalParams &P, const TC *coefs, const TWW (&xl)[N], bool need_grad,
                                              TAO &val_out, TAO (&grad_out)[N], int (&first_idx)[N], unsigned &br) {
  using TW = TWW;
  using TA = typename Promote<TWW, TC>::type;
  TA val = TA(0);
  TA grad[N];
#pragma unroll
  for (int d = 0; d < N; ++d) grad[d] = TA(0);
  AxisTaps<DEG, TW> ax[N];
  if (need_grad) {
#pragma unroll
    for (int d = 0; d < N; ++d) axis_setup<DEG, TW, true>(P, d, xl[d], ax[d]);
    Reduce<N, DEG, TC, TW, true, 0>::run(P, ax, coefs, val, grad);
  } else {
#pragma unroll
    for (int d = 0; d < N; ++d) axis_setup<DEG, TW, false>(P, d, xl[d], ax[d]);
    bool done = false;
    if constexpr (sizeof(TC) == 4 && N <= 3 && (DEG == 2 || DEG == 3)) {
      if (ax[0].contiguous && P.vec_ok) {
        // rows of axis-1 taps through aligned 128-bit loads, reduction in the reference order
        constexpr int L = DEG + 1;
        constexpr int ROWS = N == 1 ? 1 : (N == 2 ? L : L * L);
        float c[ROWS][4];
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
          long long lin = ax[0].pos[0];
          if (N >= 2) lin += (long long)ax[1].pos[N == 2 ? r : r % L] * P.stride[1];
          if (N >= 3) lin += (long long)ax[2].pos[r / L] * P.stride[2];
          load_row4<N, DEG, TW>(reinterpret_cast<const float *>(coefs), lin, P.total, c[r]);
        }
        // value = sum_k0 w0[k0] * ( sum_k1 w1[k1] * ( sum_k2 w2[k2] * c[k2][k1][k0] ) )
#pragma unroll
        for (int k0 = 0; k0 < L; ++k0) {
          TA v0;
          if (N == 1) {
            v0 = (TA)c[0][k0];
          } else {
#pragma unroll
            for (int k1 = 0; k1 < L; ++k1) {
              TA v1;
              if (N == 2) {
                v1 = (TA)c[k1][k0];
              } else {
#pragma unroll
                for (int k2 = 0; k2 < L; ++k2) {
                  const TA t2 = (TA)ax[2].wv[k2] * (TA)c[k2 * L + k1][k0];
                  v1 = k2 == 0 ? t2 : t2 + v1;
                }
              }
              const TA t1 = (TA)ax[1].wv[k1] * v1;
              v0 = k1 == 0 ? t1 : t1 + v0;
            }
          }
          const TA t0 = (TA)ax[0].wv[k0] * v0;
          val = k0 == 0 ? t0 : t0 + val;
        }
        done = true;
      }
    }
    if (!done) Reduce<N, DEG, TC, TW, false, 0>::run(P, ax, coefs, val, grad);
  }
#pragma unroll
  for (int d = 0; d < N; ++d) {
    first_idx[d] = ax[d].pos[0] + P.coef_first[d] + (DEG == 0 ? P.xshift[d] : 0);
    if (ax[d].edge) br |= B200ITP_BR_EDGE << (4 * d);
    grad_out[d] = (TAO)grad[d];
  }
  val_out = (TAO)val;
}

// MODE 0: values; MODE 1: gradients; MODE 2: hessians (in bounds only: indexing.jl:37-41, scaling.jl:150-155,
// extrapolation.jl:84-90)
template <int N, int DEG, typename TC, typename TX, typename TW, int MODE>
// 3-D / 4-D stencils: capped at 128 registers (two CTAs per SM; measured +35 % tricubic values, +60 % from the L2);
// 1-D / 2-D kernels: three CTAs per SM (<= 85 registers), the allocation they had without a cap (a cap of one CTA lets
// ptxas take 104-123 registers and costs 25-45 %)
__global__ void __launch_bounds__(256, N >= 3 ? 2 : (MODE == 0 ? 0 : 3)) eval_kernel(const __grid_constant__ EvalParams P) {
  using TA = typename Promote<TW, TC>::type;
  const TC *coefs = reinterpret_cast<const TC *>(P.coefs);
  const long long stride_q = (long long)gridDim.x * blockDim.x;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < P.nq; t += stride_q) {
    const long long q = P.perm ? (long long)P.perm[t] : t;
    TX x[N];
    unsigned br = 0;
    bool inb = true;
#pragma unroll
    for (int d = 0; d < N; ++d) {
      x[d] = reinterpret_cast<const TX *>(P.coords[d])[q];
      const double xd = (double)x[d];
      if (!(P.lb[d] <= xd)) {
        br |= B200ITP_BR_BELOW << (4 * d);
        inb = false;
      }
      if (!(xd <= P.ub[d])) {
        br |= B200ITP_BR_ABOVE << (4 * d);
        inb = false;
      }
      if constexpr (DEG == 0) {  // NoInterp: Int(x) of a non-integer throws (nointerp.jl:17)
        if (P.degree[d] == B200ITP_NOINTERP && !(x[d] == s_floor<TX>(x[d]))) inb = false;
      }
    }
    // ---- extrapolation stage: xs (in TW; exact widening when the stage ran in TX)
    TW xs[N];
    bool threw = false, filled = false, moved = false;
    bool use_x = false;  // gradient of an in-bounds query ignores the mapping (extrapolation.jl:73-74)
    if (P.etp_kind == B200ITP_ETP_FLAGS) {
#pragma unroll
      for (int d = 0; d < N; ++d) {
        bool th;
        if (sizeof(TX) == sizeof(TW) || !P.e_wide) {
          TX y;
          th = etp_inbounds<TX>(P, d, x[d], y);
          xs[d] = (TW)y;
          if (y != x[d]) {
            moved = true;
            br |= B200ITP_BR_MOVED << (4 * d);
          }
        } else {
          TW y;
          th = etp_inbounds<TW>(P, d, (TW)x[d], y);
          xs[d] = y;
          if (y != (TW)x[d]) {
            moved = true;
            br |= B200ITP_BR_MOVED << (4 * d);
          }
        }
        threw = threw || th;
      }
      if (MODE >= 1 && inb) use_x = true;
      if (MODE == 2) threw = !inb;  // "extrapolation of hessian not yet implemented": reported like a BoundsError
    } else {
#pragma unroll
      for (int d = 0; d < N; ++d) xs[d] = (TW)x[d];
      if (!inb) {
        if (P.etp_kind == B200ITP_ETP_FILL)
          filled = true;
        else
          threw = true;
      }
    }

    TA val = TA(0);
    TA grad[N];
#pragma unroll
    for (int d = 0; d < N; ++d) grad[d] = TA(0);
    int first_idx[N];
#pragma unroll
    for (int d = 0; d < N; ++d) first_idx[d] = INT_MIN;

    if (!threw && !filled) {
      // ---- scale stage: coordlookup + clamp (scaling.jl:78-83,102-115)
      TW xl[N];
#pragma unroll
      for (int d = 0; d < N; ++d) {
        TW xv = use_x ? (TW)x[d] : xs[d];
        if (P.has_scale) {
          if (sizeof(TX) == sizeof(TW) || (use_x ? P.s_wide_x : P.s_wide)) {
            xv = (xv - (TW)P.rfirst[d]) / (TW)P.rstep[d] + TW(1);
          } else {
            const TX y = ((TX)xv - (TX)P.rfirst[d]) / (TX)P.rstep[d] + TX(1);
            xv = (TW)y;
          }
          const TW lo = (TW)P.inner_lb[d], hi = (TW)P.inner_ub[d];
          xv = xv > hi ? hi : (xv < lo ? lo : xv);
        }
        // interpolate!: parent axes 2:n-1 over coefficients 1:n == the padded layout one index further (x - 1 is exact)
        if constexpr (DEG == 0) xv = xv - (TW)P.xshift[d];
        xl[d] = xv;
      }
      // ---- B-spline stage (weights in TW; the in-bounds gradient of an extrapolated Float32 query whose
      //      only widening came from the extrapolation bounds is computed in TX, as the reference does)
      const bool need_grad = MODE == 1 || (moved && P.any_line);
      bool narrow = false;
      if constexpr (MODE >= 1 && sizeof(TX) != sizeof(TW)) narrow = use_x && !P.w_wide_x;
      if constexpr (MODE == 2) {
        TA h[N * N];
        if (narrow) {
          if constexpr (sizeof(TX) != sizeof(TW)) {
            TX xn[N];
#pragma unroll
            for (int d = 0; d < N; ++d) xn[d] = (TX)xl[d];
            bspline_hessian<N, DEG, TC, TX, TA>(P, coefs, xn, h, first_idx, br);
          }
        } else {
          bspline_hessian<N, DEG, TC, TW, TA>(P, coefs, xl, h, first_idx, br);
        }
#pragma unroll
        for (int i = 0; i < N; ++i)
#pragma unroll
          for (int j = 0; j < N; ++j) {
            TA v = h[i * N + j];
            if (P.has_scale) {  // rescale_hessian_components scaling.jl:157-160: h ./ (steps .* steps')
              const double prod = P.range_f32[i] ? (double)((float)P.rstep[i] * (float)P.rstep[j]) : P.rstep[i] * P.rstep[j];
              v = v / (TA)prod;
            }
            store_out<TA>(P.out_hess, (q * N + i) * N + j, v, P.out_f64);
          }
      } else if (narrow) {
        if constexpr (MODE == 1 && sizeof(TX) != sizeof(TW)) {
          TX xn[N];
#pragma unroll
          for (int d = 0; d < N; ++d) xn[d] = (TX)xl[d];
          bspline_stage<N, DEG, TC, TX, TA>(P, coefs, xn, true, val, grad, first_idx, br);
        }
      } else {
        bspline_stage<N, DEG, TC, TW, TA>(P, coefs, xl, need_grad, val, grad, first_idx, br);
      }
      if (need_grad && P.has_scale) {
#pragma unroll
        for (int d = 0; d < N; ++d) grad[d] = grad[d] / (TA)P.rstep[d];  // rescale_gradient scaling.jl:116,137
      }
      // ---- extrapolate_value / extrapolate_gradient (extrapolation.jl:166-185)
      if (P.etp_kind == B200ITP_ETP_FLAGS) {
        if (MODE == 0) {
          if (moved) {
            br |= B200ITP_BR_SLOWPATH;
            if (P.any_line) {
#pragma unroll
              for (int d = 0; d < N; ++d) {
                const bool pair = (P.pair_mask >> d) & 1;
                TW dxs;  // x - xs in the type the extrapolation stage ran in
                if (sizeof(TX) == sizeof(TW) || !P.e_wide)
                  dxs = (TW)(x[d] - (TX)xs[d]);
                else
                  dxs = (TW)x[d] - xs[d];
                const bool lo_line = P.etp_lo[d] == B200ITP_BC_LINE, hi_line = P.etp_hi[d] == B200ITP_BC_LINE;
                if (!pair) {
                  if (lo_line) val = val + (TA)dxs * grad[d];
                } else {
                  if (lo_line && (TW)x[d] < xs[d]) val = val + (TA)dxs * grad[d];
                  if (hi_line && (TW)x[d] > xs[d]) val = val + (TA)dxs * grad[d];
                }
              }
            }
          }
        } else if (!inb) {
#pragma unroll
          for (int d = 0; d < N; ++d)
            if (P.etp_lo[d] == B200ITP_BC_FLAT && xs[d] != (TW)x[d]) grad[d] = TA(0);
        }
      }
    }
    if (filled) {
      br |= B200ITP_BR_FILLED;
    }
    if (threw) {
      br |= B200ITP_BR_THROW;
      atomicMin(P.first_oob, (unsigned long long)q);
    }
    // ---- outputs
    if (MODE == 2) {
      if (threw) {
#pragma unroll
        for (int k = 0; k < N * N; ++k) store_out<double>(P.out_hess, q * N * N + k, nan(""), P.out_f64);
      }
    } else if (MODE == 0) {
      if (threw) {
        store_out<double>(P.out_value, q, nan(""), P.out_f64);
      } else if (filled) {
        store_out<double>(P.out_value, q, P.fill, P.out_f64);
      } else {
        store_out<TA>(P.out_value, q, val, P.out_f64);
      }
    } else if (DEG == 0) {  // mixed specs: NoInterp axes have no gradient component (indexing.jl:108)
      const long long G = P.ngrad;
#pragma unroll
      for (int d = 0; d < N; ++d) {
        const int slot = P.grad_slot[d];
        if (slot < 0) continue;
        if (threw)
          store_out<double>(P.out_grad, q * G + slot, nan(""), P.out_f64);
        else
          store_out<TA>(P.out_grad, q * G + slot, filled ? TA(0) : grad[d], P.out_f64);
      }
    } else {
#pragma unroll
      for (int d = 0; d < N; ++d) {
        if (threw)
          store_out<double>(P.out_grad, q * N + d, nan(""), P.out_f64);
        else
          store_out<TA>(P.out_grad, q * N + d, filled ? TA(0) : grad[d], P.out_f64);
      }
    }
    if (P.dbg_index) {
#pragma unroll
      for (int d = 0; d < N; ++d) P.dbg_index[q * N + d] = first_idx[d];
    }
    if (P.dbg_branch) P.dbg_branch[q] = br;
  }
}

// one translation unit per (TC, TX, TW) combination instantiates and launches its kernels
template <typename TC, typename TX, typename TW>
int launch_eval_combo(const EvalParams &P, int uniform_degree, int mode, int blocks, cudaStream_t st);

}  // namespace b200itp
