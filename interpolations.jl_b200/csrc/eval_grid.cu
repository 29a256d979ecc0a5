// Separable tensor-product (grid) evaluation: `itp(xvec, yvec, ...)` (src/b-splines/indexing.jl:16-23,
// src/scaling/scaling.jl:98-100, src/extrapolation/extrapolation.jl:59-70) and `eachvalue(sitp)`
// (src/scaling/scaling.jl:174-256).
//
// The reference evaluates `coefs[wis...]` for every tuple of `Iterators.product(wis...)`: the per-axis positions and
// weights depend on ONE coordinate vector each, and the nested reduction (src/Interpolations.jl:304-316) is
//   value = sum_k1 w1[k1] * ( sum_k2 w2[k2] * ( ... sum_kN wN[kN] * c[k1, k2, ..., kN] ) ),   taps accumulated w*v + acc.
// The innermost sums do not depend on the outer coordinates, so the SAME operations in the SAME order can be run as N
// one-dimensional resampling passes, last axis first:
//   U_N[i1..i_{N-1}, oN]      = sum_kN wN(oN)[kN] * c[i1..i_{N-1}, posN(oN)[kN]]
//   U_d[i1..i_{d-1}, od, ...] = sum_kd wd(od)[kd] * U_{d+1}[i1..i_{d-1}, posd(od)[kd], ...]
// Every output element is bit-identical to the pointwise evaluation of its coordinate tuple (no FMA: this file is
// compiled -fmad=false like the other evaluation kernels), but each coefficient is read a handful of times instead of
// 4^N times per output and every pass is a coalesced stream: resampling a 512^3 volume onto 1000^3 moves ~10 GB of
// HBM traffic instead of running 10^9 scattered 64-tap gathers.
//
// Per-axis positions / weights / bounds / extrapolation mapping / coordlookup come from the device functions of the
// direct kernel (eval.cuh: etp_inbounds, axis_setup), evaluated once per coordinate by grid_axis_prep_kernel.
// Scope: Linear / Quadratic / Cubic / Constant axes (uniform or mixed), bare or scaled, extrapolation flags Throw /
// Flat / Periodic / Reflect (per-axis coordinate maps).  Line extrapolation (value + gradient * dx is not separable),
// fill values and NoInterp axes take the tuple-expansion path of eval_sorted.cu.
#include "eval.cuh"

#include <algorithm>
#include <cmath>

namespace b200itp {

int fill_eval_params_grid(const b200itp_desc *D, const void *coefs, const void *const *coords, EvalParams &P,
                          bool *w_wide, bool *homogeneous);

constexpr int kGridTaps = 4;

// per coordinate o of one axis: tap positions (0-based along the coefficient axis), weights (NaN when the coordinate
// throws: the NaN then reaches every output that uses it, as in the pointwise kernel) and the tap count
template <typename TW>
struct AxisTable {
  int *pos;   // len * 4
  TW *w;      // len * 4
  int *ntaps; // 1: number of taps of this axis (uniform along the axis)
};

template <typename TX, typename TW>
__global__ void __launch_bounds__(256) grid_axis_prep_kernel(const __grid_constant__ EvalParams P, int d, long long len,
                                                             AxisTable<TW> tab, unsigned long long *first_throw) {
  const long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= len) return;
  const TX x = reinterpret_cast<const TX *>(P.coords[d])[o];
  const double xd = (double)x;
  const bool inb = (P.lb[d] <= xd) && (xd <= P.ub[d]);
  bool threw = false;
  TW xs;
  if (P.etp_kind == B200ITP_ETP_FLAGS) {  // eval_kernel, extrapolation stage (extrapolation.jl:106-151)
    if (sizeof(TX) == sizeof(TW) || !P.e_wide) {
      TX y;
      threw = etp_inbounds<TX>(P, d, x, y);
      xs = (TW)y;
    } else {
      TW y;
      threw = etp_inbounds<TW>(P, d, (TW)x, y);
      xs = y;
    }
  } else {
    xs = (TW)x;
    threw = !inb;
  }
  TW xv = xs;
  if (P.has_scale) {  // coordlookup + clamp (scaling.jl:78-83,102-115)
    if (sizeof(TX) == sizeof(TW) || P.s_wide) {
      xv = (xv - (TW)P.rfirst[d]) / (TW)P.rstep[d] + TW(1);
    } else {
      const TX y = ((TX)xv - (TX)P.rfirst[d]) / (TX)P.rstep[d] + TX(1);
      xv = (TW)y;
    }
    const TW lo = (TW)P.inner_lb[d], hi = (TW)P.inner_ub[d];
    xv = xv > hi ? hi : (xv < lo ? lo : xv);
  }
  xv = xv - (TW)P.xshift[d];
  AxisTaps<0, TW> A;
  axis_setup<0, TW, false>(P, d, threw ? TW(1) : xv, A);
  if (o == 0) *tab.ntaps = A.L;
  const TW nanw = (TW)nan("");
#pragma unroll
  for (int k = 0; k < kGridTaps; ++k) {
    tab.pos[o * kGridTaps + k] = A.pos[k];
    tab.w[o * kGridTaps + k] = threw ? nanw : A.wv[k];
  }
  if (threw) atomicMin(first_throw + d, (unsigned long long)o);
}

template <typename TO, typename TA>
__device__ __forceinline__ TO cast_out(TA v) {
  return (TO)v;
}

// 16-byte vectors of the element types (Float32 x 4, Float64 x 2) for the row passes
template <typename T>
struct Vec16;
template <>
struct Vec16<float> {
  using type = float4;
  static constexpr int N = 4;
};
template <>
struct Vec16<double> {
  using type = double2;
  static constexpr int N = 2;
};
template <typename T, int V>
struct Pack {
  T v[V];
};
template <typename T, int V>
__device__ __forceinline__ Pack<T, V> load_pack(const T *p) {
  Pack<T, V> r;
  if constexpr (V == 1) {
    r.v[0] = __ldg(p);
  } else if constexpr (sizeof(T) * V == 16) {
    using V16 = typename Vec16<T>::type;
    const V16 q = __ldg(reinterpret_cast<const V16 *>(p));
    if constexpr (V == 4) {
      r.v[0] = q.x; r.v[1] = q.y; r.v[2] = q.z; r.v[3] = q.w;
    } else {
      r.v[0] = q.x; r.v[1] = q.y;
    }
  } else {  // 4 doubles: two 16-byte loads
    const double2 a = __ldg(reinterpret_cast<const double2 *>(p)), b = __ldg(reinterpret_cast<const double2 *>(p) + 1);
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = b.x; r.v[3] = b.y;
  }
  return r;
}
template <typename T, int V>
__device__ __forceinline__ void store_pack(T *p, const Pack<T, V> &r) {
  if constexpr (V == 1) {
    *p = r.v[0];
  } else if constexpr (sizeof(T) * V == 16) {
    using V16 = typename Vec16<T>::type;
    V16 q;
    if constexpr (V == 4) {
      q.x = r.v[0]; q.y = r.v[1]; q.z = r.v[2]; q.w = r.v[3];
    } else {
      q.x = r.v[0]; q.y = r.v[1];
    }
    *reinterpret_cast<V16 *>(p) = q;
  } else {
    reinterpret_cast<double2 *>(p)[0] = make_double2(r.v[0], r.v[1]);
    reinterpret_cast<double2 *>(p)[1] = make_double2(r.v[2], r.v[3]);
  }
}

// One resampling pass along axis d >= 1: rows of R1 contiguous elements.
//   in  [i + R1 * (p + Sd * h)]   i < R1, p < Sd, h < H
//   out [i + R1 * (o + Od * h)]
// threadIdx.x walks the row in vectors of V elements (16-byte accesses when the row length and the pointers allow),
// threadIdx.y / blockIdx.y walk the rows (o, h); a row's taps and weights are uniform across its threads.
template <typename TI, typename TW, typename TA, typename TO, int L, int V, bool TABLE_IN_SMEM>
__global__ void __launch_bounds__(256) grid_pass_rows_kernel(const TI *__restrict__ in, TO *__restrict__ out, long long R1,
                                                             long long Sd, long long Od, long long H,
                                                             const int *__restrict__ pos, const TW *__restrict__ w) {
  // the axis' tap table in shared memory when it fits (a row's taps are then an LDS away instead of a dependent global
  // load in front of every row: with short rows — one vector per thread — that latency was the bound)
  extern __shared__ unsigned char grid_smem[];
  int *spos = reinterpret_cast<int *>(grid_smem);
  TW *sw = reinterpret_cast<TW *>(grid_smem + (size_t)Od * kGridTaps * sizeof(int));
  if (TABLE_IN_SMEM) {
    const int nt = blockDim.x * blockDim.y, t = threadIdx.y * blockDim.x + threadIdx.x;
    for (long long i = t; i < Od * kGridTaps; i += nt) {
      spos[i] = pos[i];
      sw[i] = w[i];
    }
    __syncthreads();
  }
  const long long rows = Od * H;
  const long long nvec = R1 / V;
  const long long row0 = (long long)blockIdx.y * blockDim.y + threadIdx.y;
  const long long rstep = (long long)gridDim.y * blockDim.y;
  long long o = row0 % Od, h = row0 / Od;
  const long long ostep = rstep % Od, hstep = rstep / Od;
  for (long long row = row0; row < rows; row += rstep) {
    int p[L];
    TA wk[L];
#pragma unroll
    for (int k = 0; k < L; ++k) {
      p[k] = TABLE_IN_SMEM ? spos[o * kGridTaps + k] : __ldg(pos + o * kGridTaps + k);
      wk[k] = (TA)(TABLE_IN_SMEM ? sw[o * kGridTaps + k] : __ldg(w + o * kGridTaps + k));
    }
    const TI *src = in + R1 * Sd * h;
    TO *dst = out + R1 * (o + Od * h);
    for (long long iv = (long long)blockIdx.x * blockDim.x + threadIdx.x; iv < nvec; iv += (long long)gridDim.x * blockDim.x) {
      Pack<TI, V> c[L];
#pragma unroll
      for (int k = 0; k < L; ++k) c[k] = load_pack<TI, V>(src + iv * V + R1 * (long long)p[k]);
      Pack<TO, V> r;
#pragma unroll
      for (int e = 0; e < V; ++e) {
        TA acc = TA(0);
#pragma unroll
        for (int k = 0; k < L; ++k) {
          const TA t = wk[k] * (TA)c[k].v[e];
          acc = k == 0 ? t : t + acc;
        }
        r.v[e] = cast_out<TO, TA>(acc);
      }
      store_pack<TO, V>(dst + iv * V, r);
    }
    o += ostep;
    h += hstep;
    if (o >= Od) {
      o -= Od;
      ++h;
    }
  }
}

// The pass along axis 0 (contiguous): out[o + O0 * h] = sum_k w(o)[k] * in[pos(o)[k] + S0 * h].
// Thread <-> output coordinate o (its taps and weights stay in registers), rows h walked by blockIdx.y, U rows in flight.
template <typename TI, typename TW, typename TA, typename TO, int L>
__global__ void __launch_bounds__(256) grid_pass_axis0_kernel(const TI *__restrict__ in, TO *__restrict__ out, long long S0,
                                                              long long O0, long long H, const int *__restrict__ pos,
                                                              const TW *__restrict__ w) {
  constexpr int U = 4;
  const long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= O0) return;
  int p[L];
  TA wk[L];
#pragma unroll
  for (int k = 0; k < L; ++k) {
    p[k] = pos[o * kGridTaps + k];
    wk[k] = (TA)w[o * kGridTaps + k];
  }
  const long long hs = gridDim.y;
  long long h = blockIdx.y;
  for (; h + (U - 1) * hs < H; h += U * hs) {
    TI c[U][L];
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int k = 0; k < L; ++k) c[u][k] = __ldg(in + S0 * (h + u * hs) + p[k]);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      TA acc = TA(0);
#pragma unroll
      for (int k = 0; k < L; ++k) {
        const TA t = wk[k] * (TA)c[u][k];
        acc = k == 0 ? t : t + acc;
      }
      out[o + O0 * (h + u * hs)] = cast_out<TO, TA>(acc);
    }
  }
  for (; h < H; h += hs) {
    TA acc = TA(0);
#pragma unroll
    for (int k = 0; k < L; ++k) {
      const TA t = wk[k] * (TA)__ldg(in + S0 * h + p[k]);
      acc = k == 0 ? t : t + acc;
    }
    out[o + O0 * h] = cast_out<TO, TA>(acc);
  }
}

template <typename TI, typename TW, typename TA, typename TO, int L>
static int launch_rows(const TI *in, TO *out, long long R1, long long Sd, long long Od, long long H, const int *pos,
                       const TW *w, int sms, cudaStream_t st) {
  // vector width: 16-byte accesses on the wider of the two element types when rows and pointers are aligned
  constexpr int VMAX = 16 / (sizeof(TI) > sizeof(TO) ? sizeof(TO) : sizeof(TI));  // 4 for float, 2 for double (narrower type)
  const bool aligned = ((uintptr_t)in % 16) == 0 && ((uintptr_t)out % 16) == 0;
  const long long rows = Od * H;
  auto go = [&](auto vtag) -> int {
    constexpr int V = decltype(vtag)::value;
    const long long nvec = R1 / V;
    unsigned tx = 32;
    while (tx < 256 && tx < nvec) tx <<= 1;
    const unsigned ty = 256 / tx;
    const unsigned gx = (unsigned)std::min<long long>((nvec + tx - 1) / tx, 32);
    const long long want = std::max<long long>(1, (long long)sms * 8 / gx);
    const unsigned gy = (unsigned)std::min<long long>(std::min<long long>((rows + ty - 1) / ty, want), 65535LL);
    const size_t tab = (size_t)Od * kGridTaps * (sizeof(int) + sizeof(TW));
    if (tab <= 40 * 1024)
      grid_pass_rows_kernel<TI, TW, TA, TO, L, V, true><<<dim3(gx, gy), dim3(tx, ty), tab, st>>>(in, out, R1, Sd, Od, H, pos, w);
    else
      grid_pass_rows_kernel<TI, TW, TA, TO, L, V, false><<<dim3(gx, gy), dim3(tx, ty), 0, st>>>(in, out, R1, Sd, Od, H, pos, w);
    return 0;
  };
  if (aligned && R1 % VMAX == 0) return go(std::integral_constant<int, VMAX>{});
  return go(std::integral_constant<int, 1>{});
}

template <typename TI, typename TW, typename TA, typename TO>
static int launch_pass(const TI *in, TO *out, int d, long long R1, long long Sd, long long Od, long long H, int L,
                       const int *pos, const TW *w, int sms, cudaStream_t st) {
  if (Od == 0 || H == 0 || R1 == 0) return 0;
  if (d == 0) {
    // enough rows per CTA column to amortise the tap registers, enough CTAs to fill the machine
    const unsigned gx0 = (unsigned)((Od + 255) / 256);
    dim3 grid(gx0, (unsigned)std::min<long long>(std::min<long long>(H, 65535LL), std::max<long long>(1, (long long)sms * 32 / gx0)));
    switch (L) {
      case 1: grid_pass_axis0_kernel<TI, TW, TA, TO, 1><<<grid, 256, 0, st>>>(in, out, Sd, Od, H, pos, w); break;
      case 2: grid_pass_axis0_kernel<TI, TW, TA, TO, 2><<<grid, 256, 0, st>>>(in, out, Sd, Od, H, pos, w); break;
      case 3: grid_pass_axis0_kernel<TI, TW, TA, TO, 3><<<grid, 256, 0, st>>>(in, out, Sd, Od, H, pos, w); break;
      default: grid_pass_axis0_kernel<TI, TW, TA, TO, 4><<<grid, 256, 0, st>>>(in, out, Sd, Od, H, pos, w); break;
    }
  } else {
    switch (L) {
      case 1: launch_rows<TI, TW, TA, TO, 1>(in, out, R1, Sd, Od, H, pos, w, sms, st); break;
      case 2: launch_rows<TI, TW, TA, TO, 2>(in, out, R1, Sd, Od, H, pos, w, sms, st); break;
      case 3: launch_rows<TI, TW, TA, TO, 3>(in, out, R1, Sd, Od, H, pos, w, sms, st); break;
      default: launch_rows<TI, TW, TA, TO, 4>(in, out, R1, Sd, Od, H, pos, w, sms, st); break;
    }
  }
  count_launch();
  B200_CUDA_OK(cudaGetLastError());
  return 0;
}

static inline size_t up256(size_t v) { return (v + 255) / 256 * 256; }

static int taps_of(int degree) {
  switch (degree) {
    case B200ITP_LINEAR: return 2;
    case B200ITP_QUADRATIC: return 3;
    case B200ITP_CUBIC: return 4;
    default: return 1;  // Constant
  }
}

bool grid_separable_scope(const b200itp_desc *d) {
  if (d->etp_kind == B200ITP_ETP_FILL) return false;
  for (int i = 0; i < d->ndims; ++i) {
    if (d->degree[i] == B200ITP_NOINTERP) return false;
    if (d->etp_kind == B200ITP_ETP_FLAGS && (d->etp_lo[i] == B200ITP_BC_LINE || d->etp_hi[i] == B200ITP_BC_LINE)) return false;
  }
  return true;
}

// bytes of scratch the separable path needs: per-axis tables + the intermediates U_N .. U_2 (accumulator type: 8 bytes
// is the upper bound)
size_t grid_separable_scratch_bytes(const b200itp_desc *d, const int64_t *axis_len) {
  size_t total = 256;
  for (int i = 0; i < d->ndims; ++i) total += up256((size_t)axis_len[i] * kGridTaps * 4) + up256((size_t)axis_len[i] * kGridTaps * 8) + 256;
  total += 512;  // first-throw slots, tap counts
  for (int p = d->ndims - 1; p >= 1; --p) {  // U_p: axes < p still have the coefficient sizes, axes >= p the output sizes
    size_t n = 1;
    for (int a = 0; a < p; ++a) n *= (size_t)d->size[a];
    for (int a = p; a < d->ndims; ++a) n *= (size_t)axis_len[a];
    total += up256(n * 8);
  }
  return total;
}

template <typename TC, typename TX, typename TW, typename TO>
static int run_separable(int dev, const b200itp_desc *D, const EvalParams &P, const int64_t *axis_len, void *out_value,
                         int64_t *first_oob, unsigned char *sp, cudaStream_t st) {
  using TA = typename Promote<TW, TC>::type;
  const int N = D->ndims;
  const int sms = sm_count(dev);
  // ---- carve the scratch
  size_t off = 0;
  auto take = [&](size_t bytes) {
    unsigned char *p = sp + off;
    off += up256(bytes);
    return p;
  };
  unsigned long long *first_throw = reinterpret_cast<unsigned long long *>(take(64));
  int *ntaps = reinterpret_cast<int *>(take(64));
  AxisTable<TW> tab[4];
  for (int d = 0; d < N; ++d) {
    tab[d].pos = reinterpret_cast<int *>(take((size_t)axis_len[d] * kGridTaps * 4));
    tab[d].w = reinterpret_cast<TW *>(take((size_t)axis_len[d] * kGridTaps * sizeof(TW)));
    tab[d].ntaps = ntaps + d;
  }
  B200_CUDA_OK(cudaMemsetAsync(first_throw, 0xFF, 64, st));
  {
    StageScope sc("grid_axis_tables", st);
    for (int d = 0; d < N; ++d) {
      grid_axis_prep_kernel<TX, TW><<<(unsigned)((axis_len[d] + 255) / 256), 256, 0, st>>>(P, d, axis_len[d], tab[d], first_throw);
      count_launch();
    }
    B200_CUDA_OK(cudaGetLastError());
  }
  // ---- passes, last axis first
  const void *cur = P.coefs;
  bool cur_is_coefs = true;
  for (int d = N - 1; d >= 0; --d) {
    long long R1 = 1, H = 1;
    for (int a = 0; a < d; ++a) R1 *= D->size[a];
    for (int a = d + 1; a < N; ++a) H *= axis_len[a];
    const long long Sd = D->size[d], Od = axis_len[d];
    const int L = taps_of(D->degree[d]);
    const bool last = d == 0;
    void *dst = last ? out_value : (void *)take((size_t)R1 * Od * H * sizeof(TA));
    StageScope sc(d == 0 ? "grid_pass_axis0" : (d == 1 ? "grid_pass_axis1" : "grid_pass_axis2+"), st);
    int rc;
    if (cur_is_coefs) {
      if (last)
        rc = launch_pass<TC, TW, TA, TO>((const TC *)cur, (TO *)dst, d, R1, Sd, Od, H, L, tab[d].pos, tab[d].w, sms, st);
      else
        rc = launch_pass<TC, TW, TA, TA>((const TC *)cur, (TA *)dst, d, R1, Sd, Od, H, L, tab[d].pos, tab[d].w, sms, st);
    } else {
      if (last)
        rc = launch_pass<TA, TW, TA, TO>((const TA *)cur, (TO *)dst, d, R1, Sd, Od, H, L, tab[d].pos, tab[d].w, sms, st);
      else
        rc = launch_pass<TA, TW, TA, TA>((const TA *)cur, (TA *)dst, d, R1, Sd, Od, H, L, tab[d].pos, tab[d].w, sms, st);
    }
    if (rc) return rc;
    cur = dst;
    cur_is_coefs = false;
  }
  if (first_oob) {
    unsigned long long h[4] = {~0ULL, ~0ULL, ~0ULL, ~0ULL};
    B200_CUDA_OK(cudaMemcpyAsync(h, first_throw, sizeof(unsigned long long) * 4, cudaMemcpyDeviceToHost, st));
    B200_CUDA_OK(cudaStreamSynchronize(st));
    // first tuple (first axis fastest) with a throwing coordinate: the smallest o_d * prod(len[a < d]) over the axes
    unsigned long long best = ~0ULL, stride = 1;
    for (int d = 0; d < N; ++d) {
      if (h[d] != ~0ULL) best = std::min(best, h[d] * stride);
      stride *= (unsigned long long)axis_len[d];
    }
    *first_oob = best == ~0ULL ? -1 : (int64_t)best;
  }
  return 0;
}

// Returns 0 and *done = true when the separable path served the call.
int grid_separable(int dev, const b200itp_desc *D, const void *coefs, const void *const *axis_coords,
                   const int64_t *axis_len, void *out_value, int64_t *first_oob, unsigned char *sp, size_t sbytes,
                   cudaStream_t st, bool *done) {
  *done = false;
  if (!grid_separable_scope(D)) return 0;
  if (sbytes < grid_separable_scratch_bytes(D, axis_len)) return 0;
  EvalParams P;
  bool w_wide = false, homogeneous = true;
  int rc = fill_eval_params_grid(D, coefs, axis_coords, P, &w_wide, &homogeneous);
  if (rc) return rc;
  if (!homogeneous) return 0;  // reported as unsupported by the pointwise path
  const bool cf32 = D->coef_dtype == B200ITP_F32, xf32 = D->coord_dtype == B200ITP_F32, wf32 = !w_wide;
  const bool of32 = !P.out_f64;
  if (!xf32 && wf32) return 0;  // (Float64 coordinates always compute Float64 weights)
#define B200_GRID_CASE(TC, TX, TW)                                                                                     \
  rc = of32 ? run_separable<TC, TX, TW, float>(dev, D, P, axis_len, out_value, first_oob, sp, st)                      \
            : run_separable<TC, TX, TW, double>(dev, D, P, axis_len, out_value, first_oob, sp, st)
  if (cf32 && xf32 && wf32) {
    B200_GRID_CASE(float, float, float);
  } else if (cf32 && xf32 && !wf32) {
    B200_GRID_CASE(float, float, double);
  } else if (cf32 && !xf32) {
    B200_GRID_CASE(float, double, double);
  } else if (!cf32 && !xf32) {
    B200_GRID_CASE(double, double, double);
  } else if (!cf32 && xf32 && wf32) {
    B200_GRID_CASE(double, float, float);
  } else {
    B200_GRID_CASE(double, float, double);
  }
#undef B200_GRID_CASE
  if (rc) return rc;
  *done = true;
  return 0;
}

}  // namespace b200itp
