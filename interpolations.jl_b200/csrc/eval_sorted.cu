// b200itp_eval_sorted: scattered evaluation with on-device query ordering.
//
// A uniformly scattered tricubic query touches 16 (y,z) rows of the coefficient array: 16-32 DRAM sectors
// when the array does not fit the L2, and 32 different L1 lines per warp instruction even when it does,
// so the direct kernel (eval.cuh) is bound by the L1 tag stage (~9 Gevals/s for 64 taps), not by HBM.
// Random 4-byte scatters of the results are no better (measured 23 G/s over a 4 GB window).  This path
// therefore moves the QUERIES instead, with streaming passes only (every pass reads and writes whole
// 128-byte lines; measured in tools/ubench_scatter.cu):
//
//   count    key(q) = (brick, bank class) of the query's first tap; histogram with L2 reductions
//   scan     exclusive prefix sum -> start offset of every (brick, class) stream
//   split 1  CTA-local multisplit of {x,y,z,q} records by brick GROUP   (<= 512 bins, runs of >= 8 records)
//   split 2  CTA-local multisplit inside each group by (brick, class)   (<= 512 bins)
//   evaluate persistent CTAs, one brick tile (+halo, periodic wrap resolved) in shared memory at a time;
//            lane c of every warp consumes the records of bank class c: the 64 shared-memory gathers of a
//            warp instruction then fall on 32 different banks by construction (conflict-free without
//            sorting the records inside the brick).  Results {q, value} leave in stream order.
//   unsort A multisplit of {q, value} by q >> 23, unsort B by q >> 14 (bins have exact, known sizes)
//   place    one CTA per window of 2^14 results: permute in shared memory, store coalesced to out[q]
//
// The per-query arithmetic is the device code of the direct kernel (axis_setup + the reference's reduction
// order, -fmad=false), so both paths return bit-identical values.
//
// Scope: Float32 coefficients and coordinates, 2-D / 3-D, uniform Quadratic or Cubic degree, at least 2^20 queries; bare
// BSplineInterpolation (values and gradients), and — values only — scale / extrapolate wrappers whose coordinate chain
// stays Float32 (Int or Float32 ranges, OnGrid bounds, flags Throw / Flat / Periodic / Reflect): their per-axis maps are
// applied when the records are made.  Everything else is forwarded to the direct kernel.
#include "eval.cuh"

#include <cuda.h>  // CUtensorMap and its enums only: the encoder is resolved at run time, libcuda is not linked
#include <algorithm>
#include <climits>
#include <cmath>
#include <type_traits>

namespace b200itp {

int fill_eval_params_public(const b200itp_desc *D, const void *coefs, const void *const *coords, int64_t nq,
                            void *out_value, unsigned long long *first_oob_dev, EvalParams &P, int *uniform_degree,
                            bool *plain_f32);

// ------------------------------------------------------------------------------------------ geometry
template <int N>
struct Brick;
template <>
struct Brick<3> {
  static constexpr int BX = 32, BY = 32, BZ = 16;
};
template <>
struct Brick<2> {
  static constexpr int BX = 128, BY = 128, BZ = 1;
};

template <int N, int DEG>
struct TileCfg {
  using Bk = Brick<N>;
  static constexpr int L = DEG + 1;
  static constexpr int TX = Bk::BX + DEG;
  static constexpr int TY = Bk::BY + DEG;
  static constexpr int TZ = N == 3 ? Bk::BZ + DEG : 1;
  // row pitch: a multiple of 4 floats with room for a shift of up to 3, so that the tile columns whose global addresses
  // are 16-byte aligned sit on 16-byte aligned shared-memory addresses (OrdGeom::tshift) and rows are staged with
  // 16-byte asynchronous copies.  Any pitch keeps the gathers conflict-free: lane c's 64 taps sit at fixed offsets from
  // a base whose bank is c.
  static constexpr int PX = (TX + 3 + 3) / 4 * 4;
  static constexpr int FLOATS = PX * TY * TZ;
  static constexpr int MAXT = TX > TY ? (TX > TZ ? TX : TZ) : (TY > TZ ? TY : TZ);
};

constexpr int kClasses = 32;        // bank classes == lanes
constexpr int kGroupBricks = 16;    // bricks per split-1 group: split 2 has 16 * 32 = 512 bins
constexpr int kMaxBins = 512;       // bins per multisplit level (== threads of the split kernels)
#ifndef ORD_SPLIT_THREADS
#define ORD_SPLIT_THREADS 512
#endif
constexpr int kSplitThreads = ORD_SPLIT_THREADS;  // >= kMaxBins (threads 0..kMaxBins-1 own one bin each)
#ifndef ORD_PREFETCH
#define ORD_PREFETCH 1  // fetch the next chunk into registers while the current one goes through shared memory
#endif
constexpr bool kPrefetch = ORD_PREFETCH != 0;
constexpr int kSplitMinCtas = (ORD_PREFETCH || ORD_SPLIT_THREADS > 512) ? 1 : 2;
constexpr int kChunk16 = 4096;      // records per multisplit chunk, 16-byte records
constexpr int kChunk8 = 8192;       // 8-byte records
// Unsort of the results: values travel as {q, value} (8 bytes), gradients as {q, g1, g2, g3} (16 bytes).  Final windows
// of 2^SB results are permuted in shared memory (values: 2^14 x 4 B = 64 KB; gradients: 2^13 x N x 4 B <= 96 KB); level A
// bins hold 512 windows each.
#ifndef ORD_SB_VAL
#define ORD_SB_VAL 14
#endif
#ifndef ORD_SAB
#define ORD_SAB 8  // level-B bins per level-A bin: 2^8 (measured 9 -> 8: unsort A + B 8.25 -> 8.04 ms per 1e9 results)
#endif
template <bool GRAD>
struct Uns {
  using Rec = typename std::conditional<GRAD, uint4, uint2>::type;
  static constexpr int S = GRAD ? kChunk16 : kChunk8;
  static constexpr int SB = GRAD ? 13 : ORD_SB_VAL;
  static constexpr int SA = SB + (GRAD ? 9 : ORD_SAB);
};
#ifndef ORD_EVAL_THREADS
#define ORD_EVAL_THREADS 768
#endif
// (measured, round 2: 512 / 640 / 768 threads and a variant that fetches the 16 taps of the next x-slice ahead of the
//  arithmetic all land at 11.7-12.2 ms per 1e9 tricubic queries — 64 LDS at half issue rate + 263 other instructions
//  per record put the floor at ~3.1 SM cycles per record, the kernel runs at 3.3)
constexpr int kEvalThreads = ORD_EVAL_THREADS;
#ifndef ORD_TILE_PREFETCH
#define ORD_TILE_PREFETCH 0  // measured: no gain (ord_eval 2.13 -> 2.18 ms at 1.25e8 queries), see DESIGN 4.3
#endif
constexpr int kRB = 8;              // rows (records per class) a warp stages and evaluates at a time
#ifndef ORD_TAIL_RB
#define ORD_TAIL_RB 8
#endif
#ifndef ORD_TAIL_ROWS
#define ORD_TAIL_ROWS 0
#endif
constexpr int kTailRB = ORD_TAIL_RB;      // rows per block at the end of an item (8 == no finer tail)
constexpr int kTailRows = ORD_TAIL_ROWS;  // rows of an item handed out in tail blocks
constexpr int kRBP = kRB + 1;       // pitch in records: lane c reads slot c * 9 + r -> conflict-free 128-bit accesses
// Evaluation work items: a brick's records are cut into parts of at most `item target` records.  Large items keep the
// per-item overhead and the end-of-item tail small (measured: 32 Ki -> 128 Ki records per item = -10 % evaluation time);
// the target shrinks when there are few records per SM so that every CTA still gets several items.
constexpr unsigned kItemTargetMax = 262144, kItemTargetMin = 16384;

struct OrdGeom {
  float lbf[3], ubf[3];  // bounds as Float32 thresholds: (double)x >= lb <=> x >= lbf, (double)x <= ub <=> x <= ubf
  int nb[3];     // bricks per axis
  int cells[3];  // distinct first-tap indices per axis
  int fmin[3];   // smallest first-tap index (parent index space): -1 for periodic cubic axes (OnCell, x < 1), else 0
  int nbricks;
  int ngroups;   // groups of kGroupBricks bricks (split 1 bins)
  int nsuper;    // super-groups of kMaxBins groups: > 1 adds a split level in front (arrays beyond 8192 bricks)
  int nfinal;    // nbricks * kClasses
  int wrapped;   // scale / extrapolation maps are applied when the records are made (check_query)
  int tshift;    // shared-memory column of tile coordinate 0 on axis 1 (0..3), see TileCfg::PX
  int tma;       // brick tiles that need no periodic wrap are fetched by ONE tensor-map bulk copy (run_ordered)
};

// First tap index of one axis: the `first` of axis_setup (eval.cuh; cubic.jl:40-57, quadratic.jl:39-43,
// indexing.jl:172-187) computed with ONE float->int conversion (F2I.FLOOR / F2I.CEIL run on the quarter-rate
// conversion pipe; floorf + (int) would take two trips).  For every x inside the bounds the integer results equal
// (int)floorf(.) / (int)ceilf(.) of the same float expression, so the index is identical to the evaluation's.
template <int DEG>
__device__ __forceinline__ int first_tap(const EvalParams &P, int d, float x) {
  const int n = P.n[d];
  int f;
  if (DEG == B200ITP_CUBIC) {
    if (P.periodic[d]) {   // uniform branch
      f = __float2int_rd(x);
    } else {
      f = __float2int_rd(x + (x < 1.f ? 0.5f : 0.f));
      f = f > n - 1 ? f - 1 : f;
    }
  } else {  // quadratic: n + 0.5 is exact in Float32 (n < 2^22), so the Float64 comparison of axis_setup is this one
    const float xh = x + 0.5f;
    f = x < (float)n + 0.5f ? __float2int_rd(xh) : __float2int_ru(xh) - 1;
  }
  return (x == x) ? f - 1 : 0;  // NaN record -> 0
}

// (brick, class) of a query
template <int N, int DEG>
__device__ __forceinline__ void locate(const EvalParams &P, const OrdGeom &G, const float (&x)[3], int &brick, int &cls) {
  using Bk = Brick<N>;
  using TC = TileCfg<N, DEG>;
  int c[3] = {0, 0, 0};
#pragma unroll
  for (int d = 0; d < N; ++d) c[d] = min(max(first_tap<DEG>(P, d, x[d]) - G.fmin[d], 0), G.cells[d] - 1);
  const int bx = c[0] / Bk::BX, by = c[1] / Bk::BY, bz = N == 3 ? c[2] / Bk::BZ : 0;
  brick = (bz * G.nb[1] + by) * G.nb[0] + bx;
  const int lx = c[0] - bx * Bk::BX, ly = c[1] - by * Bk::BY, lz = N == 3 ? c[2] - bz * Bk::BZ : 0;
  cls = ((lz * TC::TY + ly) * TC::PX + lx) & (kClasses - 1);
}

// A query outside the bounds (or NaN) is a BoundsError (indexing.jl:6): it travels through the pipeline as
// an all-NaN record (brick 0, class 0, NaN result) so that every q appears exactly once in the unsort.
template <int N, bool WRAP>
__device__ __forceinline__ void check_query(const EvalParams &P, const OrdGeom &G, float (&x)[3], bool &valid) {
  valid = true;
  if constexpr (WRAP) {  // compile-time: the bare interpolant's count / split kernels carry none of this
    // extrapolate(scale(itp, ranges...), flags) whose whole coordinate chain stays Float32 (Int / Float32 ranges, OnGrid
    // bounds, no Line): the per-axis maps of the direct kernel (eval.cuh: etp_inbounds, coordlookup + clamp;
    // extrapolation.jl:106-163, scaling.jl:102-115) are applied HERE, once, and the record carries the coordinate in
    // the B-spline's own index space — from then on the pipeline is the bare interpolant's
#pragma unroll
    for (int d = 0; d < N; ++d) {
      float xs = x[d];
      if (P.etp_kind == B200ITP_ETP_FLAGS) {
        float y;
        if (etp_inbounds<float>(P, d, x[d], y)) valid = false;
        xs = y;
      } else {
        valid = valid && (G.lbf[d] <= x[d]) && (x[d] <= G.ubf[d]);
      }
      if (P.has_scale) {
        xs = (xs - (float)P.rfirst[d]) / (float)P.rstep[d] + 1.f;
        const float lo = (float)P.inner_lb[d], hi = (float)P.inner_ub[d];
        xs = xs > hi ? hi : (xs < lo ? lo : xs);
      }
      x[d] = xs;
    }
  } else {
#pragma unroll
    for (int d = 0; d < N; ++d) valid = valid && (G.lbf[d] <= x[d]) && (x[d] <= G.ubf[d]);  // false for NaN
  }
  if (!valid) {
    const float nn = __int_as_float(0x7fc00000);
    x[0] = nn;
    x[1] = nn;
    if (N == 3) x[2] = nn;
  }
}
// ------------------------------------------------------------------------------------------ count
template <int N, int DEG, bool WRAP>
__global__ void __launch_bounds__(256) ord_count_kernel(const __grid_constant__ EvalParams P, const OrdGeom G,
                                                        unsigned *__restrict__ hist) {
  constexpr int U = 4;  // queries per thread per iteration: their loads are in flight together
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long q0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; q0 < P.nq; q0 += U * stride) {
    float x[U][3];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long q = q0 + u * stride;
#pragma unroll
      for (int d = 0; d < 3; ++d)
        x[u][d] = (d < N && q < P.nq) ? reinterpret_cast<const float *>(P.coords[d])[q] : 0.f;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long q = q0 + u * stride;
      if (q < P.nq) {
        bool valid;
        check_query<N, WRAP>(P, G, x[u], valid);
        if (!valid) atomicMin(P.first_oob, (unsigned long long)q);
        int brick, cls;
        locate<N, DEG>(P, G, x[u], brick, cls);
        atomicAdd(hist + (brick * kClasses + cls), 1u);  // result unused: compiles to RED
      }
    }
  }
}

// Two-level geometry (at most kMaxBins groups): the first split only needs the group sizes, and a group histogram fits
// shared memory, so this pass runs at the HBM rate (12 bytes per query) instead of the L2's atomic rate (one global
// reduction per query: 6.2 ms per 1e9).  The (brick, class) histogram is accumulated by the first split itself
// (ord_split1_kernel<.., COUNT = true>), whose shared-memory-bound chunks hide the reductions.
#ifndef ORD_CG_U
#define ORD_CG_U 1
#endif
#ifndef ORD_RED_LATE
#define ORD_RED_LATE 0  // 1: the first split issues its reductions after the chunk's copy-out (measured: 9.3 -> 10.5 ms)
#endif
#ifndef ORD_FINE_SHARE
#define ORD_FINE_SHARE 2
#endif
constexpr int kFineShare = ORD_FINE_SHARE;  // of every 8 queries: (brick, class) reduction issued by the group-count pass
#ifndef ORD_CG_THREADS
#define ORD_CG_THREADS 256
#endif
#ifndef ORD_CG_MULT
#define ORD_CG_MULT 32
#endif
template <int N, int DEG, bool WRAP>
__global__ void __launch_bounds__(ORD_CG_THREADS) ord_count_groups_kernel(const __grid_constant__ EvalParams P, const OrdGeom G,
                                                               unsigned *__restrict__ ghist,
                                                               unsigned *__restrict__ fine_hist) {
  __shared__ unsigned sh[kMaxBins];
  for (int i = threadIdx.x; i < kMaxBins; i += blockDim.x) sh[i] = 0;
  __syncthreads();
  auto one = [&](float x0, float x1, float x2, long long q) {
    float x[3] = {x0, x1, x2};
    bool valid;
    check_query<N, WRAP>(P, G, x, valid);
    if (!valid) atomicMin(P.first_oob, (unsigned long long)q);
    int brick, cls;
    locate<N, DEG>(P, G, x, brick, cls);
    atomicAdd(&sh[brick / kGroupBricks], 1u);
    // this pass is issue-bound with an idle LSU: it takes the (brick, class) reductions of kFineShare of every 8
    // queries off the first split, whose shared-memory traffic they would queue behind
    if (kFineShare > 0 && ((unsigned)q & 7u) < (unsigned)kFineShare) atomicAdd(fine_hist + (brick * kClasses + cls), 1u);
  };
  const float *c0 = reinterpret_cast<const float *>(P.coords[0]);
  const float *c1 = reinterpret_cast<const float *>(P.coords[1]);
  const float *c2 = N == 3 ? reinterpret_cast<const float *>(P.coords[2]) : c0;
  const bool vec = ((reinterpret_cast<uintptr_t>(c0) | reinterpret_cast<uintptr_t>(c1) | reinterpret_cast<uintptr_t>(c2)) & 15) == 0;
  const long long nvec = vec ? P.nq / 4 : 0;  // four consecutive queries per thread: 128-bit loads
  constexpr int U = ORD_CG_U;                 // vectors in flight per thread
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long v0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; v0 < nvec; v0 += U * stride) {
    float4 a[U], b[U], c[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long v = v0 + u * stride;
      if (v < nvec) {
        a[u] = reinterpret_cast<const float4 *>(c0)[v];
        b[u] = reinterpret_cast<const float4 *>(c1)[v];
        c[u] = N == 3 ? reinterpret_cast<const float4 *>(c2)[v] : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long v = v0 + u * stride;
      if (v < nvec) {
        one(a[u].x, b[u].x, c[u].x, 4 * v);
        one(a[u].y, b[u].y, c[u].y, 4 * v + 1);
        one(a[u].z, b[u].z, c[u].z, 4 * v + 2);
        one(a[u].w, b[u].w, c[u].w, 4 * v + 3);
      }
    }
  }
  for (long long q = 4 * nvec + (long long)blockIdx.x * blockDim.x + threadIdx.x; q < P.nq; q += stride)
    one(c0[q], c1[q], N == 3 ? c2[q] : 0.f, q);
  __syncthreads();
  for (int i = threadIdx.x; i < G.ngroups; i += blockDim.x)
    if (sh[i]) atomicAdd(ghist + i, sh[i]);
}

// ------------------------------------------------------------------------------------------ scan (3 kernels)
constexpr int kScanBlock = 1024;
constexpr int kScanPer = 4;
__global__ void __launch_bounds__(kScanBlock) scan_partial_kernel(const unsigned *__restrict__ hist, long long n,
                                                                   unsigned *__restrict__ sums) {
  __shared__ unsigned red[32];
  const long long base = (long long)blockIdx.x * kScanBlock * kScanPer;
  unsigned s = 0;
#pragma unroll
  for (int k = 0; k < kScanPer; ++k) {
    const long long i = base + (long long)k * kScanBlock + threadIdx.x;
    if (i < n) s += hist[i];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    unsigned t = red[threadIdx.x];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_down_sync(0xffffffffu, t, o);
    if (threadIdx.x == 0) sums[blockIdx.x] = t;
  }
}
__global__ void __launch_bounds__(1024) scan_sums_kernel(unsigned *sums, int nblocks) {
  __shared__ unsigned buf[1024];
  __shared__ unsigned carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < nblocks; base += 1024) {
    const int i = base + threadIdx.x;
    const unsigned v = i < nblocks ? sums[i] : 0u;
    buf[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
      const unsigned t = threadIdx.x >= o ? buf[threadIdx.x - o] : 0u;
      __syncthreads();
      buf[threadIdx.x] += t;
      __syncthreads();
    }
    if (i < nblocks) sums[i] = carry + buf[threadIdx.x] - v;
    __syncthreads();
    if (threadIdx.x == 0) carry += buf[1023];
    __syncthreads();
  }
}
// start[i] = exclusive prefix (n + 1 entries); hist[i] becomes the split-2 cursor of stream i
__global__ void __launch_bounds__(kScanBlock) scan_final_kernel(unsigned *hist, long long n,
                                                                 const unsigned *__restrict__ sums,
                                                                 unsigned *__restrict__ start) {
  __shared__ unsigned wsum[32];
  const long long base = (long long)blockIdx.x * kScanBlock * kScanPer;
  unsigned v[kScanPer];
  unsigned tsum = 0;
#pragma unroll
  for (int k = 0; k < kScanPer; ++k) {
    const long long i = base + (long long)threadIdx.x * kScanPer + k;
    v[k] = i < n ? hist[i] : 0u;
    tsum += v[k];
  }
  unsigned inc = tsum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned t = __shfl_up_sync(0xffffffffu, inc, o);
    if ((threadIdx.x & 31) >= o) inc += t;
  }
  if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = inc;
  __syncthreads();
  if (threadIdx.x < 32) {
    unsigned w = wsum[threadIdx.x];
    unsigned winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned t = __shfl_up_sync(0xffffffffu, winc, o);
      if (threadIdx.x >= o) winc += t;
    }
    wsum[threadIdx.x] = winc - w;
  }
  __syncthreads();
  unsigned run = sums[blockIdx.x] + wsum[threadIdx.x >> 5] + inc - tsum;
#pragma unroll
  for (int k = 0; k < kScanPer; ++k) {
    const long long i = base + (long long)threadIdx.x * kScanPer + k;
    if (i < n) {
      start[i] = run;
      hist[i] = run;
    }
    run += v[k];
    if (i == n - 1) start[n] = run;
  }
}

// ------------------------------------------------------------------------------------------ plan
// One CTA: cursors of split 1 (per group), chunk table of split 2 (per group), work items of the
// evaluation (per brick), cursors of the two unsort levels (exact bin sizes).
struct OrdPlan {
  unsigned *cursor0;      // nsuper  (three-level geometry only)
  unsigned *chunk_mid;    // nsuper + 1 : chunks of the middle split, per super-group
  unsigned *cursor1;      // ngroups
  unsigned *chunk_start;  // ngroups + 1 : split-2 chunks
  unsigned *item_start;   // nbricks + 1 : evaluation work items
  unsigned *cursorA;      // ceil(nq / 2^SA)
  unsigned *cursorB;      // ceil(nq / 2^SB)
};

// Two-level geometry, before the first split: group cursors and the split-2 chunk table from the group histogram
// (ngroups <= kMaxBins <= 1024: one tile), plus the unsort cursors.
__global__ void __launch_bounds__(1024) ord_plan_groups_kernel(const unsigned *__restrict__ ghist, const OrdGeom G,
                                                               long long nq, int shiftA, int shiftB, OrdPlan pl) {
  __shared__ unsigned buf[1024];
  const int t = threadIdx.x;
  const long long na = (nq + (1LL << shiftA) - 1) >> shiftA, nbw = (nq + (1LL << shiftB) - 1) >> shiftB;
  for (long long i = t; i < na; i += 1024) pl.cursorA[i] = (unsigned)(i << shiftA);
  for (long long i = t; i < nbw; i += 1024) pl.cursorB[i] = (unsigned)(i << shiftB);
  const unsigned cnt = t < G.ngroups ? ghist[t] : 0u;
  for (int pass = 0; pass < 2; ++pass) {
    const unsigned v = pass == 0 ? cnt : (cnt + kChunk16 - 1) / kChunk16;
    __syncthreads();
    buf[t] = v;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
      const unsigned u = t >= o ? buf[t - o] : 0u;
      __syncthreads();
      buf[t] += u;
      __syncthreads();
    }
    unsigned *dst = pass == 0 ? pl.cursor1 : pl.chunk_start;
    if (t < G.ngroups) dst[t] = buf[t] - v;
    if (pass == 1 && t == 1023) dst[G.ngroups] = buf[t];
  }
}

// items_only: the two-level path calls this after the first split, for the work-item table alone
__global__ void __launch_bounds__(1024) ord_plan_kernel(const unsigned *__restrict__ start, const OrdGeom G,
                                                        long long nq, int ctas, int shiftA, int shiftB, OrdPlan pl,
                                                        bool items_only) {
  __shared__ unsigned buf[1024];
  __shared__ unsigned carry;
  const int t = threadIdx.x;
  constexpr long long FPG = (long long)kGroupBricks * kClasses;  // final bins per group
  if (!items_only) {
    // unsort cursors
    const long long na = (nq + (1LL << shiftA) - 1) >> shiftA, nbw = (nq + (1LL << shiftB) - 1) >> shiftB;
    for (long long i = t; i < na; i += 1024) pl.cursorA[i] = (unsigned)(i << shiftA);
    for (long long i = t; i < nbw; i += 1024) pl.cursorB[i] = (unsigned)(i << shiftB);
    // split cursors: first record of every group / super-group
    for (int g = t; g < G.ngroups; g += 1024) pl.cursor1[g] = start[min((long long)g * FPG, (long long)G.nfinal)];
    if (G.nsuper > 1)
      for (int sg = t; sg < G.nsuper; sg += 1024) pl.cursor0[sg] = start[min((long long)sg * kMaxBins * FPG, (long long)G.nfinal)];
  }
  // exclusive scans: chunks per group, items per brick, chunks per super-group
  for (int pass = items_only ? 1 : 0; pass < (items_only ? 2 : (G.nsuper > 1 ? 3 : 2)); ++pass) {
    const int n = pass == 0 ? G.ngroups : (pass == 1 ? G.nbricks : G.nsuper);
    unsigned *dst = pass == 0 ? pl.chunk_start : (pass == 1 ? pl.item_start : pl.chunk_mid);
    if (t == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < n; base += 1024) {
      const int i = base + t;
      unsigned v = 0;
      if (i < n) {
        if (pass == 0) {
          const long long a = (long long)i * kGroupBricks * kClasses;
          const long long b = min((long long)(i + 1) * kGroupBricks * kClasses, (long long)G.nfinal);
          const unsigned cnt = start[b] - start[a];
          v = (cnt + kChunk16 - 1) / kChunk16;
        } else if (pass == 1) {
          const unsigned cnt = start[(long long)(i + 1) * kClasses] - start[(long long)i * kClasses];
          const unsigned target = min(kItemTargetMax, max(kItemTargetMin, (unsigned)(nq / (8LL * ctas))));
          v = (cnt + target - 1) / target;
        } else {
          const long long a = min((long long)i * kMaxBins * FPG, (long long)G.nfinal);
          const long long b = min((long long)(i + 1) * kMaxBins * FPG, (long long)G.nfinal);
          const unsigned cnt = start[b] - start[a];
          v = (cnt + kChunk16 - 1) / kChunk16;
        }
      }
      buf[t] = v;
      __syncthreads();
      for (int o = 1; o < 1024; o <<= 1) {
        const unsigned u = t >= o ? buf[t - o] : 0u;
        __syncthreads();
        buf[t] += u;
        __syncthreads();
      }
      if (i < n) dst[i] = carry + buf[t] - v;
      __syncthreads();
      if (t == 0) carry += buf[1023];
      __syncthreads();
    }
    if (t == 0) dst[n] = carry;
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------ multisplit
// One chunk of S records, held in registers (RPT per thread), is counting-sorted by key in shared memory
// and appended bin by bin to the global bins: one global atomic per (chunk, non-empty bin), stores of whole
// runs.  key == 0xffffffff marks an absent record.
// Ranking inside a chunk: shared-memory atomics (default) or warp votes (ORD_RANK_BALLOT=1: one ballot per key bit gives
// every lane the mask of its peers, the lowest peer bumps a warp-private counter).  Measured on B200
// (tools/ubench_misc.cu, profiles/r2_ubench_misc.txt): ATOMS ranking costs 2.6 SM cycles per 32 records, the vote
// variant 25.6 (VOTE + SHFL are the slow instructions here), and the whole pipeline runs 42.7 vs 53.2 ms per 1e9
// queries.  The split kernels are bound by the barrier-separated phases of a chunk, not by the atomics.
#ifndef ORD_RANK_BALLOT
#define ORD_RANK_BALLOT 0
#endif
constexpr int kKeyBits = 9;  // kMaxBins == 512
static_assert((1 << kKeyBits) == kMaxBins, "key bits");

template <typename Rec, int S>
struct SplitSmem {
  Rec rec[S];  // records of the chunk in bin order
  unsigned short bin[S];
  unsigned gbase[kMaxBins];
#if ORD_RANK_BALLOT
  unsigned short whist[kSplitThreads / 32][kMaxBins];  // per warp: count, then first slot, of every bin
#else
  unsigned hist[kMaxBins];
  unsigned start[kMaxBins];
#endif
  unsigned wtot[32];
};

#if ORD_RANK_BALLOT
// FULL: the chunk holds exactly S records (every chunk but the last of a bin/array): no per-record predicates, and the
// copy-out loop has a constant trip count, so its shared-memory look-ups are issued back to back.
template <bool FULL, typename Rec, int S>
__device__ __forceinline__ void block_split(SplitSmem<Rec, S> &sm, const Rec (&r)[S / kSplitThreads],
                                            const unsigned (&key)[S / kSplitThreads], int nbins, unsigned cnt,
                                            unsigned *__restrict__ cursor, Rec *__restrict__ out) {
  // sm.whist[warp][*] is all zero on entry (zeroed by split_init / by the warp itself at the end of the previous call)
  constexpr int RPT = S / kSplitThreads;
  constexpr int NW = kSplitThreads / 32;
  static_assert(S <= 65536, "ranks are 16-bit");
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const unsigned lt = (1u << lane) - 1u;
  unsigned short *wh = sm.whist[warp];
  unsigned kr[RPT];  // key | (rank among the warp's records of that key) << 16
#pragma unroll
  for (int j = 0; j < RPT; ++j) {
    const unsigned k = key[j];
    const bool valid = FULL || k != 0xffffffffu;
    unsigned peers = FULL ? 0xffffffffu : __ballot_sync(0xffffffffu, valid);
#pragma unroll
    for (int b = 0; b < kKeyBits; ++b) {
      const unsigned bit = (k >> b) & 1u;
      const unsigned m = __ballot_sync(0xffffffffu, bit);
      peers &= m ^ (bit - 1u);  // bit set: lanes with the bit; clear: lanes without
    }
    const unsigned below = __popc(peers & lt);
    unsigned base = 0;
    if (valid && below == 0) {  // lowest peer: bump the warp's counter of this key
      base = wh[k];
      wh[k] = (unsigned short)(base + __popc(peers));
    }
    __syncwarp();
    base = __shfl_sync(0xffffffffu, base, __ffs(peers) - 1);  // (invalid lanes: peers may be 0 -> lane 31's value, unused)
    kr[j] = valid ? (k | ((base + below) << 16)) : 0xffffffffu;
  }
  __syncthreads();  // (1) warp counters complete; the previous chunk's copy-out has finished in every warp
  // bin owner t: total of its bin, exclusive scan over the bins, then the first slot of every warp's share
  const bool owner = t < kMaxBins;
  unsigned v = 0;
  if (owner && t < nbins) {
#pragma unroll
    for (int w = 0; w < NW; ++w) v += sm.whist[w][t];
  }
  unsigned inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned u = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += u;
  }
  if (lane == 31) sm.wtot[warp] = inc;
  // the bin reservation's round trip to the L2 overlaps the scan and the shared-memory scatter below
  unsigned gb = 0;
  if (v) gb = atomicAdd(cursor + t, v);
  __syncthreads();  // (2)
  unsigned wv = lane < NW ? sm.wtot[lane] : 0u, winc = wv;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned u = __shfl_up_sync(0xffffffffu, winc, o);
    if (lane >= o) winc += u;
  }
  const unsigned wbase = __shfl_sync(0xffffffffu, winc - wv, warp);
  const unsigned st = wbase + inc - v;
  if (owner && t < nbins) {
    unsigned run = st;  // (re-read rather than 16 live registers across the barrier)
#pragma unroll
    for (int w = 0; w < NW; ++w) {
      const unsigned cw = sm.whist[w][t];
      sm.whist[w][t] = (unsigned short)run;
      run += cw;
    }
  }
  __syncthreads();  // (3) first slots visible
#pragma unroll
  for (int j = 0; j < RPT; ++j)
    if (FULL || kr[j] != 0xffffffffu) {
      const unsigned k = kr[j] & 0xffffu;
      const unsigned slot = wh[k] + (kr[j] >> 16);
      sm.rec[slot] = r[j];
      sm.bin[slot] = (unsigned short)k;
    }
  asm volatile("" : "+r"(gb));
  if (owner) sm.gbase[t] = gb - st;
  __syncthreads();  // (4) records in bin order
  // the warp's counters are only touched by the warp itself until barrier (1) of the next call: zero them now
  {
    unsigned *w32 = reinterpret_cast<unsigned *>(wh);
#pragma unroll
    for (int i = 0; i < kMaxBins / 64; ++i) w32[i * 32 + lane] = 0u;
  }
  // stores of whole runs, consecutive threads on consecutive records; no barrier after them: the next call touches
  // only its own warp's counters before its barrier (1)
  if (FULL) {
#pragma unroll 8
    for (int j = 0; j < RPT; ++j) {
      const unsigned i = (unsigned)j * kSplitThreads + t;
      out[sm.gbase[sm.bin[i]] + i] = sm.rec[i];
    }
  } else {
    for (unsigned i = t; i < cnt; i += kSplitThreads) out[sm.gbase[sm.bin[i]] + i] = sm.rec[i];
  }
  __syncwarp();
}

template <typename Rec, int S>
__device__ __forceinline__ void split_init(SplitSmem<Rec, S> &sm) {
  unsigned *w32 = reinterpret_cast<unsigned *>(&sm.whist[0][0]);
  for (int i = threadIdx.x; i < (kSplitThreads / 32) * kMaxBins / 2; i += kSplitThreads) w32[i] = 0u;
  __syncthreads();
}
#else
// FULL: the chunk holds exactly S records (every chunk but the last of a bin/array): no per-record predicates, and the
// copy-out loop has a constant trip count, so its shared-memory look-ups are issued back to back.
template <bool FULL, typename Rec, int S>
__device__ __forceinline__ void block_split(SplitSmem<Rec, S> &sm, const Rec (&r)[S / kSplitThreads],
                                            const unsigned (&key)[S / kSplitThreads], int nbins, unsigned cnt,
                                            unsigned *__restrict__ cursor, Rec *__restrict__ out) {
  // sm.hist is all zero on entry (zeroed by split_init / by the previous call right after it was read)
  constexpr int RPT = S / kSplitThreads;
  constexpr int NW = kSplitThreads / 32;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  // 16-byte records: key | rank << 16 in one register across the barriers (ranks < S <= 2^16, keys < 2^16); 8-byte
  // records (16 per thread) were measured faster with the two values in separate registers
  constexpr bool PACK = sizeof(Rec) == 16;
  unsigned kr[RPT], rk[PACK ? 1 : RPT];
#pragma unroll
  for (int j = 0; j < RPT; ++j) {
    kr[j] = 0xffffffffu;
    if (FULL || key[j] != 0xffffffffu) {
      const unsigned rank = atomicAdd(&sm.hist[key[j]], 1u);
      if (PACK) {
        kr[j] = key[j] | (rank << 16);
      } else {
        kr[j] = key[j];
        rk[j] = rank;
      }
    }
  }
  __syncthreads();  // (1) histogram complete; the previous chunk's copy-out has finished in every warp
  // exclusive scan of the counters, one per thread: warp scan, then every warp scans the NW warp totals itself
  const bool owner = t < kMaxBins;
  const unsigned v = (owner && t < nbins) ? sm.hist[t] : 0u;
  if (owner) sm.hist[t] = 0;  // ready for the next chunk (its atomics come after barrier (3))
  unsigned inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned u = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += u;
  }
  if (lane == 31) sm.wtot[warp] = inc;
  __syncthreads();  // (2)
  unsigned wv = lane < NW ? sm.wtot[lane] : 0u, winc = wv;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned u = __shfl_up_sync(0xffffffffu, winc, o);
    if (lane >= o) winc += u;
  }
  const unsigned wbase = __shfl_sync(0xffffffffu, winc - wv, warp);
  const unsigned st = wbase + inc - v;
  if (owner) sm.start[t] = st;
  // the bin reservation's round trip to the L2 overlaps the shared-memory scatter below; the empty asm keeps the
  // compiler from waiting for it earlier
  unsigned gb = 0;
  if (v) gb = atomicAdd(cursor + t, v);
  __syncthreads();  // (3) start[] visible
#pragma unroll
  for (int j = 0; j < RPT; ++j)
    if (FULL || kr[j] != 0xffffffffu) {
      const unsigned k = PACK ? (kr[j] & 0xffffu) : kr[j];
      const unsigned slot = sm.start[k] + (PACK ? (kr[j] >> 16) : rk[PACK ? 0 : j]);
      sm.rec[slot] = r[j];
      sm.bin[slot] = (unsigned short)k;
    }
  asm volatile("" : "+r"(gb));
  if (owner) sm.gbase[t] = gb - st;
  __syncthreads();  // (4) records in bin order
  // stores of whole runs, consecutive threads on consecutive records; no barrier after them: the next call touches
  // only hist[] before its barrier (1)
  if (FULL) {
#pragma unroll 8
    for (int j = 0; j < RPT; ++j) {
      const unsigned i = (unsigned)j * kSplitThreads + t;
      out[sm.gbase[sm.bin[i]] + i] = sm.rec[i];
    }
  } else {
    for (unsigned i = t; i < cnt; i += kSplitThreads) out[sm.gbase[sm.bin[i]] + i] = sm.rec[i];
  }
}

template <typename Rec, int S>
__device__ __forceinline__ void split_init(SplitSmem<Rec, S> &sm) {
  if (threadIdx.x < kMaxBins) sm.hist[threadIdx.x] = 0;
  __syncthreads();
}
#endif

// split 1: coordinates -> {x,y,z,q} records grouped by brick group
// SUPER: three-level geometry, the first pass splits by super-group (kMaxBins groups)
// COUNT: also accumulate the (brick, class) histogram (two-level geometry, see ord_count_groups_kernel)
template <int N, int DEG, bool SUPER, bool WRAP, bool COUNT>
__global__ void __launch_bounds__(kSplitThreads, kSplitMinCtas) ord_split1_kernel(const __grid_constant__ EvalParams P,
                                                                   const OrdGeom G, unsigned *__restrict__ cursor1,
                                                                   float4 *__restrict__ rec1,
                                                                   unsigned *__restrict__ fine_hist) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  auto &sm = *reinterpret_cast<SplitSmem<float4, kChunk16> *>(smem_raw);
  split_init(sm);
  constexpr int RPT = kChunk16 / kSplitThreads;
  const unsigned nq = (unsigned)P.nq;  // < 2^32 (checked by the host)
  const unsigned nchunks = (nq + kChunk16 - 1) / kChunk16;
  const float *cx = reinterpret_cast<const float *>(P.coords[0]);
  const float *cy = reinterpret_cast<const float *>(P.coords[1]);
  const float *cz = N == 3 ? reinterpret_cast<const float *>(P.coords[2]) : cx;
  // the coordinates of the next chunk are fetched before the current chunk goes through shared memory
  float nx[RPT][3];
  auto fetch = [&](unsigned ch) {
    if (ch >= nchunks) return;
    const unsigned q0 = ch * kChunk16 + threadIdx.x;
    if (q0 - threadIdx.x + kChunk16 <= nq) {
#pragma unroll
      for (int j = 0; j < RPT; ++j) {
        nx[j][0] = cx[q0 + j * kSplitThreads];
        nx[j][1] = cy[q0 + j * kSplitThreads];
        nx[j][2] = N == 3 ? cz[q0 + j * kSplitThreads] : 0.f;
      }
    } else {
#pragma unroll
      for (int j = 0; j < RPT; ++j) {
        const unsigned q = q0 + j * kSplitThreads;
        nx[j][0] = q < nq ? cx[q] : 0.f;
        nx[j][1] = q < nq ? cy[q] : 0.f;
        nx[j][2] = (N == 3 && q < nq) ? cz[q] : 0.f;
      }
    }
  };
  auto body = [&](auto full_tag, unsigned ch) {
    constexpr bool FULL = decltype(full_tag)::value;
    const unsigned base = ch * kChunk16;
    const unsigned cnt = FULL ? (unsigned)kChunk16 : nq - base;
    float4 r[RPT];
    unsigned key[RPT];
#if ORD_RED_LATE
    unsigned fk[RPT];
#pragma unroll
    for (int j = 0; j < RPT; ++j) fk[j] = 0xffffffffu;
#endif
#pragma unroll
    for (int j = 0; j < RPT; ++j) {
      const unsigned q = base + j * kSplitThreads + threadIdx.x;
      key[j] = 0xffffffffu;
      if (FULL || q < nq) {
        float x[3] = {nx[j][0], nx[j][1], nx[j][2]};
        bool valid;
        check_query<N, WRAP>(P, G, x, valid);
        int brick, cls;
        locate<N, DEG>(P, G, x, brick, cls);
        if constexpr (COUNT) {  // result unused: RED (the group-count pass took the first kFineShare of every 8)
#if ORD_RED_LATE
          fk[j] = ((q & 7u) >= (unsigned)kFineShare) ? (unsigned)(brick * kClasses + cls) : 0xffffffffu;
#else
          if ((q & 7u) >= (unsigned)kFineShare) atomicAdd(fine_hist + (brick * kClasses + cls), 1u);
#endif
        }
        key[j] = SUPER ? (unsigned)(brick / (kGroupBricks * kMaxBins)) : (unsigned)(brick / kGroupBricks);
        r[j] = make_float4(x[0], x[1], x[2], __uint_as_float(q));
      }
    }
    if (kPrefetch) fetch(ch + gridDim.x);
    block_split<FULL, float4, kChunk16>(sm, r, key, SUPER ? G.nsuper : G.ngroups, cnt, cursor1, rec1);
#if ORD_RED_LATE
    if constexpr (COUNT) {
#pragma unroll
      for (int j = 0; j < RPT; ++j)
        if (fk[j] != 0xffffffffu) atomicAdd(fine_hist + fk[j], 1u);
    }
#endif
  };
  if (kPrefetch) fetch(blockIdx.x);
  for (unsigned ch = blockIdx.x; ch < nchunks; ch += gridDim.x) {
    if (!kPrefetch) fetch(ch);
    if (ch * kChunk16 + kChunk16 <= nq)
      body(std::true_type{}, ch);
    else
      body(std::false_type{}, ch);
  }
}

// split 2: records of one group -> (brick, class) streams
// Measured per kernel (1e9 queries): split 2 is fastest WITHOUT the register prefetch at two CTAs per SM (7.5 vs 8.6 ms);
// split 1 and the unsort levels are fastest with it at one CTA per SM (7.6 vs 8.3 ms, 3.8 vs 4.2 ms).
constexpr bool kPrefetchSplit2 = false;
template <int N, int DEG, int LEVEL>
__global__ void __launch_bounds__(kSplitThreads, (kPrefetchSplit2 || ORD_SPLIT_THREADS > 512) ? 1 : 2) ord_split2_kernel(const __grid_constant__ EvalParams P,
                                                                   const OrdGeom G,
                                                                   const unsigned *__restrict__ start,
                                                                   const unsigned *__restrict__ chunk_start,
                                                                   unsigned *__restrict__ cursor2,
                                                                   const float4 *__restrict__ rec1,
                                                                   float4 *__restrict__ rec2) {
  // LEVEL 2: the parent bins are the groups, the keys (brick, class) inside the group, the cursors the (brick, class)
  //          streams.  LEVEL 1 (three-level geometry): parents = super-groups, keys = groups inside them.
  extern __shared__ __align__(16) unsigned char smem_raw[];
  auto &sm = *reinterpret_cast<SplitSmem<float4, kChunk16> *>(smem_raw);
  split_init(sm);
  __shared__ unsigned s_chunk[kMaxBins + 1];
  constexpr int RPT = kChunk16 / kSplitThreads;
  constexpr unsigned FPG = kGroupBricks * kClasses;                     // final bins per group
  constexpr unsigned FPP = LEVEL == 2 ? FPG : FPG * (unsigned)kMaxBins;  // final bins per parent bin
  const int nparents = LEVEL == 2 ? G.ngroups : G.nsuper;
  const bool table_in_smem = nparents <= kMaxBins;
  if (table_in_smem)
    for (int i = threadIdx.x; i <= nparents; i += kSplitThreads) s_chunk[i] = chunk_start[i];
  __syncthreads();
  const unsigned nchunks = chunk_start[nparents];
  int cursor_parent = 0;  // large tables: the CTA's chunks come in increasing order, the parent only moves forward
  // chunk descriptor: parent bin, first record, count; the records of the next chunk may be fetched early
  auto describe = [&](unsigned ch, int &g, unsigned &base, unsigned &cnt, unsigned &fa, unsigned &fb) {
    unsigned c0;
    if (table_in_smem) {
      int lo = 0, hi = nparents - 1;  // last parent with chunk_start <= ch
      while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (s_chunk[mid] <= ch) lo = mid; else hi = mid - 1;
      }
      g = lo;
      c0 = s_chunk[g];
    } else {
      if (cursor_parent == 0) {
        int lo = 0, hi = nparents - 1;
        while (lo < hi) {
          const int mid = (lo + hi + 1) >> 1;
          if (chunk_start[mid] <= ch) lo = mid; else hi = mid - 1;
        }
        cursor_parent = lo;
      }
      while (cursor_parent + 1 < nparents && chunk_start[cursor_parent + 1] <= ch) ++cursor_parent;
      g = cursor_parent;
      c0 = chunk_start[g];
    }
    fa = min((unsigned)g * FPP, (unsigned)G.nfinal);
    fb = min(fa + FPP, (unsigned)G.nfinal);
    const unsigned gbeg = start[fa], gend = start[fb];
    base = gbeg + (ch - c0) * kChunk16;
    cnt = min((unsigned)kChunk16, gend - base);
  };
  float4 nr[RPT];
  int g = 0, ng = 0;
  unsigned base = 0, cnt = 0, nbase = 0, ncnt = 0, fa = 0, fb = 0, nfa = 0, nfb = 0;
  auto fetch = [&](unsigned ch) {
    ncnt = 0;
    if (ch >= nchunks) return;
    describe(ch, ng, nbase, ncnt, nfa, nfb);
    const float4 *src = rec1 + nbase + threadIdx.x;
    if (ncnt == kChunk16) {
#pragma unroll
      for (int j = 0; j < RPT; ++j) nr[j] = src[j * kSplitThreads];
    } else {
#pragma unroll
      for (int j = 0; j < RPT; ++j)
        if (j * kSplitThreads + threadIdx.x < ncnt) nr[j] = src[j * kSplitThreads];
    }
  };
  if (kPrefetchSplit2) fetch(blockIdx.x);
  for (unsigned ch = blockIdx.x; ch < nchunks; ch += gridDim.x) {
    if (!kPrefetchSplit2) fetch(ch);
    g = ng; base = nbase; cnt = ncnt; fa = nfa; fb = nfb;
    // keys are relative to the parent; the cursor array is indexed by final bin (LEVEL 2) or by group (LEVEL 1)
    const int nbins = LEVEL == 2 ? (int)(fb - fa) : min(kMaxBins, G.ngroups - g * kMaxBins);
    unsigned *cur = LEVEL == 2 ? cursor2 + fa : cursor2 + (size_t)g * kMaxBins;
    auto run = [&](auto full_tag) {
      constexpr bool FULL = decltype(full_tag)::value;
      float4 r[RPT];
      unsigned key[RPT];
#pragma unroll
      for (int j = 0; j < RPT; ++j) {
        const unsigned i = j * kSplitThreads + threadIdx.x;
        key[j] = 0xffffffffu;
        if (FULL || i < cnt) {
          r[j] = nr[j];
          const float x[3] = {r[j].x, r[j].y, r[j].z};
          int brick, cls;
          locate<N, DEG>(P, G, x, brick, cls);
          key[j] = LEVEL == 2 ? (unsigned)((brick - g * kGroupBricks) * kClasses + cls)
                              : (unsigned)(brick / kGroupBricks - g * kMaxBins);
        }
      }
      if (kPrefetchSplit2) fetch(ch + gridDim.x);
      block_split<FULL, float4, kChunk16>(sm, r, key, nbins, cnt, cur, rec2);
    };
    if (cnt == kChunk16)
      run(std::true_type{});
    else
      run(std::false_type{});
  }
}

// unsort: result records by (q >> shift) relative to the chunk's parent bin; q is the first word of the record
template <int LEVEL, bool GRAD>
__global__ void __launch_bounds__(kSplitThreads, kSplitMinCtas) ord_unsort_kernel(const typename Uns<GRAD>::Rec *__restrict__ in,
                                                                   long long nq_, unsigned *__restrict__ cursor,
                                                                   typename Uns<GRAD>::Rec *__restrict__ out) {
  using U = Uns<GRAD>;
  using Rec = typename U::Rec;
  constexpr int S = U::S, SA = U::SA, SB = U::SB;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  auto &sm = *reinterpret_cast<SplitSmem<Rec, S> *>(smem_raw);
  split_init(sm);
  constexpr int RPT = S / kSplitThreads;
  const unsigned nq = (unsigned)nq_;
  const unsigned nchunks = (nq + S - 1) / S;
  const unsigned nwin = (nq + (1u << SB) - 1) >> SB;
  Rec nr[RPT];
  auto fetch = [&](unsigned ch) {
    if (ch >= nchunks) return;
    const Rec *src = in + ch * S + threadIdx.x;
    if (ch * S + S <= nq) {
#pragma unroll
      for (int j = 0; j < RPT; ++j) nr[j] = src[j * kSplitThreads];
    } else {
#pragma unroll
      for (int j = 0; j < RPT; ++j)
        if (ch * S + j * kSplitThreads + threadIdx.x < nq) nr[j] = src[j * kSplitThreads];
    }
  };
  if (kPrefetch) fetch(blockIdx.x);
  for (unsigned ch = blockIdx.x; ch < nchunks; ch += gridDim.x) {
    if (!kPrefetch) fetch(ch);
    const unsigned base = ch * S;
    // level A: bins over the whole array; level B: the chunk lies inside one level-A bin (2^SA records)
    const unsigned bin0 = LEVEL == 0 ? 0u : (base >> SA) << (SA - SB);
    const int nbins = LEVEL == 0 ? (int)((nq + (1u << SA) - 1) >> SA) : (int)min(1u << (SA - SB), nwin - bin0);
    auto run = [&](auto full_tag) {
      constexpr bool FULL = decltype(full_tag)::value;
      const unsigned cnt = FULL ? (unsigned)S : nq - base;
      Rec r[RPT];
      unsigned key[RPT];
#pragma unroll
      for (int j = 0; j < RPT; ++j) {
        const unsigned i = base + j * kSplitThreads + threadIdx.x;
        key[j] = 0xffffffffu;
        if (FULL || i < nq) {
          r[j] = nr[j];
          key[j] = LEVEL == 0 ? (r[j].x >> SA) : ((r[j].x >> SB) - bin0);
        }
      }
      if (kPrefetch) fetch(ch + gridDim.x);
      block_split<FULL, Rec, S>(sm, r, key, nbins, cnt, cursor + bin0, out);
    };
    if (base + S <= nq)
      run(std::true_type{});
    else
      run(std::false_type{});
  }
}

// place: window w holds exactly the results q in [w * 2^SB, ...): permute in shared memory, store coalesced.
// Gradients: NG = N components per query, contiguous in the output (SVector layout).
template <bool GRAD, int NG>
__global__ void __launch_bounds__(1024) ord_place_kernel(const typename Uns<GRAD>::Rec *__restrict__ in, long long nq,
                                                         float *__restrict__ out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float *win = reinterpret_cast<float *>(smem_raw);
  constexpr int W = 1 << Uns<GRAD>::SB;
  const long long base = (long long)blockIdx.x * W;
  const int cnt = (int)min((long long)W, nq - base);
  for (int i = threadIdx.x; i < cnt; i += 1024) {
    const auto r = in[base + i];
    const unsigned slot = r.x & (W - 1);
    if constexpr (GRAD) {
      win[slot * NG + 0] = __uint_as_float(r.y);
      win[slot * NG + 1] = __uint_as_float(r.z);
      if (NG == 3) win[slot * NG + 2] = __uint_as_float(r.w);
    } else {
      win[slot] = __uint_as_float(r.y);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < cnt * NG; i += 1024) out[base * NG + i] = win[i];
}

// ------------------------------------------------------------------------------------------ evaluate
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}

// Nested reduction with gradients from the shared-memory tile: the same operations in the same order as Reduce<...,true,0>
// of eval.cuh (Interpolations.jl:304-316 with one axis' weights swapped for the gradient weights, partial sums shared).
template <int N, int DEG, int D>
struct TileReduce {
  using TC = TileCfg<N, DEG>;
  __device__ __forceinline__ static void run(const AxisTaps<DEG, float> (&ax)[N], const float *base, float &val,
                                             float (&g)[N]) {
    constexpr int L = DEG + 1;
    constexpr int stride = D == 0 ? 1 : (D == 1 ? TC::PX : TC::PX * TC::TY);
    const AxisTaps<DEG, float> &A = ax[D];
    float gv = 0.f;
#pragma unroll
    for (int k = 0; k < L; ++k) {
      float vk;
      float gk[N];
      const float *b = base + k * stride;
      if constexpr (D == N - 1) {
        vk = *b;
      } else {
        TileReduce<N, DEG, D + 1>::run(ax, b, vk, gk);
      }
      const float w = A.wv[k];
      val = k == 0 ? w * vk : w * vk + val;
      const float wgk = A.wg[k];
      gv = k == 0 ? wgk * vk : wgk * vk + gv;
      if constexpr (D < N - 1) {
#pragma unroll
        for (int j = D + 1; j < N; ++j) g[j] = k == 0 ? w * gk[j] : w * gk[j] + g[j];
      }
    }
    g[D] = gv;
  }
};

__device__ __forceinline__ void cp_async16_zfill(void *smem_dst, const void *gsrc, bool valid) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(sz) : "memory");
}

__device__ __forceinline__ void cp_async4_zfill(void *smem_dst, const void *gsrc, bool valid) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  const int sz = valid ? 4 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d), "l"(gsrc), "r"(sz) : "memory");
}

// ---- tensor-map (TMA) tile copy: one instruction moves the brick + halo box, rows outside the array arrive as zeros
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tile_bar_init(unsigned long long *bar) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void tile_bar_expect(unsigned long long *bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tile_bar_wait(unsigned long long *bar, unsigned parity) {
  unsigned done = 0;
  while (!done) {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(done)
                 : "r"(smem_u32(bar)), "r"(parity)
                 : "memory");
  }
}
template <int N>
__device__ __forceinline__ void tile_tma_load(void *dst, const CUtensorMap *map, int x, int y, int z,
                                              unsigned long long *bar) {
  if constexpr (N == 3) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar))
        : "memory");
  } else {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(x), "r"(y), "r"(smem_u32(bar))
        : "memory");
  }
}

template <int N, int DEG>
struct EvalSmem {
  float tile[TileCfg<N, DEG>::FLOATS];
  float4 wstage[kEvalThreads / 32][kClasses * kRBP];  // per-warp staging: kRB rows of every class
  long long axoff[3][TileCfg<N, DEG>::MAXT];  // element offset of tile coordinate l on each axis, -1: outside
  // work-item descriptors, double buffered: warp 0 fetches the next item's while the current one is evaluated
  struct Item {
    unsigned cbeg[kClasses];  // first record of the item's share of each class stream
    unsigned ccnt[kClasses];
    int brick, max_cnt;
  } meta[2];
  int next_blk;
  unsigned long long tile_bar;  // mbarrier of the tensor-map tile copy
};

template <int N, int DEG, bool GRAD>
__global__ void __launch_bounds__(kEvalThreads, 1) ord_eval_kernel(const __grid_constant__ EvalParams P,
                                                                   const OrdGeom G,
                                                                   const unsigned *__restrict__ start,
                                                                   const unsigned *__restrict__ item_start,
                                                                   const float4 *__restrict__ rec,
                                                                   typename Uns<GRAD>::Rec *__restrict__ res,
                                                                   const __grid_constant__ CUtensorMap tmap) {
  using TC = TileCfg<N, DEG>;
  using Bk = Brick<N>;
  constexpr int L = TC::L;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  auto &sm = *reinterpret_cast<EvalSmem<N, DEG> *>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NW = kEvalThreads / 32;
  const float *coefs = reinterpret_cast<const float *>(P.coefs);
  // CTA c takes the work items that begin in its share [R*c/G, R*(c+1)/G) of the R ordered records: balanced by
  // records, whatever the sizes of the items (the items of one brick have equal sizes n_b / P_b)
  auto first_item = [&](unsigned cta) -> unsigned {
    const unsigned nitems = item_start[G.nbricks];
    if (cta >= gridDim.x) return nitems;
    if (cta == 0) return 0u;
    const unsigned long long total = start[G.nfinal];
    const unsigned target = (unsigned)(total * cta / gridDim.x);
    int lo = 0, hi = G.nbricks - 1;  // last brick whose first record is <= target
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (start[(long long)mid * kClasses] <= target) lo = mid; else hi = mid - 1;
    }
    const unsigned s0 = start[(long long)lo * kClasses], nb = start[(long long)(lo + 1) * kClasses] - s0;
    const unsigned parts = item_start[lo + 1] - item_start[lo];
    if (nb == 0 || parts == 0) return item_start[lo + 1];
    // first part whose first record (s0 + nb * p / parts) is >= target
    const unsigned long long num = (unsigned long long)(target - s0) * parts;
    const unsigned p = (unsigned)((num + nb - 1) / nb);
    return item_start[lo] + min(p, parts);
  };
  const unsigned it_lo = first_item(blockIdx.x);
  const unsigned it_hi = first_item(blockIdx.x + 1);
  int cur_brick = -1;
  int walk_brick = -1;  // warp 0: brick of the previous item
  int o[3] = {0, 0, 0};
#ifdef B200ITP_EVAL_TRACE
  const long long t_begin = clock64();
  unsigned long long n_rec = 0, n_rounds = 0, n_tiles = 0;
  long long c_tile = 0, c_setup = 0, c_loop = 0, c_mark = t_begin, c_axoff = 0, c_issue = 0, c_tail = 0;
#endif
  // Descriptor of one work item (warp 0, all lanes): its brick — the last brick with item_start <= item (skips empty
  // bricks): a binary search (13 dependent global loads) for the CTA's first item only, a short forward walk afterwards
  // (the CTA's items are consecutive) — and its share of each of the brick's 32 class streams.
  auto load_meta = [&](unsigned item, int slot) {
    int lo = walk_brick;
    if (lo < 0) {
      lo = 0;
      int hi = G.nbricks - 1;
      while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (item_start[mid] <= item) lo = mid; else hi = mid - 1;
      }
    } else {
      while (lo + 1 < G.nbricks && item_start[lo + 1] <= item) ++lo;
    }
    walk_brick = lo;
    const unsigned is0 = item_start[lo];
    const unsigned part = item - is0, parts = item_start[lo + 1] - is0;
    const unsigned s0 = start[(long long)lo * kClasses + lane];
    const unsigned long long n = start[(long long)lo * kClasses + lane + 1] - s0;
    const unsigned a = (unsigned)(n * part / parts);
    const unsigned b = (unsigned)(n * (part + 1) / parts);
    auto &M = sm.meta[slot];
    M.cbeg[lane] = s0 + a;
    M.ccnt[lane] = b - a;
    const unsigned mx = __reduce_max_sync(0xffffffffu, b - a);
    if (lane == 0) {
      M.brick = lo;
      M.max_cnt = (int)mx;
    }
  };
  if (warp == 0 && it_lo < it_hi) load_meta(it_lo, 0);
  unsigned tile_phase = 0;
  if (G.tma && tid == 0) tile_bar_init(&sm.tile_bar);  // visible to all after the first barrier of the item loop

  for (unsigned item = it_lo; item < it_hi; ++item) {
    const int slot = (int)((item - it_lo) & 1u);
    __syncthreads();  // previous item fully consumed (tile, block counter); this item's descriptor is visible
#ifdef B200ITP_EVAL_TRACE
    { const long long now = clock64(); c_tail += now - c_mark; c_mark = now; }
#endif
    if (tid == 0) sm.next_blk = 0;
    const auto &M = sm.meta[slot];
    const int brick = M.brick;
#ifdef B200ITP_EVAL_TRACE
    { const long long now = clock64(); c_setup += now - c_mark; c_mark = now; }
#endif
    if (brick != cur_brick) {  // stage the brick + halo; tile element l <-> parent index o + l on each axis
      cur_brick = brick;
#ifdef B200ITP_EVAL_TRACE
      ++n_tiles;
#endif
      const int bx = brick % G.nb[0];
      const int by = (brick / G.nb[0]) % G.nb[1];
      const int bz = brick / (G.nb[0] * G.nb[1]);
      o[0] = bx * Bk::BX;
      o[1] = by * Bk::BY;
      o[2] = bz * Bk::BZ;
      constexpr int TDIM[3] = {TC::TX, TC::TY, TC::TZ};
      // one tensor-map copy when no axis of this tile wraps around a periodic boundary (C3: 72 % of the bricks): the
      // box starts tshift columns early so that both staging paths share the tile layout; outside the array -> zeros
      bool by_tma = G.tma != 0;
      int org[3] = {0, 0, 0};
#pragma unroll
      for (int d = 0; d < N; ++d) {
        org[d] = o[d] + G.fmin[d] - (P.periodic[d] ? 1 : P.coef_first[d]);
        if (P.periodic[d] && (org[d] < 0 || org[d] + TDIM[d] > P.n[d])) by_tma = false;
      }
      if (by_tma) {
        if (tid == 0) {
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the previous tile's reads are behind the barrier
          tile_bar_expect(&sm.tile_bar, (unsigned)(TC::FLOATS * sizeof(float)));
          tile_tma_load<N>(sm.tile, &tmap, org[0] - G.tshift, org[1], org[2], &sm.tile_bar);
        }
        tile_bar_wait(&sm.tile_bar, tile_phase);
        tile_phase ^= 1u;
      } else {
      for (int e = tid; e < 3 * TC::MAXT; e += kEvalThreads) {  // per-axis offset tables, once per brick
        const int d = e / TC::MAXT, l = e % TC::MAXT;
        if (d < N && l < TDIM[d]) {
          int idx = o[d] + l + G.fmin[d];  // parent index of the tap
          long long off = -1;
          if (P.periodic[d]) {
            const int n = P.n[d];
            idx = idx - 1;  // modrange1(idx, n) - 1 without a division in the common case
            if (idx < 0) idx += n;
            if (idx >= n) idx -= n;
            if (idx >= n) idx %= n;  // axes shorter than the tile
            off = (long long)idx * P.stride[d];
          } else {
            const int pos = idx - P.coef_first[d];
            const int size_d = P.n[d] + (P.coef_first[d] == 0 ? 2 : 0);
            if (pos >= 0 && pos < size_d) off = (long long)pos * P.stride[d];
          }
          sm.axoff[d][l] = off;
        }
      }
      __syncthreads();
#ifdef B200ITP_EVAL_TRACE
      const long long t_ax = clock64();
      c_axoff += t_ax - c_mark;
#endif
      // asynchronous copies (zero fill outside the array), no register round trip: a row is staged in groups of four
      // shared-memory columns; a group whose four tile columns are consecutive array elements at a 16-byte aligned
      // address is one 16-byte copy (8 of the 10 groups of a 35-wide row), the rest goes element by element (row ends,
      // periodic wrap inside a group, rows of padded arrays that start off alignment)
      constexpr int NG = TC::PX / 4;
      const float *coefs16 = reinterpret_cast<const float *>(reinterpret_cast<uintptr_t>(coefs) & ~(uintptr_t)15);
      const bool base16 = coefs16 == coefs;
      for (int t = tid; t < NG * TC::TY * TC::TZ; t += kEvalThreads) {
        const int gx = t % NG;
        const int ly = (t / NG) % TC::TY;
        const int lz = t / (NG * TC::TY);
        const long long oy = sm.axoff[1][ly], oz = N == 3 ? sm.axoff[2][lz] : 0;
        const bool rowok = oy >= 0 && oz >= 0;
        const int l0 = 4 * gx - G.tshift;  // tile coordinate of the group's first column
        float *dst = &sm.tile[(lz * TC::TY + ly) * TC::PX + 4 * gx];
        bool whole = base16 && l0 >= 0 && l0 + 3 < TC::TX;
        long long src = 0;
        if (whole) {
          const long long ox0 = sm.axoff[0][l0];
          src = ox0 + oy + oz;
          whole = !rowok || (ox0 >= 0 && sm.axoff[0][l0 + 3] == ox0 + 3 && (src & 3) == 0);
        }
        if (whole) {
          cp_async16_zfill(dst, rowok ? coefs + src : coefs16, rowok);
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int l = l0 + e;
            if (l >= 0 && l < TC::TX) {
              const long long ox = sm.axoff[0][l];
              const bool ok = rowok && ox >= 0;
              cp_async4_zfill(dst + e, coefs + (ok ? ox + oy + oz : 0), ok);
            }
          }
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
#ifdef B200ITP_EVAL_TRACE
      c_issue += clock64() - t_ax;
#endif
#if ORD_TILE_PREFETCH
      // Staging a tile is ~1300 scattered 128-byte lines (a 140-byte row every 2 KB): latency-bound, 15 000 cycles from
      // DRAM.  While this tile is consumed, ask the L2 for the lines of the next brick (the CTA's bricks are consecutive;
      // a hint only: an empty or foreign next brick costs nothing but the requests).
      if (brick + 1 < G.nbricks) {
        const int nbk = brick + 1;
        const int no[3] = {(nbk % G.nb[0]) * Bk::BX, ((nbk / G.nb[0]) % G.nb[1]) * Bk::BY, (nbk / (G.nb[0] * G.nb[1])) * Bk::BZ};
        auto aoff = [&](int d, int l) -> long long {
          int idx = no[d] + l + G.fmin[d];
          if (P.periodic[d]) {
            const int n = P.n[d];
            idx = (idx - 1) % n;
            if (idx < 0) idx += n;
            return (long long)idx * P.stride[d];
          }
          const int size_d = P.n[d] + (P.coef_first[d] == 0 ? 2 : 0);
          const int pos = min(max(idx - P.coef_first[d], 0), size_d - 1);
          return (long long)pos * P.stride[d];
        };
        for (int r = tid; r < TC::TY * TC::TZ; r += kEvalThreads) {
          const long long row = aoff(1, r % TC::TY) + (N == 3 ? aoff(2, r / TC::TY) : 0);
#pragma unroll
          for (int l = 0; l < TC::TX + 31; l += 32) {  // one request per 128 bytes of the row (and its end)
            const float *a = coefs + row + aoff(0, min(l, TC::TX - 1));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(a));
          }
        }
      }
#endif
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      }  // !by_tma
    }
    __syncthreads();
    if (warp == 0 && item + 1 < it_hi) load_meta(item + 1, slot ^ 1);  // its latency hides behind the other warps' blocks
    const int rounds = (M.max_cnt + kRB - 1) / kRB;
#ifdef B200ITP_EVAL_TRACE
    { const long long now = clock64(); c_tile += now - c_mark; c_mark = now; }
    if (tid == 0) {
      for (int c = 0; c < kClasses; ++c) n_rec += M.ccnt[c];
      n_rounds += rounds;
    }
#endif

    // Every warp streams its own blocks of kRB rows (row r of class c = record cbeg[c] + r): it stages the 32 x kRB
    // records in its private buffer (lanes cover 4 classes x 8 consecutive records = four 128-byte segments per
    // instruction), evaluates them with lane c on class c, and stores the {q, value} results with the same mapping.
    // Only __syncwarp() is needed: no CTA-wide barrier inside an item.
    float4 *wst = sm.wstage[warp];
    const unsigned mycnt = M.ccnt[lane];
    // blocks are handed out dynamically (one shared counter per item): warps finish an item within one block of
    // each other whatever their pace; the last kTailRows rows go out in blocks of kTailRB rows
    const int rows_all = M.max_cnt;
    const int nbig = rows_all > kTailRows ? (rows_all - kTailRows) / kRB : 0;
    const int nblk = nbig + (rows_all - nbig * kRB + kTailRB - 1) / kTailRB;
    for (;;) {
      int blk = 0;
      if (lane == 0) blk = atomicAdd(&sm.next_blk, 1);
      blk = __shfl_sync(0xffffffffu, blk, 0);
      if (blk >= nblk) break;
      const unsigned r0 = blk < nbig ? (unsigned)blk * kRB : (unsigned)(nbig * kRB + (blk - nbig) * kTailRB);
      const int rows = blk < nbig ? kRB : kTailRB;
#pragma unroll
      for (int j = 0; j < kClasses / 4; ++j) {
        const int c = 4 * j + (lane >> 3), r = lane & 7;
        if (r < rows && r0 + r < M.ccnt[c]) cp_async16(&wst[c * kRBP + r], rec + M.cbeg[c] + r0 + r);
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      __syncwarp();
#pragma unroll 1
      for (int r = 0; r < rows; ++r) {
        if (r0 + r < mycnt) {
          const float4 q4 = wst[lane * kRBP + r];
          const float x[3] = {q4.x, q4.y, q4.z};
          AxisTaps<DEG, float> ax[N];
          int lpos[3] = {0, 0, 0};
#pragma unroll
          for (int d = 0; d < N; ++d) {
            axis_setup<DEG, float, GRAD>(P, d, x[d], ax[d]);
            lpos[d] = min(max(ax[d].first - G.fmin[d], 0), G.cells[d] - 1) - o[d];
          }
          const float *tb = sm.tile + (lpos[2] * TC::TY + lpos[1]) * TC::PX + lpos[0] + G.tshift;
          if constexpr (GRAD) {
            float val = 0.f, g[N];
            TileReduce<N, DEG, 0>::run(ax, tb, val, g);
            wst[lane * kRBP + r] = make_float4(q4.w, g[0], g[1], N == 3 ? g[N - 1] : 0.f);
          } else {
            // value = sum_k0 w0[k0] * ( sum_k1 w1[k1] * ( sum_k2 w2[k2] * c[k2][k1][k0] ) )  (Interpolations.jl:304-316)
            float val = 0.f;
#pragma unroll
            for (int k0 = 0; k0 < L; ++k0) {
              float v0 = 0.f;
#pragma unroll
              for (int k1 = 0; k1 < L; ++k1) {
                float v1;
                if (N == 2) {
                  v1 = tb[k1 * TC::PX + k0];
                } else {
                  v1 = 0.f;
#pragma unroll
                  for (int k2 = 0; k2 < L; ++k2) {
                    const float t2 = ax[N - 1].wv[k2] * tb[(k2 * TC::TY + k1) * TC::PX + k0];
                    v1 = k2 == 0 ? t2 : t2 + v1;
                  }
                }
                const float t1 = ax[1].wv[k1] * v1;
                v0 = k1 == 0 ? t1 : t1 + v0;
              }
              const float t0 = ax[0].wv[k0] * v0;
              val = k0 == 0 ? t0 : t0 + val;
            }
            *reinterpret_cast<float2 *>(&wst[lane * kRBP + r]) = make_float2(q4.w, val);
          }
        }
      }
      __syncwarp();
#pragma unroll
      for (int j = 0; j < kClasses / 4; ++j) {
        const int c = 4 * j + (lane >> 3), r = lane & 7;
        if (r < rows && r0 + r < M.ccnt[c]) {
          if constexpr (GRAD) {
            const float4 v = wst[c * kRBP + r];
            res[M.cbeg[c] + r0 + r] = make_uint4(__float_as_uint(v.x), __float_as_uint(v.y), __float_as_uint(v.z),
                                                  __float_as_uint(v.w));
          } else {
            const float2 v = *reinterpret_cast<const float2 *>(&wst[c * kRBP + r]);
            res[M.cbeg[c] + r0 + r] = make_uint2(__float_as_uint(v.x), __float_as_uint(v.y));
          }
        }
      }
      __syncwarp();
    }
#ifdef B200ITP_EVAL_TRACE
    { const long long now = clock64(); c_loop += now - c_mark; c_mark = now; }
#endif
  }
#ifdef B200ITP_EVAL_TRACE
  if (tid == 0) {
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    printf("evaltrace cta %d sm %u items %u records %llu rounds %llu tiles %llu cycles %lld setup %lld tile %lld loop(warp0) %lld axoff %lld issue %lld tail %lld\n",
           blockIdx.x, smid, it_hi - it_lo, n_rec, n_rounds, n_tiles, clock64() - t_begin, c_setup, c_tile, c_loop, c_axoff, c_issue, c_tail);
  }
#endif
}

// ------------------------------------------------------------------------------------------ host
// cuTensorMapEncodeTiled through the runtime's driver entry point (the library does not link libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn tensor_map_encoder() {
  static EncodeTiledFn fn = [] {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      p = nullptr;
    (void)cudaGetLastError();
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}
static std::atomic<uint64_t> g_tma_calls{0};  // ordered calls whose tiles went through the tensor map
static int g_ord_tma = 1;  // b200itp_eval_sorted_set_tma(0): stage every tile with cp.async (tests compare the two)

// Tensor map of the coefficient array for boxes of one brick + halo (TileCfg); false: the array does not qualify
// (rows not 16-byte multiples, unaligned base, encoder unavailable) and every tile is staged with cp.async.
template <int N, int DEG>
static bool make_tile_map(const EvalParams &P, const b200itp_desc *D, CUtensorMap *map) {
  using TC = TileCfg<N, DEG>;
  EncodeTiledFn enc = tensor_map_encoder();
  if (!enc || !g_ord_tma) return false;
  if (reinterpret_cast<uintptr_t>(P.coefs) % 16 != 0) return false;
  cuuint64_t dims[3] = {1, 1, 1}, strides[2] = {0, 0};
  cuuint32_t box[3] = {(cuuint32_t)TC::PX, (cuuint32_t)TC::TY, (cuuint32_t)TC::TZ}, est[3] = {1, 1, 1};
  for (int d = 0; d < N; ++d) {
    dims[d] = (cuuint64_t)D->size[d];
    if (box[d] > 256) return false;
    if (d > 0) {
      if (P.stride[d] % 4 != 0) return false;  // byte strides must be multiples of 16
      strides[d - 1] = (cuuint64_t)P.stride[d] * sizeof(float);
    }
  }
  if (P.stride[0] != 1) return false;
  const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)N, const_cast<void *>(P.coefs), dims, strides, box,
                         est, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

struct OrdLayout {
  size_t rec1, rec2, hist, start, sums, cursor0, chunk_mid, cursor1, chunk_start, item_start, cursorA, cursorB, total;
};

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static OrdLayout ord_layout(long long nq, long long nfinal, long long nbricks, long long ngroups) {
  OrdLayout Lo;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    const size_t at = off;
    off = align_up(off + bytes, 256);
    return at;
  };
  Lo.rec1 = take((size_t)nq * 16);
  Lo.rec2 = take((size_t)nq * 16);
  Lo.hist = take((size_t)(nfinal + 1) * 4);
  Lo.start = take((size_t)(nfinal + 1) * 4);
  Lo.sums = take(65536 * 4);
  Lo.cursor0 = take((size_t)(kMaxBins + 1) * 4);
  Lo.chunk_mid = take((size_t)(kMaxBins + 2) * 4);
  Lo.cursor1 = take((size_t)(ngroups + 1) * 4);
  Lo.chunk_start = take((size_t)(ngroups + 1) * 4);
  Lo.item_start = take((size_t)(nbricks + 1) * 4);
  Lo.cursorA = take((size_t)(((nq + (1LL << Uns<true>::SA) - 1) >> Uns<true>::SA) + 1) * 4);  // gradient mode: the
  Lo.cursorB = take((size_t)(((nq + (1LL << Uns<true>::SB) - 1) >> Uns<true>::SB) + 1) * 4);  // smaller shifts

  Lo.total = off;
  return Lo;
}

template <int N, int DEG>
static bool ord_geometry(const EvalParams &P, const b200itp_desc *D, OrdGeom &G) {
  using Bk = Brick<N>;
  const int bdim[3] = {Bk::BX, Bk::BY, Bk::BZ};
  long long bricks = 1;
  for (int d = 0; d < 3; ++d) {
    G.nb[d] = 1;
    G.cells[d] = 1;
    G.lbf[d] = 0.f;
    G.ubf[d] = 0.f;
    G.fmin[d] = 0;
  }
  for (int d = 0; d < N; ++d) {
    // smallest Float32 >= lb and largest Float32 <= ub: the Float64 comparisons of the direct kernel, in Float32
    float lo = (float)P.lb[d], hi = (float)P.ub[d];
    if ((double)lo < P.lb[d]) lo = nextafterf(lo, INFINITY);
    if ((double)hi > P.ub[d]) hi = nextafterf(hi, -INFINITY);
    G.lbf[d] = lo;
    G.ubf[d] = hi;
  }
  for (int d = 0; d < N; ++d) {
    // first-tap indices: periodic cubic on OnCell grids -1..n-1 (floor(x) - 1 with x in [0.5, n + 0.5]), other periodic
    // axes 0..n-1, otherwise 0..size-1-DEG
    G.fmin[d] = (P.periodic[d] && DEG == B200ITP_CUBIC && P.inner_lb[d] < 1.0) ? -1 : 0;  // OnGrid: x >= 1, first >= 0
    G.cells[d] = P.periodic[d] ? P.n[d] - G.fmin[d] : (int)D->size[d] - DEG;
    if (G.cells[d] < 1) return false;
    G.nb[d] = (G.cells[d] + bdim[d] - 1) / bdim[d];
    bricks *= G.nb[d];
  }
  // two split levels reach kMaxBins groups of kGroupBricks bricks (8192 bricks: 512^3 cells); a third level in front
  // (super-groups of kMaxBins groups) reaches kMaxBins^2 groups; the stream table (4 bytes per (brick, class)) is the
  // practical limit
  if (bricks > (long long)kGroupBricks * kMaxBins * kMaxBins || bricks * kClasses > (1LL << 28)) return false;
  G.nbricks = (int)bricks;
  G.ngroups = (int)((bricks + kGroupBricks - 1) / kGroupBricks);
  G.nsuper = G.ngroups <= kMaxBins ? 1 : (G.ngroups + kMaxBins - 1) / kMaxBins;
  G.nfinal = (int)bricks * kClasses;
  G.wrapped = (P.has_scale || P.etp_kind == B200ITP_ETP_FLAGS) ? 1 : 0;
  // array column (0-based) of tile coordinate 0 on axis 1, modulo 4 (brick origins are multiples of 4)
  const int c0 = G.fmin[0] - (P.periodic[0] ? 1 : P.coef_first[0]);
  G.tshift = ((c0 % 4) + 4) % 4;
  G.tma = 0;
  return true;
}

template <int N, int DEG, bool GRAD>
static int run_ordered(int dev, EvalParams &P, const b200itp_desc *D, void *out, void *scratch, size_t scratch_bytes,
                       cudaStream_t st, bool *done) {
  using U = Uns<GRAD>;
  using Rec = typename U::Rec;
  *done = false;
  OrdGeom G;
  if (!ord_geometry<N, DEG>(P, D, G)) return 0;
  // level A has at most kMaxBins bins of 2^SA results (2^31 for both record types); larger calls: the direct kernel
  if (P.nq > std::min<long long>(0xfffffff0LL, ((long long)kMaxBins << U::SA) - 16)) return 0;
  const OrdLayout Lo = ord_layout(P.nq, G.nfinal, G.nbricks, G.ngroups);
  if (!scratch || scratch_bytes < Lo.total || ((uintptr_t)scratch % 256) != 0) return 0;
  unsigned char *sb = reinterpret_cast<unsigned char *>(scratch);
  float4 *rec1 = reinterpret_cast<float4 *>(sb + Lo.rec1);
  float4 *rec2 = reinterpret_cast<float4 *>(sb + Lo.rec2);
  unsigned *hist = reinterpret_cast<unsigned *>(sb + Lo.hist);
  unsigned *start = reinterpret_cast<unsigned *>(sb + Lo.start);
  unsigned *sums = reinterpret_cast<unsigned *>(sb + Lo.sums);
  OrdPlan pl;
  pl.cursor0 = reinterpret_cast<unsigned *>(sb + Lo.cursor0);
  pl.chunk_mid = reinterpret_cast<unsigned *>(sb + Lo.chunk_mid);
  pl.cursor1 = reinterpret_cast<unsigned *>(sb + Lo.cursor1);
  pl.chunk_start = reinterpret_cast<unsigned *>(sb + Lo.chunk_start);
  pl.item_start = reinterpret_cast<unsigned *>(sb + Lo.item_start);
  pl.cursorA = reinterpret_cast<unsigned *>(sb + Lo.cursorA);
  pl.cursorB = reinterpret_cast<unsigned *>(sb + Lo.cursorB);
  // result records: rec1 is dead after split 2, rec2 after the evaluation.  Values (8 bytes): res and resA share rec1,
  // resB is rec2.  Gradients (16 bytes): res = rec1, resA = rec2, resB = rec1 again.
  Rec *res = reinterpret_cast<Rec *>(rec1);
  Rec *resA = GRAD ? reinterpret_cast<Rec *>(rec2) : res + P.nq;
  Rec *resB = GRAD ? reinterpret_cast<Rec *>(rec1) : reinterpret_cast<Rec *>(rec2);

  const int sms = sm_count(dev);
  const long long per = (long long)kScanBlock * kScanPer;
  const long long nsb = (G.nfinal + per - 1) / per;
  constexpr size_t smem16 = sizeof(SplitSmem<float4, kChunk16>), smemU = sizeof(SplitSmem<Rec, U::S>);
  constexpr size_t smem_eval = sizeof(EvalSmem<N, DEG>);
  constexpr int smem_place = (4 << U::SB) * (GRAD ? N : 1);
  static_assert(smem_eval <= 227 * 1024, "evaluation tile does not fit shared memory");
  static_assert(smem_place <= 227 * 1024, "placement window does not fit shared memory");
  // once per device and instantiation (the attribute belongs to the function in that device's context)
  static std::atomic<unsigned long long> attr_done{0};
  if (dev >= 64 || !(attr_done.load(std::memory_order_acquire) >> dev & 1ULL)) {
    B200_CUDA_OK(cudaFuncSetAttribute(ord_split1_kernel<N, DEG, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem16));
    B200_CUDA_OK(cudaFuncSetAttribute(ord_split1_kernel<N, DEG, true, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem16));
    B200_CUDA_OK(cudaFuncSetAttribute(ord_split1_kernel<N, DEG, false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem16));
    B200_CUDA_OK(cudaFuncSetAttribute(ord_split1_kernel<N, DEG, true, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem16));
    B200_CUDA_OK(cudaFuncSetAttribute(ord_split2_kernel<N, DEG, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem16));
    B200_CUDA_OK(cudaFuncSetAttribute(ord_split2_kernel<N, DEG, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem16));
    B200_CUDA_OK(cudaFuncSetAttribute(ord_unsort_kernel<0, GRAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemU));
    B200_CUDA_OK(cudaFuncSetAttribute(ord_unsort_kernel<1, GRAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemU));
    B200_CUDA_OK(cudaFuncSetAttribute(ord_eval_kernel<N, DEG, GRAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_eval));
    B200_CUDA_OK(cudaFuncSetAttribute(ord_place_kernel<GRAD, GRAD ? N : 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_place));
    if (dev < 64) attr_done.fetch_or(1ULL << dev, std::memory_order_release);
  }

  B200_CUDA_OK(cudaMemsetAsync(hist, 0, (size_t)(G.nfinal + 1) * 4, st));
  const int cblocks = (int)std::min<long long>((P.nq + 255) / 256, (long long)sms * 16);
  const int sblocks = sms * 2;
  auto scan_fine = [&]() {
    scan_partial_kernel<<<(unsigned)nsb, kScanBlock, 0, st>>>(hist, G.nfinal, sums);
    scan_sums_kernel<<<1, 1024, 0, st>>>(sums, (int)nsb);
    scan_final_kernel<<<(unsigned)nsb, kScanBlock, 0, st>>>(hist, G.nfinal, sums, start);
  };
  if (G.nsuper == 1) {
    // two-level geometry: group histogram (HBM rate) -> group cursors -> first split, which also accumulates the
    // (brick, class) histogram -> scan, work items -> second split
    unsigned *ghist = pl.chunk_mid;  // unused by this geometry otherwise (kMaxBins + 2 entries)
    B200_CUDA_OK(cudaMemsetAsync(ghist, 0, (size_t)(kMaxBins + 2) * 4, st));
    {
      StageScope sc("ord_count", st);
      const int gblocks = (int)std::min<long long>((P.nq + ORD_CG_THREADS - 1) / ORD_CG_THREADS, (long long)sms * ORD_CG_MULT);
      if (G.wrapped) ord_count_groups_kernel<N, DEG, true><<<gblocks, ORD_CG_THREADS, 0, st>>>(P, G, ghist, hist);
      else ord_count_groups_kernel<N, DEG, false><<<gblocks, ORD_CG_THREADS, 0, st>>>(P, G, ghist, hist);
    }
    {
      StageScope sc("ord_scan_plan", st);
      ord_plan_groups_kernel<<<1, 1024, 0, st>>>(ghist, G, P.nq, U::SA, U::SB, pl);
    }
    {
      StageScope sc("ord_split1", st);
      if (G.wrapped) ord_split1_kernel<N, DEG, false, true, true><<<sblocks, kSplitThreads, smem16, st>>>(P, G, pl.cursor1, rec1, hist);
      else ord_split1_kernel<N, DEG, false, false, true><<<sblocks, kSplitThreads, smem16, st>>>(P, G, pl.cursor1, rec1, hist);
    }
    {
      StageScope sc("ord_scan_plan", st);
      scan_fine();
      ord_plan_kernel<<<1, 1024, 0, st>>>(start, G, P.nq, sms, U::SA, U::SB, pl, true);
    }
    {
      StageScope sc("ord_split2", st);
      ord_split2_kernel<N, DEG, 2><<<sblocks, kSplitThreads, smem16, st>>>(P, G, start, pl.chunk_start, hist, rec1, rec2);
    }
    count_launch();
  } else {  // arrays beyond 8192 bricks: super-groups -> groups -> (brick, class) streams; the records end in rec2 again
    {
      StageScope sc("ord_count", st);
      if (G.wrapped) ord_count_kernel<N, DEG, true><<<cblocks, 256, 0, st>>>(P, G, hist);
      else ord_count_kernel<N, DEG, false><<<cblocks, 256, 0, st>>>(P, G, hist);
    }
    {
      StageScope sc("ord_scan_plan", st);
      scan_fine();
      ord_plan_kernel<<<1, 1024, 0, st>>>(start, G, P.nq, sms, U::SA, U::SB, pl, false);
    }
    {
      StageScope sc("ord_split0", st);
      if (G.wrapped) ord_split1_kernel<N, DEG, true, true, false><<<sblocks, kSplitThreads, smem16, st>>>(P, G, pl.cursor0, rec2, nullptr);
      else ord_split1_kernel<N, DEG, true, false, false><<<sblocks, kSplitThreads, smem16, st>>>(P, G, pl.cursor0, rec2, nullptr);
    }
    {
      StageScope sc("ord_split1", st);
      ord_split2_kernel<N, DEG, 1><<<sblocks, kSplitThreads, smem16, st>>>(P, G, start, pl.chunk_mid, pl.cursor1, rec2, rec1);
    }
    {
      StageScope sc("ord_split2", st);
      ord_split2_kernel<N, DEG, 2><<<sblocks, kSplitThreads, smem16, st>>>(P, G, start, pl.chunk_start, hist, rec1, rec2);
    }
    count_launch();
  }
  // the count / scan / split kernels above never touch the coefficient array: a host that replicates or prefilters it
  // on another stream (b200itp_eval_set_coefs_event) overlaps that with them; the gather kernel waits here
  if (cudaEvent_t ev = take_coefs_event()) B200_CUDA_OK(cudaStreamWaitEvent(st, ev, 0));
  {
    StageScope sc(GRAD ? "ord_eval_gradient" : "ord_eval", st);
    CUtensorMap tmap;
    memset(&tmap, 0, sizeof(tmap));
    G.tma = make_tile_map<N, DEG>(P, D, &tmap) ? 1 : 0;
    if (G.tma) g_tma_calls.fetch_add(1);
    ord_eval_kernel<N, DEG, GRAD><<<sms, kEvalThreads, smem_eval, st>>>(P, G, start, pl.item_start, rec2, res, tmap);
  }
  const Rec *last = res;
  if (P.nq > (1LL << U::SA)) {
    StageScope sc("ord_unsortA", st);
    ord_unsort_kernel<0, GRAD><<<sblocks, kSplitThreads, smemU, st>>>(res, P.nq, pl.cursorA, resA);
    count_launch();
    last = resA;
  } else if (GRAD) {
    resB = reinterpret_cast<Rec *>(rec2);  // level A skipped: res (rec1) is the input of level B
  }
  {
    StageScope sc("ord_unsortB", st);
    ord_unsort_kernel<1, GRAD><<<sblocks, kSplitThreads, smemU, st>>>(last, P.nq, pl.cursorB, resB);
  }
  const long long nwin = (P.nq + (1LL << U::SB) - 1) >> U::SB;
  {
    StageScope sc("ord_place", st);
    ord_place_kernel<GRAD, GRAD ? N : 1><<<(unsigned)nwin, 1024, smem_place, st>>>(resB, P.nq, reinterpret_cast<float *>(out));
  }
  count_launch(10);
  B200_CUDA_OK(cudaGetLastError());
  *done = true;
  return 0;
}

// Query count from which the ordered pipeline is used (tests lower it to exercise the pipeline on
// small inputs; results do not depend on it).
static int64_t g_sorted_min_queries = 1 << 20;
static long long g_sorted_min_coef_bytes = 128LL << 20;
static std::atomic<uint64_t> g_ordered_calls{0};  // calls served by the ordered pipeline (not the direct kernel)

}  // namespace b200itp

using namespace b200itp;

extern "C" int64_t b200itp_eval_sorted_min_queries(void) { return g_sorted_min_queries; }
extern "C" uint64_t b200itp_eval_sorted_tma_calls(void) { return g_tma_calls.load(); }
extern "C" int b200itp_eval_sorted_set_tma(int on) {
  const int prev = g_ord_tma;
  g_ord_tma = on ? 1 : 0;
  return prev;
}
extern "C" uint64_t b200itp_eval_sorted_ordered_calls(void) { return g_ordered_calls.load(); }
extern "C" void b200itp_eval_sorted_set_min_queries(int64_t n) {
  g_sorted_min_queries = n < 1 ? 1 : n;
  g_sorted_min_coef_bytes = n <= 1 ? 0 : (128LL << 20);  // n == 1 (tests): every supported call takes the ordered path
}

// Descriptor-only test of the ordered pipeline's scope (the complete test, with the derived promotion chain, is in
// eval_sorted_common; a candidate that fails it is served by the direct kernel and its scratch stays unused).
static bool ordered_candidate(const b200itp_desc *d, int64_t nq) {
  if (d->ndims != 2 && d->ndims != 3) return false;
  if (d->coef_dtype != B200ITP_F32 || d->coord_dtype != B200ITP_F32) return false;
  // wrappers: only chains that stay Float32 (no Float64 bound or range), no Line extrapolation, no fill value
  if (d->etp_kind == B200ITP_ETP_FILL) return false;
  for (int i = 0; i < d->ndims; ++i) {
    if (d->has_scale && (d->range_tag[i] == B200ITP_TAG_F64 || d->outer_tag[i] == B200ITP_TAG_F64 || d->inner_tag[i] == B200ITP_TAG_F64))
      return false;
    if (d->etp_kind == B200ITP_ETP_FLAGS) {
      if (d->etp_lo[i] == B200ITP_BC_LINE || d->etp_hi[i] == B200ITP_BC_LINE) return false;
      if ((d->has_scale ? d->outer_tag[i] : d->inner_tag[i]) == B200ITP_TAG_F64) return false;
    }
  }
  const int deg = d->degree[0];
  if (deg != B200ITP_CUBIC && deg != B200ITP_QUADRATIC) return false;
  long long coef_bytes = 4;
  for (int i = 0; i < d->ndims; ++i) {
    if (d->degree[i] != deg || d->parent_first[i] > 1 || d->bc[i] >= 6) return false;  // mixed, interpolate!, InPlace
    coef_bytes *= d->size[i];
  }
  if (nq < g_sorted_min_queries) return false;
  return d->ndims == 3 || coef_bytes >= g_sorted_min_coef_bytes;
}

extern "C" size_t b200itp_eval_sorted_scratch_bytes(const b200itp_desc *d, int64_t nq) {
  if (!d || nq <= 0) return 0;
  if (!ordered_candidate(d, nq)) return 0;  // the call is served by the direct kernel: no scratch needed
  // upper bound over the brick geometries of this descriptor (cells <= size + 1 per axis)
  long long nbricks = 1;
  const int bdim3[3] = {Brick<3>::BX, Brick<3>::BY, Brick<3>::BZ}, bdim2[3] = {Brick<2>::BX, Brick<2>::BY, 1};
  for (int i = 0; i < d->ndims && i < 3; ++i) {
    const long long cells = (d->size[i] > 0 ? d->size[i] : 1) + 1;
    const int b = d->ndims == 3 ? bdim3[i] : bdim2[i];
    nbricks *= (cells + b - 1) / b;
  }
  nbricks = std::max<long long>(nbricks, (long long)kGroupBricks * kMaxBins);
  if (nbricks * kClasses > (1LL << 28)) nbricks = (1LL << 28) / kClasses;  // beyond: the call falls back anyway
  const long long ngroups = (nbricks + kGroupBricks - 1) / kGroupBricks;
  return ord_layout(nq, nbricks * kClasses, nbricks, ngroups).total + 256;
}

static int eval_sorted_common(int dev, const b200itp_desc *d, const void *coefs, const void *const *coords, int64_t nq,
                              void *out_value, void *out_grad, int64_t *first_oob, void *scratch, size_t scratch_bytes,
                              void *stream) {
  void *out = out_value ? out_value : out_grad;
  const bool grad = out_value == nullptr;
  if (!d || !coefs || !coords || !out) B200_FAIL(B200ITP_E_BADARG, "eval_sorted: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  int deg = 0;
  bool plain = false;
  EvalParams P;
  DeviceGuard g(dev);
  if (!g.ok) B200_FAIL(-(int)cudaErrorInvalidDevice, "eval_sorted: cannot select device %d", dev);
  void *flag = nullptr;
  int rc = scratch_get(dev, 2, st, 256, &flag);
  if (rc) return rc;
  rc = fill_eval_params_public(d, coefs, coords, nq, out, (unsigned long long *)flag, P, &deg, &plain);
  if (rc) return rc;
  bool done = false;
  // The ordered pipeline pays off when the coefficient array does not fit the L2 (measured: 4096^2 Float32, 64 MiB:
  // direct 36 vs ordered 20 Gevals/s; 8192^2, 256 MiB: 20 vs 21; 512^3, 512 MiB: 2.1 vs 23).
  long long coef_bytes = 4;
  for (int i = 0; i < d->ndims && i < B200ITP_MAX_DIMS; ++i) coef_bytes *= d->size[i];
  // (3-D: the direct tricubic kernel stays at 2-4 Gevals/s even from the L2 — 128^3: 4.0 vs 16.3, 256^3: 3.1 vs 20.5 — so
  // every 3-D array takes the ordered path; the size condition applies to 2-D arrays only)
  // (gradients of wrapped objects — rescale_gradient, the in-bounds / Flat rules of extrapolation.jl:71-82 — stay with the
  //  direct kernel; values of all-Float32 scaled / extrapolated objects take the ordered pipeline)
  if (grad && (d->has_scale || d->etp_kind != B200ITP_ETP_NONE)) plain = false;
  if (plain && nq >= g_sorted_min_queries && (d->ndims == 3 || coef_bytes >= g_sorted_min_coef_bytes) &&
      (deg == B200ITP_CUBIC || deg == B200ITP_QUADRATIC) && (d->ndims == 2 || d->ndims == 3)) {
    B200_CUDA_OK(cudaMemsetAsync(flag, 0xFF, sizeof(unsigned long long), st));
    // the scratch pointer is aligned up to 256 bytes inside the caller's buffer
    unsigned char *sp = reinterpret_cast<unsigned char *>(scratch);
    size_t sbytes = scratch_bytes;
    if (sp) {
      const size_t mis = (256 - ((uintptr_t)sp % 256)) % 256;
      if (mis <= sbytes) {
        sp += mis;
        sbytes -= mis;
      }
    }
    const int key = d->ndims * 100 + deg * 10 + (grad ? 1 : 0);
    switch (key) {
      case 330: rc = run_ordered<3, 3, false>(dev, P, d, out, sp, sbytes, st, &done); break;
      case 320: rc = run_ordered<3, 2, false>(dev, P, d, out, sp, sbytes, st, &done); break;
      case 230: rc = run_ordered<2, 3, false>(dev, P, d, out, sp, sbytes, st, &done); break;
      case 220: rc = run_ordered<2, 2, false>(dev, P, d, out, sp, sbytes, st, &done); break;
      case 331: rc = run_ordered<3, 3, true>(dev, P, d, out, sp, sbytes, st, &done); break;
      case 321: rc = run_ordered<3, 2, true>(dev, P, d, out, sp, sbytes, st, &done); break;
      case 231: rc = run_ordered<2, 3, true>(dev, P, d, out, sp, sbytes, st, &done); break;
      case 221: rc = run_ordered<2, 2, true>(dev, P, d, out, sp, sbytes, st, &done); break;
      default: break;
    }
    if (rc) return rc;
    if (done) {
      g_ordered_calls.fetch_add(1, std::memory_order_relaxed);
      if (first_oob) {
        unsigned long long h = ~0ULL;
        B200_CUDA_OK(cudaMemcpyAsync(&h, flag, sizeof h, cudaMemcpyDeviceToHost, st));
        B200_CUDA_OK(cudaStreamSynchronize(st));
        *first_oob = h == ~0ULL ? -1 : (int64_t)h;
      }
      return 0;
    }
  }
  // outside the ordered path's scope: same results from the direct kernel
  return b200itp_eval(dev, d, coefs, coords, nq, out_value, out_grad, first_oob, nullptr, nullptr, stream);
}

extern "C" int b200itp_eval_sorted(int dev, const b200itp_desc *d, const void *coefs, const void *const *coords,
                                   int64_t nq, void *out_value, int64_t *first_oob, void *scratch,
                                   size_t scratch_bytes, void *stream) {
  if (!out_value) B200_FAIL(B200ITP_E_BADARG, "eval_sorted: null output pointer");
  return eval_sorted_common(dev, d, coefs, coords, nq, out_value, nullptr, first_oob, scratch, scratch_bytes, stream);
}

extern "C" int b200itp_gradient_sorted(int dev, const b200itp_desc *d, const void *coefs, const void *const *coords,
                                       int64_t nq, void *out_grad, int64_t *first_oob, void *scratch,
                                       size_t scratch_bytes, void *stream) {
  if (!out_grad) B200_FAIL(B200ITP_E_BADARG, "gradient_sorted: null output pointer");
  return eval_sorted_common(dev, d, coefs, coords, nq, nullptr, out_grad, first_oob, scratch, scratch_bytes, stream);
}

// ------------------------------------------------------------------------------------------ grid evaluation
// itp(xvec, yvec, ...) (indexing.jl:16-23, scaling.jl:98-100, extrapolation.jl:59-70) is the pointwise evaluation over
// Iterators.product(x...), first axis fastest.  The coordinate tuples are expanded on the device into the caller's
// scratch (one streaming write of N words per output, ~1 % of the evaluation time) and take the usual paths: the ordered
// pipeline for bare Float32 2-D / 3-D interpolants (a coherent query set is just a very friendly distribution for it),
// the direct kernel otherwise.
namespace b200itp {
template <typename T>
__global__ void __launch_bounds__(256) grid_expand_kernel(const T *a0, const T *a1, const T *a2, const T *a3, long long l0,
                                                          long long l1, long long l2, int ndims, long long nq, T *o0, T *o1,
                                                          T *o2, T *o3) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < nq; q += stride) {
    long long r = q;
    const long long i0 = r % l0;
    r /= l0;
    o0[q] = a0[i0];
    if (ndims > 1) {
      const long long i1 = r % l1;
      r /= l1;
      o1[q] = a1[i1];
    }
    if (ndims > 2) {
      const long long i2 = r % l2;
      r /= l2;
      o2[q] = a2[i2];
    }
    if (ndims > 3) o3[q] = a3[r];
  }
}
}  // namespace b200itp

namespace b200itp {
bool grid_separable_scope(const b200itp_desc *d);
size_t grid_separable_scratch_bytes(const b200itp_desc *d, const int64_t *axis_len);
int grid_separable(int dev, const b200itp_desc *D, const void *coefs, const void *const *axis_coords,
                   const int64_t *axis_len, void *out_value, int64_t *first_oob, unsigned char *sp, size_t sbytes,
                   cudaStream_t st, bool *done);
std::atomic<int> g_grid_separable{1};
std::atomic<uint64_t> g_grid_separable_calls{0};
}  // namespace b200itp

// verification switch: 0 forces the tuple-expansion path (every element through the pointwise kernels); returns the
// previous setting
extern "C" int b200itp_eval_grid_set_separable(int enable) { return b200itp::g_grid_separable.exchange(enable ? 1 : 0); }
extern "C" uint64_t b200itp_eval_grid_separable_calls(void) { return b200itp::g_grid_separable_calls.load(); }

static size_t grid_coord_bytes(const b200itp_desc *d, int64_t nq) {
  const size_t es = d->coord_dtype == B200ITP_F64 ? 8 : 4;
  return ((size_t)nq * es + 255) / 256 * 256;
}

extern "C" size_t b200itp_eval_grid_scratch_bytes(const b200itp_desc *d, const int64_t *axis_len) {
  if (!d || !axis_len || d->ndims < 1 || d->ndims > B200ITP_MAX_DIMS) return 0;
  int64_t nq = 1;
  for (int i = 0; i < d->ndims; ++i) nq *= axis_len[i] > 0 ? axis_len[i] : 0;
  if (nq <= 0) return 0;
  // the separable path needs its axis tables and intermediates only; the expansion path the expanded tuples (+ the
  // ordered pipeline's scratch)
  if (b200itp::g_grid_separable.load() && grid_separable_scope(d)) return grid_separable_scratch_bytes(d, axis_len) + 512;
  return grid_coord_bytes(d, nq) * d->ndims + b200itp_eval_sorted_scratch_bytes(d, nq) + 512;
}

extern "C" int b200itp_eval_grid(int dev, const b200itp_desc *d, const void *coefs, const void *const *axis_coords,
                                 const int64_t *axis_len, void *out_value, int64_t *first_oob, void *scratch,
                                 size_t scratch_bytes, void *stream) {
  if (!d || !coefs || !axis_coords || !axis_len || !out_value) B200_FAIL(B200ITP_E_BADARG, "eval_grid: null pointer");
  if (d->ndims < 1 || d->ndims > B200ITP_MAX_DIMS) B200_FAIL(B200ITP_E_BADARG, "eval_grid: ndims must be 1..4");
  if (d->coord_dtype != B200ITP_F32 && d->coord_dtype != B200ITP_F64) B200_FAIL(B200ITP_E_DTYPE, "eval_grid: bad coordinate dtype");
  int64_t nq = 1;
  for (int i = 0; i < d->ndims; ++i) {
    if (axis_len[i] < 0) B200_FAIL(B200ITP_E_BADARG, "eval_grid: negative axis length");
    if (axis_len[i] > 0 && !axis_coords[i]) B200_FAIL(B200ITP_E_BADARG, "eval_grid: null coordinate vector %d", i);
    nq *= axis_len[i];
  }
  if (first_oob) *first_oob = -1;
  if (nq == 0) return 0;
  const size_t cb = grid_coord_bytes(d, nq);
  unsigned char *sp = reinterpret_cast<unsigned char *>(scratch);
  size_t sbytes = scratch_bytes;
  if (sp) {
    const size_t mis = (256 - ((uintptr_t)sp % 256)) % 256;
    if (mis > sbytes) B200_FAIL(B200ITP_E_SIZE, "eval_grid: scratch too small");
    sp += mis;
    sbytes -= mis;
  }
  DeviceGuard g(dev);
  if (!g.ok) B200_FAIL(-(int)cudaErrorInvalidDevice, "eval_grid: cannot select device %d", dev);
  cudaStream_t st = (cudaStream_t)stream;
  if (cudaEvent_t ev = take_coefs_event()) B200_CUDA_OK(cudaStreamWaitEvent(st, ev, 0));  // coefficients produced elsewhere
  if (sp && g_grid_separable.load()) {  // N one-dimensional resampling passes (eval_grid.cu), same bits as pointwise
    bool done = false;
    const int rc = grid_separable(dev, d, coefs, axis_coords, axis_len, out_value, first_oob, sp, sbytes, st, &done);
    if (rc) return rc;
    if (done) {
      g_grid_separable_calls.fetch_add(1, std::memory_order_relaxed);
      return 0;
    }
  }
  if (!sp || sbytes < cb * d->ndims) B200_FAIL(B200ITP_E_SIZE, "eval_grid: scratch too small (b200itp_eval_grid_scratch_bytes)");
  const void *cptr[B200ITP_MAX_DIMS] = {nullptr, nullptr, nullptr, nullptr};
  void *optr[B200ITP_MAX_DIMS] = {nullptr, nullptr, nullptr, nullptr};
  for (int i = 0; i < d->ndims; ++i) {
    optr[i] = sp + cb * i;
    cptr[i] = optr[i];
  }
  const int blocks = (int)std::min<long long>((nq + 255) / 256, (long long)sm_count(dev) * 16);
  {
    StageScope sc("grid_expand", st);
    const long long l0 = axis_len[0], l1 = d->ndims > 1 ? axis_len[1] : 1, l2 = d->ndims > 2 ? axis_len[2] : 1;
    if (d->coord_dtype == B200ITP_F32)
      grid_expand_kernel<float><<<blocks, 256, 0, st>>>((const float *)axis_coords[0], (const float *)(d->ndims > 1 ? axis_coords[1] : nullptr),
                                                        (const float *)(d->ndims > 2 ? axis_coords[2] : nullptr),
                                                        (const float *)(d->ndims > 3 ? axis_coords[3] : nullptr), l0, l1, l2, d->ndims, nq,
                                                        (float *)optr[0], (float *)optr[1], (float *)optr[2], (float *)optr[3]);
    else
      grid_expand_kernel<double><<<blocks, 256, 0, st>>>((const double *)axis_coords[0], (const double *)(d->ndims > 1 ? axis_coords[1] : nullptr),
                                                         (const double *)(d->ndims > 2 ? axis_coords[2] : nullptr),
                                                         (const double *)(d->ndims > 3 ? axis_coords[3] : nullptr), l0, l1, l2, d->ndims, nq,
                                                         (double *)optr[0], (double *)optr[1], (double *)optr[2], (double *)optr[3]);
    count_launch();
    B200_CUDA_OK(cudaGetLastError());
  }
  return eval_sorted_common(dev, d, coefs, cptr, nq, out_value, nullptr, first_oob, sp + cb * d->ndims, sbytes - cb * d->ndims,
                            stream);
}
