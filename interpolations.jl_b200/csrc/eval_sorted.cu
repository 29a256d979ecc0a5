// b200itp_eval_sorted: scattered evaluation with L2/shared-memory friendly query ordering.
//
// A uniformly scattered tricubic query touches 16 (y,z) rows of the coefficient array, i.e. 16-32 DRAM
// sectors when the array does not fit the L2, and 64 different L1 lines per warp instruction even when
// it does: the direct kernel (eval.cuh) is bound by that, not by arithmetic.  This path reorders the
// queries on the device, inside the call, so that the coefficients are read once:
//   1. count   : key(q) = brick of 32 x BY x BZ cells holding the query's first tap; histogram in L2
//   2. scan    : exclusive prefix sum of the histogram
//   3. scatter : records {x, y, z, q} (16 B) are appended to their brick's slot (one open 32-byte tail
//                per brick: all tails stay L2-resident, so DRAM sees full-sector streaming writes)
//   4. evaluate: one CTA per (brick, slice of its records): the brick plus its 3-cell halo is staged in
//                shared memory once (periodic wrap resolved while staging); the slice is counting-sorted
//                in shared memory by the (y,z) cell inside the brick, so that every warp serves 32 records
//                of (nearly) one cell column and its shared-memory gathers fall on consecutive words of
//                one row: conflict-free.  Results go to out[q]: the caller's order.
// The per-query arithmetic is the same device code as the direct kernel (axis_setup + the reference's
// reduction order), so both paths return bit-identical values.
//
// Scope: BSplineInterpolation without wrappers, Float32 coefficients and coordinates, 2-D / 3-D,
// uniform Quadratic or Cubic degree.  Everything else is forwarded to the direct kernel.
#include "eval.cuh"

#include <climits>

namespace b200itp {

int fill_eval_params_public(const b200itp_desc *D, const void *coefs, const void *const *coords, int64_t nq,
                            void *out_value, unsigned long long *first_oob_dev, EvalParams &P, int *uniform_degree,
                            bool *plain_f32);

template <int N>
struct Brick;
template <>
struct Brick<3> {
  static constexpr int BX = 32, BY = 8, BZ = 8;
};
template <>
struct Brick<2> {
  static constexpr int BX = 32, BY = 16, BZ = 1;
};

struct SortGeom {
  int nb[3];       // bricks per axis
  int cells[3];    // number of distinct first-tap indices per axis (first in [0, cells))
  long long nbins; // bricks
};

// first tap index per axis (bit-identical to what the evaluation will compute) and in-bounds test
template <int N, int DEG>
__device__ __forceinline__ bool query_key(const EvalParams &P, const SortGeom &G, const float (&x)[N], long long &key) {
  bool inb = true;
#pragma unroll
  for (int d = 0; d < N; ++d) {
    const double xd = (double)x[d];
    inb = inb && (P.lb[d] <= xd) && (xd <= P.ub[d]);
  }
  if (!inb) return false;
  int c[3] = {0, 0, 0};
#pragma unroll
  for (int d = 0; d < N; ++d) {
    AxisTaps<DEG, float> A;
    axis_setup<DEG, float, false>(P, d, x[d], A);
    c[d] = min(max(A.first, 0), G.cells[d] - 1);
  }
  using Bk = Brick<N>;
  const int bx = c[0] / Bk::BX, by = c[1] / Bk::BY, bz = N == 3 ? c[2] / Bk::BZ : 0;
  key = ((long long)bz * G.nb[1] + by) * G.nb[0] + bx;
  return true;
}

template <int N, int DEG>
__global__ void __launch_bounds__(256) sort_count_kernel(const __grid_constant__ EvalParams P, const SortGeom G,
                                                         unsigned *__restrict__ hist) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < P.nq; q += stride) {
    float x[N];
#pragma unroll
    for (int d = 0; d < N; ++d) x[d] = reinterpret_cast<const float *>(P.coords[d])[q];
    long long key;
    if (query_key<N, DEG>(P, G, x, key)) {
      atomicAdd(hist + key, 1u);
    } else {
      // BoundsError (indexing.jl:6): NaN result, smallest offending index reported
      reinterpret_cast<float *>(P.out_value)[q] = nanf("");
      atomicMin(P.first_oob, (unsigned long long)q);
    }
  }
}

template <int N, int DEG>
__global__ void __launch_bounds__(256) sort_scatter_kernel(const __grid_constant__ EvalParams P, const SortGeom G,
                                                           unsigned *__restrict__ cursor, float4 *__restrict__ rec) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < P.nq; q += stride) {
    float x[N];
#pragma unroll
    for (int d = 0; d < N; ++d) x[d] = reinterpret_cast<const float *>(P.coords[d])[q];
    long long key;
    if (query_key<N, DEG>(P, G, x, key)) {
      const unsigned pos = atomicAdd(cursor + key, 1u);
      float4 r;
      r.x = x[0];
      r.y = x[1];
      r.z = N == 3 ? x[2] : 0.f;
      r.w = __uint_as_float((unsigned)q);
      rec[pos] = r;
    }
  }
}

// ---- exclusive scan of the histogram (in place), three small kernels
constexpr int kScanBlock = 1024;
constexpr int kScanPer = 4;  // elements per thread
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
  // single CTA: sequential chunks of 1024 block sums (nblocks is at most a few 10^4)
  __shared__ unsigned buf[1024];
  __shared__ unsigned carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < nblocks; base += 1024) {
    const int i = base + threadIdx.x;
    const unsigned v = i < nblocks ? sums[i] : 0u;
    buf[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {  // Hillis-Steele inclusive scan
      const unsigned t = threadIdx.x >= o ? buf[threadIdx.x - o] : 0u;
      __syncthreads();
      buf[threadIdx.x] += t;
      __syncthreads();
    }
    if (i < nblocks) sums[i] = carry + buf[threadIdx.x] - v;  // exclusive
    __syncthreads();
    if (threadIdx.x == 0) carry += buf[1023];
    __syncthreads();
  }
}
__global__ void __launch_bounds__(kScanBlock) scan_final_kernel(unsigned *hist, long long n,
                                                                 const unsigned *__restrict__ sums,
                                                                 unsigned *__restrict__ start /* n+1 */) {
  // exclusive scan inside the block's kScanBlock*kScanPer elements; element order: i = base + t*kScanPer + k
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
      hist[i] = run;  // becomes the scatter cursor
    }
    run += v[k];
    if (i == n - 1) start[n] = run;
  }
}

// ---- evaluation of one brick from shared memory
template <int N, int DEG>
struct TileCfg {
  using Bk = Brick<N>;
  static constexpr int L = DEG + 1;
  static constexpr int TX = Bk::BX + DEG;
  static constexpr int TY = Bk::BY + DEG;
  static constexpr int TZ = N == 3 ? Bk::BZ + DEG : 1;
  static constexpr int PX = TX | 1;  // odd row pitch: rows of neighbouring (y,z) start on different banks
  static constexpr int FLOATS = PX * TY * TZ;
};

constexpr int kSplit = 4;      // CTAs per brick (each takes every kSplit-th slice of the brick's records)
constexpr int kSlice = 1536;   // records per slice (static shared memory: tile + slice < 48 KB)
constexpr int kEvalThreads = 256;
constexpr int kRPT = kSlice / kEvalThreads;  // records per thread per slice

template <int N, int DEG>
__global__ void __launch_bounds__(kEvalThreads) sort_eval_kernel(const __grid_constant__ EvalParams P,
                                                                 const SortGeom G,
                                                                 const unsigned *__restrict__ start,
                                                                 const float4 *__restrict__ rec) {
  using TC = TileCfg<N, DEG>;
  using Bk = Brick<N>;
  constexpr int L = TC::L;
  constexpr int FPB = Bk::BY * Bk::BZ;  // (y,z) cell columns per brick
  __shared__ float tile[TC::FLOATS];
  __shared__ float4 srec[kSlice];
  __shared__ unsigned shist[FPB];
  __shared__ unsigned sstart[FPB];
  const long long brick = blockIdx.x;
  const int split = blockIdx.y;
  const unsigned r_begin = start[brick], r_end = start[brick + 1];
  if (r_begin + (unsigned)split * kSlice >= r_end) return;

  // brick origin in first-tap index space
  const int bx = (int)(brick % G.nb[0]);
  const int by = (int)((brick / G.nb[0]) % G.nb[1]);
  const int bz = (int)(brick / ((long long)G.nb[0] * G.nb[1]));
  const int o[3] = {bx * Bk::BX, by * Bk::BY, bz * Bk::BZ};
  const float *coefs = reinterpret_cast<const float *>(P.coefs);

  // ---- stage the brick + halo; tile element (lx,ly,lz) <-> parent index o + l on each axis
  for (int t = threadIdx.x; t < TC::TX * TC::TY * TC::TZ; t += blockDim.x) {
    const int lx = t % TC::TX;
    const int ly = (t / TC::TX) % TC::TY;
    const int lz = t / (TC::TX * TC::TY);
    const int l[3] = {lx, ly, lz};
    long long off = 0;
    bool ok = true;
#pragma unroll
    for (int d = 0; d < N; ++d) {
      const int idx = o[d] + l[d];
      const int pos = P.periodic[d] ? modrange1(idx, P.n[d]) - 1 : idx - P.coef_first[d];
      const int size_d = P.n[d] + (P.coef_first[d] == 0 ? 2 : 0);
      ok = ok && pos >= 0 && pos < size_d;
      off += (long long)pos * P.stride[d];
    }
    tile[(lz * TC::TY + ly) * TC::PX + lx] = ok ? __ldg(coefs + off) : 0.f;
  }

  for (unsigned sbase = r_begin + (unsigned)split * kSlice; sbase < r_end; sbase += (unsigned)kSplit * kSlice) {
    const unsigned cnt = min(r_end - sbase, (unsigned)kSlice);
    if (threadIdx.x < FPB) shist[threadIdx.x] = 0;
    __syncthreads();  // also: tile staged / previous slice fully consumed
    // ---- counting sort of the slice by (y,z) cell column, in shared memory
    float4 r[kRPT];
    unsigned keyrank[kRPT];
#pragma unroll
    for (int j = 0; j < kRPT; ++j) {
      const unsigned i = (unsigned)j * kEvalThreads + threadIdx.x;
      keyrank[j] = 0xffffffffu;
      if (i < cnt) {
        r[j] = rec[sbase + i];
        int col = 0;
        {
          AxisTaps<DEG, float> A;
          axis_setup<DEG, float, false>(P, 1, r[j].y, A);
          col = min(max(A.first, 0), G.cells[1] - 1) - o[1];
        }
        if (N == 3) {
          AxisTaps<DEG, float> A;
          axis_setup<DEG, float, false>(P, N - 1, r[j].z, A);
          col += (min(max(A.first, 0), G.cells[N - 1] - 1) - o[2]) * Bk::BY;
        }
        const unsigned rank = atomicAdd(&shist[col], 1u);
        keyrank[j] = ((unsigned)col << 16) | rank;
      }
    }
    __syncthreads();
    if (threadIdx.x < 32) {  // exclusive scan of FPB <= 64 counters by one warp
      unsigned a = threadIdx.x < FPB ? shist[threadIdx.x] : 0u;
      unsigned b = threadIdx.x + 32 < FPB ? shist[threadIdx.x + 32] : 0u;
      unsigned ia = a, ib = b;
#pragma unroll
      for (int w = 1; w < 32; w <<= 1) {
        const unsigned ta = __shfl_up_sync(0xffffffffu, ia, w);
        const unsigned tb = __shfl_up_sync(0xffffffffu, ib, w);
        if (threadIdx.x >= w) {
          ia += ta;
          ib += tb;
        }
      }
      const unsigned tot_a = __shfl_sync(0xffffffffu, ia, 31);
      if (threadIdx.x < FPB) sstart[threadIdx.x] = ia - a;
      if (threadIdx.x + 32 < FPB) sstart[threadIdx.x + 32] = tot_a + ib - b;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < kRPT; ++j)
      if (keyrank[j] != 0xffffffffu) srec[sstart[keyrank[j] >> 16] + (keyrank[j] & 0xffffu)] = r[j];
    __syncthreads();
    // ---- evaluate in sorted order: a warp's 32 records share (nearly) one cell column
    for (unsigned i = threadIdx.x; i < cnt; i += kEvalThreads) {
      const float4 q4 = srec[i];
      const float x[3] = {q4.x, q4.y, q4.z};
      AxisTaps<DEG, float> ax[N];
      int lpos[3] = {0, 0, 0};
#pragma unroll
      for (int d = 0; d < N; ++d) {
        axis_setup<DEG, float, false>(P, d, x[d], ax[d]);
        lpos[d] = min(max(ax[d].first, 0), G.cells[d] - 1) - o[d];
      }
      // value = sum_k0 w0[k0] * ( sum_k1 w1[k1] * ( sum_k2 w2[k2] * c[k2][k1][k0] ) )   (Interpolations.jl:304-316)
      const float *tb = tile + (lpos[2] * TC::TY + lpos[1]) * TC::PX + lpos[0];
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
      reinterpret_cast<float *>(P.out_value)[__float_as_uint(q4.w)] = val;
    }
  }
}

template <int N, int DEG>
static int run_sorted(int dev, EvalParams &P, const b200itp_desc *D, void *scratch, size_t scratch_bytes,
                      cudaStream_t st, bool *done) {
  using Bk = Brick<N>;
  *done = false;
  SortGeom G;
  long long bricks = 1;
  for (int d = 0; d < 3; ++d) {
    G.nb[d] = 1;
    G.cells[d] = 1;
  }
  const int bdim[3] = {Bk::BX, Bk::BY, Bk::BZ};
  for (int d = 0; d < N; ++d) {
    // first-tap indices: periodic 0..n-1, otherwise 0..size-1-DEG
    G.cells[d] = P.periodic[d] ? P.n[d] : (int)D->size[d] - DEG;
    if (G.cells[d] < 1) return 0;
    G.nb[d] = (G.cells[d] + bdim[d] - 1) / bdim[d];
    bricks *= G.nb[d];
  }
  G.nbins = bricks;
  if (bricks > 0x7fffffffLL || G.nbins > (1LL << 27) || P.nq > 0xfffffff0LL) return 0;
  const size_t need = (size_t)P.nq * 16 + (size_t)(G.nbins + 1) * 8 + 65536 * 4 + 1024;
  if (!scratch || scratch_bytes < need || ((uintptr_t)scratch % 16) != 0) return 0;
  float4 *rec = reinterpret_cast<float4 *>(scratch);
  unsigned *hist = reinterpret_cast<unsigned *>(rec + P.nq);
  unsigned *start = hist + (G.nbins + 1);
  unsigned *sums = start + (G.nbins + 1);
  const long long per = (long long)kScanBlock * kScanPer;
  const long long nsb = (G.nbins + per - 1) / per;
  if (nsb > 65536) return 0;

  B200_CUDA_OK(cudaMemsetAsync(hist, 0, (size_t)(G.nbins + 1) * 4, st));
  const int blocks = (int)std::min<long long>((P.nq + 255) / 256, (long long)sm_count(dev) * 16);
  sort_count_kernel<N, DEG><<<blocks, 256, 0, st>>>(P, G, hist);
  scan_partial_kernel<<<(unsigned)nsb, kScanBlock, 0, st>>>(hist, G.nbins, sums);
  scan_sums_kernel<<<1, 1024, 0, st>>>(sums, (int)nsb);
  scan_final_kernel<<<(unsigned)nsb, kScanBlock, 0, st>>>(hist, G.nbins, sums, start);
  sort_scatter_kernel<N, DEG><<<blocks, 256, 0, st>>>(P, G, hist, rec);
  dim3 grid((unsigned)bricks, kSplit, 1);
  sort_eval_kernel<N, DEG><<<grid, kEvalThreads, 0, st>>>(P, G, start, rec);
  count_launch(6);
  B200_CUDA_OK(cudaGetLastError());
  *done = true;
  return 0;
}

}  // namespace b200itp

using namespace b200itp;

extern "C" size_t b200itp_eval_sorted_scratch_bytes(const b200itp_desc *d, int64_t nq) {
  if (!d || nq <= 0) return 0;
  long long cells = 1;
  for (int i = 0; i < d->ndims && i < B200ITP_MAX_DIMS; ++i) cells *= (d->size[i] > 0 ? d->size[i] : 1);
  // records + histogram/start arrays (one bin per run of 32 cells, generously rounded) + scan sums
  return (size_t)nq * 16 + (size_t)(cells / 16 + 4096) * 8 + 65536 * 4 + 4096;
}

extern "C" int b200itp_eval_sorted(int dev, const b200itp_desc *d, const void *coefs, const void *const *coords,
                                   int64_t nq, void *out_value, int64_t *first_oob, void *scratch,
                                   size_t scratch_bytes, void *stream) {
  if (!d || !coefs || !coords || !out_value) B200_FAIL(B200ITP_E_BADARG, "eval_sorted: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  int deg = 0;
  bool plain = false;
  EvalParams P;
  DeviceGuard g(dev);
  if (!g.ok) B200_FAIL(-(int)cudaErrorInvalidDevice, "eval_sorted: cannot select device %d", dev);
  void *flag = nullptr;
  int rc = scratch_get(dev, 2, 256, &flag);
  if (rc) return rc;
  rc = fill_eval_params_public(d, coefs, coords, nq, out_value, (unsigned long long *)flag, P, &deg, &plain);
  if (rc) return rc;
  bool done = false;
  if (plain && nq > 0 && (deg == B200ITP_CUBIC || deg == B200ITP_QUADRATIC) && (d->ndims == 2 || d->ndims == 3)) {
    B200_CUDA_OK(cudaMemsetAsync(flag, 0xFF, sizeof(unsigned long long), st));
    if (d->ndims == 3 && deg == 3) rc = run_sorted<3, 3>(dev, P, d, scratch, scratch_bytes, st, &done);
    if (d->ndims == 3 && deg == 2) rc = run_sorted<3, 2>(dev, P, d, scratch, scratch_bytes, st, &done);
    if (d->ndims == 2 && deg == 3) rc = run_sorted<2, 3>(dev, P, d, scratch, scratch_bytes, st, &done);
    if (d->ndims == 2 && deg == 2) rc = run_sorted<2, 2>(dev, P, d, scratch, scratch_bytes, st, &done);
    if (rc) return rc;
    if (done) {
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
  return b200itp_eval(dev, d, coefs, coords, nq, out_value, nullptr, first_oob, nullptr, nullptr, stream);
}
