// Micro-benchmarks behind the query-ordering design: how does the B200 L2 treat
//   T1  appends of 16-byte records to B open bin tails (atomic cursor + 16 B store)
//   T2  random 4-byte writes into a window of W bytes
//   T3  CTA-local multisplit (shared-memory counting sort of a chunk, run-wise coalesced append)
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o build/ubench_scatter tools/ubench_scatter.cu
// Each test prints time, bytes/record algorithmic and GB/s; run under ncu for dram__bytes_*.
#include <cuda_runtime.h>

#include <cstdint>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                  \
  do {                                                                         \
    cudaError_t e = (x);                                                       \
    if (e != cudaSuccess) {                                                    \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
      exit(1);                                                                 \
    }                                                                          \
  } while (0)

__device__ __forceinline__ unsigned hash32(unsigned x) {
  x ^= x >> 16;
  x *= 0x7feb352du;
  x ^= x >> 15;
  x *= 0x846ca68bu;
  x ^= x >> 16;
  return x;
}

// ---------------------------------------------------------------- T1
template <int MODE>
__global__ void __launch_bounds__(256) append16(const float *__restrict__ xs, unsigned n, unsigned nbins,
                                                unsigned *__restrict__ cursor, float4 *__restrict__ rec, size_t cap) {
  const unsigned stride = gridDim.x * blockDim.x;
  for (unsigned q = blockIdx.x * blockDim.x + threadIdx.x; q < n; q += stride) {
    const float x = xs[q];
    const unsigned key = hash32(q) % nbins;
    const unsigned pos = atomicAdd(cursor + key, 1u);
    if (pos >= cap) continue;
    float4 r = make_float4(x, x, x, __uint_as_float(q));
    if (MODE == 0) {
      rec[pos] = r;
    } else if (MODE == 1) {
      __stcg(rec + pos, r);
    } else if (MODE == 2) {
      __stcs(rec + pos, r);
    } else if (MODE == 3) {
      __stwt(rec + pos, r);
    }
  }
}

// 32-byte records (full sector) for comparison
__global__ void __launch_bounds__(256) append32(const float *__restrict__ xs, unsigned n, unsigned nbins,
                                                unsigned *__restrict__ cursor, float4 *__restrict__ rec, size_t cap) {
  const unsigned stride = gridDim.x * blockDim.x;
  for (unsigned q = blockIdx.x * blockDim.x + threadIdx.x; q < n; q += stride) {
    const float x = xs[q];
    const unsigned key = hash32(q) % nbins;
    const unsigned pos = atomicAdd(cursor + key, 1u);
    if (2 * (size_t)pos + 1 >= cap) continue;
    float4 r = make_float4(x, x, x, __uint_as_float(q));
    rec[2 * (size_t)pos] = r;
    rec[2 * (size_t)pos + 1] = r;
  }
}

// ---------------------------------------------------------------- T2
__global__ void __launch_bounds__(256) scatter4(unsigned n, unsigned wmask, float *__restrict__ out) {
  const unsigned stride = gridDim.x * blockDim.x;
  for (unsigned q = blockIdx.x * blockDim.x + threadIdx.x; q < n; q += stride) {
    out[hash32(q) & wmask] = (float)q;
  }
}
// window moves: chunk c of `chunk` consecutive writes goes to window c (bijective inside the window)
__global__ void __launch_bounds__(256) scatter4_windowed(unsigned n, unsigned wbits, float *__restrict__ out) {
  const unsigned stride = gridDim.x * blockDim.x;
  const unsigned wmask = (1u << wbits) - 1u;
  for (unsigned q = blockIdx.x * blockDim.x + threadIdx.x; q < n; q += stride) {
    const unsigned win = q >> wbits;
    const unsigned local = (hash32(q & wmask) * 2654435761u) & wmask;  // not bijective but uniform
    out[((size_t)win << wbits) + local] = (float)q;
  }
}

// ---------------------------------------------------------------- T3: CTA-local multisplit
// chunk of S records per CTA; keys in [0, NB); shared histogram, scan, local scatter, run-wise append.
template <int NB, int S>
__global__ void __launch_bounds__(512) multisplit(const float *__restrict__ xs, const float *__restrict__ ys,
                                                  const float *__restrict__ zs, unsigned n,
                                                  unsigned *__restrict__ cursor, float4 *__restrict__ rec, size_t cap) {
  extern __shared__ float4 sm[];
  float4 *srec = sm;                                   // S
  unsigned *shist = reinterpret_cast<unsigned *>(srec + S);  // NB
  unsigned *sbase = shist + NB;                        // NB (global base of each run)
  unsigned *sstart = sbase + NB;                       // NB
  constexpr int T = 512;
  constexpr int RPT = S / T;
  for (unsigned base = blockIdx.x * S; base < n; base += gridDim.x * S) {
    for (int i = threadIdx.x; i < NB; i += T) shist[i] = 0;
    __syncthreads();
    float4 r[RPT];
    unsigned kr[RPT];
#pragma unroll
    for (int j = 0; j < RPT; ++j) {
      const unsigned q = base + j * T + threadIdx.x;
      kr[j] = 0xffffffffu;
      if (q < n) {
        r[j] = make_float4(xs[q], ys[q], zs[q], __uint_as_float(q));
        const unsigned key = hash32(q) % NB;
        const unsigned rank = atomicAdd(&shist[key], 1u);
        kr[j] = (key << 16) | rank;
      }
    }
    __syncthreads();
    // per-bin: reserve global space, compute local start (simple: thread i handles bins i, i+T, ...)
    // local exclusive scan over NB bins by warp 0.. (NB <= 1024): do a block scan
    {
      // Hillis-Steele over NB entries using sstart as buffer
      for (int i = threadIdx.x; i < NB; i += T) sstart[i] = shist[i];
      __syncthreads();
      for (int o = 1; o < NB; o <<= 1) {
        unsigned v[(NB + T - 1) / T];
        int c = 0;
        for (int i = threadIdx.x; i < NB; i += T, ++c) v[c] = i >= o ? sstart[i - o] : 0u;
        __syncthreads();
        c = 0;
        for (int i = threadIdx.x; i < NB; i += T, ++c) sstart[i] += v[c];
        __syncthreads();
      }
      for (int i = threadIdx.x; i < NB; i += T) {
        const unsigned cnt = shist[i];
        sstart[i] -= cnt;  // exclusive
        sbase[i] = cnt ? atomicAdd(cursor + i, cnt) : 0u;
      }
      __syncthreads();
    }
#pragma unroll
    for (int j = 0; j < RPT; ++j)
      if (kr[j] != 0xffffffffu) srec[sstart[kr[j] >> 16] + (kr[j] & 0xffffu)] = r[j];
    __syncthreads();
    // write out: thread i writes sorted record i to its bin's run
    const unsigned cnt_total = min((unsigned)S, n - base);
    for (unsigned i = threadIdx.x; i < cnt_total; i += T) {
      // find the bin of sorted position i: binary search in sstart
      int lo = 0, hi = NB - 1;
      while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (sstart[mid] <= i) lo = mid; else hi = mid - 1;
      }
      const size_t dst = (size_t)sbase[lo] + (i - sstart[lo]);
      if (dst < cap) rec[dst] = srec[i];
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------- plain streaming copy for reference
__global__ void __launch_bounds__(256) stream_copy(const float4 *__restrict__ a, float4 *__restrict__ b, size_t n) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) b[i] = a[i];
}

struct Timer {
  cudaEvent_t a, b;
  Timer() {
    CK(cudaEventCreate(&a));
    CK(cudaEventCreate(&b));
  }
  void start() { CK(cudaEventRecord(a)); }
  float stop() {
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms;
    CK(cudaEventElapsedTime(&ms, a, b));
    return ms;
  }
};

static void set_cursors(unsigned *cursor, unsigned nbins, unsigned per_bin, std::vector<unsigned> &h) {
  h.resize(nbins);
  for (unsigned i = 0; i < nbins; ++i) h[i] = i * per_bin;
  CK(cudaMemcpy(cursor, h.data(), nbins * 4, cudaMemcpyHostToDevice));
}

int main(int argc, char **argv) {
  const unsigned n = argc > 1 ? (unsigned)atol(argv[1]) : (1u << 27);
  const int only = argc > 2 ? atoi(argv[2]) : 0;
  int sms = 148;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  const int blocks = sms * 16;
  float *xs, *ys, *zs;
  CK(cudaMalloc(&xs, (size_t)n * 4));
  CK(cudaMalloc(&ys, (size_t)n * 4));
  CK(cudaMalloc(&zs, (size_t)n * 4));
  CK(cudaMemset(xs, 0, (size_t)n * 4));
  CK(cudaMemset(ys, 0, (size_t)n * 4));
  CK(cudaMemset(zs, 0, (size_t)n * 4));
  float4 *rec;
  const size_t rec_bytes = (size_t)n * 48 + ((size_t)1 << 30);
  CK(cudaMalloc(&rec, rec_bytes));
  CK(cudaMemset(rec, 0, rec_bytes));
  unsigned *cursor;
  CK(cudaMalloc(&cursor, (4u << 20) * 4));
  std::vector<unsigned> h;
  Timer t;

  if (!only || only == 9) {
    stream_copy<<<blocks, 256>>>(rec, rec + (size_t)n / 2, (size_t)n / 2);
    t.start();
    stream_copy<<<blocks, 256>>>(rec, rec + (size_t)n / 2, (size_t)n / 2);
    const float ms = t.stop();
    printf("copy   %8.3f ms  %7.1f GB/s (read+write)\n", ms, (double)n / 2 * 32 / ms / 1e6);
  }

  if (!only || only == 1) {
    const unsigned bins_list[] = {1024, 8192, 65536, 524288};
    for (unsigned nb : bins_list) {
      const unsigned per_bin = (unsigned)((size_t)n / nb + 8 * (unsigned)sqrt((double)n / nb) + 16) & ~1u;
      for (int mode = 0; mode < 4; ++mode) {
        float best = 1e9f;
        for (int rep = 0; rep < 3; ++rep) {
          set_cursors(cursor, nb, per_bin, h);
          t.start();
          if (mode == 0) append16<0><<<blocks, 256>>>(xs, n, nb, cursor, rec, rec_bytes / 16);
          if (mode == 1) append16<1><<<blocks, 256>>>(xs, n, nb, cursor, rec, rec_bytes / 16);
          if (mode == 2) append16<2><<<blocks, 256>>>(xs, n, nb, cursor, rec, rec_bytes / 16);
          if (mode == 3) append16<3><<<blocks, 256>>>(xs, n, nb, cursor, rec, rec_bytes / 16);
          best = fminf(best, t.stop());
        }
        CK(cudaGetLastError());
        printf("append16 bins=%7u mode=%d  %8.3f ms  %6.2f Grec/s  alg %6.1f GB/s\n", nb, mode, best, n / best / 1e6,
               (double)n * 20 / best / 1e6);
      }
      {
        float best = 1e9f;
        for (int rep = 0; rep < 3; ++rep) {
          set_cursors(cursor, nb, per_bin, h);
          t.start();
          append32<<<blocks, 256>>>(xs, n, nb, cursor, rec, rec_bytes / 16);
          best = fminf(best, t.stop());
        }
        CK(cudaGetLastError());
        printf("append32 bins=%7u         %8.3f ms  %6.2f Grec/s  alg %6.1f GB/s\n", nb, best, n / best / 1e6,
               (double)n * 36 / best / 1e6);
      }
    }
  }

  if (!only || only == 2) {
    float *out = reinterpret_cast<float *>(rec);
    for (int wb = 20; wb <= 30; ++wb) {  // window of 2^wb floats = 4 MiB .. 4 GiB
      if (((size_t)4 << wb) > rec_bytes) break;
      const unsigned wmask = (1u << wb) - 1u;
      scatter4<<<blocks, 256>>>(n, wmask, out);
      t.start();
      scatter4<<<blocks, 256>>>(n, wmask, out);
      const float ms = t.stop();
      printf("scatter4 window=%6u MiB  %8.3f ms  %6.2f Gwrites/s\n", (4u << wb) >> 20, ms, n / ms / 1e6);
    }
    for (int wb = 20; wb <= 26; ++wb) {
      scatter4_windowed<<<blocks, 256>>>(n, wb, out);
      t.start();
      scatter4_windowed<<<blocks, 256>>>(n, wb, out);
      const float ms = t.stop();
      printf("scatter4 moving window=%6u MiB  %8.3f ms  %6.2f Gwrites/s\n", (4u << wb) >> 20, ms, n / ms / 1e6);
    }
  }

  if (!only || only == 3) {
    {
      constexpr int NB = 512, S = 8192;
      const size_t smem = (size_t)S * 16 + NB * 12;
      CK(cudaFuncSetAttribute(multisplit<NB, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      const unsigned per_bin = (unsigned)((size_t)n / NB + 8 * (unsigned)sqrt((double)n / NB) + 16) & ~1u;
      float best = 1e9f;
      for (int rep = 0; rep < 3; ++rep) {
        set_cursors(cursor, NB, per_bin, h);
        t.start();
        multisplit<NB, S><<<sms, 512, smem>>>(xs, ys, zs, n, cursor, rec, rec_bytes / 16);
        best = fminf(best, t.stop());
      }
      CK(cudaGetLastError());
      printf("multisplit NB=%d S=%d  %8.3f ms  %6.2f Grec/s  alg %6.1f GB/s\n", NB, S, best, n / best / 1e6,
             (double)n * 28 / best / 1e6);
    }
    {
      constexpr int NB = 256, S = 4096;
      const size_t smem = (size_t)S * 16 + NB * 12;
      CK(cudaFuncSetAttribute(multisplit<NB, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      const unsigned per_bin = (unsigned)((size_t)n / NB + 8 * (unsigned)sqrt((double)n / NB) + 16) & ~1u;
      float best = 1e9f;
      for (int rep = 0; rep < 3; ++rep) {
        set_cursors(cursor, NB, per_bin, h);
        t.start();
        multisplit<NB, S><<<sms * 3, 512, smem>>>(xs, ys, zs, n, cursor, rec, rec_bytes / 16);
        best = fminf(best, t.stop());
      }
      CK(cudaGetLastError());
      printf("multisplit NB=%d S=%d  %8.3f ms  %6.2f Grec/s  alg %6.1f GB/s\n", NB, S, best, n / best / 1e6,
             (double)n * 28 / best / 1e6);
    }
  }
  printf("done\n");
  return 0;
}
