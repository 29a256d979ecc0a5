// Micro-benchmark of the CTA-local multisplit (the streaming pass of the ordered evaluation):
// chunk size S, threads T, bins NB, CTAs per SM, register prefetch of the next chunk.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o build/ubench_split tools/ubench_split.cu
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                        \
  do {                                                                               \
    cudaError_t e = (x);                                                             \
    if (e != cudaSuccess) {                                                          \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
      exit(1);                                                                       \
    }                                                                                \
  } while (0)

__device__ __forceinline__ unsigned hash32(unsigned x) {
  x ^= x >> 16;
  x *= 0x7feb352du;
  x ^= x >> 15;
  x *= 0x846ca68bu;
  x ^= x >> 16;
  return x;
}

template <int S, int NBMAX>
struct Smem {
  float4 rec[S];
  unsigned short bin[S];
  unsigned hist[NBMAX];
  unsigned start[NBMAX];
  unsigned gbase[NBMAX];
  unsigned wtot[32];
};

// bins per thread = NBMAX / T (>= 1)
template <int S, int T, int NBMAX, bool PREFETCH, int MINB>
__global__ void __launch_bounds__(T, MINB) split_kernel(const float4 *__restrict__ in, unsigned n, int nbins,
                                                        unsigned *__restrict__ cursor, float4 *__restrict__ out,
                                                        size_t cap) {
  extern __shared__ __align__(16) unsigned char raw[];
  auto &sm = *reinterpret_cast<Smem<S, NBMAX> *>(raw);
  constexpr int RPT = S / T;
  constexpr int BPT = NBMAX / T;
  static_assert(BPT >= 1, "bins per thread");
  const int t = threadIdx.x;
  const unsigned nchunks = (n + S - 1) / S;
  float4 nr[RPT];
  auto fetch = [&](unsigned ch) {
#pragma unroll
    for (int j = 0; j < RPT; ++j) {
      const unsigned i = ch * S + j * T + t;
      if (ch < nchunks && i < n) nr[j] = in[i];
    }
  };
  if (PREFETCH) fetch(blockIdx.x);
  for (unsigned ch = blockIdx.x; ch < nchunks; ch += gridDim.x) {
    const unsigned base = ch * S;
    const unsigned cnt = min((unsigned)S, n - base);
    float4 r[RPT];
    unsigned key[RPT];
    if (!PREFETCH) fetch(ch);
#pragma unroll
    for (int j = 0; j < RPT; ++j) {
      const unsigned i = base + j * T + t;
      key[j] = 0xffffffffu;
      if (i < n) {
        r[j] = nr[j];
        key[j] = hash32(__float_as_uint(r[j].w)) % (unsigned)nbins;
      }
    }
    if (PREFETCH) fetch(ch + gridDim.x);
#pragma unroll
    for (int b = 0; b < BPT; ++b) sm.hist[b * T + t] = 0;
    __syncthreads();
    unsigned rank[RPT];
#pragma unroll
    for (int j = 0; j < RPT; ++j)
      if (key[j] != 0xffffffffu) rank[j] = atomicAdd(&sm.hist[key[j]], 1u);
    __syncthreads();
    unsigned gb[BPT];
    {
      unsigned v[BPT], tsum = 0;
#pragma unroll
      for (int b = 0; b < BPT; ++b) {
        v[b] = sm.hist[t * BPT + b];
        tsum += v[b];
      }
      unsigned inc = tsum;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned u = __shfl_up_sync(0xffffffffu, inc, o);
        if ((t & 31) >= o) inc += u;
      }
      if ((t & 31) == 31) sm.wtot[t >> 5] = inc;
      __syncthreads();
      if (t < 32) {
        const unsigned w = t < T / 32 ? sm.wtot[t] : 0u;
        unsigned winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const unsigned u = __shfl_up_sync(0xffffffffu, winc, o);
          if (t >= o) winc += u;
        }
        sm.wtot[t] = winc - w;
      }
      __syncthreads();
      unsigned st = sm.wtot[t >> 5] + inc - tsum;
#pragma unroll
      for (int b = 0; b < BPT; ++b) {
        sm.start[t * BPT + b] = st;
        gb[b] = v[b] ? atomicAdd(cursor + t * BPT + b, v[b]) - st : 0u;
        st += v[b];
      }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < RPT; ++j)
      if (key[j] != 0xffffffffu) {
        const unsigned slot = sm.start[key[j]] + rank[j];
        sm.rec[slot] = r[j];
        sm.bin[slot] = (unsigned short)key[j];
      }
#pragma unroll
    for (int b = 0; b < BPT; ++b) sm.gbase[t * BPT + b] = gb[b];
    __syncthreads();
    for (unsigned i = t; i < cnt; i += T) {
      const size_t dst = (size_t)(sm.gbase[sm.bin[i]] + i);
      if (dst < cap) out[dst] = sm.rec[i];
    }
    __syncthreads();
  }
}


// ---- variant: warp-private histograms + match.any ranking (no shared-memory atomics)
template <int S, int T, int NB>
struct SmemM {
  float4 rec[S];
  unsigned short bin[S];
  unsigned whist[T / 32][NB];  // per-warp counters, then per-warp exclusive offsets inside the bin
  unsigned char tag[T / 32][NB];
  unsigned start[NB];
  unsigned gbase[NB];
  unsigned wtot[32];
};

template <int S, int T, int NB, bool PREFETCH, bool TAG>
__global__ void __launch_bounds__(T, 1) split_match_kernel(const float4 *__restrict__ in, unsigned n, int nbins,
                                                           unsigned *__restrict__ cursor, float4 *__restrict__ out,
                                                           size_t cap) {
  extern __shared__ __align__(16) unsigned char raw[];
  auto &sm = *reinterpret_cast<SmemM<S, T, NB> *>(raw);
  constexpr int RPT = S / T;
  constexpr int NW = T / 32;
  static_assert(NB == T, "one bin per thread");
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const unsigned nchunks = (n + S - 1) / S;
  float4 nr[RPT];
  auto fetch = [&](unsigned ch) {
#pragma unroll
    for (int j = 0; j < RPT; ++j) {
      const unsigned i = ch * S + j * T + t;
      if (ch < nchunks && i < n) nr[j] = in[i];
    }
  };
  if (PREFETCH) fetch(blockIdx.x);
  for (unsigned ch = blockIdx.x; ch < nchunks; ch += gridDim.x) {
    const unsigned base = ch * S;
    const unsigned cnt = min((unsigned)S, n - base);
    float4 r[RPT];
    unsigned key[RPT];
    if (!PREFETCH) fetch(ch);
#pragma unroll
    for (int j = 0; j < RPT; ++j) {
      const unsigned i = base + j * T + t;
      key[j] = 0xffffffffu;
      if (i < n) {
        r[j] = nr[j];
        key[j] = hash32(__float_as_uint(r[j].w)) % (unsigned)nbins;
      }
    }
    if (PREFETCH) fetch(ch + gridDim.x);
#pragma unroll
    for (int w = 0; w < NW; ++w) sm.whist[w][t] = 0;
    __syncthreads();
    unsigned rank[RPT];
#pragma unroll
    for (int j = 0; j < RPT; ++j) {
      const unsigned k = key[j];
      if (TAG) {
        // lock-free ranking on the warp's private counters: every pending lane writes its id into the tag of its
        // bin, the lane whose id survives owns the bin for this round and bumps the counter with plain accesses
        bool pend = k != 0xffffffffu;
        while (true) {
          if (pend) sm.tag[warp][k] = (unsigned char)lane;
          __syncwarp();
          const bool win = pend && sm.tag[warp][k] == (unsigned char)lane;
          if (win) {
            const unsigned c = sm.whist[warp][k];
            sm.whist[warp][k] = c + 1;
            rank[j] = c;
            pend = false;
          }
          __syncwarp();
          if (!__any_sync(0xffffffffu, pend)) break;
        }
      } else {
        const unsigned peers = __match_any_sync(0xffffffffu, k);
        const int leader = __ffs(peers) - 1;
        unsigned b = 0;
        if (k != 0xffffffffu && lane == leader) {
          b = sm.whist[warp][k];
          sm.whist[warp][k] = b + __popc(peers);
        }
        b = __shfl_sync(0xffffffffu, b, leader);
        rank[j] = b + __popc(peers & ((1u << lane) - 1u));
      }
    }
    __syncthreads();
    unsigned gb;
    {
      // bin t: exclusive offsets of the warps inside the bin, total count
      unsigned v = 0;
#pragma unroll
      for (int w = 0; w < NW; ++w) {
        const unsigned c = sm.whist[w][t];
        sm.whist[w][t] = v;
        v += c;
      }
      unsigned inc = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned u = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += u;
      }
      if (lane == 31) sm.wtot[warp] = inc;
      __syncthreads();
      if (t < 32) {
        const unsigned w = t < NW ? sm.wtot[t] : 0u;
        unsigned winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const unsigned u = __shfl_up_sync(0xffffffffu, winc, o);
          if (t >= o) winc += u;
        }
        sm.wtot[t] = winc - w;
      }
      __syncthreads();
      const unsigned st = sm.wtot[warp] + inc - v;
      sm.start[t] = st;
      gb = v ? atomicAdd(cursor + t, v) - st : 0u;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < RPT; ++j)
      if (key[j] != 0xffffffffu) {
        const unsigned slot = sm.start[key[j]] + sm.whist[warp][key[j]] + rank[j];
        sm.rec[slot] = r[j];
        sm.bin[slot] = (unsigned short)key[j];
      }
    sm.gbase[t] = gb;
    __syncthreads();
    for (unsigned i = t; i < cnt; i += T) {
      const size_t dst = (size_t)(sm.gbase[sm.bin[i]] + i);
      if (dst < cap) out[dst] = sm.rec[i];
    }
    __syncthreads();
  }
}

template <int S, int T, int NB, bool PF, bool TAG>
static void run_match(const char *name, int nbins, int ctas_per_sm);

__global__ void init_kernel(float4 *in, unsigned n) {
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    in[i] = make_float4(1.f, 2.f, 3.f, __uint_as_float(i));
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

static int g_sms = 148;
static unsigned g_n;
static float4 *g_in, *g_out;
static unsigned *g_cursor;
static size_t g_cap;

template <int S, int T, int NBMAX, bool PF, int MINB>
static void run(const char *name, int nbins, int ctas_per_sm) {
  const size_t smem = sizeof(Smem<S, NBMAX>);
  CK(cudaFuncSetAttribute(split_kernel<S, T, NBMAX, PF, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int occ = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, split_kernel<S, T, NBMAX, PF, MINB>, T, smem));
  const unsigned per_bin = (unsigned)((size_t)g_n / nbins + 8 * (unsigned)sqrt((double)g_n / nbins) + 16);
  std::vector<unsigned> h(NBMAX, 0);
  for (int i = 0; i < nbins; ++i) h[i] = i * per_bin;
  float best = 1e9f;
  for (int rep = 0; rep < 3; ++rep) {
    CK(cudaMemcpy(g_cursor, h.data(), NBMAX * 4, cudaMemcpyHostToDevice));
    Timer t;
    t.start();
    split_kernel<S, T, NBMAX, PF, MINB><<<g_sms * ctas_per_sm, T, smem>>>(g_in, g_n, nbins, g_cursor, g_out, g_cap);
    best = fminf(best, t.stop());
  }
  CK(cudaGetLastError());
  printf("%-34s nbins=%4d occ=%d grid=%d/SM  %8.3f ms  %7.2f Grec/s  %7.1f GB/s\n", name, nbins, occ, ctas_per_sm, best,
         g_n / best / 1e6, (double)g_n * 32 / best / 1e6);
}

template <int S, int T, int NB, bool PF, bool TAG>
static void run_match(const char *name, int nbins, int ctas_per_sm) {
  const size_t smem = sizeof(SmemM<S, T, NB>);
  CK(cudaFuncSetAttribute(split_match_kernel<S, T, NB, PF, TAG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int occ = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, split_match_kernel<S, T, NB, PF, TAG>, T, smem));
  const unsigned per_bin = (unsigned)((size_t)g_n / nbins + 8 * (unsigned)sqrt((double)g_n / nbins) + 16);
  std::vector<unsigned> h(NB, 0);
  for (int i = 0; i < nbins; ++i) h[i] = i * per_bin;
  float best = 1e9f;
  for (int rep = 0; rep < 3; ++rep) {
    CK(cudaMemcpy(g_cursor, h.data(), NB * 4, cudaMemcpyHostToDevice));
    Timer t;
    t.start();
    split_match_kernel<S, T, NB, PF, TAG><<<g_sms * ctas_per_sm, T, smem>>>(g_in, g_n, nbins, g_cursor, g_out, g_cap);
    best = fminf(best, t.stop());
  }
  CK(cudaGetLastError());
  printf("%-34s nbins=%4d occ=%d grid=%d/SM  %8.3f ms  %7.2f Grec/s  %7.1f GB/s\n", name, nbins, occ, ctas_per_sm, best,
         g_n / best / 1e6, (double)g_n * 32 / best / 1e6);
}

int main(int argc, char **argv) {
  g_n = argc > 1 ? (unsigned)atol(argv[1]) : (1u << 27);
  CK(cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, 0));
  CK(cudaMalloc(&g_in, (size_t)g_n * 16));
  g_cap = (size_t)g_n + (size_t)g_n / 4 + (1u << 22);
  CK(cudaMalloc(&g_out, g_cap * 16));
  CK(cudaMalloc(&g_cursor, 4096 * 4));
  init_kernel<<<g_sms * 8, 256>>>(g_in, g_n);
  CK(cudaDeviceSynchronize());
  const int bins[] = {16, 128, 512};
  for (int nb : bins) {
    run<4096, 512, 512, false, 2>("S4096 T512 2/SM", nb, 2);
    run<4096, 512, 512, true, 1>("S4096 T512 1/SM prefetch", nb, 1);
    run<4096, 512, 512, true, 2>("S4096 T512 2/SM prefetch(64reg)", nb, 2);
    run<8192, 1024, 1024, false, 1>("S8192 T1024 1/SM", nb, 1);
    run<2048, 256, 512, false, 4>("S2048 T256 4/SM", nb, 4);
    run<2048, 256, 512, true, 4>("S2048 T256 4/SM prefetch", nb, 4);
    run<4096, 256, 512, false, 2>("S4096 T256 2/SM", nb, 2);
    run<6144, 768, 768, false, 1>("S6144 T768 1/SM", nb, 1);
    run_match<4096, 512, 512, true, false>("match S4096 T512 prefetch", nb, 1);
    run_match<4096, 512, 512, true, true>("tag S4096 T512 prefetch", nb, 1);
    run_match<4096, 512, 512, false, true>("tag S4096 T512", nb, 1);
    run_match<8192, 512, 512, true, true>("tag S8192 T512 prefetch", nb, 1);
  }
  printf("done\n");
  return 0;
}
