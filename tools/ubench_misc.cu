// Micro-benchmarks behind the round-2 pipeline decisions (run on a B200: nvcc -arch=sm_100a -O3 -fmad=false).
//   fp      issue rate of unfused FMUL / FADD / FFMA chains (cycles per warp instruction per SMSP)
//   mix     the evaluation loop's mix: 64 LDS + 108 FMUL + 90 FADD per record
//   window  permutation of {q, value} records into out[q] when q is a bijection inside windows of 2^wb results
//           (the last level of the unsort): scattered 4-byte stores into an L2-resident window
//   atoms   shared-memory atomic ranking vs ballot ranking of 512 keys
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

struct Timer {
  cudaEvent_t a, b;
  Timer() { CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b)); }
  void start() { CK(cudaEventRecord(a)); }
  float stop() { CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b)); float ms; CK(cudaEventElapsedTime(&ms, a, b)); return ms; }
};

// ---------------------------------------------------------------- fp
template <int MODE>
__global__ void fp_kernel(float *out, int iters, float a, float b) {
  float x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = threadIdx.x * 0.001f + i;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) x[i] = fmaf(x[i], a, b);
      if (MODE == 1) x[i] = x[i] * a;
      if (MODE == 2) x[i] = x[i] + b;
      if (MODE == 3) { x[i] = x[i] * a; x[i] = x[i] + b; }
      if (MODE == 4) { x[i] = x[i] * x[(i + 1) & 7]; }           // two distinct register sources
      if (MODE == 5) { x[i] = x[i] * x[(i + 1) & 7]; x[i] = x[i] + x[(i + 3) & 7]; }
    }
  }
  const long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) reinterpret_cast<long long *>(out)[1 << 20] = t1 - t0;
}

// ---------------------------------------------------------------- mix
__global__ void mix_kernel(float *out, int iters, const float *w) {
  extern __shared__ float tile[];
  for (int i = threadIdx.x; i < 24000; i += blockDim.x) tile[i] = i * 0.5f;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  float acc = 0.f;
  unsigned off = lane + (threadIdx.x >> 5) * 97;
  const float w0 = w[0], w1 = w[1], w2 = w[2], w3 = w[3];
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    const float *tb = tile + (off % 20000u);
    float val = 0.f;
#pragma unroll
    for (int k0 = 0; k0 < 4; ++k0) {
      float v0 = 0.f;
#pragma unroll
      for (int k1 = 0; k1 < 4; ++k1) {
        float v1 = 0.f;
#pragma unroll
        for (int k2 = 0; k2 < 4; ++k2) {
          const float wk = k2 == 0 ? w0 : (k2 == 1 ? w1 : (k2 == 2 ? w2 : w3));
          const float t2 = wk * tb[(k2 * 35 + k1) * 35 + k0];
          v1 = k2 == 0 ? t2 : t2 + v1;
        }
        const float wk = k1 == 0 ? w1 : (k1 == 1 ? w2 : (k1 == 2 ? w3 : w0));
        const float t1 = wk * v1;
        v0 = k1 == 0 ? t1 : t1 + v0;
      }
      const float wk = k0 == 0 ? w2 : (k0 == 1 ? w3 : (k0 == 2 ? w0 : w1));
      const float t0_ = wk * v0;
      val = k0 == 0 ? t0_ : t0_ + val;
    }
    acc += val;
    off += 1237u + (__float_as_uint(val) & 32u);
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) reinterpret_cast<long long *>(out)[1 << 20] = t1 - t0;
}

// ---------------------------------------------------------------- window scatter
__global__ void make_records(uint2 *rec, unsigned n, unsigned wb) {
  const unsigned mask = (1u << wb) - 1u;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const unsigned win = i >> wb;
    const unsigned local = ((i & mask) * 2654435761u + 12345u * win) & mask;  // odd multiplier: a bijection of the window
    unsigned q = (win << wb) + local;
    if (q >= n) q = i;  // (last partial window: identity)
    rec[i] = make_uint2(q, i);
  }
}
template <int U>
__global__ void __launch_bounds__(256) window_scatter(const uint2 *__restrict__ rec, unsigned n, float *__restrict__ out) {
  const unsigned stride = gridDim.x * blockDim.x;
  unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + (U - 1) * stride < n; i += U * stride) {
    uint2 r[U];
#pragma unroll
    for (int u = 0; u < U; ++u) r[u] = __ldcs(rec + i + u * stride);
#pragma unroll
    for (int u = 0; u < U; ++u) out[r[u].x] = __uint_as_float(r[u].y);
  }
  for (; i < n; i += stride) {
    const uint2 r = rec[i];
    out[r.x] = __uint_as_float(r.y);
  }
}
// one CTA walks one window at a time (all CTAs inside a span of `span` consecutive windows)
__global__ void __launch_bounds__(512) window_scatter_cta(const uint2 *__restrict__ rec, unsigned n, unsigned wb,
                                                          float *__restrict__ out) {
  const unsigned nwin = (n + (1u << wb) - 1) >> wb;
  for (unsigned w = blockIdx.x; w < nwin; w += gridDim.x) {
    const unsigned b = w << wb, e = min(n, b + (1u << wb));
    for (unsigned i = b + threadIdx.x; i < e; i += 4 * blockDim.x) {
      uint2 r[4];
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (i + u * blockDim.x < e) r[u] = __ldcs(rec + i + u * blockDim.x);
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (i + u * blockDim.x < e) out[r[u].x] = __uint_as_float(r[u].y);
    }
  }
}

// ---------------------------------------------------------------- atoms vs ballot ranking
template <int MODE>
__global__ void __launch_bounds__(512) rank_kernel(unsigned *out, int iters) {
  __shared__ unsigned hist[512];
  __shared__ unsigned short whist[16][512];
  for (int i = threadIdx.x; i < 512; i += blockDim.x) hist[i] = 0;
  for (int i = threadIdx.x; i < 16 * 512; i += blockDim.x) (&whist[0][0])[i] = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned h = threadIdx.x * 2654435761u + blockIdx.x;
  unsigned acc = 0;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    h = h * 1664525u + 1013904223u;
    const unsigned k = h >> 23;  // 9 bits
    if (MODE == 0) {
      acc += atomicAdd(&hist[k], 1u);
    } else {
      unsigned peers = 0xffffffffu;
#pragma unroll
      for (int b = 0; b < 9; ++b) {
        const unsigned bit = (k >> b) & 1u;
        const unsigned m = __ballot_sync(0xffffffffu, bit);
        peers &= m ^ (bit - 1u);
      }
      const unsigned below = __popc(peers & ((1u << lane) - 1u));
      unsigned base = 0;
      if (below == 0) {
        base = whist[warp][k];
        whist[warp][k] = (unsigned short)(base + __popc(peers));
      }
      __syncwarp();
      base = __shfl_sync(0xffffffffu, base, __ffs(peers) - 1);
      acc += base + below;
    }
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) reinterpret_cast<long long *>(out)[1 << 20] = t1 - t0;
}

int main(int argc, char **argv) {
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  Timer t;
  float *out;
  CK(cudaMalloc(&out, (size_t)(1 << 23) + 64));
  long long cyc;
  const char *names[6] = {"FFMA r,c,c", "FMUL r,c", "FADD r,c", "FMUL+FADD r,c", "FMUL r,r", "FMUL+FADD r,r"};
  const int per[6] = {8, 8, 8, 16, 8, 16};
  for (int threads : {128, 256, 512, 1024}) {
    for (int mode = 0; mode < 6; ++mode) {
      const int iters = 4096;
      auto launch = [&]() {
        switch (mode) {
          case 0: fp_kernel<0><<<sms, threads>>>(out, iters, 1.0001f, 0.5f); break;
          case 1: fp_kernel<1><<<sms, threads>>>(out, iters, 1.0001f, 0.5f); break;
          case 2: fp_kernel<2><<<sms, threads>>>(out, iters, 1.0001f, 0.5f); break;
          case 3: fp_kernel<3><<<sms, threads>>>(out, iters, 1.0001f, 0.5f); break;
          case 4: fp_kernel<4><<<sms, threads>>>(out, iters, 1.0001f, 0.5f); break;
          case 5: fp_kernel<5><<<sms, threads>>>(out, iters, 1.0001f, 0.5f); break;
        }
      };
      launch();
      launch();
      CK(cudaDeviceSynchronize());
      CK(cudaMemcpy(&cyc, reinterpret_cast<long long *>(out) + (1 << 20), 8, cudaMemcpyDeviceToHost));
      const double winstr = (double)iters * per[mode] * (threads / 32) / 4.0;  // warp instructions per SMSP
      printf("fp %-16s threads=%4d  cycles/warp-instr/SMSP = %.3f\n", names[mode], threads, cyc / winstr);
    }
  }
  {
    float *w;
    CK(cudaMalloc(&w, 16));
    float hw[4] = {0.1f, 0.4f, 0.4f, 0.1f};
    CK(cudaMemcpy(w, hw, 16, cudaMemcpyHostToDevice));
    CK(cudaFuncSetAttribute(mix_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    for (int threads : {256, 384, 512, 768, 1024}) {
      const int iters = 2000;
      mix_kernel<<<sms, threads, 96 * 1024>>>(out, iters, w);
      mix_kernel<<<sms, threads, 96 * 1024>>>(out, iters, w);
      CK(cudaDeviceSynchronize());
      CK(cudaMemcpy(&cyc, reinterpret_cast<long long *>(out) + (1 << 20), 8, cudaMemcpyDeviceToHost));
      printf("mix (64 LDS + 84 FMUL + 63 FADD per record) threads=%4d  cycles/record/SM = %.3f  (1e9 records on %d SMs at 1.9 GHz: %.2f ms)\n",
             threads, (double)cyc / ((double)iters * threads), sms, (double)cyc / ((double)iters * threads) * 1e9 / sms / 1.9e9 * 1e3);
    }
  }
  for (int mode = 0; mode < 2; ++mode) {
    const int iters = 4096;
    if (mode == 0) rank_kernel<0><<<sms, 512>>>((unsigned *)out, iters); else rank_kernel<1><<<sms, 512>>>((unsigned *)out, iters);
    if (mode == 0) rank_kernel<0><<<sms, 512>>>((unsigned *)out, iters); else rank_kernel<1><<<sms, 512>>>((unsigned *)out, iters);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(&cyc, reinterpret_cast<long long *>(out) + (1 << 20), 8, cudaMemcpyDeviceToHost));
    printf("rank %-7s 512 threads: cycles per warp-row (32 records) per SM = %.2f\n", mode ? "ballot" : "ATOMS", (double)cyc / ((double)iters * 16));
  }
  {
    const unsigned n = argc > 1 ? (unsigned)atol(argv[1]) : (1u << 28);
    uint2 *rec;
    float *o;
    CK(cudaMalloc(&rec, (size_t)n * 8));
    CK(cudaMalloc(&o, (size_t)n * 4));
    for (unsigned wb : {14u, 18u, 20u, 21u, 22u, 23u, 24u, 25u}) {
      make_records<<<sms * 8, 256>>>(rec, n, wb);
      CK(cudaDeviceSynchronize());
      window_scatter<4><<<sms * 8, 256>>>(rec, n, o);
      t.start();
      window_scatter<4><<<sms * 8, 256>>>(rec, n, o);
      const float ms = t.stop();
      window_scatter_cta<<<sms * 2, 512>>>(rec, n, wb, o);
      t.start();
      window_scatter_cta<<<sms * 2, 512>>>(rec, n, wb, o);
      const float ms2 = t.stop();
      printf("window scatter wb=%2u (%6.1f MiB window): grid-stride %7.3f ms %6.1f G/s (%6.1f GB/s alg) | cta-per-window %7.3f ms %6.1f G/s\n", wb,
             (4u << wb) / 1048576.0, ms, n / ms / 1e6, (double)n * 12 / ms / 1e6, ms2, n / ms2 / 1e6);
    }
  }
  printf("done\n");
  return 0;
}
