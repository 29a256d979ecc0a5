// How much DRAM traffic does a scattered 16-byte gather cost on B200, per load flavour?  (C5: 8 rows of 2 doubles per query
// measured 1070 B of DRAM reads per query = a full 128-byte line per row.)  Run under ncu for dram__bytes_read.sum.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)
__device__ __forceinline__ unsigned hash32(unsigned x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }
template <int MODE>
__device__ __forceinline__ float4 ld16(const float4 *p) {
  float4 r;
  if (MODE == 0) r = __ldg(p);
  if (MODE == 1) asm volatile("ld.global.ca.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  if (MODE == 2) asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  if (MODE == 3) asm volatile("ld.global.cs.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  if (MODE == 4) asm volatile("ld.global.lu.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  if (MODE == 5) asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  if (MODE == 6) asm volatile("ld.global.nc.L2::64B.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  if (MODE == 7) asm volatile("ld.volatile.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
template <int MODE>
__global__ void __launch_bounds__(256) gather(const float4 *__restrict__ a, unsigned mask, unsigned n, float *out) {
  float acc = 0.f;
  const unsigned stride = gridDim.x * blockDim.x;
  for (unsigned q = blockIdx.x * blockDim.x + threadIdx.x; q < n; q += 4 * stride) {
    float4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = ld16<MODE>(a + (hash32(q + u * stride) & mask));
#pragma unroll
    for (int u = 0; u < 4; ++u) acc += v[u].x + v[u].w;
  }
  if (acc == 12345.678f) out[0] = acc;
}
int main() {
  const size_t bytes = 2ull << 30;
  float4 *a; float *out;
  CK(cudaMalloc(&a, bytes)); CK(cudaMalloc(&out, 64)); CK(cudaMemset(a, 0, bytes));
  const unsigned mask = (unsigned)(bytes / 16 - 1), n = 1u << 28;
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const char *names[8] = {"__ldg (ld.global.nc)", "ld.global.ca", "ld.global.cg", "ld.global.cs", "ld.global.lu", "ld.global.nc.L1::no_allocate", "ld.global.nc.L2::64B", "ld.volatile.global"};
  for (int mode = 0; mode < 8; ++mode) {
    float best = 1e9f;
    for (int rep = 0; rep < 2; ++rep) {
      CK(cudaEventRecord(e0));
      switch (mode) {
        case 0: gather<0><<<148 * 8, 256>>>(a, mask, n, out); break;
        case 1: gather<1><<<148 * 8, 256>>>(a, mask, n, out); break;
        case 2: gather<2><<<148 * 8, 256>>>(a, mask, n, out); break;
        case 3: gather<3><<<148 * 8, 256>>>(a, mask, n, out); break;
        case 4: gather<4><<<148 * 8, 256>>>(a, mask, n, out); break;
        case 5: gather<5><<<148 * 8, 256>>>(a, mask, n, out); break;
        case 6: gather<6><<<148 * 8, 256>>>(a, mask, n, out); break;
        case 7: gather<7><<<148 * 8, 256>>>(a, mask, n, out); break;
      }
      CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); best = ms < best ? ms : best;
    }
    printf("%-32s %8.3f ms  %6.2f G gathers/s  (%.0f GB/s of 16-byte payload)\n", names[mode], best, n / best / 1e6, n * 16.0 / best / 1e6);
  }
  return 0;
}
