// Shared helpers of libb200itp (sm_100a only; no CPU fallback anywhere in this library).
#pragma once

#include <cuda_runtime.h>

#include <nvtx3/nvToolsExt.h>
#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>

#include "../../include/b200itp.h"

namespace b200itp {

// ---- error reporting ------------------------------------------------------------------------
extern thread_local char g_err[1024];
extern std::atomic<uint64_t> g_launches;

inline void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
}

#define B200_CUDA_OK(expr)                                                                         \
  do {                                                                                             \
    cudaError_t e__ = (expr);                                                                      \
    if (e__ != cudaSuccess) {                                                                      \
      ::b200itp::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
      return -(int)e__;                                                                            \
    }                                                                                              \
  } while (0)

#define B200_FAIL(code, ...)            \
  do {                                  \
    ::b200itp::set_error(__VA_ARGS__);  \
    return (code);                      \
  } while (0)

inline void count_launch(uint64_t n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

struct DeviceGuard {
  int prev = -1;
  bool ok = true;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) ok = false;
    if (ok && prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

int sm_count(int dev);

// ---- optional per-stage timing (b200itp_profile_*): CUDA events around each launch on the caller's stream
bool stage_timing_enabled();
void stage_begin(const char *name, cudaStream_t st);
void stage_end(cudaStream_t st);
// Every stage is also an NVTX range on the launching host thread (header-only NVTX 3: no-ops unless a profiler
// injected itself), so an nsys / ncu timeline shows the same names as b200itp_profile_get.
struct StageScope {
  cudaStream_t st;
  bool on;
  StageScope(const char *name, cudaStream_t s) : st(s), on(stage_timing_enabled()) {
    nvtxRangePushA(name);
    if (on) stage_begin(name, st);
  }
  ~StageScope() {
    if (on) stage_end(st);
    nvtxRangePop();
  }
};

// "coefficients ready" event of the next evaluation call (b200itp_eval_set_coefs_event): taken (and cleared) by the
// evaluation entry points, waited for on the caller's stream right before the first kernel that READS coefficients
cudaEvent_t take_coefs_event();
// per-device grow-only scratch buffer (guarded by a mutex; see api.cu)
int scratch_get(int dev, int slot, cudaStream_t stream, size_t bytes, void **out);

// ---- cache-hinted 128-bit global accesses -----------------------------------------------------
// Streaming data that is touched exactly once per pass: do not allocate in L1.
__device__ __forceinline__ float4 ldg_stream(const float4 *p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void stg_stream(float4 *p, const float4 &v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
}

}  // namespace b200itp
