// Library-level entry points of the C ABI (include/b200itp.h): version, errors, launch counter,
// device-memory helpers for hosts without their own CUDA binding, and the per-device scratch cache.
#include "common.cuh"

#include <map>
#include <mutex>

namespace b200itp {

thread_local char g_err[1024] = {0};
std::atomic<uint64_t> g_launches{0};

int sm_count(int dev) {
  static std::mutex mu;
  static std::map<int, int> cache;
  std::lock_guard<std::mutex> lk(mu);
  auto it = cache.find(dev);
  if (it != cache.end()) return it->second;
  int n = 148;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) n = 148;
  cache[dev] = n;
  return n;
}

namespace {
struct Scratch {
  void *p = nullptr;
  size_t bytes = 0;
};
std::mutex g_scratch_mu;
std::map<std::pair<int, int>, Scratch> g_scratch;
}  // namespace

// Grow-only scratch, one buffer per (device, slot).  Callers use it on one stream at a time
// (the ABI enqueues everything on the caller's stream; growth synchronises the device).
int scratch_get(int dev, int slot, size_t bytes, void **out) {
  std::lock_guard<std::mutex> lk(g_scratch_mu);
  Scratch &s = g_scratch[{dev, slot}];
  if (s.bytes < bytes) {
    if (s.p) {
      B200_CUDA_OK(cudaDeviceSynchronize());
      B200_CUDA_OK(cudaFree(s.p));
      s.p = nullptr;
      s.bytes = 0;
    }
    size_t want = bytes + bytes / 8;
    B200_CUDA_OK(cudaMalloc(&s.p, want));
    s.bytes = want;
  }
  *out = s.p;
  return 0;
}

}  // namespace b200itp

using namespace b200itp;

extern "C" int b200itp_version(void) { return B200ITP_VERSION; }

extern "C" size_t b200itp_last_error(char *buf, size_t buflen) {
  const size_t len = strlen(g_err);
  if (buf && buflen) {
    const size_t c = len < buflen - 1 ? len : buflen - 1;
    memcpy(buf, g_err, c);
    buf[c] = 0;
  }
  return len;
}

extern "C" uint64_t b200itp_launch_count(void) { return g_launches.load(); }

extern "C" int b200itp_malloc(int dev, size_t bytes, void **out) {
  if (!out) B200_FAIL(B200ITP_E_BADARG, "malloc: null out pointer");
  DeviceGuard g(dev);
  if (!g.ok) B200_FAIL(-(int)cudaErrorInvalidDevice, "malloc: cannot select device %d", dev);
  B200_CUDA_OK(cudaMalloc(out, bytes));
  return 0;
}

extern "C" int b200itp_free(int dev, void *p) {
  DeviceGuard g(dev);
  if (!g.ok) B200_FAIL(-(int)cudaErrorInvalidDevice, "free: cannot select device %d", dev);
  B200_CUDA_OK(cudaFree(p));
  return 0;
}

extern "C" int b200itp_memcpy_h2d(int dev, void *dst, const void *src, size_t bytes, void *stream) {
  DeviceGuard g(dev);
  if (!g.ok) B200_FAIL(-(int)cudaErrorInvalidDevice, "memcpy_h2d: cannot select device %d", dev);
  B200_CUDA_OK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream));
  return 0;
}

extern "C" int b200itp_memcpy_d2h(int dev, void *dst, const void *src, size_t bytes, void *stream) {
  DeviceGuard g(dev);
  if (!g.ok) B200_FAIL(-(int)cudaErrorInvalidDevice, "memcpy_d2h: cannot select device %d", dev);
  B200_CUDA_OK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  return 0;
}

extern "C" int b200itp_sync(int dev, void *stream) {
  DeviceGuard g(dev);
  if (!g.ok) B200_FAIL(-(int)cudaErrorInvalidDevice, "sync: cannot select device %d", dev);
  B200_CUDA_OK(cudaStreamSynchronize((cudaStream_t)stream));
  return 0;
}
