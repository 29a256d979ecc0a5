// Library-level entry points of the C ABI (include/b200itp.h): version, errors, launch counter,
// device-memory helpers for hosts without their own CUDA binding, and the per-device scratch cache.
#include "common.cuh"

#include <map>
#include <mutex>
#include <string>
#include <tuple>
#include <vector>

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
std::map<std::tuple<int, int, cudaStream_t>, Scratch> g_scratch;
}  // namespace

// Grow-only scratch, one buffer per (device, slot, stream): all work of a call is enqueued on the caller's stream, so
// calls on different streams (or host threads driving different streams) never share a flag or a ping-pong array, and
// calls on one stream are ordered by the stream.  Growth synchronises the device before the old buffer is freed.
// (Two host threads issuing calls on the SAME stream concurrently are not supported; include/b200itp.h says so.)
int scratch_get(int dev, int slot, cudaStream_t stream, size_t bytes, void **out) {
  std::lock_guard<std::mutex> lk(g_scratch_mu);
  Scratch &s = g_scratch[std::make_tuple(dev, slot, stream)];
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

// ---- coefficient-ready event: lets a host replicate / produce the coefficient array on another stream (NCCL
//      broadcast, prefilter) while the ordered pipeline already counts and splits the queries, which needs no
//      coefficients; only the kernels that gather coefficients wait.
// (thread-local: the event belongs to the NEXT evaluation call of the thread that set it, so concurrent host threads
//  driving different streams cannot take each other's event)
static thread_local cudaEvent_t g_coefs_event = nullptr;
cudaEvent_t take_coefs_event() {
  cudaEvent_t e = g_coefs_event;
  g_coefs_event = nullptr;
  return e;
}

// ---- per-stage timing
namespace {
struct StageRec {
  std::string name;
  cudaEvent_t a = nullptr, b = nullptr;
};
std::mutex g_prof_mu;
std::atomic<bool> g_prof_on{false};
std::vector<StageRec> g_prof_recs;
std::vector<cudaEvent_t> g_prof_pool;
cudaEvent_t prof_event() {
  if (!g_prof_pool.empty()) {
    cudaEvent_t e = g_prof_pool.back();
    g_prof_pool.pop_back();
    return e;
  }
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  return e;
}
}  // namespace

bool stage_timing_enabled() { return g_prof_on.load(std::memory_order_relaxed); }
void stage_begin(const char *name, cudaStream_t st) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  StageRec r;
  r.name = name;
  r.a = prof_event();
  r.b = prof_event();
  if (r.a) cudaEventRecord(r.a, st);
  g_prof_recs.push_back(r);
}
void stage_end(cudaStream_t st) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  if (!g_prof_recs.empty() && g_prof_recs.back().b) cudaEventRecord(g_prof_recs.back().b, st);
}

}  // namespace b200itp

using namespace b200itp;

extern "C" int b200itp_profile_enable(int on) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  if (on) {  // a new recording starts: recycle the events of the previous one
    for (auto &r : g_prof_recs) {
      if (r.a) g_prof_pool.push_back(r.a);
      if (r.b) g_prof_pool.push_back(r.b);
    }
    g_prof_recs.clear();
  }
  g_prof_on.store(on != 0);
  return 0;
}

extern "C" int b200itp_profile_count(void) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  return (int)g_prof_recs.size();
}

extern "C" int b200itp_profile_get(int i, char *name, size_t namelen, double *ms) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  if (i < 0 || i >= (int)g_prof_recs.size() || !ms) B200_FAIL(B200ITP_E_BADARG, "profile_get: bad index");
  const StageRec &r = g_prof_recs[i];
  B200_CUDA_OK(cudaEventSynchronize(r.b));
  float t = 0.f;
  B200_CUDA_OK(cudaEventElapsedTime(&t, r.a, r.b));
  *ms = t;
  if (name && namelen) {
    const size_t c = r.name.size() < namelen - 1 ? r.name.size() : namelen - 1;
    memcpy(name, r.name.data(), c);
    name[c] = 0;
  }
  return 0;
}

extern "C" int b200itp_set_l2_fetch_granularity(int dev, int bytes) {
  // cudaLimitMaxL2FetchGranularity (32 / 64 / 128): how much the L2 fetches from HBM around a missing sector.  The
  // scattered-gather kernels want 32 (a 16-byte tap row otherwise drags in a 128-byte line: measured 1070 B of DRAM
  // reads per 4-D linear query), the streaming kernels do not care.  A context-wide hint; returns the previous value.
  DeviceGuard g(dev);
  if (!g.ok) B200_FAIL(-(int)cudaErrorInvalidDevice, "set_l2_fetch_granularity: cannot select device %d", dev);
  size_t prev = 0;
  B200_CUDA_OK(cudaDeviceGetLimit(&prev, cudaLimitMaxL2FetchGranularity));
  if (bytes > 0) B200_CUDA_OK(cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)bytes));
  return (int)prev;
}

extern "C" int b200itp_eval_set_coefs_event(void *cuda_event) {
  g_coefs_event = (cudaEvent_t)cuda_event;
  return 0;
}

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
