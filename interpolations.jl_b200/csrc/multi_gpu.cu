// Multi-GPU entry points of the C ABI (include/b200itp.h): one host process (the Julia session) drives
// the B200s of one box; NCCL over NVLink 5 / NVSwitch moves the data, the library's own kernels do every flop.
//
//   b200itp_comm_init / _destroy     ncclCommInitAll over the given devices, one stream per device
//   b200itp_bcast_coefs              replicate a coefficient array (evaluation shards the QUERIES: no other collective)
//   b200itp_prefilter_sharded        large 3-D prefilter: device g owns a z-slab; x and y passes are local (the first
//                                    reads the samples, so they are never copied), ONE all-to-all re-shards the array
//                                    along y (ncclSend/ncclRecv group; the send buffer is the slab in peer-major order,
//                                    written by one strided copy kernel), the z pass is local, and one all-gather plus
//                                    one strided copy leaves the full coefficient array on every device, ready for
//                                    query-sharded evaluation.  Bit-identical to the single-GPU prefilter (line
//                                    segmentation is switched off around the slab solves).
//
// libnccl.so.2 is loaded with dlopen at the first b200itp_comm_init: the library has no link-time NCCL dependency, and
// single-GPU hosts never touch it.
#include "common.cuh"

#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <mutex>
#include <vector>

namespace b200itp {

namespace {
struct Nccl {
  void *handle = nullptr;
  decltype(&ncclCommInitAll) CommInitAll = nullptr;
  decltype(&ncclCommDestroy) CommDestroy = nullptr;
  decltype(&ncclGetErrorString) GetErrorString = nullptr;
  decltype(&ncclBroadcast) Broadcast = nullptr;
  decltype(&ncclAllGather) AllGather = nullptr;
  decltype(&ncclSend) Send = nullptr;
  decltype(&ncclRecv) Recv = nullptr;
  decltype(&ncclGroupStart) GroupStart = nullptr;
  decltype(&ncclGroupEnd) GroupEnd = nullptr;
};
Nccl g_nccl;
std::mutex g_nccl_mu;

int load_nccl() {
  std::lock_guard<std::mutex> lk(g_nccl_mu);
  if (g_nccl.handle) return 0;
  void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) B200_FAIL(B200ITP_E_UNSUPPORTED, "comm_init: cannot load libnccl.so.2 (%s)", dlerror());
#define B200_SYM(field, name)                                                                  \
  g_nccl.field = reinterpret_cast<decltype(g_nccl.field)>(dlsym(h, name));                      \
  if (!g_nccl.field) B200_FAIL(B200ITP_E_UNSUPPORTED, "comm_init: libnccl has no symbol %s", name)
  B200_SYM(CommInitAll, "ncclCommInitAll");
  B200_SYM(CommDestroy, "ncclCommDestroy");
  B200_SYM(GetErrorString, "ncclGetErrorString");
  B200_SYM(Broadcast, "ncclBroadcast");
  B200_SYM(AllGather, "ncclAllGather");
  B200_SYM(Send, "ncclSend");
  B200_SYM(Recv, "ncclRecv");
  B200_SYM(GroupStart, "ncclGroupStart");
  B200_SYM(GroupEnd, "ncclGroupEnd");
#undef B200_SYM
  g_nccl.handle = h;
  return 0;
}

#define B200_NCCL_OK(expr)                                                                       \
  do {                                                                                           \
    ncclResult_t r_ = (expr);                                                                    \
    if (r_ != ncclSuccess) B200_FAIL(-1000 - (int)r_, "NCCL: %s (%s)", g_nccl.GetErrorString(r_), #expr); \
  } while (0)

struct DevBuf {
  void *p = nullptr;
  size_t bytes = 0;
};
}  // namespace

}  // namespace b200itp

using namespace b200itp;

struct b200itp_comm {
  int n = 0;
  std::vector<int> dev;
  std::vector<ncclComm_t> comm;
  std::vector<cudaStream_t> stream;
  std::vector<cudaEvent_t> ev;            // per device: "phase finished"
  std::vector<DevBuf> work, send, recv, gath;  // grow-only scratch per device
  cudaEvent_t t0 = nullptr, t1 = nullptr, t2 = nullptr;
};

namespace b200itp {
namespace {

int grow(int dev, DevBuf &b, size_t bytes) {
  if (b.bytes >= bytes) return 0;
  DeviceGuard g(dev);
  if (b.p) {
    B200_CUDA_OK(cudaDeviceSynchronize());
    B200_CUDA_OK(cudaFree(b.p));
    b.p = nullptr;
    b.bytes = 0;
  }
  B200_CUDA_OK(cudaMalloc(&b.p, bytes));
  b.bytes = bytes;
  return 0;
}

// contiguous, order-preserving split of range(n) into `world` parts whose sizes differ by at most one
inline void split(int64_t n, int world, int r, int64_t &lo, int64_t &hi) {
  const int64_t base = n / world, rem = n % world;
  lo = r * base + std::min<int64_t>(r, rem);
  hi = lo + base + (r < rem ? 1 : 0);
}

// All peers' blocks in ONE launch per device (the host thread drives every device: a launch per peer made the
// 8-GPU prefilter launch-bound).  y-ranges y0[h] .. y0[h+1] tile [0, ny); block h of the peer-major / rank-major buffer
// starts at element off[h] and is the dense box (nx, yl_h, nzb).
constexpr int kMaxPeers = 16;
struct PeerSplit {
  int n;
  long long y0[kMaxPeers + 1];
  long long off[kMaxPeers + 1];
};

// PACK:   slab (nx, ny, nzb) -> peer-major buffer;   UNPACK: rank-major buffer -> array (nx, ny, nzb)
template <typename V, bool PACK>
__global__ void __launch_bounds__(256) slab_pack_kernel(const V *__restrict__ src, V *__restrict__ dst, long long nxv,
                                                        long long ny, long long nzb, const PeerSplit ps) {
  const long long rows = ny * nzb;
  for (long long row = blockIdx.y; row < rows; row += gridDim.y) {
    const long long y = row % ny, z = row / ny;
    int h = 0;
    while (h + 1 < ps.n && y >= ps.y0[h + 1]) ++h;
    const long long yl = ps.y0[h + 1] - ps.y0[h];
    const long long dense = ps.off[h] + nxv * ((y - ps.y0[h]) + yl * z);  // position in the blocked buffer
    const long long full = nxv * (y + ny * z);                             // position in the (x, y, z) array
    const V *s = src + (PACK ? full : dense);
    V *d = dst + (PACK ? dense : full);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nxv; i += (long long)gridDim.x * blockDim.x)
      d[i] = s[i];
  }
}

template <bool PACK>
int slab_pack(int dev, const void *src, void *dst, int es, int64_t nx, int64_t ny, int64_t nzb, const int64_t *y0, int G,
              cudaStream_t st) {
  if (nx == 0 || ny == 0 || nzb == 0) return 0;
  const bool v16 = ((nx * es) % 16 == 0) && ((uintptr_t)src % 16 == 0) && ((uintptr_t)dst % 16 == 0);
  const long long k = v16 ? 16 / es : 1;
  PeerSplit ps;
  ps.n = G;
  long long off = 0;
  for (int h = 0; h <= G; ++h) {
    ps.y0[h] = y0[h];
    ps.off[h] = off;
    if (h < G) off += (nx / k) * (y0[h + 1] - y0[h]) * nzb;
  }
  const long long nxv = nx / k, rows = ny * nzb;
  dim3 grid((unsigned)std::min<long long>((nxv + 255) / 256, 8), (unsigned)std::min<long long>(rows, (long long)sm_count(dev) * 16));
  if (v16)
    slab_pack_kernel<float4, PACK><<<grid, 256, 0, st>>>((const float4 *)src, (float4 *)dst, nxv, ny, nzb, ps);
  else if (es == 4)
    slab_pack_kernel<float, PACK><<<grid, 256, 0, st>>>((const float *)src, (float *)dst, nxv, ny, nzb, ps);
  else
    slab_pack_kernel<double, PACK><<<grid, 256, 0, st>>>((const double *)src, (double *)dst, nxv, ny, nzb, ps);
  count_launch();
  B200_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace
}  // namespace b200itp

extern "C" int b200itp_comm_init(int ngpu, const int *devs, b200itp_comm **out) {
  if (ngpu < 1 || !devs || !out) B200_FAIL(B200ITP_E_BADARG, "comm_init: bad arguments");
  int rc = load_nccl();
  if (rc) return rc;
  int have = 0;
  B200_CUDA_OK(cudaGetDeviceCount(&have));
  for (int i = 0; i < ngpu; ++i)
    if (devs[i] < 0 || devs[i] >= have) B200_FAIL(B200ITP_E_BADARG, "comm_init: device %d does not exist (%d visible)", devs[i], have);
  auto *c = new b200itp_comm();
  c->n = ngpu;
  c->dev.assign(devs, devs + ngpu);
  c->comm.resize(ngpu);
  c->stream.resize(ngpu);
  c->ev.resize(ngpu);
  c->work.resize(ngpu);
  c->send.resize(ngpu);
  c->recv.resize(ngpu);
  c->gath.resize(ngpu);
  ncclResult_t r = g_nccl.CommInitAll(c->comm.data(), ngpu, devs);
  if (r != ncclSuccess) {
    delete c;
    B200_FAIL(-1000 - (int)r, "comm_init: ncclCommInitAll failed: %s", g_nccl.GetErrorString(r));
  }
  for (int i = 0; i < ngpu; ++i) {
    DeviceGuard g(devs[i]);
    B200_CUDA_OK(cudaStreamCreateWithFlags(&c->stream[i], cudaStreamNonBlocking));
    B200_CUDA_OK(cudaEventCreateWithFlags(&c->ev[i], cudaEventDisableTiming));
  }
  {
    DeviceGuard g(devs[0]);
    B200_CUDA_OK(cudaEventCreate(&c->t0));
    B200_CUDA_OK(cudaEventCreate(&c->t1));
    B200_CUDA_OK(cudaEventCreate(&c->t2));
  }
  *out = c;
  return 0;
}

extern "C" int b200itp_comm_size(const b200itp_comm *c) { return c ? c->n : 0; }

extern "C" int b200itp_comm_destroy(b200itp_comm *c) {
  if (!c) return 0;
  for (int i = 0; i < c->n; ++i) {
    DeviceGuard g(c->dev[i]);
    cudaDeviceSynchronize();
    for (DevBuf *b : {&c->work[i], &c->send[i], &c->recv[i], &c->gath[i]})
      if (b->p) cudaFree(b->p);
    if (c->stream[i]) cudaStreamDestroy(c->stream[i]);
    if (c->ev[i]) cudaEventDestroy(c->ev[i]);
    g_nccl.CommDestroy(c->comm[i]);
  }
  if (c->t0) cudaEventDestroy(c->t0);
  if (c->t1) cudaEventDestroy(c->t1);
  if (c->t2) cudaEventDestroy(c->t2);
  delete c;
  return 0;
}

extern "C" int b200itp_bcast_coefs(b200itp_comm *c, int root, void *const *coefs_per_dev, size_t bytes) {
  if (!c || !coefs_per_dev || root < 0 || root >= c->n) B200_FAIL(B200ITP_E_BADARG, "bcast_coefs: bad arguments");
  for (int i = 0; i < c->n; ++i)
    if (!coefs_per_dev[i]) B200_FAIL(B200ITP_E_BADARG, "bcast_coefs: null buffer for device %d", c->dev[i]);
  if (bytes == 0 || c->n == 1) return 0;
  B200_NCCL_OK(g_nccl.GroupStart());
  for (int i = 0; i < c->n; ++i)
    B200_NCCL_OK(g_nccl.Broadcast(coefs_per_dev[root], coefs_per_dev[i], bytes, ncclChar, root, c->comm[i], c->stream[i]));
  B200_NCCL_OK(g_nccl.GroupEnd());
  for (int i = 0; i < c->n; ++i) {
    DeviceGuard g(c->dev[i]);
    B200_CUDA_OK(cudaStreamSynchronize(c->stream[i]));
  }
  return 0;
}

extern "C" int b200itp_prefilter_sharded(b200itp_comm *c, const void *const *samples_slab, void *const *coefs_full, int dtype,
                                         const int64_t *size, const int32_t *degree, const int32_t *bc,
                                         const int32_t *gridtype, double *ms_out) {
  if (!c || !samples_slab || !coefs_full || !size || !degree || !bc || !gridtype)
    B200_FAIL(B200ITP_E_BADARG, "prefilter_sharded: null pointer");
  if (dtype != B200ITP_F32 && dtype != B200ITP_F64) B200_FAIL(B200ITP_E_DTYPE, "prefilter_sharded: dtype must be F32 or F64");
  const int G = c->n;
  const int es = dtype == B200ITP_F32 ? 4 : 8;
  const int64_t nx = size[0], ny = size[1], nz = size[2];
  if (nx < 1 || ny < G || nz < G) B200_FAIL(B200ITP_E_SIZE, "prefilter_sharded: every device needs at least one y row and one z plane");
  for (int a = 0; a < 3; ++a) {
    const bool filtered = degree[a] == B200ITP_CUBIC || degree[a] == B200ITP_QUADRATIC;
    if (filtered && bc[a] != B200ITP_BC_PERIODIC && !B200ITP_IS_INPLACE_BC(bc[a]))
      B200_FAIL(B200ITP_E_UNSUPPORTED,
                "prefilter_sharded: axis %d pads its array (cubic.jl:86, quadratic.jl:64): pass the padded array as samples "
                "(copy_with_padding on the host side) — only unpadded specs are sharded here", a);
  }
  const ncclDataType_t nt = dtype == B200ITP_F32 ? ncclFloat : ncclDouble;
  const int32_t lin = B200ITP_LINEAR;  // "skip this axis" (prefiltering.jl:30)
  const int32_t deg_xy[3] = {degree[0], degree[1], lin}, deg_z[3] = {lin, lin, degree[2]};
  if (G > kMaxPeers) B200_FAIL(B200ITP_E_UNSUPPORTED, "prefilter_sharded: at most %d devices", kMaxPeers);
  std::vector<int64_t> z0(G), z1(G), y0(G), y1(G), ybound(G + 1);
  for (int g = 0; g < G; ++g) {
    split(nz, G, g, z0[g], z1[g]);
    split(ny, G, g, y0[g], y1[g]);
    ybound[g] = y0[g];
    ybound[g + 1] = y1[g];
    if (!samples_slab[g] || !coefs_full[g]) B200_FAIL(B200ITP_E_BADARG, "prefilter_sharded: null buffer for device %d", c->dev[g]);
  }
  const int prev_seg = b200itp_prefilter_set_segmentation(0);  // same bits whatever the slab sizes
  int rc = 0;
  auto fail = [&](int code) {
    b200itp_prefilter_set_segmentation(prev_seg);
    return code;
  };
  // ---- scratch
  for (int g = 0; g < G && !rc; ++g) {
    const size_t slab = (size_t)nx * ny * (z1[g] - z0[g]) * es, ysl = (size_t)nx * (y1[g] - y0[g]) * nz * es;
    rc = grow(c->dev[g], c->work[g], slab);
    if (!rc && G > 1) rc = grow(c->dev[g], c->send[g], slab);
    if (!rc) rc = grow(c->dev[g], c->recv[g], ysl);
    if (!rc && G > 1) rc = grow(c->dev[g], c->gath[g], (size_t)nx * ny * nz * es);
  }
  if (rc) return fail(rc);
  {
    DeviceGuard g0(c->dev[0]);
    if (cudaEventRecord(c->t0, c->stream[0]) != cudaSuccess) return fail(-1);
  }
  // ---- x and y passes on the z-slab (the first pass reads the samples), then the slab in peer-major order
  for (int g = 0; g < G; ++g) {
    DeviceGuard dg(c->dev[g]);
    const int64_t nzl = z1[g] - z0[g];
    const int64_t ssz[3] = {nx, ny, nzl};
    rc = b200itp_prefilter_from(c->dev[g], samples_slab[g], c->work[g].p, dtype, 3, ssz, deg_xy, bc, gridtype, c->stream[g]);
    if (rc) return fail(rc);
    if (G == 1) continue;
    // the slab in peer-major order (block h = the box (nx, yl_h, nzl), dense): one launch
    rc = slab_pack<true>(c->dev[g], c->work[g].p, c->send[g].p, es, nx, ny, nzl, ybound.data(), G, c->stream[g]);
    if (rc) return fail(rc);
  }
  // ---- one all-to-all: z-slabs -> y-slabs.  Device g receives from h the box (nx, yl_g, nzl_h), which is the z-range of
  //      h inside g's y-slab (nx, yl_g, nz): blocks land in rank order == z order, the receive buffer IS the y-slab.
  if (G > 1) {
    if (g_nccl.GroupStart() != ncclSuccess) return fail(-1001);
    for (int g = 0; g < G; ++g) {
      size_t soff = 0;
      const int64_t ylg = y1[g] - y0[g], nzl = z1[g] - z0[g];
      for (int h = 0; h < G; ++h) {
        const int64_t ylh = y1[h] - y0[h], nzh = z1[h] - z0[h];
        ncclResult_t r1 = g_nccl.Send((const char *)c->send[g].p + soff, (size_t)nx * ylh * nzl, nt, h, c->comm[g], c->stream[g]);
        ncclResult_t r2 = g_nccl.Recv((char *)c->recv[g].p + (size_t)nx * ylg * z0[h] * es, (size_t)nx * ylg * nzh, nt, h, c->comm[g],
                                      c->stream[g]);
        if (r1 != ncclSuccess || r2 != ncclSuccess) {
          g_nccl.GroupEnd();
          b200itp_prefilter_set_segmentation(prev_seg);
          B200_FAIL(-1002, "prefilter_sharded: ncclSend/ncclRecv failed");
        }
        soff += (size_t)nx * ylh * nzl * es;
      }
    }
    if (g_nccl.GroupEnd() != ncclSuccess) return fail(-1003);
  } else {
    DeviceGuard dg(c->dev[0]);
    if (cudaMemcpyAsync(c->recv[0].p, c->work[0].p, (size_t)nx * ny * nz * es, cudaMemcpyDeviceToDevice, c->stream[0]) != cudaSuccess)
      return fail(-1);
  }
  // ---- z pass on the y-slab
  for (int g = 0; g < G; ++g) {
    DeviceGuard dg(c->dev[g]);
    const int64_t ysz[3] = {nx, y1[g] - y0[g], nz};
    rc = b200itp_prefilter(c->dev[g], c->recv[g].p, dtype, 3, ysz, deg_z, bc, gridtype, c->stream[g]);
    if (rc) return fail(rc);
  }
  b200itp_prefilter_set_segmentation(prev_seg);
  {
    DeviceGuard g0(c->dev[0]);
    B200_CUDA_OK(cudaEventRecord(c->t1, c->stream[0]));
  }
  // ---- replicate: all-gather of the y-slabs (rank-major), then one strided copy into the (x, y, z) layout
  if (G > 1) {
    std::vector<size_t> goff(G + 1, 0);
    for (int h = 0; h < G; ++h) goff[h + 1] = goff[h] + (size_t)nx * (y1[h] - y0[h]) * nz * es;
    const bool equal = ny % G == 0;
    B200_NCCL_OK(g_nccl.GroupStart());
    for (int g = 0; g < G; ++g) {
      if (equal) {
        B200_NCCL_OK(g_nccl.AllGather(c->recv[g].p, c->gath[g].p, (size_t)nx * (ny / G) * nz, nt, c->comm[g], c->stream[g]));
      } else {  // uneven shards: one broadcast per shard
        for (int h = 0; h < G; ++h)
          B200_NCCL_OK(g_nccl.Broadcast(c->recv[h].p, (char *)c->gath[g].p + goff[h], (size_t)nx * (y1[h] - y0[h]) * nz, nt, h,
                                        c->comm[g], c->stream[g]));
      }
    }
    B200_NCCL_OK(g_nccl.GroupEnd());
    for (int g = 0; g < G; ++g) {  // rank-major y-slabs -> the (x, y, z) array: one launch per device
      DeviceGuard dg(c->dev[g]);
      rc = slab_pack<false>(c->dev[g], c->gath[g].p, coefs_full[g], es, nx, ny, nz, ybound.data(), G, c->stream[g]);
      if (rc) return rc;
    }
  } else {
    DeviceGuard dg(c->dev[0]);
    B200_CUDA_OK(cudaMemcpyAsync(coefs_full[0], c->recv[0].p, (size_t)nx * ny * nz * es, cudaMemcpyDeviceToDevice, c->stream[0]));
  }
  {
    DeviceGuard g0(c->dev[0]);
    B200_CUDA_OK(cudaEventRecord(c->t2, c->stream[0]));
  }
  for (int g = 0; g < G; ++g) {
    DeviceGuard dg(c->dev[g]);
    B200_CUDA_OK(cudaStreamSynchronize(c->stream[g]));
  }
  if (ms_out) {
    float a = 0.f, b = 0.f;
    DeviceGuard g0(c->dev[0]);
    B200_CUDA_OK(cudaEventElapsedTime(&a, c->t0, c->t1));
    B200_CUDA_OK(cudaEventElapsedTime(&b, c->t0, c->t2));
    ms_out[0] = a;  // x/y passes + pack + all-to-all + z pass, as seen by device 0's stream
    ms_out[1] = b;  // ... + all-gather + unpack
  }
  return 0;
}
