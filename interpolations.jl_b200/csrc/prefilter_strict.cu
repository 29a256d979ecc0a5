// Strict prefilter: A_ldiv_B_md! exactly as the reference runs it (filter1d.jl:21-85), operation by operation.
//
// The production kernels (prefilter.cu) stream every line once: they use fused multiply-adds, restart the backward
// recurrence one block further down (an algorithmic truncation, |error| <= 0.268^16) and collapse the Woodbury
// correction into `x -= z1*c1 + zn*cn`.  This file keeps the reference's operation order instead — forward sweep with
// the zero offset added (`+ b[i]`, filter1d.jl:78), scaling by 1/d, backward sweep, then the explicit six-step Woodbury
// correction with a FULL-SIZE temporary and a second full sweep (filter1d.jl:77-85) — with no contraction of a*b+c
// (compiled -fmad=false; Julia never fuses).  One thread per line, three passes over the array plus the temporary:
// slow by design.  Given the same factors (`b200itp_axis_system`, e.g. computed by the reference's own
// `prefiltering_system` on the host), the result equals the reference's CPU arithmetic bit for bit, which separates the two error
// sources of the fast path: strict == reference (rounding identical), |fast - strict| = truncation + re-association.
//
// Switched on with b200itp_prefilter_set_strict(1); verification only.
#include "common.cuh"

#include <atomic>
#include <vector>

namespace b200itp {

std::atomic<int> g_prefilter_strict{0};

struct StrictWoodbury {
  int k;
  int urow[B200ITP_MAX_WOODBURY];  // 1-based, as in the reference
  int vcol[B200ITP_MAX_WOODBURY];
};

// x: the array (in place); tmp: full-size temporary (only touched when k > 0); fac = [dl (n-1, padded to n) | dinv (n) |
// du (n-1, padded to n) | Cp (k*k, column-major)]
template <typename T>
__global__ void __launch_bounds__(128) prefilter_strict_kernel(T *__restrict__ x, T *__restrict__ tmp, long long n,
                                                               long long R1, long long lines,
                                                               const T *__restrict__ fac, const StrictWoodbury wb) {
  const long long line = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (line >= lines) return;
  const T *dl = fac, *dinv = fac + n, *du = fac + 2 * n, *Cp = fac + 3 * n;
  const long long I1 = line % R1, I2 = line / R1;
  T *p = x + I2 * R1 * n + I1;
  const T zero = T(0);
  // ---- step 1: dest = A \ (src + b), b == 0  (filter1d.jl:21-73: forward, scale, backward)
  T prev = p[0] + zero;
  p[0] = prev;
  for (long long i = 1; i < n; ++i) {
    const T t = dl[i - 1] * prev;
    prev = (p[i * R1] + zero) - t;
    p[i * R1] = prev;
  }
  prev = prev * dinv[n - 1];
  p[(n - 1) * R1] = prev;
  for (long long i = n - 2; i >= 0; --i) {
    const T t = du[i] * prev;
    prev = (p[i * R1] - t) * dinv[i];
    p[i * R1] = prev;
  }
  if (wb.k == 0) return;
  // ---- steps 2, 3: tmp1 = V*dest, tmp2 = Cp*tmp1  (filter1d.jl:79-80)
  T tmp1[B200ITP_MAX_WOODBURY], tmp2[B200ITP_MAX_WOODBURY];
  for (int j = 0; j < wb.k; ++j) {
    const T t = T(1) * p[(long long)(wb.vcol[j] - 1) * R1];
    tmp1[j] = zero + t;
  }
  for (int i = 0; i < wb.k; ++i) tmp2[i] = zero;
  for (int j = 0; j < wb.k; ++j)
    for (int i = 0; i < wb.k; ++i) {
      const T t = Cp[i + j * wb.k] * tmp1[j];
      tmp2[i] = tmp2[i] + t;
    }
  // ---- step 4: tmp3 = U*tmp2 (full length, zero except the rows of U)  (filter1d.jl:81)
  T *t3 = tmp + I2 * R1 * n + I1;
  for (long long i = 0; i < n; ++i) t3[i * R1] = zero;
  for (int j = 0; j < wb.k; ++j) {
    const long long r = (long long)(wb.urow[j] - 1) * R1;
    const T t = T(1) * tmp2[j];
    t3[r] = t3[r] + t;
  }
  // ---- step 5: tmp3 = A \ tmp3, no offset  (filter1d.jl:83)
  prev = t3[0];
  for (long long i = 1; i < n; ++i) {
    const T t = dl[i - 1] * prev;
    prev = t3[i * R1] - t;
    t3[i * R1] = prev;
  }
  prev = prev * dinv[n - 1];
  t3[(n - 1) * R1] = prev;
  for (long long i = n - 2; i >= 0; --i) {
    const T t = du[i] * prev;
    prev = (t3[i * R1] - t) * dinv[i];
    t3[i * R1] = prev;
  }
  // ---- step 6: dest -= tmp3  (filter1d.jl:84)
  for (long long i = 0; i < n; ++i) p[i * R1] = p[i * R1] - t3[i * R1];
}

template <typename T>
int prefilter_strict_axis(int dev, T *coefs, int ndims, const int64_t *size, int axis, const T *dl, const T *d,
                          const T *du, int k, const int32_t *urow, const int32_t *vcol, const T *Cp, cudaStream_t st) {
  const int64_t n = size[axis];
  int64_t R1 = 1, R2 = 1, total = 1;
  for (int i = 0; i < axis; ++i) R1 *= size[i];
  for (int i = axis + 1; i < ndims; ++i) R2 *= size[i];
  for (int i = 0; i < ndims; ++i) total *= size[i];
  if (total == 0) return 0;
  std::vector<T> h((size_t)3 * n + (size_t)k * k, T(0));
  for (int64_t i = 0; i + 1 < n; ++i) {
    h[i] = dl[i];
    h[2 * n + i] = du[i];
  }
  for (int64_t i = 0; i < n; ++i) h[n + i] = T(1) / d[i];  // dinv (filter1d.jl:31,51)
  for (int i = 0; i < k * k; ++i) h[3 * n + i] = Cp[i];
  void *fac = nullptr, *tmp = nullptr;
  int rc = scratch_get(dev, 1, st, sizeof(T) * h.size(), &fac);
  if (rc) return rc;
  if (k > 0) {
    rc = scratch_get(dev, 0, st, sizeof(T) * (size_t)total, &tmp);
    if (rc) return rc;
  }
  B200_CUDA_OK(cudaMemcpyAsync(fac, h.data(), sizeof(T) * h.size(), cudaMemcpyHostToDevice, st));
  B200_CUDA_OK(cudaStreamSynchronize(st));  // h is a function-local staging buffer
  StrictWoodbury wb;
  wb.k = k;
  for (int j = 0; j < B200ITP_MAX_WOODBURY; ++j) {
    wb.urow[j] = j < k ? urow[j] : 1;
    wb.vcol[j] = j < k ? vcol[j] : 1;
  }
  const int64_t lines = R1 * R2;
  StageScope sc("prefilter_strict", st);
  prefilter_strict_kernel<T><<<(unsigned)((lines + 127) / 128), 128, 0, st>>>(coefs, (T *)tmp, n, R1, lines, (const T *)fac, wb);
  count_launch();
  B200_CUDA_OK(cudaGetLastError());
  return 0;
}

template int prefilter_strict_axis<float>(int, float *, int, const int64_t *, int, const float *, const float *,
                                          const float *, int, const int32_t *, const int32_t *, const float *, cudaStream_t);
template int prefilter_strict_axis<double>(int, double *, int, const int64_t *, int, const double *, const double *,
                                           const double *, int, const int32_t *, const int32_t *, const double *, cudaStream_t);

}  // namespace b200itp

extern "C" int b200itp_prefilter_set_strict(int enable) { return b200itp::g_prefilter_strict.exchange(enable ? 1 : 0); }
