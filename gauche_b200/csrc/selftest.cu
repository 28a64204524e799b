// selftest.cu — device-side self checks exported through the C ABI (used by tests/ only).
#include "common.cuh"

namespace gk {
__global__ void fdiv_selftest_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                     float* __restrict__ fast, float* __restrict__ ieee, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    fast[i] = fdiv_rn_normal(a[i], b[i]);
    ieee[i] = __fdiv_rn(a[i], b[i]);
  }
}
__global__ void ddiv_selftest_kernel(const double* __restrict__ a, const double* __restrict__ b,
                                     double* __restrict__ fast, double* __restrict__ ieee, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    fast[i] = fdiv_rn_normal(a[i], b[i]);
    ieee[i] = __ddiv_rn(a[i], b[i]);
  }
}
}  // namespace gk

extern "C" int gk_selftest_fdiv_f64(const double* a, const double* b, double* out_fast, double* out_ieee,
                                    int64_t n, void* stream) {
  using namespace gk;
  GK_CHECK_ARG(n >= 0, GK_EINVAL, "gk_selftest_fdiv_f64: negative size");
  if (n == 0) return 0;
  GK_CHECK_ARG(a && b && out_fast && out_ieee, GK_EINVAL, "gk_selftest_fdiv_f64: null pointer");
  int64_t blocks = (n + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  ddiv_selftest_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(a, b, out_fast, out_ieee, n);
  return cuda_err(cudaGetLastError(), "ddiv_selftest_kernel launch");
}

extern "C" int gk_selftest_fdiv_f32(const float* a, const float* b, float* out_fast, float* out_ieee,
                                    int64_t n, void* stream) {
  using namespace gk;
  GK_CHECK_ARG(n >= 0, GK_EINVAL, "gk_selftest_fdiv_f32: negative size");
  if (n == 0) return 0;
  GK_CHECK_ARG(a && b && out_fast && out_ieee, GK_EINVAL, "gk_selftest_fdiv_f32: null pointer");
  int64_t blocks = (n + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  fdiv_selftest_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(a, b, out_fast, out_ieee, n);
  return cuda_err(cudaGetLastError(), "fdiv_selftest_kernel launch");
}
