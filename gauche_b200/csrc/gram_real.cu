// gram_real.cu — Tanimoto / MinMax Gram on real-valued (non-integer) dense inputs.
//
// Correctness path for inputs that are not bit or small-count fingerprints (warped
// inputs, learnable inducing points, the reference's own 0.1*ones KAT —
// tests/test_kernels/test_fingerprint_kernels/test_tanimoto_kernel.py:18-31).
// Evaluates tanimoto_kernel.py:36-46 / minmax_kernel.py:33-44 directly in the input
// dtype: dot product (or L1 distance) plus both row norms accumulate in one pass over
// K, and the epilogue is fused so the tile is written once.  Summation order differs
// from MKL/cuBLAS, so parity for this path is tolerance-based, not bit-exact.
#include "common.cuh"

namespace gk {

constexpr int RT = 64;   // tile edge
constexpr int RK = 16;   // K chunk

template <typename T, bool kMinMax>
__global__ void __launch_bounds__(256)
gram_real_kernel(const T* __restrict__ X1, int64_t ld1, int64_t n, const T* __restrict__ X2,
                 int64_t ld2, int64_t m, int64_t D, T* __restrict__ out, int64_t ldo, T eps) {
  __shared__ T sA[RK][RT + 1];
  __shared__ T sB[RK][RT + 1];
  __shared__ T sNa[RT], sNb[RT];

  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int64_t row0 = (int64_t)blockIdx.y * RT;
  const int64_t col0 = (int64_t)blockIdx.x * RT;

  // loader mapping: thread -> (row lr, 4 consecutive k starting at lk)
  const int lr = tid >> 2, lk = (tid & 3) * 4;
  const int64_t ga = row0 + lr, gb = col0 + lr;
  T na = 0, nb = 0;  // partial row norms (sum x^2 for Tanimoto, sum x for MinMax)

  T acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0;

  for (int64_t k0 = 0; k0 < D; k0 += RK) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int64_t k = k0 + lk + q;
      const T a = (ga < n && k < D) ? X1[ga * ld1 + k] : T(0);
      const T b = (gb < m && k < D) ? X2[gb * ld2 + k] : T(0);
      sA[lk + q][lr] = a;
      sB[lk + q][lr] = b;
      if (kMinMax) {
        na += a;
        nb += b;
      } else {
        na += a * a;
        nb += b * b;
      }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < RK; ++k) {
      T a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = sA[k][ty + 16 * i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = sB[k][tx + 16 * j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (kMinMax) {
            acc[i][j] += fabs(a[i] - b[j]);
          } else {
            acc[i][j] += a[i] * b[j];
          }
        }
    }
    __syncthreads();
  }

  // reduce the 4 partial norms of each row (4 consecutive lanes)
  na += __shfl_xor_sync(0xffffffffu, na, 1);
  na += __shfl_xor_sync(0xffffffffu, na, 2);
  nb += __shfl_xor_sync(0xffffffffu, nb, 1);
  nb += __shfl_xor_sync(0xffffffffu, nb, 2);
  if ((tid & 3) == 0) {
    sNa[lr] = na;
    sNb[lr] = nb;
  }
  __syncthreads();

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t gr = row0 + ty + 16 * i;
    if (gr >= n) continue;
    const T n1 = sNa[ty + 16 * i];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t gc = col0 + tx + 16 * j;
      if (gc >= m) continue;
      const T n2 = sNb[tx + 16 * j];
      T num, den;
      if (kMinMax) {
        const T s = n1 + n2;
        const T d = acc[i][j];
        num = (s - d) + eps;
        den = (s + d) + eps;
      } else {
        const T dot = acc[i][j];
        num = dot + eps;
        den = ((eps + n1) + n2) - dot;
      }
      const T q = num / den;
      out[gr * ldo + gc] = q < T(0) ? T(0) : q;
    }
  }
}

template <bool kMinMax>
static int dispatch_real(const char* name, const void* x1, int64_t ld1, int64_t n, const void* x2,
                         int64_t ld2, int64_t m, int64_t D, void* out, int64_t ldo, int dtype,
                         double eps, void* stream) {
  GK_CHECK_ARG(n >= 0 && m >= 0 && D >= 0, GK_EINVAL, "%s: negative size", name);
  if (n == 0 || m == 0) return 0;
  GK_CHECK_ARG(x1 && x2 && out, GK_EINVAL, "%s: null pointer", name);
  GK_CHECK_ARG(ld1 >= D && ld2 >= D && ldo >= m, GK_EINVAL, "%s: leading dimension too small", name);
  const int64_t gx = (m + RT - 1) / RT, gy = (n + RT - 1) / RT;
  GK_CHECK_ARG(gy <= 65535, GK_ERANGE, "%s: n too large for the real-valued path (%lld rows)", name,
               (long long)n);
  dim3 grid((unsigned)gx, (unsigned)gy);
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == GK_F32) {
    gram_real_kernel<float, kMinMax><<<grid, 256, 0, s>>>((const float*)x1, ld1, n, (const float*)x2,
                                                          ld2, m, D, (float*)out, ldo, (float)eps);
  } else if (dtype == GK_F64) {
    gram_real_kernel<double, kMinMax><<<grid, 256, 0, s>>>((const double*)x1, ld1, n,
                                                           (const double*)x2, ld2, m, D,
                                                           (double*)out, ldo, eps);
  } else {
    return set_err(GK_EINVAL, "%s: dtype must be GK_F32 or GK_F64 (got %d)", name, dtype);
  }
  return cuda_err(cudaGetLastError(), name);
}

}  // namespace gk

extern "C" int gk_tanimoto_gram_real(const void* x1, int64_t ld1, int64_t n, const void* x2,
                                     int64_t ld2, int64_t m, int64_t D, void* out, int64_t ldo,
                                     int dtype, double eps, void* stream) {
  return gk::dispatch_real<false>("gk_tanimoto_gram_real", x1, ld1, n, x2, ld2, m, D, out, ldo,
                                  dtype, eps, stream);
}

extern "C" int gk_minmax_gram_real(const void* x1, int64_t ld1, int64_t n, const void* x2,
                                   int64_t ld2, int64_t m, int64_t D, void* out, int64_t ldo,
                                   int dtype, double eps, void* stream) {
  return gk::dispatch_real<true>("gk_minmax_gram_real", x1, ld1, n, x2, ld2, m, D, out, ldo,
                                 dtype, eps, stream);
}
