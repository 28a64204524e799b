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

// Backward coefficients of the Tanimoto kernel for real-valued inputs.
// With s = <x1_i,x2_j>, a = |x1_i|^2, b = |x2_j|^2, num = s+eps, den = eps+a+b-s, K = num/den:
//   dK/ds = (den+num)/den^2,  dK/da = dK/db = -num/den^2   (zero where the forward clamp is active)
// Given the upstream gradient G this kernel recomputes s/a/b tile by tile and emits
//   W[i,j] = G[i,j]*dK/ds,  ra[i] += sum_j G[i,j]*dK/da,  cb[j] += sum_i G[i,j]*dK/db
// so that grad_x1 = W @ x2 + 2*x1*ra[:,None] and grad_x2 = W^T @ x1 + 2*x2*cb[:,None] (two plain GEMMs).
template <typename T>
__global__ void __launch_bounds__(256)
tanimoto_bwd_coeff_kernel(const T* __restrict__ X1, int64_t ld1, int64_t n, const T* __restrict__ X2,
                          int64_t ld2, int64_t m, int64_t D, const T* __restrict__ G, int64_t ldg,
                          T* __restrict__ W, int64_t ldw, T* __restrict__ ra, T* __restrict__ cb, T eps) {
  __shared__ T sA[RK][RT + 1];
  __shared__ T sB[RK][RT + 1];
  __shared__ T sNa[RT], sNb[RT];
  __shared__ T sCol[16][RT];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int64_t row0 = (int64_t)blockIdx.y * RT;
  const int64_t col0 = (int64_t)blockIdx.x * RT;
  const int lr = tid >> 2, lk = (tid & 3) * 4;
  const int64_t ga = row0 + lr, gb = col0 + lr;
  T na = 0, nb = 0;
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
      na += a * a;
      nb += b * b;
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
        for (int j = 0; j < 4; ++j) acc[i][j] += a[i] * b[j];
    }
    __syncthreads();
  }
  na += __shfl_xor_sync(0xffffffffu, na, 1);
  na += __shfl_xor_sync(0xffffffffu, na, 2);
  nb += __shfl_xor_sync(0xffffffffu, nb, 1);
  nb += __shfl_xor_sync(0xffffffffu, nb, 2);
  if ((tid & 3) == 0) {
    sNa[lr] = na;
    sNb[lr] = nb;
  }
  __syncthreads();
  T colsum[4] = {0, 0, 0, 0};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t gr = row0 + ty + 16 * i;
    T rowsum = 0;
    if (gr < n) {
      const T n1 = sNa[ty + 16 * i];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int64_t gc = col0 + tx + 16 * j;
        if (gc >= m) continue;
        const T s = acc[i][j];
        const T num = s + eps;
        const T den = ((eps + n1) + sNb[tx + 16 * j]) - s;
        const T g = (num / den >= T(0)) ? G[gr * ldg + gc] : T(0);   // clamp_min(0): gradient flows where K >= 0
        const T inv2 = T(1) / (den * den);
        W[gr * ldw + gc] = g * (den + num) * inv2;
        const T v = -g * num * inv2;
        rowsum += v;
        colsum[j] += v;
      }
    }
    // 16 lanes (tx) of a half-warp share the row
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) rowsum += __shfl_xor_sync(0xffffffffu, rowsum, o);
    if (tx == 0 && gr < n) atomicAdd(ra + gr, rowsum);
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) sCol[ty][tx + 16 * j] = colsum[j];
  __syncthreads();
  if (tid < RT) {
    T t = 0;
#pragma unroll
    for (int r = 0; r < 16; ++r) t += sCol[r][tid];
    const int64_t gc = col0 + tid;
    if (gc < m) atomicAdd(cb + gc, t);
  }
}

// Backward coefficients of the MinMax kernel for real-valued inputs.  With S = sum(x1_i) + sum(x2_j),
// d = |x1_i - x2_j|_1, num = (S-d)+eps, den = (S+d)+eps, K = num/den:
//   dK/dS = (den-num)/den^2,   dK/dd = -(den+num)/den^2      (zero where the forward clamp is active)
// Emits CD[i,j] = G*dK/dd and rs[i] += sum_j G*dK/dS, cs[j] += sum_i G*dK/dS, so that
//   grad_x1[i,k] = rs[i] + sum_j CD[i,j]*sgn(x1_ik - x2_jk),  grad_x2[j,k] = cs[j] + sum_i CD[i,j]*sgn(x2_jk - x1_ik)
// (the sign contractions are sign_contract_kernel below).
template <typename T>
__global__ void __launch_bounds__(256)
minmax_bwd_coeff_kernel(const T* __restrict__ X1, int64_t ld1, int64_t n, const T* __restrict__ X2, int64_t ld2,
                        int64_t m, int64_t D, const T* __restrict__ G, int64_t ldg, T* __restrict__ CD,
                        int64_t ldc, T* __restrict__ rs, T* __restrict__ cs, T eps) {
  __shared__ T sA[RK][RT + 1];
  __shared__ T sB[RK][RT + 1];
  __shared__ T sNa[RT], sNb[RT];
  __shared__ T sCol[16][RT];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int64_t row0 = (int64_t)blockIdx.y * RT;
  const int64_t col0 = (int64_t)blockIdx.x * RT;
  const int lr = tid >> 2, lk = (tid & 3) * 4;
  const int64_t ga = row0 + lr, gb = col0 + lr;
  T na = 0, nb = 0;
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
      na += a;
      nb += b;
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
        for (int j = 0; j < 4; ++j) acc[i][j] += fabs(a[i] - b[j]);
    }
    __syncthreads();
  }
  na += __shfl_xor_sync(0xffffffffu, na, 1);
  na += __shfl_xor_sync(0xffffffffu, na, 2);
  nb += __shfl_xor_sync(0xffffffffu, nb, 1);
  nb += __shfl_xor_sync(0xffffffffu, nb, 2);
  if ((tid & 3) == 0) {
    sNa[lr] = na;
    sNb[lr] = nb;
  }
  __syncthreads();
  T colsum[4] = {0, 0, 0, 0};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t gr = row0 + ty + 16 * i;
    T rowsum = 0;
    if (gr < n) {
      const T n1 = sNa[ty + 16 * i];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int64_t gc = col0 + tx + 16 * j;
        if (gc >= m) continue;
        const T S = n1 + sNb[tx + 16 * j];
        const T d = acc[i][j];
        const T num = (S - d) + eps;
        const T den = (S + d) + eps;
        const T g = (num / den >= T(0)) ? G[gr * ldg + gc] : T(0);
        const T inv2 = T(1) / (den * den);
        CD[gr * ldc + gc] = -g * (den + num) * inv2;
        const T v = g * (den - num) * inv2;
        rowsum += v;
        colsum[j] += v;
      }
    }
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) rowsum += __shfl_xor_sync(0xffffffffu, rowsum, o);
    if (tx == 0 && gr < n) atomicAdd(rs + gr, rowsum);
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) sCol[ty][tx + 16 * j] = colsum[j];
  __syncthreads();
  if (tid < RT) {
    T t = 0;
#pragma unroll
    for (int r = 0; r < 16; ++r) t += sCol[r][tid];
    const int64_t gc = col0 + tid;
    if (gc < m) atomicAdd(cs + gc, t);
  }
}

// out[a,k] = base[a] + sum_b C(a,b) * sgn(Xa[a,k] - Xb[b,k]),  C(a,b) = trans ? C[b*ldc+a] : C[a*ldc+b].
// 64 (a) x 64 (k) outputs per CTA, 4x4 per thread, b streamed in chunks of 16 through shared memory.
template <typename T>
__global__ void __launch_bounds__(256)
sign_contract_kernel(const T* __restrict__ C, int64_t ldc, int trans, const T* __restrict__ Xa, int64_t lda,
                     int64_t na_rows, const T* __restrict__ Xb, int64_t ldb, int64_t nb_rows, int64_t D,
                     const T* __restrict__ base, T* __restrict__ out, int64_t ldo) {
  __shared__ T sC[RK][RT + 1];   // [b][a]
  __shared__ T sX[RK][RT + 1];   // [b][k]
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int64_t a0 = (int64_t)blockIdx.y * RT;
  const int64_t k0 = (int64_t)blockIdx.x * RT;
  T xa[4][4], acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t a = a0 + ty + 16 * i, k = k0 + tx + 16 * j;
      xa[i][j] = (a < na_rows && k < D) ? Xa[a * lda + k] : T(0);
      acc[i][j] = 0;
    }
  for (int64_t b0 = 0; b0 < nb_rows; b0 += RK) {
    // 16 x 64 elements each: 4 per thread
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int idx = tid + q * 256;
      const int bb = idx >> 6, cc = idx & 63;
      const int64_t b = b0 + bb;
      const int64_t a = a0 + cc, k = k0 + cc;
      T cv = T(0), xv = T(0);
      if (b < nb_rows) {
        if (a < na_rows) cv = trans ? C[b * ldc + a] : C[a * ldc + b];
        if (k < D) xv = Xb[b * ldb + k];
      }
      sC[bb][cc] = cv;
      sX[bb][cc] = xv;
    }
    __syncthreads();
    const int lim = (int)min((int64_t)RK, nb_rows - b0);
    for (int bb = 0; bb < lim; ++bb) {
      T c[4], x[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) c[i] = sC[bb][ty + 16 * i];
#pragma unroll
      for (int j = 0; j < 4; ++j) x[j] = sX[bb][tx + 16 * j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const T df = xa[i][j] - x[j];
          const T sg = (T)((df > T(0)) - (df < T(0)));
          acc[i][j] += c[i] * sg;
        }
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t a = a0 + ty + 16 * i;
    if (a >= na_rows) continue;
    const T bs = base ? base[a] : T(0);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t k = k0 + tx + 16 * j;
      if (k < D) out[a * ldo + k] = bs + acc[i][j];
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

extern "C" int gk_tanimoto_bwd_coeff(const void* x1, int64_t ld1, int64_t n, const void* x2, int64_t ld2,
                                     int64_t m, int64_t D, const void* G, int64_t ldg, void* W, int64_t ldw,
                                     void* ra, void* cb, int dtype, double eps, void* stream) {
  using namespace gk;
  const char* name = "gk_tanimoto_bwd_coeff";
  GK_CHECK_ARG(n >= 0 && m >= 0 && D >= 0, GK_EINVAL, "%s: negative size", name);
  if (n == 0 || m == 0) return 0;
  GK_CHECK_ARG(x1 && x2 && G && W && ra && cb, GK_EINVAL, "%s: null pointer", name);
  GK_CHECK_ARG(ld1 >= D && ld2 >= D && ldg >= m && ldw >= m, GK_EINVAL, "%s: leading dimension too small", name);
  const int64_t gx = (m + RT - 1) / RT, gy = (n + RT - 1) / RT;
  GK_CHECK_ARG(gy <= 65535, GK_ERANGE, "%s: n too large (%lld rows)", name, (long long)n);
  dim3 grid((unsigned)gx, (unsigned)gy);
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == GK_F32) {
    GK_CUDA(cudaMemsetAsync(ra, 0, n * sizeof(float), s));
    GK_CUDA(cudaMemsetAsync(cb, 0, m * sizeof(float), s));
    tanimoto_bwd_coeff_kernel<float><<<grid, 256, 0, s>>>((const float*)x1, ld1, n, (const float*)x2, ld2, m, D,
                                                          (const float*)G, ldg, (float*)W, ldw, (float*)ra,
                                                          (float*)cb, (float)eps);
  } else if (dtype == GK_F64) {
    GK_CUDA(cudaMemsetAsync(ra, 0, n * sizeof(double), s));
    GK_CUDA(cudaMemsetAsync(cb, 0, m * sizeof(double), s));
    tanimoto_bwd_coeff_kernel<double><<<grid, 256, 0, s>>>((const double*)x1, ld1, n, (const double*)x2, ld2, m, D,
                                                           (const double*)G, ldg, (double*)W, ldw, (double*)ra,
                                                           (double*)cb, eps);
  } else {
    return set_err(GK_EINVAL, "%s: dtype must be GK_F32 or GK_F64 (got %d)", name, dtype);
  }
  return cuda_err(cudaGetLastError(), name);
}

extern "C" int gk_minmax_bwd(const void* x1, int64_t ld1, int64_t n, const void* x2, int64_t ld2, int64_t m,
                             int64_t D, const void* G, int64_t ldg, void* CD, int64_t ldc, void* rs, void* cs,
                             void* grad_x1, int64_t ldg1, void* grad_x2, int64_t ldg2, int dtype, double eps,
                             void* stream) {
  using namespace gk;
  const char* name = "gk_minmax_bwd";
  GK_CHECK_ARG(n >= 0 && m >= 0 && D >= 0, GK_EINVAL, "%s: negative size", name);
  if (n == 0 || m == 0 || D == 0) return 0;
  GK_CHECK_ARG(x1 && x2 && G && CD && rs && cs, GK_EINVAL, "%s: null pointer", name);
  GK_CHECK_ARG(ld1 >= D && ld2 >= D && ldg >= m && ldc >= m, GK_EINVAL, "%s: leading dimension too small", name);
  const int64_t gx = (m + RT - 1) / RT, gy = (n + RT - 1) / RT, gk_ = (D + RT - 1) / RT;
  GK_CHECK_ARG(gy <= 65535 && gx <= 65535, GK_ERANGE, "%s: too many rows for the real-valued path", name);
  cudaStream_t s = (cudaStream_t)stream;
#define GK_MM_BWD(T)                                                                                               \
  do {                                                                                                             \
    GK_CUDA(cudaMemsetAsync(rs, 0, n * sizeof(T), s));                                                              \
    GK_CUDA(cudaMemsetAsync(cs, 0, m * sizeof(T), s));                                                              \
    minmax_bwd_coeff_kernel<T><<<dim3((unsigned)gx, (unsigned)gy), 256, 0, s>>>(                                    \
        (const T*)x1, ld1, n, (const T*)x2, ld2, m, D, (const T*)G, ldg, (T*)CD, ldc, (T*)rs, (T*)cs, (T)eps);      \
    if (grad_x1)                                                                                                   \
      sign_contract_kernel<T><<<dim3((unsigned)gk_, (unsigned)gy), 256, 0, s>>>(                                    \
          (const T*)CD, ldc, 0, (const T*)x1, ld1, n, (const T*)x2, ld2, m, D, (const T*)rs, (T*)grad_x1, ldg1);    \
    if (grad_x2)                                                                                                   \
      sign_contract_kernel<T><<<dim3((unsigned)gk_, (unsigned)gx), 256, 0, s>>>(                                    \
          (const T*)CD, ldc, 1, (const T*)x2, ld2, m, (const T*)x1, ld1, n, D, (const T*)cs, (T*)grad_x2, ldg2);    \
  } while (0)
  if (dtype == GK_F32) GK_MM_BWD(float);
  else if (dtype == GK_F64) GK_MM_BWD(double);
  else return set_err(GK_EINVAL, "%s: dtype must be GK_F32 or GK_F64 (got %d)", name, dtype);
#undef GK_MM_BWD
  return cuda_err(cudaGetLastError(), name);
}
