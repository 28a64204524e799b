// siblings.cu — GAUCHE's other twelve fingerprint kernels on REAL-VALUED dense inputs (CUDA cores; correctness path,
// tolerance-level parity like gram_real.cu).  Bit / count fingerprints never come here: they run as epilogues of the
// tensor-core Gram kernel (gram_siblings.cu).  The arithmetic is shared (siblings.cuh); like there the reference's
// shape quirks arrive as data: row_a[i] (|x1_i|_1, or 2|x1_i|_1 + 2|x2_i|_1 for rogers_tanimoto / sokal_sneath),
// row_b[i] (d or the Braun-Blanquet denominator; may be NULL), col[j] = |x2_j|_1, all in the compute dtype.
#include "siblings.cuh"

namespace gk {

// out[r] = sum_k x[r, k] (torch.sum(x, dim=-1) of the reference, in the input dtype; one warp per row)
template <typename T>
__global__ void __launch_bounds__(256)
rowsum_real_kernel(const T* __restrict__ x, int64_t ld, int64_t rows, int64_t D, T* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t r = warp0; r < rows; r += nwarps) {
    T s = 0;
    for (int64_t k = lane; k < D; k += 32) s += x[r * ld + k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) out[r] = s;
  }
}

constexpr int ST = 64;   // tile edge
constexpr int SK = 16;   // K chunk

template <typename T, typename OutT>
__global__ void __launch_bounds__(256)
sibling_real_kernel(int which, const T* __restrict__ X1, int64_t ld1, int64_t n, const T* __restrict__ X2, int64_t ld2,
                    int64_t m, int64_t D, const T* __restrict__ row_a, const T* __restrict__ row_b,
                    const T* __restrict__ col, SiblingConsts<T> k, OutT* __restrict__ out, int64_t ldo) {
  __shared__ T sA[SK][ST + 1];
  __shared__ T sB[SK][ST + 1];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int64_t row0 = (int64_t)blockIdx.y * ST, col0 = (int64_t)blockIdx.x * ST;
  const int lr = tid >> 2, lk = (tid & 3) * 4;     // loader: thread -> (row lr, 4 consecutive k starting at lk)
  const int64_t ga = row0 + lr, gb = col0 + lr;
  T acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0;
  for (int64_t k0 = 0; k0 < D; k0 += SK) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int64_t kk = k0 + lk + q;
      sA[lk + q][lr] = (ga < n && kk < D) ? X1[ga * ld1 + kk] : T(0);
      sB[lk + q][lr] = (gb < m && kk < D) ? X2[gb * ld2 + kk] : T(0);
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < SK; ++kk) {
      T a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = sA[kk][ty + 16 * i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = sB[kk][tx + 16 * j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] += a[i] * b[j];
    }
    __syncthreads();
  }
  const bool doubled = which == SIB_ROGERS_TANIMOTO || which == SIB_SOKAL_SNEATH;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t gr = row0 + ty + 16 * i;
    if (gr >= n) continue;
    const T ra = doubled ? (T)2 * row_a[gr] : row_a[gr];     // fl(2 n1 + 2 n2) == 2 fl(n1 + n2)
    const T rb = row_b ? row_b[gr] : (T)0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t gc = col0 + tx + 16 * j;
      if (gc >= m) continue;
      const OutT o = (OutT)sibling_eval<T>(which, acc[i][j], ra, rb, col[gc], k);
      out[gr * ldo + gc] = o < (OutT)0 ? (OutT)0 : o;
    }
  }
}

}  // namespace gk

extern "C" int gk_rowsum_real(const void* x, int64_t ld, int64_t rows, int64_t D, int dtype, void* out, void* stream) {
  using namespace gk;
  const char* name = "gk_rowsum_real";
  GK_CHECK_ARG(rows >= 0 && D >= 0 && ld >= D, GK_EINVAL, "%s: bad size", name);
  if (rows == 0) return 0;
  GK_CHECK_ARG(x && out, GK_EINVAL, "%s: null pointer", name);
  int64_t blocks = (rows + 7) / 8;
  const int64_t cap = (int64_t)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == GK_F32) rowsum_real_kernel<float><<<(unsigned)blocks, 256, 0, s>>>((const float*)x, ld, rows, D, (float*)out);
  else if (dtype == GK_F64) rowsum_real_kernel<double><<<(unsigned)blocks, 256, 0, s>>>((const double*)x, ld, rows, D, (double*)out);
  else return set_err(GK_EINVAL, "%s: dtype must be GK_F32 or GK_F64", name);
  return cuda_err(cudaGetLastError(), name);
}

extern "C" int gk_sibling_gram_real(int which, const void* x1, int64_t ld1, const void* row_a, const void* row_b, int64_t n,
                                    const void* x2, int64_t ld2, const void* col, int64_t m, int64_t D, int64_t nfeat,
                                    int dtype, void* out, int64_t ldo, int out_dtype, double eps, void* stream) {
  using namespace gk;
  const char* name = "gk_sibling_gram_real";
  GK_CHECK_ARG(which >= 0 && which < SIB_COUNT, GK_EINVAL, "%s: unknown kernel id %d", name, which);
  GK_CHECK_ARG(n >= 0 && m >= 0 && D >= 0, GK_EINVAL, "%s: negative size", name);
  if (n == 0 || m == 0) return 0;
  GK_CHECK_ARG(x1 && x2 && row_a && col && out && ld1 >= D && ld2 >= D && ldo >= m, GK_EINVAL, "%s: bad arguments", name);
  GK_CHECK_ARG(n < (1LL << 31) * gk::ST && (m + gk::ST - 1) / gk::ST < 65536LL * 32768, GK_ERANGE, "%s: size too large", name);
  const dim3 grid((unsigned)((m + ST - 1) / ST), (unsigned)((n + ST - 1) / ST));
  GK_CHECK_ARG(grid.y <= 65535, GK_ERANGE, "%s: more than 65535 row tiles; split the row range", name);
  cudaStream_t s = (cudaStream_t)stream;
#define GK_SIB_REAL(T, OutT)                                                                                          \
  sibling_real_kernel<T, OutT><<<grid, 256, 0, s>>>(which, (const T*)x1, ld1, n, (const T*)x2, ld2, m, D, (const T*)row_a, \
                                                    (const T*)row_b, (const T*)col, SiblingConsts<T>::make((double)nfeat, eps), \
                                                    (OutT*)out, ldo)
  if (dtype == GK_F32 && out_dtype == GK_F32) GK_SIB_REAL(float, float);
  else if (dtype == GK_F32 && out_dtype == GK_F64) GK_SIB_REAL(float, double);
  else if (dtype == GK_F64 && out_dtype == GK_F64) GK_SIB_REAL(double, double);
  else return set_err(GK_EINVAL, "%s: unsupported compute/out dtype combination (%d, %d)", name, dtype, out_dtype);
#undef GK_SIB_REAL
  return cuda_err(cudaGetLastError(), name);
}
