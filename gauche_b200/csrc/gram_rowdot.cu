// gram_rowdot.cu — fused screening: scores = bias + K(a, b) @ alpha with the Tanimoto tiles consumed in the epilogue of
// the tcgen05 Gram kernel (gram_tc.cuh, TanimotoRowdotF32x2); the Gram block is never written.  Caller side of the
// path: gauche/gp.py:169-226 (posterior mean K(X*, X) @ alpha of the BO screening loop).
#include "gram_tc.cuh"

namespace gk {
namespace gtc {

// scores[i] = bias + sum_p partials[p, i].  Block = 32 rows x 8 slab lanes: lane y sums slabs y, y+8, ...
// (coalesced 128-byte reads), then the 8 partial sums are added in a fixed order -> deterministic.
__global__ void __launch_bounds__(256)
reduce_partials_kernel(const float* __restrict__ partials, int64_t slabs, int64_t n, int64_t ldp, float bias,
                       float* __restrict__ scores) {
  __shared__ float red[8][32];
  const int x = threadIdx.x & 31, y = threadIdx.x >> 5;
  for (int64_t r0 = (int64_t)blockIdx.x * 32; r0 < n; r0 += (int64_t)gridDim.x * 32) {
    const int64_t i = r0 + x;
    float acc = 0.f;
    if (i < n)
      for (int64_t p = y; p < slabs; p += 8) acc += __ldg(partials + p * ldp + i);
    red[y][x] = acc;
    __syncthreads();
    if (y == 0 && i < n) {
      float t = red[0][x];
#pragma unroll
      for (int k = 1; k < 8; ++k) t += red[k][x];
      scores[i] = bias + t;
    }
    __syncthreads();
  }
}

static int reduce_partials(const float* partials, int64_t slabs, int64_t n, int64_t ldp, float bias, float* scores,
                           cudaStream_t s) {
  int64_t blocks = (n + 31) / 32;
  const int64_t cap = (int64_t)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  reduce_partials_kernel<<<(unsigned)blocks, 256, 0, s>>>(partials, slabs, n, ldp, bias, scores);
  return cuda_err(cudaGetLastError(), "reduce_partials_kernel launch");
}

static inline int tile_n(int operand_kind) {
  return operand_kind == 0 ? KindI8::BN : operand_kind == 3 ? KindMxf4N192::BN : KindMxf4::BN;
}

}  // namespace gtc
}  // namespace gk

extern "C" int gk_reduce_partials(const float* partials, int64_t slabs, int64_t n, int64_t ldp, float bias,
                                  float* scores, void* stream) {
  using namespace gk;
  GK_CHECK_ARG(slabs >= 0 && n >= 0, GK_EINVAL, "gk_reduce_partials: negative size");
  if (n == 0) return 0;
  GK_CHECK_ARG(partials && scores && ldp >= n, GK_EINVAL, "gk_reduce_partials: bad arguments");
  return gtc::reduce_partials(partials, slabs, n, ldp, bias, scores, (cudaStream_t)stream);
}

// operand_kind 3 = the hybrid operands of gk_tanimoto_rowdot_hybrid (128 x 192 tiles)
extern "C" int64_t gk_tanimoto_rowdot_slabs(int64_t m, int operand_kind) {
  const int bn = gk::gtc::tile_n(operand_kind);
  return 2 * ((m + bn - 1) / bn);   // one slab per column tile and epilogue role
}

extern "C" int gk_tanimoto_rowdot(const void* a, int64_t lda, const int32_t* a_sq, int64_t n,
                                  const void* b, int64_t ldb, const int32_t* b_sq, int64_t m,
                                  int64_t k_bytes, int operand_kind, const float* alpha, double eps, float bias,
                                  float* partials, int64_t ldp, float* scores, void* stream) {
  using namespace gk;
  using namespace gk::gtc;
  const char* name = "gk_tanimoto_rowdot";
  GK_CHECK_ARG(operand_kind >= 0 && operand_kind <= 2, GK_EINVAL,
               "%s: operand_kind must be 0 (u8), 1 (packed e2m1) or 2 (packed bits)", name);
  GK_CHECK_ARG(n >= 0 && m > 0, GK_EINVAL, "%s: bad size", name);
  if (n == 0) return 0;
  GK_CHECK_ARG(alpha && partials && scores, GK_EINVAL, "%s: null pointer", name);
  GK_CHECK_ARG(ldp >= n && ldp % 4 == 0 && ((uintptr_t)partials) % 16 == 0, GK_EALIGN,
               "%s: partials must be 16-byte aligned with ldp >= n and a multiple of 4", name);
  GK_CHECK_ARG(eps >= 1e-12 && eps <= 1.0, GK_ERANGE, "%s: eps outside [1e-12, 1]", name);
  if (int e = check_args(name, a, lda, a_sq, n, b, ldb, b_sq, m, k_bytes, operand_kind == 2 ? 4 : BK, partials, m, GK_F32))
    return e;
  if (operand_kind != 0)
    GK_CHECK_ARG(k_bytes * (operand_kind == 2 ? 8 : 2) < (1LL << 24), GK_ERANGE,
                 "%s: fingerprint too long for exact f32 accumulation", name);
  cudaStream_t s = (cudaStream_t)stream;
  TanimotoRowdotF32x2 epi;
  epi.eps = (float)eps;
  epi.alpha = alpha;
  epi.partials = partials;
  epi.ldp = ldp;
  int rc;
  if (operand_kind == 2)
    rc = launch<1, KindMxf4, true>(a, lda, a_sq, n, b, ldb, b_sq, m, k_bytes, partials, ldp, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, epi, s);
  else if (operand_kind == 1)   // single CTAs: CTA pairs measured 17 % slower here (1370 vs 1650 Gpairs/s), unlike for MinMax
    rc = launch<1, KindMxf4, false>(a, lda, a_sq, n, b, ldb, b_sq, m, k_bytes, partials, ldp, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, epi, s);
  else
    rc = launch<2, KindI8, false>(a, lda, a_sq, n, b, ldb, b_sq, m, k_bytes, partials, ldp, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, epi, s);
  if (rc) return rc;
  const int bn = tile_n(operand_kind);
  return reduce_partials(partials, 2 * ((m + bn - 1) / bn), n, ldp, bias, scores, s);
}

// Hybrid operands (gk_tanimoto_gram_hybrid): a = pre-permuted packed bits (gk_permute_bits_kblock; lda and `words` in 32-bit
// words), b = packed e2m1 (ldb, k_bytes in bytes).  Workspace: gk_tanimoto_rowdot_slabs(m, 3) x ldp floats.
extern "C" int gk_tanimoto_rowdot_hybrid(const uint32_t* a_bitsp, int64_t lda, const int32_t* a_sq, int64_t n,
                                         const uint8_t* b4, int64_t ldb, const int32_t* b_sq, int64_t m,
                                         int64_t words, int64_t k_bytes, const float* alpha, double eps, float bias,
                                         float* partials, int64_t ldp, float* scores, void* stream) {
  using namespace gk;
  using namespace gk::gtc;
  const char* name = "gk_tanimoto_rowdot_hybrid";
  GK_CHECK_ARG(n >= 0 && m > 0, GK_EINVAL, "%s: bad size", name);
  if (n == 0) return 0;
  GK_CHECK_ARG(alpha && partials && scores, GK_EINVAL, "%s: null pointer", name);
  GK_CHECK_ARG(ldp >= n && ldp % 4 == 0 && ((uintptr_t)partials) % 16 == 0, GK_EALIGN,
               "%s: partials must be 16-byte aligned with ldp >= n and a multiple of 4", name);
  GK_CHECK_ARG(eps >= 1e-12 && eps <= 1.0, GK_ERANGE, "%s: eps outside [1e-12, 1]", name);
  GK_CHECK_ARG(words > 0 && words % 8 == 0 && lda >= words && a_bitsp && ((uintptr_t)a_bitsp) % 16 == 0 && lda % 4 == 0, GK_EALIGN,
               "%s: the permuted bit rows hold whole 256-bit k-blocks, 16-byte aligned", name);
  GK_CHECK_ARG(k_bytes <= words * 16 && k_bytes > (words - 8) * 16, GK_EINVAL,
               "%s: k_bytes (%lld) does not match %lld bit words", name, (long long)k_bytes, (long long)words);
  if (int e = check_args(name, b4, ldb, a_sq, n, b4, ldb, b_sq, m, k_bytes, BK, partials, m, GK_F32)) return e;
  GK_CHECK_ARG(words * 32 < (1LL << 24), GK_ERANGE, "%s: fingerprint too long for exact f32 accumulation", name);
  cudaStream_t s = (cudaStream_t)stream;
  TanimotoRowdotF32x2 epi;
  epi.eps = (float)eps;
  epi.alpha = alpha;
  epi.partials = partials;
  epi.ldp = ldp;
  if (int rc = launch<1, KindMxf4N192, 3>(a_bitsp, lda * 4, a_sq, n, b4, ldb, b_sq, m, words * 4, partials, ldp,
                                          CU_TENSOR_MAP_DATA_TYPE_FLOAT32, epi, s, k_bytes))
    return rc;
  return reduce_partials(partials, 2 * ((m + KindMxf4N192::BN - 1) / KindMxf4N192::BN), n, ldp, bias, scores, s);
}
