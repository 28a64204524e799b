// gram_tanimoto.cu — C entry points of the Tanimoto Gram on the tcgen05 kernel template (gram_tc.cuh).
// Replaces gauche/kernels/fingerprint_kernels/tanimoto_kernel.py:36-46.
//   gk_tanimoto_gram_bits   packed-bit operands, expanded to e2m1 inside the SM (default for bit fingerprints)
//   gk_tanimoto_gram_mxf4   ready-made packed-e2m1 operands (gk_expand_bits_e2m1)
//   gk_tanimoto_gram_hybrid A = pre-permuted packed bits expanded into tensor memory, B = ready-made e2m1 by TMA
//   gk_tanimoto_gram_umma2  u8 operands on kind::i8 (count fingerprints), CTA pairs or single CTAs
#include "gram_tc.cuh"

namespace gk {
namespace gtc {

template <int kCG, typename Kind, int kBX>
static int dispatch(const void* a, int64_t lda, const int32_t* a_n, int64_t n, const void* b, int64_t ldb,
                    const int32_t* b_n, int64_t m, int64_t k, void* out, int64_t ldo, int out_dtype, double eps,
                    cudaStream_t s, int64_t k_b = 0) {
  const bool fast_div = eps >= 1e-12 && eps <= 1.0;   // normal operands, normal quotient: branch-free exact division
  switch (out_dtype) {
    case GK_F32:
      if (fast_div)
        return launch<kCG, Kind, kBX>(a, lda, a_n, n, b, ldb, b_n, m, k, out, ldo, CU_TENSOR_MAP_DATA_TYPE_FLOAT32,
                                      TanimotoF32x2{(float)eps}, s, k_b);
      return launch<kCG, Kind, kBX>(a, lda, a_n, n, b, ldb, b_n, m, k, out, ldo, CU_TENSOR_MAP_DATA_TYPE_FLOAT32,
                                    Elementwise<TanimotoEpilogue<float>, float>{TanimotoEpilogue<float>{(float)eps}}, s, k_b);
    case GK_F64:
      if (fast_div)
        return launch<kCG, Kind, kBX>(a, lda, a_n, n, b, ldb, b_n, m, k, out, ldo, CU_TENSOR_MAP_DATA_TYPE_FLOAT64,
                                      Elementwise<TanimotoEpilogueInt<double>, double>{TanimotoEpilogueInt<double>{{eps}}}, s, k_b);
      return launch<kCG, Kind, kBX>(a, lda, a_n, n, b, ldb, b_n, m, k, out, ldo, CU_TENSOR_MAP_DATA_TYPE_FLOAT64,
                                    Elementwise<TanimotoEpilogue<double>, double>{TanimotoEpilogue<double>{eps}}, s, k_b);
    case GK_I32:
      return launch<kCG, Kind, kBX>(a, lda, a_n, n, b, ldb, b_n, m, k, out, ldo, CU_TENSOR_MAP_DATA_TYPE_INT32,
                                    Elementwise<RawEpilogue, int32_t>{RawEpilogue{}}, s, k_b);
    default:
      return set_err(GK_EINVAL, "tanimoto gram: unsupported out_dtype %d", out_dtype);
  }
}

}  // namespace gtc
}  // namespace gk

extern "C" int gk_tanimoto_gram_umma2(const uint8_t* a, int64_t lda, const int32_t* a_sq, int64_t n,
                                      const uint8_t* b, int64_t ldb, const int32_t* b_sq, int64_t m,
                                      int64_t k, void* out, int64_t ldo, int out_dtype, double eps,
                                      int cta_group, void* stream) {
  using namespace gk;
  using namespace gk::gtc;
  const char* name = "gk_tanimoto_gram_umma2";
  if (n == 0 || m == 0) return 0;
  GK_CHECK_ARG(cta_group == 1 || cta_group == 2, GK_EINVAL, "%s: cta_group must be 1 or 2 (CTA pair)", name);
  if (int e = check_args(name, a, lda, a_sq, n, b, ldb, b_sq, m, k, BK, out, ldo, out_dtype)) return e;
  cudaStream_t s = (cudaStream_t)stream;
  if (cta_group == 2) return dispatch<2, KindI8, 0>(a, lda, a_sq, n, b, ldb, b_sq, m, k, out, ldo, out_dtype, eps, s);
  return dispatch<1, KindI8, 0>(a, lda, a_sq, n, b, ldb, b_sq, m, k, out, ldo, out_dtype, eps, s);
}

extern "C" int gk_tanimoto_gram_mxf4(const uint8_t* a4, int64_t lda, const int32_t* a_n, int64_t n,
                                     const uint8_t* b4, int64_t ldb, const int32_t* b_n, int64_t m,
                                     int64_t k_bytes, void* out, int64_t ldo, int out_dtype, double eps,
                                     void* stream) {
  using namespace gk;
  using namespace gk::gtc;
  const char* name = "gk_tanimoto_gram_mxf4";
  if (n == 0 || m == 0) return 0;
  if (int e = check_args(name, a4, lda, a_n, n, b4, ldb, b_n, m, k_bytes, BK, out, ldo, out_dtype)) return e;
  // f32 accumulation of 0/1 products is exact while every count stays below 2^24
  GK_CHECK_ARG(k_bytes * 2 < (1LL << 24), GK_ERANGE, "%s: %lld bits per fingerprint exceed exact f32 accumulation", name,
               (long long)(k_bytes * 2));
  return dispatch<1, KindMxf4, 0>(a4, lda, a_n, n, b4, ldb, b_n, m, k_bytes, out, ldo, out_dtype, eps,
                                      (cudaStream_t)stream);
}

extern "C" int gk_tanimoto_gram_bits(const uint32_t* a_bits, int64_t lda, const int32_t* a_n, int64_t n,
                                     const uint32_t* b_bits, int64_t ldb, const int32_t* b_n, int64_t m,
                                     int64_t words, void* out, int64_t ldo, int out_dtype, double eps,
                                     void* stream) {
  using namespace gk;
  using namespace gk::gtc;
  const char* name = "gk_tanimoto_gram_bits";
  if (n == 0 || m == 0) return 0;
  if (int e = check_args(name, a_bits, lda * 4, a_n, n, b_bits, ldb * 4, b_n, m, words * 4, 4, out, ldo, out_dtype)) return e;
  GK_CHECK_ARG(words * 32 < (1LL << 24), GK_ERANGE, "%s: %lld bits per fingerprint exceed exact f32 accumulation", name,
               (long long)(words * 32));
  return dispatch<1, KindMxf4, 1>(a_bits, lda * 4, a_n, n, b_bits, ldb * 4, b_n, m, words * 4, out, ldo, out_dtype, eps,
                                     (cudaStream_t)stream);
}

extern "C" int gk_tanimoto_gram_hybrid(const uint32_t* a_bitsp, int64_t lda, const int32_t* a_n, int64_t n,
                                       const uint8_t* b4, int64_t ldb, const int32_t* b_n, int64_t m,
                                       int64_t words, int64_t k_bytes, void* out, int64_t ldo, int out_dtype, double eps,
                                       void* stream) {
  using namespace gk;
  using namespace gk::gtc;
  const char* name = "gk_tanimoto_gram_hybrid";
  if (n == 0 || m == 0) return 0;
  GK_CHECK_ARG(words > 0 && words % 8 == 0 && lda >= words, GK_EALIGN,
               "%s: the permuted bit rows hold whole 256-bit k-blocks (words %% 8 == 0, lda >= words)", name);
  GK_CHECK_ARG(a_bitsp && ((uintptr_t)a_bitsp) % 16 == 0 && lda % 4 == 0, GK_EALIGN,
               "%s: a_bitsp must be 16-byte aligned with a row pitch that is a multiple of 16 bytes", name);
  GK_CHECK_ARG(k_bytes <= words * 16 && k_bytes > (words - 8) * 16, GK_EINVAL,
               "%s: k_bytes (%lld) does not match %lld bit words", name, (long long)k_bytes, (long long)words);
  if (int e = check_args(name, b4, ldb, a_n, n, b4, ldb, b_n, m, k_bytes, BK, out, ldo, out_dtype)) return e;
  GK_CHECK_ARG(words * 32 < (1LL << 24), GK_ERANGE, "%s: %lld bits per fingerprint exceed exact f32 accumulation", name,
               (long long)(words * 32));
  // fp64 / int32 outputs gain nothing from the deeper ring (their epilogues bound them: 443 vs 449 Gpairs/s for fp64,
  // profiles/r2_bxh_hybrid_kernel.log), so only the packed fp32 epilogue is instantiated on the hybrid mainloop
  GK_CHECK_ARG(out_dtype == GK_F32 && eps >= 1e-12 && eps <= 1.0, GK_EINVAL,
               "%s: fp32 output with eps in [1e-12, 1] (use gk_tanimoto_gram_mxf4 for the other epilogues)", name);
  return launch<1, KindMxf4N192, 3>(a_bitsp, lda * 4, a_n, n, b4, ldb, b_n, m, words * 4, out, ldo,
                                    CU_TENSOR_MAP_DATA_TYPE_FLOAT32, TanimotoF32x2{(float)eps}, (cudaStream_t)stream, k_bytes);
}
