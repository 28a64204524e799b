// gram_minmax.cu — MinMax Gram on the FP4 tensor cores (gram_tc.cuh).  Replaces gauche/kernels/fingerprint_kernels/
// minmax_kernel.py:33-44 through the thermometer identity sum_c min(a_c, b_c) = sum_t <[a >= t], [b >= t]>: the
// operands are `planes` indicator planes of packed e2m1 nibbles (gk_expand_thermometer_e2m1), K = planes * Kp, and
// the epilogue rebuilds the exact L1 distance d = l1_a + l1_b - 2 * acc before the reference's float ops.
// Block-sparse variant: k-blocks that are empty on either side are skipped (exact; planes t >= 5 of ECFP counts are
// > 90 % empty blocks).
#include "gram_tc.cuh"

namespace gk {
namespace gtc {

// occ[tile * words + kb / 64] bit (kb % 64) = any non-zero byte in rows [tile * tile_rows, +tile_rows) x bytes [kb * 128, +128)
__global__ void __launch_bounds__(128)
block_occupancy_kernel(const uint8_t* __restrict__ q, int64_t ld, int64_t rows, int tile_rows, int words,
                       unsigned long long* __restrict__ occ) {
  const int64_t tile = blockIdx.x;
  const int kb = blockIdx.y;
  const int64_t r0 = tile * tile_rows;
  int any = 0;
  for (int r = threadIdx.x; r < tile_rows && r0 + r < rows; r += blockDim.x) {
    const uint4* p = reinterpret_cast<const uint4*>(q + (r0 + r) * ld + (int64_t)kb * BK);
    uint32_t o = 0;
#pragma unroll
    for (int i = 0; i < BK / 16; ++i) {
      const uint4 v = __ldg(p + i);
      o |= v.x | v.y | v.z | v.w;
    }
    any |= (o != 0);
  }
  any = __syncthreads_or(any);
  if (threadIdx.x == 0 && any) atomicOr(occ + tile * words + (kb >> 6), 1ull << (kb & 63));
}

template <typename Base, typename Out>
static ElementwiseSparse<Base, Out> sparse_epi(Base base, const uint64_t* occ_a, const uint64_t* occ_b, int words) {
  ElementwiseSparse<Base, Out> epi;
  epi.base = base;
  epi.occ_a = occ_a;
  epi.occ_b = occ_b;
  epi.occ_words = words;
  return epi;
}

}  // namespace gtc
}  // namespace gk

extern "C" int64_t gk_block_occupancy_words(int64_t rows, int64_t k_bytes, int tile_rows) {
  if (rows < 0 || k_bytes <= 0 || tile_rows <= 0) return -1;
  const int64_t tiles = (rows + tile_rows - 1) / tile_rows;
  const int64_t words = (k_bytes / gk::tc::BK + 63) / 64;
  return tiles * words;
}

extern "C" int gk_block_occupancy(const uint8_t* q, int64_t ld, int64_t rows, int64_t k_bytes, int tile_rows,
                                  uint64_t* occ, void* stream) {
  using namespace gk;
  using namespace gk::gtc;
  const char* name = "gk_block_occupancy";
  GK_CHECK_ARG(rows >= 0 && k_bytes > 0 && k_bytes % BK == 0 && tile_rows > 0 && ld >= k_bytes, GK_EINVAL, "%s: bad size", name);
  if (rows == 0) return 0;
  GK_CHECK_ARG(q && occ, GK_EINVAL, "%s: null pointer", name);
  GK_CHECK_ARG(((uintptr_t)q) % 16 == 0 && ld % 16 == 0, GK_EALIGN, "%s: operand must be 16-byte aligned", name);
  const int64_t k_blocks = k_bytes / BK;
  GK_CHECK_ARG(k_blocks <= 65535, GK_ERANGE, "%s: more than 65535 k-blocks", name);
  const int64_t tiles = (rows + tile_rows - 1) / tile_rows;
  const int words = (int)((k_blocks + 63) / 64);
  cudaStream_t s = (cudaStream_t)stream;
  GK_CUDA(cudaMemsetAsync(occ, 0, (size_t)(tiles * words) * sizeof(uint64_t), s));
  block_occupancy_kernel<<<dim3((unsigned)tiles, (unsigned)k_blocks), 128, 0, s>>>(
      q, ld, rows, tile_rows, words, reinterpret_cast<unsigned long long*>(occ));
  return cuda_err(cudaGetLastError(), "block_occupancy_kernel launch");
}

extern "C" int gk_minmax_gram_mxf4_sparse(const uint8_t* a4t, int64_t lda, const int32_t* a_l1, int64_t n,
                                          const uint8_t* b4t, int64_t ldb, const int32_t* b_l1, int64_t m,
                                          int64_t k_bytes, const uint64_t* occ_a, const uint64_t* occ_b, void* out,
                                          int64_t ldo, int out_dtype, double eps, int cta_group, void* stream) {
  using namespace gk;
  using namespace gk::gtc;
  const char* name = "gk_minmax_gram_mxf4_sparse";
  if (n == 0 || m == 0) return 0;
  if (int e = check_args(name, a4t, lda, a_l1, n, b4t, ldb, b_l1, m, k_bytes, BK, out, ldo, out_dtype)) return e;
  GK_CHECK_ARG(occ_a && occ_b, GK_EINVAL, "%s: null occupancy map", name);
  GK_CHECK_ARG(cta_group == 1 || cta_group == 2, GK_EINVAL, "%s: cta_group must be 1 or 2 (CTA pair: occ_a per 256-row tile)", name);
  GK_CHECK_ARG(k_bytes * 2 < (1LL << 24), GK_ERANGE, "%s: %lld thermometer bits exceed exact f32 accumulation", name,
               (long long)(k_bytes * 2));
  GK_CHECK_ARG(eps >= 1e-12 && eps <= 1.0, GK_ERANGE, "%s: eps outside [1e-12, 1] (use gk_minmax_gram_mxf4)", name);
  cudaStream_t s = (cudaStream_t)stream;
  const int words = (int)((k_bytes / BK + 63) / 64);
  if (out_dtype == GK_F32 && k_bytes * 2 < (1LL << 23)) {   // every integer of the epilogue < 2^24: packed fp32 epilogue (exact)
    MinMaxSparseF32x2<false> ep{(float)eps, occ_a, occ_b, words};
    if (cta_group == 2)
      return launch<2, KindMxf4, 0>(a4t, lda, a_l1, n, b4t, ldb, b_l1, m, k_bytes, out, ldo, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, ep, s);
    return launch<1, KindMxf4, 0>(a4t, lda, a_l1, n, b4t, ldb, b_l1, m, k_bytes, out, ldo, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, ep, s);
  }
#define GK_SPARSE(CG)                                                                                                          \
  switch (out_dtype) {                                                                                                         \
    case GK_F32:                                                                                                               \
      return launch<CG, KindMxf4, false>(a4t, lda, a_l1, n, b4t, ldb, b_l1, m, k_bytes, out, ldo, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, \
                                         sparse_epi<MinMaxFromMinEpilogueFast<float>, float>({{(float)eps}}, occ_a, occ_b, words), s); \
    case GK_F64:                                                                                                               \
      return launch<CG, KindMxf4, false>(a4t, lda, a_l1, n, b4t, ldb, b_l1, m, k_bytes, out, ldo, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, \
                                         sparse_epi<MinMaxFromMinEpilogueFast<double>, double>({{eps}}, occ_a, occ_b, words), s);  \
    case GK_I32:                                                                                                               \
      return launch<CG, KindMxf4, false>(a4t, lda, a_l1, n, b4t, ldb, b_l1, m, k_bytes, out, ldo, CU_TENSOR_MAP_DATA_TYPE_INT32,   \
                                         sparse_epi<RawL1FromMinEpilogue, int32_t>({}, occ_a, occ_b, words), s);                 \
    default:                                                                                                                   \
      return set_err(GK_EINVAL, "%s: unsupported out_dtype %d", name, out_dtype);                                              \
  }
  if (cta_group == 2) {   // CTA pairs: 256-row tiles, each CTA stages half of the B tile -> 32 % less operand ingress per pair
    GK_SPARSE(2)
  }
  GK_SPARSE(1)
#undef GK_SPARSE
}

// Symmetric Gram K(x, x): only the tiles that touch or lie above the diagonal are computed (about half of the tensor
// work), each chunk is stored twice — as computed and mirrored.  Exact: the reference's MinMax Gram is bitwise symmetric
// (minmax_kernel.py:33-44: fl(n_i + n_j) commutes, cdist is symmetric), unlike its Tanimoto Gram.
extern "C" int gk_minmax_gram_mxf4_sym(const uint8_t* a4t, int64_t lda, const int32_t* a_l1, int64_t n, int64_t k_bytes,
                                       const uint64_t* occ_rows, const uint64_t* occ_cols, void* out, int64_t ldo,
                                       int out_dtype, double eps, int cta_group, void* stream) {
  using namespace gk;
  using namespace gk::gtc;
  const char* name = "gk_minmax_gram_mxf4_sym";
  if (n == 0) return 0;
  if (int e = check_args(name, a4t, lda, a_l1, n, a4t, lda, a_l1, n, k_bytes, BK, out, ldo, out_dtype)) return e;
  GK_CHECK_ARG(occ_rows && occ_cols, GK_EINVAL, "%s: null occupancy map", name);
  GK_CHECK_ARG(cta_group == 1 || cta_group == 2, GK_EINVAL, "%s: cta_group must be 1 or 2 (CTA pair: occ_rows per 256-row tile)", name);
  GK_CHECK_ARG(k_bytes * 2 < (1LL << 24), GK_ERANGE, "%s: %lld thermometer bits exceed exact f32 accumulation", name,
               (long long)(k_bytes * 2));
  GK_CHECK_ARG(eps >= 1e-12 && eps <= 1.0, GK_ERANGE, "%s: eps outside [1e-12, 1] (use gk_minmax_gram_mxf4)", name);
  GK_CHECK_ARG(out_dtype == GK_F32 || out_dtype == GK_I32, GK_EINVAL, "%s: fp32 or int32 output (4-byte elements)", name);
  cudaStream_t s = (cudaStream_t)stream;
  const int words = (int)((k_bytes / BK + 63) / 64);
  ElementwiseSparseSym<MinMaxFromMinEpilogueFast<float>, float> ef;
  ef.base = {{(float)eps}};
  ef.occ_a = occ_rows;
  ef.occ_b = occ_cols;
  ef.occ_words = words;
  ElementwiseSparseSym<RawL1FromMinEpilogue, int32_t> ei;
  ei.occ_a = occ_rows;
  ei.occ_b = occ_cols;
  ei.occ_words = words;
  if (out_dtype == GK_F32 && k_bytes * 2 < (1LL << 23)) {   // packed fp32 epilogue (exact in this range)
    MinMaxSparseF32x2<true> ep{(float)eps, occ_rows, occ_cols, words};
    if (cta_group == 2)
      return launch<2, KindMxf4, 0>(a4t, lda, a_l1, n, a4t, lda, a_l1, n, k_bytes, out, ldo, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, ep, s);
    return launch<1, KindMxf4, 0>(a4t, lda, a_l1, n, a4t, lda, a_l1, n, k_bytes, out, ldo, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, ep, s);
  }
  if (out_dtype == GK_F32) {
    if (cta_group == 2)
      return launch<2, KindMxf4, false>(a4t, lda, a_l1, n, a4t, lda, a_l1, n, k_bytes, out, ldo, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, ef, s);
    return launch<1, KindMxf4, false>(a4t, lda, a_l1, n, a4t, lda, a_l1, n, k_bytes, out, ldo, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, ef, s);
  }
  if (cta_group == 2)
    return launch<2, KindMxf4, false>(a4t, lda, a_l1, n, a4t, lda, a_l1, n, k_bytes, out, ldo, CU_TENSOR_MAP_DATA_TYPE_INT32, ei, s);
  return launch<1, KindMxf4, false>(a4t, lda, a_l1, n, a4t, lda, a_l1, n, k_bytes, out, ldo, CU_TENSOR_MAP_DATA_TYPE_INT32, ei, s);
}

extern "C" int gk_minmax_gram_mxf4(const uint8_t* a4t, int64_t lda, const int32_t* a_l1, int64_t n,
                                   const uint8_t* b4t, int64_t ldb, const int32_t* b_l1, int64_t m,
                                   int64_t k_bytes, void* out, int64_t ldo, int out_dtype, double eps,
                                   void* stream) {
  using namespace gk;
  using namespace gk::gtc;
  const char* name = "gk_minmax_gram_mxf4";
  if (n == 0 || m == 0) return 0;
  if (int e = check_args(name, a4t, lda, a_l1, n, b4t, ldb, b_l1, m, k_bytes, BK, out, ldo, out_dtype)) return e;
  GK_CHECK_ARG(k_bytes * 2 < (1LL << 24), GK_ERANGE, "%s: %lld thermometer bits exceed exact f32 accumulation", name,
               (long long)(k_bytes * 2));
  cudaStream_t s = (cudaStream_t)stream;
  const bool fast_div = eps >= 1e-12 && eps <= 1.0;   // branch-free exact division (common.cuh)
  switch (out_dtype) {
    case GK_F32:
      if (fast_div)
        return launch<1, KindMxf4, false>(a4t, lda, a_l1, n, b4t, ldb, b_l1, m, k_bytes, out, ldo, CU_TENSOR_MAP_DATA_TYPE_FLOAT32,
                                          Elementwise<MinMaxFromMinEpilogueFast<float>, float>{MinMaxFromMinEpilogueFast<float>{{(float)eps}}}, s);
      return launch<1, KindMxf4, false>(a4t, lda, a_l1, n, b4t, ldb, b_l1, m, k_bytes, out, ldo, CU_TENSOR_MAP_DATA_TYPE_FLOAT32,
                                        Elementwise<MinMaxFromMinEpilogue<float>, float>{MinMaxFromMinEpilogue<float>{(float)eps}}, s);
    case GK_F64:
      if (fast_div)
        return launch<1, KindMxf4, false>(a4t, lda, a_l1, n, b4t, ldb, b_l1, m, k_bytes, out, ldo, CU_TENSOR_MAP_DATA_TYPE_FLOAT64,
                                          Elementwise<MinMaxFromMinEpilogueFast<double>, double>{MinMaxFromMinEpilogueFast<double>{{eps}}}, s);
      return launch<1, KindMxf4, false>(a4t, lda, a_l1, n, b4t, ldb, b_l1, m, k_bytes, out, ldo, CU_TENSOR_MAP_DATA_TYPE_FLOAT64,
                                        Elementwise<MinMaxFromMinEpilogue<double>, double>{MinMaxFromMinEpilogue<double>{eps}}, s);
    case GK_I32:
      return launch<1, KindMxf4, false>(a4t, lda, a_l1, n, b4t, ldb, b_l1, m, k_bytes, out, ldo, CU_TENSOR_MAP_DATA_TYPE_INT32,
                                        Elementwise<RawL1FromMinEpilogue, int32_t>{RawL1FromMinEpilogue{}}, s);
    default:
      return set_err(GK_EINVAL, "%s: unsupported out_dtype %d", name, out_dtype);
  }
}
