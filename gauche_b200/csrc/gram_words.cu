// gram_words.cu — CUDA-core Gram kernels on 32-bit operand words with fused epilogues.
//
// One templated kernel serves three integer contractions over the same tiling:
//   OpPopc : packed bits,   acc += popc(a & b)          (Tanimoto on bit fingerprints)
//   OpDot4 : 4 x u8 / word, acc  = dp4a(a, b, acc)      (Tanimoto on count fingerprints)
//   OpSad4 : 4 x u8 / word, acc += sum |a_i - b_i|      (MinMax, == torch.cdist(p=1))
// and writes each Gram tile to HBM exactly once through the epilogue functor
// (reference: tanimoto_kernel.py:36-46, minmax_kernel.py:33-44).
//
// Tiling: 64x64 output tile per 256-thread CTA, 4x4 register block per thread with
// rows {ty+16i} and columns {tx+16j} so that 128-bit shared-memory reads are
// conflict-free (row pitch 36 words) and a half-warp stores 64 contiguous bytes.
// Operand K-chunks of 32 words per row are staged with cp.async, double buffered.
#include <cuda_pipeline.h>

#include "common.cuh"

namespace gk {

constexpr int TILE = 64;       // output tile edge
constexpr int KC = 32;         // words per K chunk
constexpr int PITCH = KC + 4;  // smem row pitch in words (16-byte aligned, conflict-free)

template <typename Op, typename Epi, typename OutT>
__global__ void __launch_bounds__(256)
gram_words_kernel(const uint32_t* __restrict__ A, int64_t lda, const int32_t* __restrict__ na,
                  int64_t n, const uint32_t* __restrict__ B, int64_t ldb,
                  const int32_t* __restrict__ nb, int64_t m, int64_t words,
                  OutT* __restrict__ out, int64_t ldo, Epi epi) {
  __shared__ __align__(16) uint32_t sA[2][TILE][PITCH];
  __shared__ __align__(16) uint32_t sB[2][TILE][PITCH];

  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int64_t row0 = (int64_t)blockIdx.y * TILE;
  const int64_t col0 = (int64_t)blockIdx.x * TILE;

  // stage loader: 64 rows x 8 uint4 per operand = 512 uint4, 2 per thread per operand
  auto load_stage = [&](int buf, int64_t w0) {
#pragma unroll
    for (int it = 0; it < 2; ++it) {
      const int idx = tid + it * 256;
      const int r = idx >> 3, v = idx & 7;
      const int64_t w = w0 + v * 4;
      {
        const int64_t gr = row0 + r;
        uint32_t* dst = &sA[buf][r][v * 4];
        if (gr < n && w < words) {
          __pipeline_memcpy_async(dst, A + gr * lda + w, 16);
        } else {
          *reinterpret_cast<uint4*>(dst) = make_uint4(0, 0, 0, 0);
        }
      }
      {
        const int64_t gc = col0 + r;
        uint32_t* dst = &sB[buf][r][v * 4];
        if (gc < m && w < words) {
          __pipeline_memcpy_async(dst, B + gc * ldb + w, 16);
        } else {
          *reinterpret_cast<uint4*>(dst) = make_uint4(0, 0, 0, 0);
        }
      }
    }
    __pipeline_commit();
  };

  int acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0;

  const int64_t nchunks = (words + KC - 1) / KC;
  load_stage(0, 0);
  for (int64_t c = 0; c < nchunks; ++c) {
    const int buf = (int)(c & 1);
    if (c + 1 < nchunks) {
      load_stage(buf ^ 1, (c + 1) * KC);
      __pipeline_wait_prior(1);
    } else {
      __pipeline_wait_prior(0);
    }
    __syncthreads();
#pragma unroll
    for (int v = 0; v < KC / 4; ++v) {
      uint4 a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = *reinterpret_cast<const uint4*>(&sA[buf][ty + 16 * i][v * 4]);
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = *reinterpret_cast<const uint4*>(&sB[buf][tx + 16 * j][v * 4]);
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          int t = acc[i][j];
          t = Op::apply(a[i].x, b[j].x, t);
          t = Op::apply(a[i].y, b[j].y, t);
          t = Op::apply(a[i].z, b[j].z, t);
          t = Op::apply(a[i].w, b[j].w, t);
          acc[i][j] = t;
        }
    }
    __syncthreads();
  }

  // fused epilogue: the Gram tile goes to HBM once
  using TermT = decltype(epi.row_term(0));
  TermT cterm[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int64_t gc = col0 + tx + 16 * j;
    cterm[j] = epi.col_term(gc < m ? nb[gc] : 0);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t gr = row0 + ty + 16 * i;
    if (gr >= n) continue;
    const TermT rterm = epi.row_term(na[gr]);
    OutT* orow = out + gr * ldo;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t gc = col0 + tx + 16 * j;
      if (gc < m) orow[gc] = (OutT)epi(acc[i][j], rterm, cterm[j]);
    }
  }
}

template <typename Op, typename Epi, typename OutT>
static int launch_words(const uint32_t* A, int64_t lda, const int32_t* na, int64_t n,
                        const uint32_t* B, int64_t ldb, const int32_t* nb, int64_t m,
                        int64_t words, void* out, int64_t ldo, Epi epi, cudaStream_t stream) {
  const int64_t gx = (m + TILE - 1) / TILE, gy = (n + TILE - 1) / TILE;
  // grid.y is limited to 65535: walk the row range in slabs
  const int64_t max_gy = 65535;
  for (int64_t y0 = 0; y0 < gy; y0 += max_gy) {
    const int64_t cur = (gy - y0 < max_gy) ? (gy - y0) : max_gy;
    const int64_t roff = y0 * TILE;
    dim3 grid((unsigned)gx, (unsigned)cur);
    gram_words_kernel<Op, Epi, OutT><<<grid, 256, 0, stream>>>(
        A + roff * lda, lda, na + roff, n - roff, B, ldb, nb, m, words,
        (OutT*)out + roff * ldo, ldo, epi);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_err(e, "gram_words_kernel launch");
  }
  return 0;
}

template <typename Op, template <typename> class EpiT>
static int dispatch_words(const char* name, const uint32_t* A, int64_t lda, const int32_t* na,
                          int64_t n, const uint32_t* B, int64_t ldb, const int32_t* nb, int64_t m,
                          int64_t words, void* out, int64_t ldo, int out_dtype, double eps,
                          void* stream) {
  GK_CHECK_ARG(n >= 0 && m >= 0 && words >= 0, GK_EINVAL, "%s: negative size", name);
  if (n == 0 || m == 0) return 0;
  GK_CHECK_ARG(A && B && na && nb && out, GK_EINVAL, "%s: null pointer", name);
  GK_CHECK_ARG(lda >= words && ldb >= words && ldo >= m, GK_EINVAL, "%s: leading dimension too small", name);
  GK_CHECK_ARG(((uintptr_t)A) % 16 == 0 && ((uintptr_t)B) % 16 == 0 && lda % 4 == 0 && ldb % 4 == 0 &&
                   words % 4 == 0,
               GK_EALIGN, "%s: operands must be 16-byte aligned with lda/ldb/words multiples of 4 words", name);
  GK_CHECK_ARG((m + TILE - 1) / TILE <= 2147483647LL, GK_ERANGE, "%s: m too large", name);
  cudaStream_t s = (cudaStream_t)stream;
  switch (out_dtype) {
    case GK_F32:
      return launch_words<Op, EpiT<float>, float>(A, lda, na, n, B, ldb, nb, m, words, out, ldo,
                                                  EpiT<float>{(float)eps}, s);
    case GK_F64:
      return launch_words<Op, EpiT<double>, double>(A, lda, na, n, B, ldb, nb, m, words, out, ldo,
                                                    EpiT<double>{eps}, s);
    case GK_I32:
      return launch_words<Op, RawEpilogue, int32_t>(A, lda, na, n, B, ldb, nb, m, words, out, ldo,
                                                    RawEpilogue{}, s);
    default:
      return set_err(GK_EINVAL, "%s: unsupported out_dtype %d", name, out_dtype);
  }
}

}  // namespace gk

extern "C" int gk_tanimoto_gram_popc(const uint32_t* a_bits, int64_t lda, const int32_t* a_n,
                                     int64_t n, const uint32_t* b_bits, int64_t ldb,
                                     const int32_t* b_n, int64_t m, int64_t words, void* out,
                                     int64_t ldo, int out_dtype, double eps, void* stream) {
  return gk::dispatch_words<gk::OpPopc, gk::TanimotoEpilogue>(
      "gk_tanimoto_gram_popc", a_bits, lda, a_n, n, b_bits, ldb, b_n, m, words, out, ldo,
      out_dtype, eps, stream);
}

extern "C" int gk_tanimoto_gram_u8(const uint8_t* a, int64_t lda, const int32_t* a_sq, int64_t n,
                                   const uint8_t* b, int64_t ldb, const int32_t* b_sq, int64_t m,
                                   int64_t k, void* out, int64_t ldo, int out_dtype, double eps,
                                   void* stream) {
  if (k % 16 != 0 || lda % 16 != 0 || ldb % 16 != 0)
    return gk::set_err(GK_EALIGN, "gk_tanimoto_gram_u8: k/lda/ldb must be multiples of 16 bytes");
  return gk::dispatch_words<gk::OpDot4, gk::TanimotoEpilogue>(
      "gk_tanimoto_gram_u8", (const uint32_t*)a, lda / 4, a_sq, n, (const uint32_t*)b, ldb / 4,
      b_sq, m, k / 4, out, ldo, out_dtype, eps, stream);
}

extern "C" int gk_minmax_gram_u8(const uint8_t* a, int64_t lda, const int32_t* a_l1, int64_t n,
                                 const uint8_t* b, int64_t ldb, const int32_t* b_l1, int64_t m,
                                 int64_t k, void* out, int64_t ldo, int out_dtype, double eps,
                                 void* stream) {
  if (k % 16 != 0 || lda % 16 != 0 || ldb % 16 != 0)
    return gk::set_err(GK_EALIGN, "gk_minmax_gram_u8: k/lda/ldb must be multiples of 16 bytes");
  return gk::dispatch_words<gk::OpSad4, gk::MinMaxEpilogue>(
      "gk_minmax_gram_u8", (const uint32_t*)a, lda / 4, a_l1, n, (const uint32_t*)b, ldb / 4,
      b_l1, m, k / 4, out, ldo, out_dtype, eps, stream);
}
