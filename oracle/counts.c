/*
 * counts.c — plain C restatement of the integer contractions of the reference path, on the
 * PACKED operand formats the CUDA library consumes (test infrastructure only).
 *
 *   oracle_intersections : |A_i ∩ B_j| = popcount(a_i & b_j) over uint32 bit words
 *                          == torch.matmul(x1, x2ᵀ) of tanimoto_kernel.py:36 for 0/1 inputs
 *   oracle_dot_u8        : <a_i, b_j> over u8 counts (same line, count fingerprints)
 *   oracle_l1_u8         : sum_k |a_ik - b_jk| == torch.cdist(x1, x2, p=1) of minmax_kernel.py:36
 *   oracle_pack          : dense double -> bits / u8 / row norms, same layout as gk_pack_classify
 *
 * Build: make -C oracle   (gcc -O2 -shared -fPIC; output oracle/_build/liboracle_counts.so)
 */
#include <stdint.h>
#include <stdlib.h>

void oracle_intersections(const uint32_t* a, int64_t lda, int64_t n, const uint32_t* b, int64_t ldb,
                          int64_t m, int64_t words, int32_t* out, int64_t ldo) {
  for (int64_t i = 0; i < n; ++i)
    for (int64_t j = 0; j < m; ++j) {
      int32_t acc = 0;
      for (int64_t w = 0; w < words; ++w) acc += __builtin_popcount(a[i * lda + w] & b[j * ldb + w]);
      out[i * ldo + j] = acc;
    }
}

void oracle_dot_u8(const uint8_t* a, int64_t lda, int64_t n, const uint8_t* b, int64_t ldb, int64_t m,
                   int64_t k, int32_t* out, int64_t ldo) {
  for (int64_t i = 0; i < n; ++i)
    for (int64_t j = 0; j < m; ++j) {
      int32_t acc = 0;
      for (int64_t t = 0; t < k; ++t) acc += (int32_t)a[i * lda + t] * (int32_t)b[j * ldb + t];
      out[i * ldo + j] = acc;
    }
}

void oracle_l1_u8(const uint8_t* a, int64_t lda, int64_t n, const uint8_t* b, int64_t ldb, int64_t m,
                  int64_t k, int32_t* out, int64_t ldo) {
  for (int64_t i = 0; i < n; ++i)
    for (int64_t j = 0; j < m; ++j) {
      int32_t acc = 0;
      for (int64_t t = 0; t < k; ++t) acc += abs((int)a[i * lda + t] - (int)b[j * ldb + t]);
      out[i * ldo + j] = acc;
    }
}

/* returns classification flags: bit0 = some value not in {0,1}; bit1 = some value not an integer in [0,255] */
int oracle_pack(const double* x, int64_t rows, int64_t D, int64_t ld, uint32_t* bits, int64_t ld_bits,
                uint8_t* q8, int64_t ld_q8, int32_t* rowsum, int32_t* rowsq) {
  int flags = 0;
  for (int64_t r = 0; r < rows; ++r) {
    int32_t s = 0, sq = 0;
    for (int64_t w = 0; w < ld_bits; ++w) bits[r * ld_bits + w] = 0;
    for (int64_t c = 0; c < ld_q8; ++c) q8[r * ld_q8 + c] = 0;
    for (int64_t c = 0; c < D; ++c) {
      const double v = x[r * ld + c];
      const int is_bin = (v == 0.0) || (v == 1.0);
      const int is_u8 = (v >= 0.0) && (v <= 255.0) && (v == (double)(int64_t)v);
      if (!is_bin) flags |= 1;
      if (!is_u8) flags |= 2;
      const uint32_t q = is_u8 ? (uint32_t)v : 0u;
      q8[r * ld_q8 + c] = (uint8_t)q;
      if (v != 0.0) bits[r * ld_bits + (c >> 5)] |= 1u << (c & 31);
      s += (int32_t)q;
      sq += (int32_t)(q * q);
    }
    rowsum[r] = s;
    rowsq[r] = sq;
  }
  return flags;
}
