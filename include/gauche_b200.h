/*
 * gauche_b200.h — C ABI of the B200-native fingerprint-kernel Gram-matrix path.
 *
 * This is the drop-in boundary for the arithmetic under
 *   gauche/kernels/fingerprint_kernels/tanimoto_kernel.py:11-46  (batch_tanimoto_sim)
 *   gauche/kernels/fingerprint_kernels/minmax_kernel.py:10-44    (batch_minmax_sim)
 * of leojklarner/gauche.  The reference has no native code and therefore no FFI of
 * its own; every entry point below names the reference lines it replaces.
 *
 * Conventions
 *   - One shared library built for sm_100a only.  No torch types in any signature.
 *   - Every pointer is a DEVICE pointer owned by the caller (e.g. torch's caching
 *     allocator).  The library never allocates or frees device memory and never
 *     synchronises: all work is enqueued on `stream` (a cudaStream_t passed as void*).
 *   - Return value: 0 = ok, negative = argument error (GK_E*), positive = cudaError_t.
 *     gk_last_error() returns a thread-local human readable message.
 *   - Matrices are row-major; `ld*` arguments are row strides in ELEMENTS of the
 *     pointed-to type unless stated otherwise.
 *   - "n1/n2 norms" are exact integers produced by gk_pack_classify.
 */
#ifndef GAUCHE_B200_H
#define GAUCHE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GK_ABI_VERSION 3

/* element types of dense inputs / outputs */
enum gk_dtype {
  GK_F32 = 0,
  GK_F64 = 1,
  GK_I64 = 2,
  GK_I32 = 3,
  GK_U8 = 4,
  GK_F16 = 5,
  GK_BF16 = 6,
  GK_I8 = 7,
  GK_I16 = 8,
  GK_BOOL = 9
};

/* argument errors */
#define GK_EINVAL (-1)   /* bad size / null pointer / unsupported dtype */
#define GK_EALIGN (-2)   /* pointer or leading dimension violates an alignment rule */
#define GK_ERANGE (-3)   /* size exceeds what the exact-integer path can represent */
#define GK_EDEVICE (-4)  /* not an sm_100 device / driver entry point missing */

/* classification flags OR-ed into *flags by gk_pack_classify */
#define GK_FLAG_NONBINARY 1 /* some value is not exactly 0 or 1 */
#define GK_FLAG_NONU8 2     /* some value is not an integer in [0,255] (incl. NaN/neg.) */

/* gk_query selectors */
#define GK_Q_ABI_VERSION 0
#define GK_Q_SM_COUNT 1
#define GK_Q_CC 2            /* major*10+minor of the current device */
#define GK_Q_UMMA_TILE_M 3
#define GK_Q_UMMA_TILE_N 4
#define GK_Q_K_ALIGN 5       /* required multiple for the padded u8 row length (bytes) */

const char* gk_last_error(void);
int gk_query(int what, int64_t* out);

/*
 * Pack + classify + row norms in one pass over a dense fingerprint matrix.
 * Replaces torch.sum(x**2,-1) (tanimoto_kernel.py:37-38) and torch.sum(x,-1)
 * (minmax_kernel.py:33-34) and produces the operands of the integer Gram kernels.
 *   x       [rows, D] dense, row stride ld (elements), inner stride 1
 *   bits    [rows, ld_bits] uint32 words, bit i of word w <-> (x[., 32w+i] != 0); pad words zeroed
 *   q8      [rows, ld_q8]  u8 copy of x (0 where not representable); pad bytes zeroed; may be NULL (bit
 *           fingerprints never need it: pass NULL first and run a second pass only if flags say "counts");
 *           ld_q8 (the padded row length Kp) must be a multiple of 128 and >= D; ld_bits must equal ld_q8/32
 *   rowsum  [rows] int32  = sum_j q8[i,j]        (L1 norm, popcount for bits)
 *   rowsq   [rows] int32  = sum_j q8[i,j]^2      (squared L2 norm; a row whose sum exceeds int32 sets GK_FLAG_NONU8)
 *   flags   [2]   int32 (caller zeroes both): flags[0] OR-accumulated classification bits,
 *                 flags[1] = largest u8 value seen (thermometer depth of the MinMax tensor-core path)
 */
int gk_pack_classify(const void* x, int dtype, int64_t rows, int64_t D, int64_t ld,
                     uint32_t* bits, int64_t ld_bits, uint8_t* q8, int64_t ld_q8,
                     int32_t* rowsum, int32_t* rowsq, int32_t* flags, void* stream);

/*
 * Tanimoto Gram / cross-kernel tile pipeline.  Replaces tanimoto_kernel.py:36-46:
 *   out[i,j] = clamp_min((dot+eps) / (((eps+n1[i]) + n2[j]) - dot), 0)
 * evaluated in out_dtype (GK_F32 or GK_F64) in exactly the reference's operation
 * order.  out_dtype == GK_I32 writes the raw intersection count / dot product.
 *
 * _popc : CUDA-core path on packed bits (a_bits/b_bits [., words], __popc(a&b)).
 * _u8   : CUDA-core path on u8 counts (dp4a); k = padded row length in bytes.
 */
int gk_tanimoto_gram_popc(const uint32_t* a_bits, int64_t lda, const int32_t* a_n, int64_t n,
                          const uint32_t* b_bits, int64_t ldb, const int32_t* b_n, int64_t m,
                          int64_t words, void* out, int64_t ldo, int out_dtype, double eps,
                          void* stream);
int gk_tanimoto_gram_u8(const uint8_t* a, int64_t lda, const int32_t* a_sq, int64_t n,
                        const uint8_t* b, int64_t ldb, const int32_t* b_sq, int64_t m,
                        int64_t k, void* out, int64_t ldo, int out_dtype, double eps,
                        void* stream);

/*
 * Tensor-core path (csrc/gram_tc.cuh: persistent warp-specialised tcgen05 kernel, TMA -> shared memory ->
 * tcgen05.mma -> TMEM -> fused reference-order epilogue -> TMA store).  Same contract as above.  Operands
 * must be 16-byte aligned with row pitches that are multiples of 16 bytes; `out` may have any pitch (rows
 * whose pitch or base is not 16-byte aligned are written by the epilogue's LSU path instead of TMA stores).
 *
 * gk_tanimoto_gram_bits  : operands are the packed bit words themselves (the wire
 *         format below; lda/ldb/words in 32-bit words).  The bits are expanded to e2m1 nibbles INSIDE the SM by
 *         four expander warps, so only 32 B per row and 256-bit k-block cross L2 -> SM (4x less than ready-made
 *         e2m1 operands), and no expanded copy of the library exists in HBM.
 * gk_tanimoto_gram_mxf4  : operands are ready-made packed e2m1 nibbles (1.0 / 0.0) from gk_expand_bits_e2m1
 *         (a4/b4 [., k_bytes], k_bytes a multiple of 128 = 256 fingerprint bits; lda/ldb in bytes); tcgen05
 *         kind::mxf4.block_scale with unit scale factors accumulates the 0/1 products exactly in f32.
 * gk_tanimoto_gram_umma2 : u8 operands (counts, or 0/1) on kind::i8 with exact s32 accumulation; k = padded row
 *         length in bytes (multiple of 128); cta_group = 2: CTA pairs, 256x256 tiles (default for counts), 1: single CTAs.
 */
int gk_tanimoto_gram_bits(const uint32_t* a_bits, int64_t lda, const int32_t* a_n, int64_t n,
                          const uint32_t* b_bits, int64_t ldb, const int32_t* b_n, int64_t m,
                          int64_t words, void* out, int64_t ldo, int out_dtype, double eps,
                          void* stream);
/* HYBRID operands: A (the 128-row side of a tile) stays packed bits all the way into the SM and is expanded into TENSOR
 * MEMORY by expander warps; B arrives as ready-made e2m1 tiles by TMA (128 x 192 tiles).  A k-block then costs 4 + 24 KB
 * of L2 -> SM traffic and shared memory instead of 16 + 28 KB, so six to eight k-blocks are in flight per SM instead of
 * four.  a_bitsp = gk_permute_bits_kblock(bits) (lda, words in 32-bit words, words % 8 == 0), b4 = gk_expand_bits_e2m1
 * (ldb, k_bytes in bytes).  fp32 output, eps in [1e-12, 1] (the other outputs are bound by their epilogues and stay on
 * gk_tanimoto_gram_mxf4).  Results are bit-identical to the other engines. */
int gk_tanimoto_gram_hybrid(const uint32_t* a_bitsp, int64_t lda, const int32_t* a_n, int64_t n,
                            const uint8_t* b4, int64_t ldb, const int32_t* b_n, int64_t m,
                            int64_t words, int64_t k_bytes, void* out, int64_t ldo, int out_dtype, double eps,
                            void* stream);
/* bit words -> the per-k-block bit order gk_tanimoto_gram_hybrid / gk_tanimoto_rowdot_hybrid expand from: inside every
 * 256-bit k-block, bit 8j+i of word 2s+h moves to bit 4i+s of word 4h+j.  ld_out (words) is a multiple of 8; words past
 * `words` are written as zeros. */
int gk_permute_bits_kblock(const uint32_t* bits, int64_t ld_bits, int64_t rows, int64_t words,
                           uint32_t* out, int64_t ld_out, void* stream);
int gk_tanimoto_gram_mxf4(const uint8_t* a4, int64_t lda, const int32_t* a_n, int64_t n,
                          const uint8_t* b4, int64_t ldb, const int32_t* b_n, int64_t m,
                          int64_t k_bytes, void* out, int64_t ldo, int out_dtype, double eps,
                          void* stream);
int gk_tanimoto_gram_umma2(const uint8_t* a, int64_t lda, const int32_t* a_sq, int64_t n,
                           const uint8_t* b, int64_t ldb, const int32_t* b_sq, int64_t m,
                           int64_t k, void* out, int64_t ldo, int out_dtype, double eps,
                           int cta_group, void* stream);

int gk_expand_bits_e2m1(const uint32_t* bits, int64_t ld_bits, int64_t rows, int64_t words,
                        uint8_t* q4, int64_t ld_q4, void* stream);
/*
 * Packed-bit wire format (256 B per 2048-bit molecule, little-endian bit order: bit i of word w is
 * fingerprint bit 32w+i, i.e. np.packbits(x, axis=1, bitorder="little")): libraries delivered this
 * way never exist as dense floats.  gk_bits_popcount gives the exact row norms, gk_expand_bits_u8
 * the u8 operand of the int8 kernels (the e2m1 operand comes from gk_expand_bits_e2m1).
 */
/*
 * HOST function (no device work, no stream): packs a dense block of 0/1 fingerprints that lives in host memory
 * (dtype GK_F32 / GK_F64 / GK_U8, row stride `ld` elements) into the wire format above — `bits[r, w]`, row stride
 * ld_words >= ceil(d/32) (extra words are zeroed) — plus the exact row popcounts.  a persistent pool of sleeping worker threads splits the rows
 * over `threads` threads (<= 0: all cores), AVX-512 / AVX2 when the CPU has it.  *not_binary = 1 if any element is neither 0 nor 1 (the packed
 * output is then meaningless: ship the block dense and let gk_pack_classify sort it out).  Purpose: a 16 384 x 2048
 * fp32 block takes 2.7 ms over PCIe but 1.1 ms to pack on 16 cores, after which 4 MB cross the bus.  Input
 * preparation only — no similarity arithmetic runs on the host.
 */
int gk_host_pack_bits(const void* x, int dtype, int64_t rows, int64_t d, int64_t ld, uint32_t* bits,
                      int64_t ld_words, int32_t* popcnt, int32_t* not_binary, int threads);
int gk_bits_popcount(const uint32_t* bits, int64_t ld_bits, int64_t rows, int64_t words,
                     int32_t* popcnt, void* stream);
int gk_expand_bits_u8(const uint32_t* bits, int64_t ld_bits, int64_t rows, int64_t words,
                      uint8_t* q8, int64_t ld_q8, void* stream);
int gk_reduce_partials(const float* partials, int64_t slabs, int64_t n, int64_t ldp, float bias,
                       float* scores, void* stream);

/*
 * MinMax Gram on u8 counts.  Replaces minmax_kernel.py:33-44:
 *   s = l1[i]+l1[j];  d = sum_k |a[i,k]-b[j,k]|  (== torch.cdist(p=1))
 *   out[i,j] = clamp_min(((s-d)+eps) / ((s+d)+eps), 0)
 * out_dtype == GK_I32 writes the raw L1 distance d.
 */
int gk_minmax_gram_u8(const uint8_t* a, int64_t lda, const int32_t* a_l1, int64_t n,
                      const uint8_t* b, int64_t ldb, const int32_t* b_l1, int64_t m,
                      int64_t k, void* out, int64_t ldo, int out_dtype, double eps,
                      void* stream);

/*
 * MinMax on the FP4 tensor cores.  sum_c min(a_c,b_c) = sum_t <[a>=t],[b>=t]>: gk_expand_thermometer_e2m1
 * writes `planes` (= largest count value) indicator planes of packed e2m1 nibbles per row
 * (plane stride ceil(kp/256)*128 bytes), and gk_minmax_gram_mxf4 runs the kind::mxf4 Gram kernel over
 * K = planes * Kp with an epilogue that rebuilds the exact L1 distance d = l1_a + l1_b - 2*acc and then
 * evaluates minmax_kernel.py:35-44 in the reference's order.  out_dtype GK_I32 writes d.
 */
int gk_expand_thermometer_e2m1(const uint8_t* q8, int64_t ld_q8, int64_t rows, int64_t kp, int planes,
                               uint8_t* q4t, int64_t ld_q4t, void* stream);
/*
 * Block-sparse variant.  The high thermometer planes of count fingerprints are almost empty (ECFP counts: planes
 * t >= 5 have < 10 % non-empty 128-row x 256-bit blocks), and a k-block whose A or B block is all zero adds
 * nothing: gk_block_occupancy writes one bit per (row tile, 128-byte k-block) of an operand — tile_rows = 128 for
 * the left operand (256 with cta_group = 2: CTA pairs compute 256-row tiles), 224 for the right one; occ holds
 * gk_block_occupancy_words(rows, k_bytes, tile_rows) uint64 —
 * and gk_minmax_gram_mxf4_sparse skips, tile by tile, every k-block that is empty on either side.  Results are
 * identical to gk_minmax_gram_mxf4 (eps in [1e-12, 1]).
 */
int64_t gk_block_occupancy_words(int64_t rows, int64_t k_bytes, int tile_rows);
int gk_block_occupancy(const uint8_t* q, int64_t ld, int64_t rows, int64_t k_bytes, int tile_rows,
                       uint64_t* occ, void* stream);
int gk_minmax_gram_mxf4_sparse(const uint8_t* a4t, int64_t lda, const int32_t* a_l1, int64_t n,
                               const uint8_t* b4t, int64_t ldb, const int32_t* b_l1, int64_t m,
                               int64_t k_bytes, const uint64_t* occ_a, const uint64_t* occ_b, void* out,
                               int64_t ldo, int out_dtype, double eps, int cta_group, void* stream);
/* The SYMMETRIC Gram K(x, x) of gk_minmax_gram_mxf4_sparse: tiles strictly below the diagonal are not computed, every
 * computed 32 x 32 chunk is stored twice (as it is and mirrored), about half the tensor work.  Bit-identical to the full
 * computation because the reference's MinMax Gram is bitwise symmetric (minmax_kernel.py:33-44).  occ_rows: tile_rows =
 * 128 * cta_group, occ_cols: tile_rows = 224, both of the same operand.  out_dtype GK_F32 or GK_I32. */
int gk_minmax_gram_mxf4_sym(const uint8_t* a4t, int64_t lda, const int32_t* a_l1, int64_t n, int64_t k_bytes,
                            const uint64_t* occ_rows, const uint64_t* occ_cols, void* out, int64_t ldo,
                            int out_dtype, double eps, int cta_group, void* stream);
int gk_minmax_gram_mxf4(const uint8_t* a4t, int64_t lda, const int32_t* a_l1, int64_t n,
                        const uint8_t* b4t, int64_t ldb, const int32_t* b_l1, int64_t m,
                        int64_t k_bytes, void* out, int64_t ldo, int out_dtype, double eps,
                        void* stream);

/*
 * Real-valued (non-integer) inputs: same formulas evaluated directly on the dense
 * fp32/fp64 matrices (dtype == out dtype; GK_F32 or GK_F64).  Correctness path for
 * warped / learnable inputs; not the roofline-graded path.
 */
int gk_tanimoto_gram_real(const void* x1, int64_t ld1, int64_t n, const void* x2, int64_t ld2,
                          int64_t m, int64_t D, void* out, int64_t ldo, int dtype, double eps,
                          void* stream);
int gk_minmax_gram_real(const void* x1, int64_t ld1, int64_t n, const void* x2, int64_t ld2,
                        int64_t m, int64_t D, void* out, int64_t ldo, int dtype, double eps,
                        void* stream);

/*
 * Backward of the Tanimoto kernel w.r.t. real-valued inputs (learnable inducing points, input warping):
 * emits W = G*dK/ds and the row/column sums of G*dK/d|x|^2 (ra[n], cb[m], zeroed here) so that
 *   grad_x1 = W @ x2 + 2*x1*ra[:,None],   grad_x2 = W^T @ x1 + 2*x2*cb[:,None]
 * (the two GEMMs are plain library GEMMs on the caller's side).  dtype GK_F32 or GK_F64 for all arrays.
 */
int gk_tanimoto_bwd_coeff(const void* x1, int64_t ld1, int64_t n, const void* x2, int64_t ld2,
                          int64_t m, int64_t D, const void* G, int64_t ldg, void* W, int64_t ldw,
                          void* ra, void* cb, int dtype, double eps, void* stream);

/*
 * Backward of the MinMax kernel w.r.t. real-valued inputs: recomputes the L1 statistics, then
 *   grad_x1[i,k] = sum_j G*dK/dS + sum_j (G*dK/dd)[i,j]*sgn(x1_ik - x2_jk)   (and symmetrically grad_x2).
 * CD [n,m], rs [n], cs [m] are caller-owned workspaces; grad_x1 / grad_x2 may be NULL to skip a side.
 */
int gk_minmax_bwd(const void* x1, int64_t ld1, int64_t n, const void* x2, int64_t ld2, int64_t m,
                  int64_t D, const void* G, int64_t ldg, void* CD, int64_t ldc, void* rs, void* cs,
                  void* grad_x1, int64_t ldg1, void* grad_x2, int64_t ldg2, int dtype, double eps,
                  void* stream);

/*
 * The other twelve fingerprint kernels of gauche/kernels/fingerprint_kernels/ as EPILOGUES of the tensor-core Gram kernel
 * (csrc/gram_siblings.cu): the exact integer dot products stay in TMEM, the similarity tile is written once.
 * `which`: 0 dice, 1 braun_blanquet, 2 faith, 3 forbes, 4 inner_product, 5 intersection, 6 otsuka, 7 rand,
 * 8 rogers_tanimoto, 9 russell_rao, 10 sogenfrei, 11 sokal_sneath.  operand_kind 0: u8 operands (kind::i8 CTA pairs),
 * 1: packed e2m1 bits; lda/ldb/k_bytes in bytes.  Float ops follow the reference's order in `compute_dtype`; the result
 * is cast to `out_dtype` (the reference's `.to(double)`: GK_F64 for all but dice / faith / inner_product) and clamped
 * at 0.  The reference's shape quirks arrive as data:
 *   row_a [n]  |x1_i|_1 — or |x1_i|_1 + |x2_i|_1 for rogers_tanimoto / sokal_sneath (x2_norm is added UNTRANSPOSED there)
 *   row_b [n]  d = common zeros of the LAST rows (faith / intersection / rand / rogers_tanimoto) or the Braun-Blanquet
 *              denominator max(x1_norm[-1], x2_norm[-1]), one value per output row; NULL = zeros
 *   col   [m]  |x2_j|_1
 * gk_sibling_gram_real: the same kernels on real-valued dense inputs (CUDA cores, tolerance-level parity); row_a / row_b /
 * col are then arrays of `dtype` (row_a already summed, not yet doubled); gk_rowsum_real gives the L1 norms.
 */
int gk_sibling_gram(int which, int operand_kind, const void* a, int64_t lda, const int32_t* row_a,
                    const int32_t* row_b, int64_t n, const void* b, int64_t ldb, const int32_t* col,
                    int64_t m, int64_t k_bytes, int64_t nfeat, int compute_dtype, void* out, int64_t ldo,
                    int out_dtype, double eps, void* stream);
int gk_rowsum_real(const void* x, int64_t ld, int64_t rows, int64_t D, int dtype, void* out, void* stream);
int gk_sibling_gram_real(int which, const void* x1, int64_t ld1, const void* row_a, const void* row_b, int64_t n,
                         const void* x2, int64_t ld2, const void* col, int64_t m, int64_t D, int64_t nfeat,
                         int dtype, void* out, int64_t ldo, int out_dtype, double eps, void* stream);

/*
 * Screening helpers (caller side of the cross-kernel, gauche/gp.py:169-226 protocol):
 *   gk_rowdot_f32 : scores[i] = bias + sum_j K[i,j]*alpha[j]   (posterior-mean GEMV)
 *   gk_topk_f32   : k largest scores with their indices (+index_base), descending;
 *                   ties broken towards the smaller index.  workspace: gk_topk_workspace().
 */
/*
 * Fused screening: scores[i] = bias + sum_j K[i,j]*alpha[j] with K = Tanimoto(a_i, b_j) computed on the
 * tensor cores and consumed in the epilogue — the Gram block is never written.  operand_kind 0: u8
 * operands (k_bytes = padded row length), 1: packed e2m1 bits (gk_expand_bits_e2m1), 2: packed bit words expanded
 * inside the SM (the wire format; default for bits).  lda/ldb/k_bytes are in BYTES for every kind.  `partials` is a
 * caller-owned workspace of gk_tanimoto_rowdot_slabs(m, kind) x ldp floats (ldp >= n, multiple of 4):
 * one partial sum per row and (column tile, epilogue role = a fixed subset of the tile's 32-column chunks); they are
 * reduced in a fixed order (deterministic, independent of the launch geometry).  Inside the sum each K[i,j] is
 * numerator * rcp.approx(denominator): within 2 ulp of the Gram kernels' correctly rounded value (use
 * gk_tanimoto_gram_* + gk_rowdot_f32 when the exact Gram entries matter).
 */
int64_t gk_tanimoto_rowdot_slabs(int64_t m, int operand_kind);
int gk_tanimoto_rowdot(const void* a, int64_t lda, const int32_t* a_sq, int64_t n,
                       const void* b, int64_t ldb, const int32_t* b_sq, int64_t m,
                       int64_t k_bytes, int operand_kind, const float* alpha, double eps, float bias,
                       float* partials, int64_t ldp, float* scores, void* stream);
/* gk_tanimoto_rowdot on the hybrid operands of gk_tanimoto_gram_hybrid; workspace gk_tanimoto_rowdot_slabs(m, 3) x ldp. */
int gk_tanimoto_rowdot_hybrid(const uint32_t* a_bitsp, int64_t lda, const int32_t* a_sq, int64_t n,
                              const uint8_t* b4, int64_t ldb, const int32_t* b_sq, int64_t m,
                              int64_t words, int64_t k_bytes, const float* alpha, double eps, float bias,
                              float* partials, int64_t ldp, float* scores, void* stream);
int gk_rowdot_f32(const float* K, int64_t ldk, int64_t n, int64_t m, const float* alpha,
                  float bias, float* scores, void* stream);
int64_t gk_topk_workspace(int64_t n, int k);
int gk_topk_f32(const float* scores, int64_t n, int k, int64_t index_base, float* out_vals,
                int64_t* out_idx, void* workspace, int64_t workspace_bytes, void* stream);

/*
 * Self test of the epilogue's branch-free fp32 division (csrc/common.cuh fdiv_rn_normal):
 * out_fast[i] = fdiv_rn_normal(a[i], b[i]), out_ieee[i] = div.rn.f32(a[i], b[i]).
 * The two must be bit-identical on the epilogue's operand domain (tests/test_gpu_parity.py).
 */
int gk_selftest_fdiv_f32(const float* a, const float* b, float* out_fast, float* out_ieee,
                         int64_t n, void* stream);
int gk_selftest_fdiv_f64(const double* a, const double* b, double* out_fast, double* out_ieee,
                         int64_t n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GAUCHE_B200_H */
