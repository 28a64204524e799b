// umma3.cu — A-stationary FP4 Tanimoto Gram kernel (tcgen05 kind::mxf4, single CTA per SM).
//
// Why: the second-generation kernels (umma2.cu) are bound by the SM<->L2 port — operand loads and
// Gram stores together top out near 52 B/clk/SM (profiles/r1_mxf4x1_ablation_modes.log).  A 128 x N
// tile re-reads its 128-row A slab (128 KB of e2m1 for 2048 bits) for every N columns.  Here the
// A slab of a row block is loaded ONCE into shared memory and stays there while the CTA sweeps a
// long run of 128-column B tiles, so per 128 x 128 tile only the 128 KB B slab and the 64 KB of
// output cross the port (11.7 B/pair instead of 16.6).
//
//   smem   A slab  k_blocks x 16 KB (<= 128 KB, resident per work unit)
//          B ring  4 x 16 KB (TMA, 128B swizzle), epilogue staging 8 x 4 KB, column terms 2 x 1 KB
//   TMEM   3 accumulator stages x 128 columns + constant UE8M0 scale factors in columns [448,512)
//   warps  0 TMA producer | 1 MMA issuer | 2-9 epilogue (lane quadrant x 64-column half)
//   work   unit = (row block, segment of consecutive column tiles); units are ordered with the row
//          block fastest so the CTAs that run together stream the same B tiles out of L2.
//
// Same contract / epilogue / exactness as umma2.cu KindMxf4 (reference tanimoto_kernel.py:36-46).
#include "tc_common.cuh"

namespace gk {
namespace umma3 {
using namespace gk::tc;

constexpr int BN = 128;
constexpr int STAGES = 4;                 // B ring
constexpr int ACC_STAGES = 3;
constexpr int TMEM_COLS = 512;
constexpr int EPI_WARPS = 8;
constexpr int EPI_THREADS = EPI_WARPS * 32;
constexpr int THREADS = 64 + EPI_THREADS;
constexpr int EPI_COLS = BN / 2;          // 64 columns per epilogue warp
constexpr int EPI_BUF_BYTES = 32 * 128;
constexpr int TILE_BYTES = BM * BK;       // 16 KB: one k-block of A or B
constexpr int MAX_KB = 8;                 // A slab: up to 8 k-blocks = 2048 bits
constexpr int COLTERM_BYTES = 2 * BN * 4; // negated norms + alpha, per buffer

constexpr int OFF_A = 0;
constexpr int OFF_B = OFF_A + MAX_KB * TILE_BYTES;              // 131072
constexpr int OFF_EPI = OFF_B + STAGES * TILE_BYTES;            // 196608
constexpr int OFF_COL = OFF_EPI + EPI_WARPS * EPI_BUF_BYTES;    // 229376
constexpr int OFF_BAR = OFF_COL + 2 * COLTERM_BYTES;            // 231424
constexpr int NUM_BARS = 2 + 2 * STAGES + 2 * ACC_STAGES;       // 16
constexpr int OFF_TMEMPTR = OFF_BAR + NUM_BARS * 8;
constexpr int SMEM_BYTES = OFF_TMEMPTR + 16;
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget exceeded");

__host__ __device__ constexpr uint32_t idesc_mxf4() {
  return (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | (1u << 23) | ((uint32_t)(BM >> 4) << 24);
}

struct Work {
  int m_blocks, n_tiles, segs, tiles_per_seg;
  uint32_t units;
};
__host__ __device__ inline Work make_work(int64_t n, int64_t m, int ctas) {
  Work w;
  w.m_blocks = (int)((n + BM - 1) / BM);
  w.n_tiles = (int)((m + BN - 1) / BN);
  // ~8 units per CTA for balance, but segments of at least 8 tiles so the A reload stays amortised
  int segs = (ctas * 8 + w.m_blocks - 1) / w.m_blocks;
  const int max_segs = (w.n_tiles + 7) / 8;
  if (segs > max_segs) segs = max_segs;
  if (segs < 1) segs = 1;
  w.tiles_per_seg = (w.n_tiles + segs - 1) / segs;
  w.segs = (w.n_tiles + w.tiles_per_seg - 1) / w.tiles_per_seg;
  w.units = (uint32_t)w.m_blocks * (uint32_t)w.segs;
  return w;
}

template <typename Epi>
__global__ void __launch_bounds__(THREADS, 1)
tanimoto_umma3_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                      const __grid_constant__ CUtensorMap map_out, const int32_t* __restrict__ na,
                      const int32_t* __restrict__ nb, int64_t n, int64_t m, int k_blocks, Work work, Epi epi) {
  using OutT = typename Epi::OutT;
  using TermT = typename Epi::TermT;
  static_assert(sizeof(OutT) == 4 && sizeof(TermT) == 4, "umma3 serves the fp32 / int32 epilogues");
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const uint32_t bar_afull = sbase + OFF_BAR;
  const uint32_t bar_aempty = bar_afull + 8;
  const uint32_t bar_full = bar_aempty + 8;                  // [STAGES]
  const uint32_t bar_empty = bar_full + STAGES * 8;          // [STAGES]
  const uint32_t bar_tfull = bar_empty + STAGES * 8;         // [ACC_STAGES]
  const uint32_t bar_tempty = bar_tfull + ACC_STAGES * 8;    // [ACC_STAGES]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + OFF_TMEMPTR);

  if ((sbase & 1023u) != 0) {   // 128B-swizzle atoms need a 1024-byte aligned base; no slack is budgeted
    if (threadIdx.x == 0) printf("gauche_b200 umma3: dynamic shared memory base 0x%x is not 1024-byte aligned\n", sbase);
    __trap();
  }
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&map_a);
    prefetch_tmap(&map_b);
    prefetch_tmap(&map_out);
    mbar_init(bar_afull, 1);
    mbar_init(bar_aempty, 1);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int a = 0; a < ACC_STAGES; ++a) {
      mbar_init(bar_tfull + 8 * a, 1);
      mbar_init(bar_tempty + 8 * a, EPI_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<1>(smem_u32(tmem_ptr_smem), TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  if (warp >= 2 && warp < 6) {   // constant unit scale factors (see umma2.cu KindMxf4)
    const uint32_t lane_base = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    tmem_fill32(lane_base + SF_COL_A, 0x7F7F7F7Fu);
    tmem_fill32(lane_base + SF_COL_B, 0x7F7F7F7Fu);
    tmem_wait_st();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  // unit u -> row block (fastest) and column segment
  auto unit_rows = [&](uint32_t u) { return (int)(u % (uint32_t)work.m_blocks); };
  auto unit_seg = [&](uint32_t u) { return (int)(u / (uint32_t)work.m_blocks); };
  auto seg_begin = [&](int seg) { return seg * work.tiles_per_seg; };
  auto seg_end = [&](int seg) { return min(work.n_tiles, (seg + 1) * work.tiles_per_seg); };

  if (warp == 0) {
    // ================================ TMA producer ================================
    int stage = 0;
    uint32_t phase = 0, aphase = 0;
    for (uint32_t u = blockIdx.x; u < work.units; u += gridDim.x) {
      const int mb = unit_rows(u), seg = unit_seg(u);
      mbar_wait(bar_aempty, aphase ^ 1);          // every MMA of the previous unit has retired
      if (lane == 0) {
        mbar_expect_tx(bar_afull, (uint32_t)k_blocks * TILE_BYTES);
        for (int kb = 0; kb < k_blocks; ++kb)
          tma_load_2d<1>(sbase + OFF_A + kb * TILE_BYTES, &map_a, bar_afull, kb * BK, mb * BM);
      }
      __syncwarp();
      aphase ^= 1;
      for (int nt = seg_begin(seg); nt < seg_end(seg); ++nt) {
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(bar_empty + 8 * stage, phase ^ 1);
          if (lane == 0) {
            mbar_expect_tx(bar_full + 8 * stage, TILE_BYTES);
            tma_load_2d<1>(sbase + OFF_B + stage * TILE_BYTES, &map_b, bar_full + 8 * stage, kb * BK, nt * BN);
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================================= MMA issuer =================================
    int stage = 0, acc = 0;
    uint32_t phase = 0, acc_phase = 0, aphase = 0;
    constexpr uint32_t idesc = idesc_mxf4();
    for (uint32_t u = blockIdx.x; u < work.units; u += gridDim.x) {
      const int seg = unit_seg(u);
      mbar_wait(bar_afull, aphase);               // this unit's A slab has landed
      aphase ^= 1;
      tc_fence_after();
      for (int nt = seg_begin(seg); nt < seg_end(seg); ++nt) {
        mbar_wait(bar_tempty + 8 * acc, acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        const bool last_tile = (nt == seg_end(seg) - 1);
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(bar_full + 8 * stage, phase);
          tc_fence_after();
          if (lane == 0) {
            const uint64_t da = make_smem_desc(sbase + OFF_A + kb * TILE_BYTES);
            const uint64_t db = make_smem_desc(sbase + OFF_B + stage * TILE_BYTES);
#pragma unroll
            for (int kk = 0; kk < BK / UK; ++kk)
              mma_mxf4<1>(d_tmem, da + (uint64_t)(kk * (UK >> 4)), db + (uint64_t)(kk * (UK >> 4)), idesc,
                          (uint32_t)((kb | kk) != 0), tmem_base + SF_COL_A, tmem_base + SF_COL_B);
            mma_commit<1>(bar_empty + 8 * stage);
            if (kb == k_blocks - 1) {
              mma_commit<1>(bar_tfull + 8 * acc);
              if (last_tile) mma_commit<1>(bar_aempty);   // A slab may be overwritten once these retire
            }
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ================================== epilogue ==================================
    constexpr int CS = 32;                        // columns per step (fp32: 128 bytes per row)
    constexpr int STEPS = EPI_COLS / CS;          // 2
    const int ew = warp - 2;
    const int quad = warp & 3;
    const int chalf = ew >> 2;
    const int et = threadIdx.x - 64;
    uint8_t* sb = smem + OFF_EPI + ew * EPI_BUF_BYTES;
    const uint32_t sb_u32 = sbase + OFF_EPI + ew * EPI_BUF_BYTES;
    int acc = 0;
    uint32_t acc_phase = 0, tile_par = 0;

    // norms (and alpha) of the next tile are fetched one tile ahead
    int next_col = 0, next_row = 0;
    float next_alpha = 0.f;
    auto fetch_col = [&](int nt, bool valid) {
      next_col = 0;
      next_alpha = 0.f;
      const int64_t gc = (int64_t)nt * BN + et;
      if (valid && et < BN && gc < m) {
        next_col = __ldg(nb + gc);
        if constexpr (Epi::kRowdot) next_alpha = __ldg(epi.alpha + gc);
      }
    };
    for (uint32_t u = blockIdx.x; u < work.units; u += gridDim.x) {
      const int mb = unit_rows(u), seg = unit_seg(u);
      const int64_t row0 = (int64_t)mb * BM + quad * 32;
      const int64_t my_row = row0 + lane;
      next_row = my_row < n ? __ldg(na + my_row) : 0;
      const TermT rterm = epi.row_term(next_row);
      fetch_col(seg_begin(seg), true);
      for (int nt = seg_begin(seg); nt < seg_end(seg); ++nt) {
        const int64_t col0 = (int64_t)nt * BN;
        TermT* col_terms = reinterpret_cast<TermT*>(smem + OFF_COL + tile_par * COLTERM_BYTES);
        if (et < BN) {
          col_terms[et] = epi.col_term(next_col);
          if constexpr (Epi::kRowdot) reinterpret_cast<float*>(col_terms)[BN + et] = next_alpha;
        }
        fetch_col(nt + 1, nt + 1 < seg_end(seg));
        float rowacc = 0.f;
        asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");

        mbar_wait(bar_tfull + 8 * acc, acc_phase);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BN + chalf * EPI_COLS);
        const bool rows_live = row0 < n;
#pragma unroll 1
        for (int c = 0; c < STEPS; ++c) {
          const int cl = chalf * EPI_COLS + c * CS;
          const bool live = rows_live && (col0 + cl < m);
          uint32_t v[CS];
          if (live) {
            tmem_ld32(taddr + c * CS, v);
            tmem_wait_ld();
          }
          if (c == STEPS - 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_tempty + 8 * acc);
          }
          if (!live) continue;
          const float2 nr2 = make_float2(rterm, rterm);
          if constexpr (Epi::kRowdot) {
            const float* alpha_s = reinterpret_cast<const float*>(col_terms) + BN;
            float4 nc[8], al[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              nc[q] = *reinterpret_cast<const float4*>(&col_terms[cl + q * 4]);
              al[q] = *reinterpret_cast<const float4*>(&alpha_s[cl + q * 4]);
            }
            float2 part = make_float2(0.f, 0.f);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const float2 k0 = epi.template pair<true>(v[q * 4 + 0], v[q * 4 + 1], nr2, make_float2(nc[q].x, nc[q].y));
              const float2 k1 = epi.template pair<true>(v[q * 4 + 2], v[q * 4 + 3], nr2, make_float2(nc[q].z, nc[q].w));
              part = __ffma2_rn(k0, make_float2(al[q].x, al[q].y), part);
              part = __ffma2_rn(k1, make_float2(al[q].z, al[q].w), part);
            }
            rowacc += part.x + part.y;
          } else {
            uint4 res[8];
            if constexpr (Epi::kPacked) {
              float4 nc[8];
#pragma unroll
              for (int q = 0; q < 8; ++q) nc[q] = *reinterpret_cast<const float4*>(&col_terms[cl + q * 4]);
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                const float2 p0 = epi.template pair<true>(v[q * 4 + 0], v[q * 4 + 1], nr2, make_float2(nc[q].x, nc[q].y));
                const float2 p1 = epi.template pair<true>(v[q * 4 + 2], v[q * 4 + 3], nr2, make_float2(nc[q].z, nc[q].w));
                res[q] = make_uint4(__float_as_uint(p0.x), __float_as_uint(p0.y), __float_as_uint(p1.x), __float_as_uint(p1.y));
              }
            } else {
              TermT ct[CS];
#pragma unroll
              for (int j = 0; j < CS; ++j) ct[j] = col_terms[cl + j];
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                OutT t4[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) t4[e] = epi.template one<true>(v[q * 4 + e], rterm, ct[q * 4 + e]);
                res[q] = *reinterpret_cast<const uint4*>(t4);
              }
            }
            // single staging buffer: the previous TMA store must have finished reading it; the wait
            // overlaps with the math above
            if (lane == 0) tma_store_wait_read<0>();
            __syncwarp();
#pragma unroll
            for (int q = 0; q < 8; ++q)
              *reinterpret_cast<uint4*>(sb + lane * 128 + ((q ^ (lane & 7)) << 4)) = res[q];
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              tma_store_2d(&map_out, sb_u32, (int)(col0 + cl), (int)row0);
              tma_store_commit();
            }
          }
        }
        if constexpr (Epi::kRowdot) {
          if (my_row < n) epi.partials[(int64_t)(nt * 2 + chalf) * epi.ldp + my_row] = rowacc;
        }
        if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
        tile_par ^= 1;
      }
    }
    if (lane == 0) tma_store_wait_read<0>();
    __syncwarp();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<1>(tmem_base, TMEM_COLS);
  }
}

template <typename Epi>
static int launch(const uint8_t* a, int64_t lda, const int32_t* na, int64_t n, const uint8_t* b, int64_t ldb,
                  const int32_t* nb, int64_t m, int64_t k_bytes, void* out_base, int64_t out_inner,
                  int64_t out_rows, int64_t out_pitch_bytes, CUtensorMapDataType out_dt, Epi epi,
                  cudaStream_t stream) {
  CUtensorMap ma, mb, mo;
  if (int e = make_map(&ma, CU_TENSOR_MAP_DATA_TYPE_UINT8, a, k_bytes, n, lda, BK, BM)) return e;
  if (int e = make_map(&mb, CU_TENSOR_MAP_DATA_TYPE_UINT8, b, k_bytes, m, ldb, BK, BN)) return e;
  if (int e = make_map(&mo, out_dt, out_base, out_inner, out_rows, out_pitch_bytes, 32, 32)) return e;
  auto kern = tanimoto_umma3_kernel<Epi>;
  static bool attr_done[64] = {false};
  int dev = 0;
  GK_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || !attr_done[dev]) {
    GK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    if (dev >= 0 && dev < 64) attr_done[dev] = true;
  }
  const int ctas = sm_count();
  const Work w = make_work(n, m, ctas);
  int64_t grid = ctas;
  if (grid > (int64_t)w.units) grid = w.units;
  kern<<<(unsigned)grid, THREADS, SMEM_BYTES, stream>>>(ma, mb, mo, na, nb, n, m, (int)(k_bytes / BK), w, epi);
  return cuda_err(cudaGetLastError(), "tanimoto_umma3_kernel launch");
}

static int check(const char* name, const uint8_t* a, int64_t lda, const int32_t* a_n, int64_t n, const uint8_t* b,
                 int64_t ldb, const int32_t* b_n, int64_t m, int64_t k_bytes) {
  GK_CHECK_ARG(n >= 0 && m >= 0 && k_bytes > 0, GK_EINVAL, "%s: bad size", name);
  GK_CHECK_ARG(a && b && a_n && b_n, GK_EINVAL, "%s: null pointer", name);
  GK_CHECK_ARG(k_bytes % BK == 0 && k_bytes / BK <= MAX_KB, GK_ERANGE,
               "%s: k_bytes (%lld) must be a multiple of %d and at most %d (A-stationary slab)", name,
               (long long)k_bytes, BK, MAX_KB * BK);
  GK_CHECK_ARG(lda >= k_bytes && ldb >= k_bytes, GK_EINVAL, "%s: leading dimension too small", name);
  GK_CHECK_ARG(((uintptr_t)a) % 16 == 0 && ((uintptr_t)b) % 16 == 0 && lda % 16 == 0 && ldb % 16 == 0, GK_EALIGN,
               "%s: operands must be 16-byte aligned with lda/ldb multiples of 16", name);
  GK_CHECK_ARG(n < (1LL << 31) && m < (1LL << 31), GK_ERANGE, "%s: size too large", name);
  int64_t cc = 0;
  if (int e = gk_query(GK_Q_CC, &cc)) return e;
  GK_CHECK_ARG(cc / 10 == 10, GK_EDEVICE, "%s: requires an sm_100 device (found sm_%lld)", name, (long long)cc);
  return 0;
}

}  // namespace umma3
}  // namespace gk

extern "C" int gk_tanimoto_gram_mxf4_as(const uint8_t* a4, int64_t lda, const int32_t* a_n, int64_t n,
                                        const uint8_t* b4, int64_t ldb, const int32_t* b_n, int64_t m,
                                        int64_t k_bytes, void* out, int64_t ldo, int out_dtype, double eps,
                                        void* stream) {
  using namespace gk;
  using namespace gk::umma3;
  const char* name = "gk_tanimoto_gram_mxf4_as";
  if (n == 0 || m == 0) return 0;
  if (int e = check(name, a4, lda, a_n, n, b4, ldb, b_n, m, k_bytes)) return e;
  GK_CHECK_ARG(out && ldo >= m, GK_EINVAL, "%s: bad output", name);
  GK_CHECK_ARG(((uintptr_t)out) % 16 == 0 && (ldo * 4) % 16 == 0, GK_EALIGN,
               "%s: out must be 16-byte aligned with a 16-byte multiple row pitch (TMA store)", name);
  cudaStream_t s = (cudaStream_t)stream;
  if (out_dtype == GK_F32) {
    if (eps >= 1e-12 && eps <= 1.0)
      return launch(a4, lda, a_n, n, b4, ldb, b_n, m, k_bytes, out, m, n, ldo * 4, CU_TENSOR_MAP_DATA_TYPE_FLOAT32,
                    TanimotoF32x2{(float)eps}, s);
    return launch(a4, lda, a_n, n, b4, ldb, b_n, m, k_bytes, out, m, n, ldo * 4, CU_TENSOR_MAP_DATA_TYPE_FLOAT32,
                  Elementwise<TanimotoEpilogue<float>, float>{TanimotoEpilogue<float>{(float)eps}}, s);
  }
  if (out_dtype == GK_I32)
    return launch(a4, lda, a_n, n, b4, ldb, b_n, m, k_bytes, out, m, n, ldo * 4, CU_TENSOR_MAP_DATA_TYPE_INT32,
                  Elementwise<RawEpilogue, int32_t>{RawEpilogue{}}, s);
  return set_err(GK_EINVAL, "%s: out_dtype must be GK_F32 or GK_I32 (use gk_tanimoto_gram_mxf4 for fp64)", name);
}

extern "C" int64_t gk_tanimoto_rowdot_as_slabs(int64_t m) { return 2 * ((m + gk::umma3::BN - 1) / gk::umma3::BN); }

extern "C" int gk_tanimoto_rowdot_mxf4_as(const uint8_t* a4, int64_t lda, const int32_t* a_n, int64_t n,
                                          const uint8_t* b4, int64_t ldb, const int32_t* b_n, int64_t m,
                                          int64_t k_bytes, const float* alpha, double eps, float* partials,
                                          int64_t ldp, void* stream) {
  using namespace gk;
  using namespace gk::umma3;
  const char* name = "gk_tanimoto_rowdot_mxf4_as";
  if (n == 0) return 0;
  GK_CHECK_ARG(m > 0, GK_EINVAL, "%s: bad size", name);
  if (int e = check(name, a4, lda, a_n, n, b4, ldb, b_n, m, k_bytes)) return e;
  GK_CHECK_ARG(alpha && partials, GK_EINVAL, "%s: null pointer", name);
  GK_CHECK_ARG(ldp >= n && ldp % 4 == 0 && ((uintptr_t)partials) % 16 == 0, GK_EALIGN,
               "%s: partials must be 16-byte aligned with ldp >= n and a multiple of 4", name);
  GK_CHECK_ARG(eps >= 1e-12 && eps <= 1.0, GK_ERANGE, "%s: eps outside [1e-12, 1]", name);
  TanimotoRowdotF32x2 epi;
  epi.eps = (float)eps;
  epi.alpha = alpha;
  epi.partials = partials;
  epi.ldp = ldp;
  return launch(a4, lda, a_n, n, b4, ldb, b_n, m, k_bytes, partials, ldp, gk_tanimoto_rowdot_as_slabs(m), ldp * 4,
                CU_TENSOR_MAP_DATA_TYPE_FLOAT32, epi, (cudaStream_t)stream);
}
