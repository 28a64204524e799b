// umma.cu — Tanimoto Gram / cross-kernel on the 5th-gen tensor cores (tcgen05, sm_100a).
//
// Replaces torch.matmul + 6 unfused elementwise passes of the reference
// (gauche/kernels/fingerprint_kernels/tanimoto_kernel.py:36-46) by ONE persistent,
// warp-specialised kernel:
//
//   warp 0     TMA producer : cp.async.bulk.tensor (128B swizzle) of u8 operand K-slabs
//                             A[128 x 128B], B[256 x 128B] into a 4-stage smem ring
//   warp 1     MMA issuer   : tcgen05.mma.cta_group::1.kind::i8, M=128 N=256 K=32, exact
//                             s32 accumulation in TMEM (2 x 256 columns, double buffered)
//   warps 2-5  epilogue     : tcgen05.ld -> reference-order float epilogue with the row
//                             norms fused -> swizzled smem transpose -> 128-byte coalesced
//                             streaming stores.  The Gram tile is written to HBM once.
//
// Operands are 0/1 bits or small counts stored as u8, K-major ("row-major [rows, k]"),
// k a multiple of 128; out-of-range rows are zero-filled by TMA.  Intersection counts
// are exact integers (k*255^2 < 2^31 is enforced by the packer).
#include "tc_common.cuh"

namespace gk {
namespace umma {
using namespace gk::tc;   // BM/BK/UK, inline-PTX wrappers, smem descriptor, tensor-map encoder

constexpr int BN = 256;                  // tile cols   (UMMA N)
constexpr int STAGES = 4;
constexpr int A_BYTES = BM * BK;         // 16 KB
constexpr int B_BYTES = BN * BK;         // 32 KB
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int ACC_STAGES = 2;
constexpr int TMEM_COLS = ACC_STAGES * BN;  // 512
constexpr int EPI_WARPS = 8;                  // 2 per TMEM lane quadrant, each owns 128 columns
constexpr int EPI_THREADS = EPI_WARPS * 32;   // 256
constexpr int THREADS = 64 + EPI_THREADS;     // 320
constexpr int EPI_COLS = BN / (EPI_WARPS / 4);  // columns per epilogue warp (128)
constexpr int EPI_STAGE_BYTES = 32 * 64;      // per warp transpose buffer: 32 rows x 64 B
constexpr int COLTERM_BYTES = BN * 8;         // one tile's column terms (up to 8 B each), x2 buffers
constexpr int GROUP_M = 16;                   // raster: m-blocks swept together for L2 reuse

// dynamic shared memory carve-up (offsets from the 1024-aligned base)
constexpr int OFF_RING = 0;
constexpr int OFF_EPI = OFF_RING + STAGES * STAGE_BYTES;
constexpr int OFF_COL = OFF_EPI + EPI_WARPS * EPI_STAGE_BYTES;
constexpr int OFF_BAR = OFF_COL + 2 * COLTERM_BYTES;
constexpr int NUM_BARS = 2 * STAGES + 2 * ACC_STAGES;
constexpr int OFF_TMEMPTR = OFF_BAR + NUM_BARS * 8;
constexpr int SMEM_BYTES = OFF_TMEMPTR + 16;
constexpr int SMEM_ALLOC = SMEM_BYTES + 1024;  // slack for manual 1024-byte alignment
static_assert(SMEM_ALLOC <= 227 * 1024, "shared memory budget exceeded");

// Instruction descriptor for kind::i8: c=S32 (2) at [4,6), a/b format U8 (0) at [7,10)/[10,13),
// both K-major, N>>3 at [17,23), M>>4 at [24,29).
constexpr uint32_t kInstrDesc = (2u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

struct TileCoord { int mb, nb; };
__device__ __forceinline__ TileCoord tile_coord(uint32_t t, int m_tiles, int n_tiles) {
  // groups of GROUP_M row-blocks are swept along n with m fastest, so the CTAs that run
  // together share GROUP_M A-slabs and a handful of B-slabs in L2.
  const uint32_t per_group = (uint32_t)GROUP_M * (uint32_t)n_tiles;
  const uint32_t g = t / per_group;
  const int first_m = (int)g * GROUP_M;
  const uint32_t gsz = (uint32_t)min(GROUP_M, m_tiles - first_m);
  const uint32_t r = t - g * per_group;
  TileCoord c;
  c.mb = first_m + (int)(r % gsz);
  c.nb = (int)(r / gsz);
  return c;
}

template <typename Epi, typename OutT>
__global__ void __launch_bounds__(THREADS, 1)
tanimoto_umma_kernel(const __grid_constant__ CUtensorMap map_a,
                     const __grid_constant__ CUtensorMap map_b, const int32_t* __restrict__ na,
                     const int32_t* __restrict__ nb, int64_t n, int64_t m, int k_blocks,
                     OutT* __restrict__ out, int64_t ldo, int vec_ok, Epi epi) {
  extern __shared__ uint8_t smem_raw[];
  // manual 1024-byte alignment (128B-swizzle atoms) by pointer arithmetic on the __shared__ array,
  // so the compiler still knows these are shared-memory addresses (LDS/STS, not generic LD/ST)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t sbase = smem_u32(smem);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const uint32_t bar_full = sbase + OFF_BAR;                   // [STAGES]
  const uint32_t bar_empty = bar_full + STAGES * 8;            // [STAGES]
  const uint32_t bar_tfull = bar_empty + STAGES * 8;           // [ACC_STAGES]
  const uint32_t bar_tempty = bar_tfull + ACC_STAGES * 8;      // [ACC_STAGES]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + OFF_TMEMPTR);

  const int m_tiles = (int)((n + BM - 1) / BM);
  const int n_tiles = (int)((m + BN - 1) / BN);
  const uint32_t total_tiles = (uint32_t)m_tiles * (uint32_t)n_tiles;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&map_a);
    prefetch_tmap(&map_b);
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

  if (warp == 0) {
    // ================================ TMA producer ================================
    int stage = 0;
    uint32_t phase = 0;
    for (uint32_t t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      const TileCoord tc = tile_coord(t, m_tiles, n_tiles);
      for (int kb = 0; kb < k_blocks; ++kb) {
        mbar_wait(bar_empty + 8 * stage, phase ^ 1);
        {   // converged warp, leader elected inside the asm blocks (tc_common.cuh, namespace w)
          const uint32_t sa = sbase + OFF_RING + stage * STAGE_BYTES;
          const uint32_t sb = sa + A_BYTES;
          w::mbar_expect_tx(bar_full + 8 * stage, STAGE_BYTES);
          w::tma_load_2d<1>(sa, &map_a, bar_full + 8 * stage, kb * BK, tc.mb * BM);
          w::tma_load_2d<1>(sb, &map_b, bar_full + 8 * stage, kb * BK, tc.nb * BN);
        }
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ================================= MMA issuer =================================
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (uint32_t t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      mbar_wait(bar_tempty + 8 * acc, acc_phase ^ 1);  // epilogue drained this accumulator
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
      for (int kb = 0; kb < k_blocks; ++kb) {
        mbar_wait(bar_full + 8 * stage, phase);         // TMA bytes have landed
        tc_fence_after();
        {
          const uint32_t sa = sbase + OFF_RING + stage * STAGE_BYTES;
          const uint64_t da = make_smem_desc(sa);
          const uint64_t db = make_smem_desc(sa + A_BYTES);
#pragma unroll
          for (int kk = 0; kk < BK / UK; ++kk) {
            // advance 32 bytes inside the swizzle atom: +2 in the (addr >> 4) field
            w::mma_i8<1>(d_tmem, da + (uint64_t)(kk * (UK >> 4)), db + (uint64_t)(kk * (UK >> 4)),
                         kInstrDesc, (uint32_t)((kb | kk) != 0));
          }
          w::mma_commit<1>(bar_empty + 8 * stage);            // frees the smem slot when MMAs retire
          if (kb == k_blocks - 1) w::mma_commit<1>(bar_tfull + 8 * acc);  // accumulator ready
        }
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
    }
  } else {
    // ================================== epilogue ==================================
    // 8 warps: warp w may only touch TMEM lanes [32*(w%4), +32); warps w and w+4 split the 256
    // accumulator columns of their 32 rows.  Per 16-byte-chunked step each lane converts its row
    // segment, the warp transposes it through a private swizzled smem buffer and stores full
    // 64-byte row segments with streaming (evict-first) stores.
    using TermT = decltype(epi.row_term(0));
    constexpr int EPC = 16 / (int)sizeof(OutT);   // elements per 16-byte chunk
    constexpr int CS = 4 * EPC;                   // columns per transpose step (64 bytes per row)
    constexpr int LD_COLS = 32;                   // columns per tcgen05.ld
    const int ew = warp - 2;                      // 0..7
    const int quad = warp & 3;                    // TMEM lane quadrant this warp may access
    const int chalf = ew >> 2;                    // which 128-column half
    uint8_t* stage_buf = smem + OFF_EPI + ew * EPI_STAGE_BYTES;
    const int et = threadIdx.x - 64;              // 0..255 among the epilogue threads
    int acc = 0;
    uint32_t acc_phase = 0;
    uint32_t tile_par = 0;
    for (uint32_t t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      const TileCoord tc = tile_coord(t, m_tiles, n_tiles);
      const int64_t row0 = (int64_t)tc.mb * BM + quad * 32;
      const int64_t col0 = (int64_t)tc.nb * BN;
      // column terms of this tile: one per epilogue thread, double buffered by tile parity; the
      // named barrier (id 1, epilogue threads only) also orders reuse two tiles later
      TermT* col_terms = reinterpret_cast<TermT*>(smem + OFF_COL + tile_par * COLTERM_BYTES);
      {
        const int64_t gc = col0 + et;
        col_terms[et] = epi.col_term(gc < m ? __ldg(nb + gc) : 0);
      }
      const int64_t my_row = row0 + lane;
      const TermT rterm = epi.row_term(my_row < n ? __ldg(na + my_row) : 0);
      asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");

      mbar_wait(bar_tfull + 8 * acc, acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BN + chalf * EPI_COLS);
      const bool rows_live = row0 < n;
#pragma unroll 1
      for (int c = 0; c < EPI_COLS / LD_COLS; ++c) {
        const int cl = chalf * EPI_COLS + c * LD_COLS;       // first tile-local column of this load
        if (!rows_live || col0 + cl >= m) break;             // warp-uniform
        uint32_t v[LD_COLS];
        tmem_ld32(taddr + c * LD_COLS, v);
        tmem_wait_ld();
#pragma unroll
        for (int h = 0; h < LD_COLS / CS; ++h) {
          const int64_t cbase = col0 + cl + h * CS;
          // convert this lane's CS columns (independent chains -> ILP) and stage them swizzled
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            OutT tmp[EPC];
#pragma unroll
            for (int e = 0; e < EPC; ++e) {
              const int j = h * CS + q * EPC + e;
              tmp[e] = (OutT)epi((int)v[j], rterm, col_terms[cl + j]);
            }
            *reinterpret_cast<uint4*>(stage_buf + lane * 64 + ((q ^ ((lane >> 1) & 3)) << 4)) =
                *reinterpret_cast<const uint4*>(tmp);
          }
          __syncwarp();
          // transposed read-back: 4 lanes cover one 64-byte output row segment, 8 rows per pass
#pragma unroll
          for (int it = 0; it < 4; ++it) {
            const int r = it * 8 + (lane >> 2);
            const int q = lane & 3;
            const int64_t gr = row0 + r;
            const int64_t gc = cbase + q * EPC;
            if (gr < n && gc < m) {
              const uint4 val =
                  *reinterpret_cast<const uint4*>(stage_buf + r * 64 + ((q ^ ((r >> 1) & 3)) << 4));
              OutT* dst = out + gr * ldo + gc;
              if (vec_ok && gc + EPC <= m) {
                __stcs(reinterpret_cast<uint4*>(dst), val);
              } else {
                const OutT* pv = reinterpret_cast<const OutT*>(&val);
#pragma unroll
                for (int e = 0; e < EPC; ++e)
                  if (gc + e < m) dst[e] = pv[e];
              }
            }
          }
          __syncwarp();
        }
      }
      // all TMEM reads of this accumulator are complete -> hand it back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_tempty + 8 * acc);
      if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
      tile_par ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<1>(tmem_base, TMEM_COLS);
  }
}

// --------------------------------------------------------------------------------- host
template <typename Epi, typename OutT>
static int launch(const CUtensorMap& ma, const CUtensorMap& mb, const int32_t* na, const int32_t* nb,
                  int64_t n, int64_t m, int64_t k, void* out, int64_t ldo, Epi epi,
                  cudaStream_t stream) {
  auto kern = tanimoto_umma_kernel<Epi, OutT>;
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [&] {
    attr_err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_ALLOC);
  });
  if (attr_err != cudaSuccess) return cuda_err(attr_err, "cudaFuncSetAttribute(umma smem)");
  const int64_t tiles = ((n + BM - 1) / BM) * ((m + BN - 1) / BN);
  int64_t grid = sm_count();
  if (grid > tiles) grid = tiles;
  const int vec_ok = (((uintptr_t)out) % 16 == 0) && ((ldo * (int64_t)sizeof(OutT)) % 16 == 0);
  kern<<<(unsigned)grid, THREADS, SMEM_ALLOC, stream>>>(ma, mb, na, nb, n, m, (int)(k / BK),
                                                        (OutT*)out, ldo, vec_ok, epi);
  return cuda_err(cudaGetLastError(), "tanimoto_umma_kernel launch");
}

}  // namespace umma
}  // namespace gk

extern "C" int gk_tanimoto_gram_umma(const uint8_t* a, int64_t lda, const int32_t* a_sq, int64_t n,
                                     const uint8_t* b, int64_t ldb, const int32_t* b_sq, int64_t m,
                                     int64_t k, void* out, int64_t ldo, int out_dtype, double eps,
                                     void* stream) {
  using namespace gk;
  using namespace gk::umma;
  const char* name = "gk_tanimoto_gram_umma";
  GK_CHECK_ARG(n >= 0 && m >= 0 && k > 0, GK_EINVAL, "%s: bad size", name);
  if (n == 0 || m == 0) return 0;
  GK_CHECK_ARG(a && b && a_sq && b_sq && out, GK_EINVAL, "%s: null pointer", name);
  GK_CHECK_ARG(k % BK == 0, GK_EALIGN, "%s: k (%lld) must be a multiple of %d", name, (long long)k, BK);
  GK_CHECK_ARG(lda >= k && ldb >= k && ldo >= m, GK_EINVAL, "%s: leading dimension too small", name);
  GK_CHECK_ARG(((uintptr_t)a) % 16 == 0 && ((uintptr_t)b) % 16 == 0 && lda % 16 == 0 && ldb % 16 == 0,
               GK_EALIGN, "%s: operands must be 16-byte aligned with lda/ldb multiples of 16", name);
  GK_CHECK_ARG(n < (1LL << 31) && m < (1LL << 31) && k < (1LL << 31), GK_ERANGE, "%s: size too large", name);
  GK_CHECK_ARG(((n + BM - 1) / BM) * ((m + BN - 1) / BN) < (1LL << 31), GK_ERANGE,
               "%s: more than 2^31 output tiles; split the row range", name);
  int64_t cc = 0;
  if (int e = gk_query(GK_Q_CC, &cc)) return e;
  GK_CHECK_ARG(cc / 10 == 10, GK_EDEVICE, "%s: requires an sm_100 device (found sm_%lld)", name,
               (long long)cc);
  CUtensorMap ma, mb;
  if (int e = make_map(&ma, CU_TENSOR_MAP_DATA_TYPE_UINT8, a, k, n, lda, BK, BM)) return e;
  if (int e = make_map(&mb, CU_TENSOR_MAP_DATA_TYPE_UINT8, b, k, m, ldb, BK, BN)) return e;
  cudaStream_t s = (cudaStream_t)stream;
  switch (out_dtype) {
    case GK_F32:
      // default eps (1e-6) keeps num/den normal and the quotient far from the denormal range:
      // branch-free exact division, no clamp (see common.cuh).  Exotic eps -> generic IEEE path.
      if (eps >= 1e-12 && eps <= 1.0)
        return launch<TanimotoEpilogueInt<float>, float>(ma, mb, a_sq, b_sq, n, m, k, out, ldo,
                                                         TanimotoEpilogueInt<float>{{(float)eps}}, s);
      return launch<TanimotoEpilogue<float>, float>(ma, mb, a_sq, b_sq, n, m, k, out, ldo,
                                                    TanimotoEpilogue<float>{(float)eps}, s);
    case GK_F64:
      if (eps >= 1e-12 && eps <= 1.0)
        return launch<TanimotoEpilogueInt<double>, double>(ma, mb, a_sq, b_sq, n, m, k, out, ldo,
                                                           TanimotoEpilogueInt<double>{{eps}}, s);
      return launch<TanimotoEpilogue<double>, double>(ma, mb, a_sq, b_sq, n, m, k, out, ldo,
                                                      TanimotoEpilogue<double>{eps}, s);
    case GK_I32:
      return launch<RawEpilogue, int32_t>(ma, mb, a_sq, b_sq, n, m, k, out, ldo, RawEpilogue{}, s);
    default:
      return set_err(GK_EINVAL, "%s: unsupported out_dtype %d", name, out_dtype);
  }
}
