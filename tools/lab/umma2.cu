// umma2.cu — the tcgen05 Gram kernels of the path (Tanimoto Gram, fused K·alpha screening, MinMax on
// thermometer planes).  Contract: gauche/kernels/fingerprint_kernels/tanimoto_kernel.py:36-46 (and
// minmax_kernel.py:33-44 through the thermometer identity), results bit-identical to the oracle.
//
// One persistent, warp-specialised kernel template:
//   warp 0      TMA producer: K-major 128B-swizzled operand boxes into a shared-memory ring (as deep as fits)
//   warp 1      MMA issuer: tcgen05.mma SS-mode, accumulators double-buffered in TMEM
//   warps 2..   epilogue: tcgen05.ld -> reference-order float math (packed f32x2, branch-free exact division)
//               -> swizzled staging -> TMA store; or the K·alpha row sums of the fused screening path
//   all TMA / tcgen05 instructions are issued from converged warps with the leader elected inside the asm block
//   (tc::w::), see tc_common.cuh.
//
// Template parameters (defaults are the shipped configuration; the rest are measured experiments)
//   kCG        1 single-CTA tiles | 2 cta_group::2 pairs (256-row tiles, each CTA stages half of B)
//   Kind       KindI8: u8 operands, kind::i8, s32 accumulators, N = 256 (bits or counts)
//              KindMxf4T<BN>: 0/1 bits as packed e2m1 nibbles (1.0 = 0b0010), kind::mxf4.block_scale with all
//              UE8M0 scale factors = 2^0 (TMEM columns [448,512)), exact f32 accumulators, N = 224 (or 192):
//              half the operand bytes and twice the MMA rate of kind::i8
//   Epi        TanimotoF32x2 (packed fp32 Gram) | TanimotoRowdotF32x2 (fused screening) | Elementwise<...>
//              (fp64 / raw counts / MinMax / exotic eps: element-wise epilogue with the same structure)
//   kRB        2 = two accumulator row blocks per CTA (no MMA/epilogue overlap)        [experiment]
//   kDirect    registers -> global 8-byte stores, no staging                          [experiment]
//   kMC        1 = two-CTA clusters share B through TMA multicast                     [experiment]
//   kEW        8 | 16 epilogue warps (16: 16-column steps, 64-byte staging rows)       [experiment]
//   kRowStore  every warp owns 16 rows, one 1-D bulk copy per staged 512-byte row      [experiment]
// What bounds it (profiles/r1_epilogue_ablation.log, r1_store_stream_ceiling.log): operand reads and Gram writes
// serialise chip-wide; the Gram kernel sits on that empirical bound, the fused path at 0.93 of loads + MMA.
#include "tc_common.cuh"

namespace gk {
namespace umma2 {
using namespace gk::tc;

constexpr int TMEM_COLS = 512;           // whole TMEM: 2 accumulator stages (+ scale factors for mxf4)
// kEW = epilogue warps per CTA: 8 (two per TMEM lane quadrant, 32-column steps, 128-byte staging rows) or
// 16 (four per quadrant, 16-column steps dealt round-robin, 64-byte staging rows).  The epilogue is latency
// bound (TMEM load -> 10-deep float chains -> staging -> proxy fence -> TMA store, IPC ~0.15 per warp,
// profiles/r1_ncu_source_stalls_mxf4x1.txt), so twice the warps hide twice the latency.
constexpr int EPI_COLS = 128;                 // kEW = 8: columns owned by the first epilogue warp of a quadrant
constexpr int epi_row_bytes(int kEW) { return kEW == 16 ? 64 : 128; }
constexpr int epi_buf_bytes(int kEW) { return 32 * epi_row_bytes(kEW); }   // 32 rows per staging buffer
constexpr int COLTERM_BYTES = 256 * 8;
constexpr int GROUP_M = 32;   // measured best of {4..128} on 16384 x 100k (profiles/r1_raster_group_m.log)

struct KindI8 {
  static constexpr int BN = 256;
  static constexpr bool kBlockScaled = false;
  static constexpr bool kF32Acc = false;
  template <int kCG>
  __host__ __device__ static constexpr uint32_t idesc() {
    // c=S32 at [4,6); a/b = U8, K-major; N>>3 at [17,23); M>>4 at [24,29) with M = 128*kCG
    return (2u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((BM * kCG) >> 4) << 24);
  }
};
template <int kBN>
struct KindMxf4T {
  static_assert(kBN % 32 == 0 && kBN >= 128 && kBN <= 224, "two accumulator stages + scale factors must fit 512 TMEM columns");
  static constexpr int BN = kBN;
  static constexpr bool kBlockScaled = true;
  static constexpr bool kF32Acc = true;
  template <int kCG>
  __host__ __device__ static constexpr uint32_t idesc() {
    // block-scaled descriptor: a/b format E2M1 (1) at [7,10)/[10,13), K-major, N>>3 at [17,23),
    // scale format UE8M0 (1) at [23], M>>4 at [24,29), sf ids 0, dense K = 64
    return (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | (1u << 23) |
           ((uint32_t)((BM * kCG) >> 4) << 24);
  }
};
using KindMxf4 = KindMxf4T<224>;      // widest tile that leaves room for the scale factors
using KindMxf4N192 = KindMxf4T<192>;  // 768-byte row segments: 256-byte aligned Gram stores (see the row-store epilogue)

// kRB = accumulator row blocks per CTA (single-CTA only).  kRB = 2 computes a 256 x BN tile per CTA: both
// 128-row A slabs share every B k-block, which cuts the operand bytes per pair by a third; the two
// accumulators fill TMEM, so there is a single accumulator stage and the epilogue is not overlapped with
// the next tile's MMAs (the operand ring keeps prefetching meanwhile).
// kMC = 1 (single-CTA MMAs only): clusters of two CTAs work on vertically adjacent row blocks of the SAME column
// tile; each CTA loads half of the B rows and TMA-multicasts them to both, so every B line leaves L2 once per
// cluster.  A ring stage may only be refilled after BOTH CTAs consumed it (empty barriers count 2, commits are
// multicast).
constexpr int ROWSTORE_PITCH = 544;   // staged row of the row-store epilogue: 512 B + 32 B pad (pitch = 8 banks mod 32)
template <int kCG, typename Kind, int kRB = 1, bool kStaging = true, int kMC = 0, int kEW = 8, bool kRowStore = false>
struct Cfg {
  static_assert(kEW == 8 || kEW == 16, "8 or 16 epilogue warps");
  static constexpr int EPI_WARPS = kEW;
  static constexpr int EPI_THREADS = kEW * 32;
  static constexpr int THREADS = 64 + EPI_THREADS;          // 320 or 576
  static constexpr int EPI_BUF_BYTES = kRowStore ? 8 * ROWSTORE_PITCH : epi_buf_bytes(kEW);
  static_assert(kMC == 0 || (kCG == 1 && kRB == 1), "B multicast: single-CTA MMAs, one row block");
  static constexpr int PAIR = kCG * (kMC ? 2 : 1);           // CTAs that share one tile index
  static_assert(kRB == 1 || (kRB == 2 && kCG == 1), "two row blocks per CTA only in single-CTA mode");
  static constexpr int BN = Kind::BN;
  static constexpr int ACC_STAGES = (kRB == 2) ? 1 : 2;
  static constexpr int A_BYTES = BM * BK * kRB;             // 16 KB per row block
  static constexpr int B_ROWS = (kCG == 2) ? BN / 2 : BN;   // B rows staged by this CTA
  static constexpr int B_BYTES = B_ROWS * BK;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  // The operand pipeline is LATENCY bound (~2-4k clk per TMA load under load, profiles/
  // r1_ring_depth_experiment.log): throughput scales with the bytes in flight, so the ring takes
  // every byte of shared memory the epilogue does not need (one 4 KB staging buffer per warp).
  // epilogues that never stage through shared memory (direct register->global stores, fused K·alpha)
  // hand their 32 KB to the operand ring
  static constexpr int EPI_BYTES = kStaging ? EPI_WARPS * EPI_BUF_BYTES : 0;
  static constexpr int FIXED_BYTES = EPI_BYTES + 2 * COLTERM_BYTES + 256 + 1024;
  static constexpr int STAGES = (227 * 1024 - FIXED_BYTES) / STAGE_BYTES;   // i8: 3 / 5, mxf4: 4 / 6
  static constexpr int OFF_RING = 0;
  static constexpr int OFF_EPI = OFF_RING + STAGES * STAGE_BYTES;
  static constexpr int OFF_COL = OFF_EPI + EPI_BYTES;
  static constexpr int OFF_BAR = OFF_COL + 2 * COLTERM_BYTES;
  static constexpr int NUM_BARS = 2 * STAGES + 2 * ACC_STAGES;
  static constexpr int OFF_TMEMPTR = OFF_BAR + NUM_BARS * 8;
  static constexpr int SMEM_BYTES = OFF_TMEMPTR + 16;
  static constexpr int SMEM_ALLOC = SMEM_BYTES + 1024;
  static constexpr int TILE_M = BM * PAIR * kRB;            // rows per (pair-)tile
  static constexpr int B_BOX_ROWS = kMC ? BN / 2 : B_ROWS;  // rows per B TMA box
  static_assert(SMEM_ALLOC <= 227 * 1024, "shared memory budget exceeded");
};

struct TileCoord { int mb, nb; };
__device__ __forceinline__ TileCoord tile_coord(uint32_t t, int m_tiles, int n_tiles, int group_m) {
  const uint32_t per_group = (uint32_t)group_m * (uint32_t)n_tiles;
  const uint32_t g = t / per_group;
  const int first_m = (int)g * group_m;
  const uint32_t gsz = (uint32_t)min(group_m, m_tiles - first_m);
  const uint32_t r = t - g * per_group;
  TileCoord c;
  c.mb = first_m + (int)(r % gsz);
  c.nb = (int)(r / gsz);
  return c;
}

// --------------------------------------------------------------------------------- kernel
template <int kCG, typename Kind, typename Epi, int kRB = 1, bool kDirect = false, int kMC = 0, int kEW = 8,
          bool kRowStore = false>
__global__ void __launch_bounds__(64 + kEW * 32, 1)
tanimoto_umma2_kernel(const __grid_constant__ CUtensorMap map_a,
                      const __grid_constant__ CUtensorMap map_b,
                      const __grid_constant__ CUtensorMap map_out, const int32_t* __restrict__ na,
                      const int32_t* __restrict__ nb, int64_t n, int64_t m, int k_blocks, int dbg,
                      typename Epi::OutT* __restrict__ out, int64_t ldo, int lsu_store, int group_m, Epi epi) {
  static_assert(!kDirect || (kRB == 1 && Epi::kPacked && !Epi::kRowdot), "direct-store epilogue: fp32 packed, one row block");
  static_assert(kEW == 8 || (kRB == 1 && !kDirect && Epi::kPacked), "16 epilogue warps: packed fp32 epilogues, one row block");
  static_assert(!kRowStore || (kEW == 8 && kRB == 1 && !kDirect && Epi::kPacked && !Epi::kRowdot),
                "row-store epilogue: packed fp32 Gram, 8 warps, one row block");
  static_assert(!Epi::kSparse || (kCG == 1 && kRB == 1 && kMC == 0), "block-sparse k loop: single-CTA tiles, one row block");
  using C = Cfg<kCG, Kind, kRB, !(kDirect || Epi::kRowdot), kMC, kEW, kRowStore>;
  constexpr int EPI_THREADS = C::EPI_THREADS;
  constexpr int EPI_BUF_BYTES = C::EPI_BUF_BYTES;
  constexpr int BN = Kind::BN;
  constexpr int ACC_STAGES = C::ACC_STAGES;
  constexpr int kPair = C::PAIR;
  using OutT = typename Epi::OutT;
  using TermT = typename Epi::TermT;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t sbase = smem_u32(smem);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = (kPair == 2) ? cluster_ctarank() : 0u;
  const bool is_leader = (kCG == 1) || cta_rank == 0;   // only cta_group::2 has a single MMA-issuing CTA

  const uint32_t bar_full = sbase + C::OFF_BAR;                 // [STAGES]
  const uint32_t bar_empty = bar_full + C::STAGES * 8;          // [STAGES]
  const uint32_t bar_tfull = bar_empty + C::STAGES * 8;         // [ACC_STAGES]
  const uint32_t bar_tempty = bar_tfull + ACC_STAGES * 8;       // [ACC_STAGES]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + C::OFF_TMEMPTR);

  const int m_tiles = (int)((n + C::TILE_M - 1) / C::TILE_M);
  const int n_tiles = (int)((m + BN - 1) / BN);
  const uint32_t total_tiles = (uint32_t)m_tiles * (uint32_t)n_tiles;
  const uint32_t tile0 = blockIdx.x / kPair;
  const uint32_t tile_step = gridDim.x / kPair;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&map_a);
    prefetch_tmap(&map_b);
    prefetch_tmap(&map_out);
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(bar_full + 8 * s, 1);     // leader's producer arrive(+expect_tx); TMA bytes of both CTAs
      mbar_init(bar_empty + 8 * s, kMC ? 2 : 1);    // one (multicast) tcgen05.commit per MMA-issuing CTA
    }
    for (int a = 0; a < ACC_STAGES; ++a) {
      mbar_init(bar_tfull + 8 * a, 1);                   // one (multicast) tcgen05.commit
      mbar_init(bar_tempty + 8 * a, C::EPI_WARPS * kCG);    // every epilogue warp of the pair
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<kCG>(smem_u32(tmem_ptr_smem), TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  if constexpr (Kind::kBlockScaled) {
    // constant scale factors: every byte the MMA can read as SFA/SFB is UE8M0 127 = 2^0, whatever
    // the scale-factor layout; one epilogue warp per lane quadrant fills columns [448,512)
    if (warp >= 2 && warp < 6) {
      const uint32_t lane_base = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
      tmem_fill32(lane_base + SF_COL_A, 0x7F7F7F7Fu);
      tmem_fill32(lane_base + SF_COL_B, 0x7F7F7F7Fu);
      tmem_wait_st();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }
  if constexpr (kPair == 2) cluster_sync_all();   // peer barriers + scale factors ready before any remote signal

  if (warp == 0) {
    // ================================ TMA producer ================================
    int stage = 0;
    uint32_t phase = 0;
    // transaction bytes of BOTH CTAs land on the leader's full barrier
    const uint32_t full_base = (kCG == 2) ? mapa(bar_full, 0) : bar_full;
    for (uint32_t t = tile0; t < total_tiles; t += tile_step) {
      const TileCoord tc = tile_coord(t, m_tiles, n_tiles, group_m);
      const int a_row = tc.mb * C::TILE_M + (int)cta_rank * BM * kRB;   // one TMA box of BM*kRB rows
      const int b_row = tc.nb * BN + (int)cta_rank * C::B_ROWS * (kCG - 1);
      auto load_block = [&](int kb) {
          mbar_wait(bar_empty + 8 * stage, phase ^ 1);
          // all 32 lanes stay converged; tc::w:: wrappers elect the issuing lane inside the asm block
          const uint32_t sa = sbase + C::OFF_RING + stage * C::STAGE_BYTES;
          const uint32_t sb = sa + C::A_BYTES;
          if (dbg & 8) {   // diagnostic: no operand traffic, barrier protocol only (results are garbage)
            if (is_leader) w::mbar_arrive(bar_full + 8 * stage);
          } else {
            if (is_leader) w::mbar_expect_tx(bar_full + 8 * stage, C::STAGE_BYTES * kCG);
            if (kCG == 1 && (dbg & 32)) {   // experiment: keep operands in L2 against the Gram store stream
              const uint64_t pol = l2_policy_evict_last();
              w::tma_load_2d_hint(sa, &map_a, full_base + 8 * stage, kb * BK, a_row, pol);
              w::tma_load_2d_hint(sb, &map_b, full_base + 8 * stage, kb * BK, b_row, pol);
            } else {
              w::tma_load_2d<kCG>(sa, &map_a, full_base + 8 * stage, kb * BK, a_row);
              if constexpr (kMC) {
                // this CTA's half of the B rows, delivered to both CTAs of the cluster
                w::tma_load_2d_mcast(sb + cta_rank * (BN / 2) * BK, &map_b, bar_full + 8 * stage, kb * BK,
                                     tc.nb * BN + (int)cta_rank * (BN / 2), (uint16_t)3);
              } else {
                w::tma_load_2d<kCG>(sb, &map_b, full_base + 8 * stage, kb * BK, b_row);
              }
            }
          }
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
      };
      if constexpr (Epi::kSparse) {
        // only the k-blocks whose A AND B blocks hold something; block 0 always (it initialises the accumulator)
        for (int wd = 0; wd < epi.occ_words; ++wd) {
          uint64_t mk = __ldg(epi.occ_a + (int64_t)tc.mb * epi.occ_words + wd) &
                        __ldg(epi.occ_b + (int64_t)tc.nb * epi.occ_words + wd);
          if (wd == 0) mk |= 1ull;
          while (mk) {
            load_block(wd * 64 + __ffsll((long long)mk) - 1);
            mk &= mk - 1;
          }
        }
      } else {
        for (int kb = 0; kb < k_blocks; ++kb) load_block(kb);
      }
    }
  } else if (warp == 1) {
    // ========================== MMA issuer (leader CTA only) ==========================
    if (is_leader) {
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      constexpr uint32_t idesc = Kind::template idesc<kCG>();
      for (uint32_t t = tile0; t < total_tiles; t += tile_step) {
        mbar_wait(bar_tempty + 8 * acc, acc_phase ^ 1);   // both CTAs drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem0 = tmem_base + (uint32_t)(acc * BN * kRB);
        auto mma_block = [&](int kb) {
          mbar_wait(bar_full + 8 * stage, phase);          // operands of both CTAs have landed
          tc_fence_after();
          const uint32_t sa = sbase + C::OFF_RING + stage * C::STAGE_BYTES;
          const uint64_t db = make_smem_desc(sa + C::A_BYTES);
#pragma unroll
          for (int rb = 0; rb < kRB; ++rb) {
            const uint64_t da = make_smem_desc(sa + rb * (BM * BK));
            const uint32_t d_tmem = d_tmem0 + (uint32_t)(rb * BN);
#pragma unroll
            for (int kk = 0; kk < BK / UK; ++kk) {
              // k-block 0 is always the first one issued for a tile (also in sparse mode): it overwrites
              if constexpr (Kind::kBlockScaled)
                w::mma_mxf4<kCG>(d_tmem, da + (uint64_t)(kk * (UK >> 4)), db + (uint64_t)(kk * (UK >> 4)), idesc,
                                 (uint32_t)((kb | kk) != 0), tmem_base + SF_COL_A, tmem_base + SF_COL_B);
              else
                w::mma_i8<kCG>(d_tmem, da + (uint64_t)(kk * (UK >> 4)), db + (uint64_t)(kk * (UK >> 4)), idesc,
                               (uint32_t)((kb | kk) != 0));
            }
          }
          if constexpr (kMC) w::mma_commit_mcast(bar_empty + 8 * stage, (uint16_t)3);   // frees the slot in both CTAs
          else w::mma_commit<kCG>(bar_empty + 8 * stage);
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        };
        if constexpr (Epi::kSparse) {
          // same k-block list as the producer: both A and B blocks non-empty, block 0 always
          const TileCoord tc = tile_coord(t, m_tiles, n_tiles, group_m);
          for (int wd = 0; wd < epi.occ_words; ++wd) {
            uint64_t mk = __ldg(epi.occ_a + (int64_t)tc.mb * epi.occ_words + wd) &
                          __ldg(epi.occ_b + (int64_t)tc.nb * epi.occ_words + wd);
            if (wd == 0) mk |= 1ull;
            while (mk) {
              mma_block(wd * 64 + __ffsll((long long)mk) - 1);
              mk &= mk - 1;
            }
          }
        } else {
          for (int kb = 0; kb < k_blocks; ++kb) mma_block(kb);
        }
        w::mma_commit<kCG>(bar_tfull + 8 * acc);   // accumulator complete once every MMA issued so far retires
        if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if constexpr (kDirect) {
    // ======================= epilogue, direct-store variant (fp32) =======================
    // tcgen05.ld.16x256b hands four neighbouring lanes 32 contiguous bytes of one output row, so the
    // results go from registers straight to global memory as sector-aligned 8-byte streaming stores:
    // no staging buffer, no proxy fence, no TMA store — and the 32 KB of staging become a fifth ring stage.
    const int ew = warp - 2;
    const int quad = warp & 3;
    const int chalf = ew >> 2;
    const int et = threadIdx.x - 64;
    const int t0 = lane & 3, t1 = lane >> 2;      // this lane: rows t1 + 8k (k = 0..3), columns 8g + 2*t0 + {0,1}
    const uint32_t tempty_leader = (kCG == 2) ? mapa(bar_tempty, 0) : bar_tempty;
    int acc = 0;
    uint32_t acc_phase = 0, tile_par = 0;
    int next_col = 0, next_row[4] = {0, 0, 0, 0};
    auto fetch_norms = [&](uint32_t t) {
      next_col = 0;
#pragma unroll
      for (int k = 0; k < 4; ++k) next_row[k] = 0;
      if (t < total_tiles) {
        const TileCoord tc = tile_coord(t, m_tiles, n_tiles, group_m);
        const int64_t gc = (int64_t)tc.nb * BN + et;
        if (et < BN && gc < m) next_col = __ldg(nb + gc);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int64_t gr = (int64_t)tc.mb * C::TILE_M + (int64_t)cta_rank * BM + quad * 32 + t1 + 8 * k;
          if (gr < n) next_row[k] = __ldg(na + gr);
        }
      }
    };
    fetch_norms(tile0);
    for (uint32_t t = tile0; t < total_tiles; t += tile_step) {
      const TileCoord tc = tile_coord(t, m_tiles, n_tiles, group_m);
      const int64_t row_base = (int64_t)tc.mb * C::TILE_M + (int64_t)cta_rank * BM + quad * 32;
      const int64_t col0 = (int64_t)tc.nb * BN;
      float* col_terms = reinterpret_cast<float*>(smem + C::OFF_COL + tile_par * COLTERM_BYTES);
      if (et < BN) col_terms[et] = epi.col_term(next_col);
      float2 nr2[4];
      OutT* row_ptr[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float r = epi.row_term(next_row[k]);
        nr2[k] = make_float2(r, r);
        row_ptr[k] = out + (row_base + t1 + 8 * k) * ldo + col0 + 2 * t0;
      }
      fetch_norms(t + tile_step);
      asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");

      mbar_wait(bar_tfull + 8 * acc, acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BN + chalf * EPI_COLS);
      const bool rows_live = row_base < n;
      const int STEPS = (chalf == 0 ? EPI_COLS : BN - EPI_COLS) / 32;
#pragma unroll 1
      for (int c = 0; c < STEPS; ++c) {
        const int cl = chalf * EPI_COLS + c * 32;
        const bool live = rows_live && (col0 + cl < m) && !(dbg & 1);
        uint32_t v[32];
        if (live) {
          tmem_ld_16x256b_x4(taddr + c * 32, v);                    // TMEM lanes 0-15 of the quadrant
          tmem_ld_16x256b_x4(taddr + c * 32 + (16u << 16), v + 16); // TMEM lanes 16-31
          tmem_wait_ld();
        }
        if (c == STEPS - 1) {
          tc_fence_before();
          __syncwarp();
          if constexpr (kCG == 2) w::mbar_arrive_cluster(tempty_leader + 8 * acc);
          else w::mbar_arrive(bar_tempty + 8 * acc);
        }
        if (!live) continue;
        if (dbg & 2) {
          if (v[0] == 0x7fffffffu && v[31] == 0x7ffffffeu) col_terms[0] = 0.f;
          continue;
        }
        float2 nc[4];
#pragma unroll
        for (int g = 0; g < 4; ++g) nc[g] = *reinterpret_cast<const float2*>(&col_terms[cl + 8 * g + 2 * t0]);
        float2 res[4][4];
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const int idx = 16 * (k >> 1) + 4 * g + 2 * (k & 1);
            res[k][g] = epi.template pair<Kind::kF32Acc>(v[idx], v[idx + 1], nr2[k], nc[g]);
          }
        if (!(dbg & 4)) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            if (row_base + t1 + 8 * k >= n) continue;
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const int64_t gc = col0 + cl + 8 * g + 2 * t0;
              OutT* dst = row_ptr[k] + cl + 8 * g;
              if (gc + 1 < m) __stcs(reinterpret_cast<float2*>(dst), res[k][g]);
              else if (gc < m) dst[0] = res[k][g].x;
            }
          }
        }
      }
      if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
      tile_par ^= 1;
    }
  } else if constexpr (kRowStore) {
    // ============ epilogue, row-contiguous stores (packed fp32 Gram) ============
    // Tile-wise fp32 stores are bound chip-wide by their granularity: 32-row x 128-byte TMA boxes reach 4.6 TB/s,
    // stores that cover >= 512 contiguous bytes of ONE row 5.2 TB/s (896-byte segments) to 5.8 TB/s (768-byte,
    // 256-byte aligned segments) — tools/store_microbench.cu, profiles/r1_store_stream_ceiling.log.  So here every
    // warp owns 16 rows of the tile (balanced for any BN), reads them with tcgen05.ld.16x256b (thread (t0,t1): rows
    // t1 and t1+8, column pairs 8g+2*t0), stages 8 rows x 128 columns in a padded buffer (544-byte pitch: conflict
    // free for the STS.64 writes) and hands each staged row to the TMA engine as one 1-D bulk copy
    // (cp.async.bulk.global.shared::cta, up to 512 contiguous bytes, asynchronous).  Warp-wide STG.128 stores of the
    // same rows were slower (838 Gpairs/s): they block the issuing warp on the SM's ~32 B/clk store path.
    constexpr int PITCH = ROWSTORE_PITCH;
    const int ew = warp - 2;                      // 0..7
    const int quad = warp & 3;                    // TMEM lane quadrant
    const int half = ew >> 2;                     // rows [16*half, 16*half+16) of the quadrant; also the barrier group
    const int gt = (ew & 3) * 32 + lane;          // thread index inside the group, 0..127
    const int t0 = lane & 3, t1 = lane >> 2;
    uint8_t* sb = smem + C::OFF_EPI + ew * EPI_BUF_BYTES;
    const uint32_t tempty_leader = (kCG == 2) ? mapa(bar_tempty, 0) : bar_tempty;
    // column terms: [group][tile parity][256] floats
    float* terms_base = reinterpret_cast<float*>(smem + C::OFF_COL) + half * 512;
    int acc = 0;
    uint32_t acc_phase = 0, it = 0;
    int next_col[2] = {0, 0}, next_row[2] = {0, 0};
    TileCoord next_tc = tile_coord(tile0 < total_tiles ? tile0 : 0, m_tiles, n_tiles, group_m);
    auto fetch_norms = [&](uint32_t t) {
      next_col[0] = next_col[1] = 0;
      next_row[0] = next_row[1] = 0;
      if (t < total_tiles) {
        next_tc = tile_coord(t, m_tiles, n_tiles, group_m);
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const int lc = gt + 128 * i;
          const int64_t gc = (int64_t)next_tc.nb * BN + lc;
          if (lc < BN && gc < m) next_col[i] = __ldg(nb + gc);
          const int64_t gr = (int64_t)next_tc.mb * C::TILE_M + (int64_t)cta_rank * BM + quad * 32 + half * 16 + t1 + 8 * i;
          if (gr < n) next_row[i] = __ldg(na + gr);
        }
      }
    };
    fetch_norms(tile0);
    for (uint32_t t = tile0; t < total_tiles; t += tile_step, ++it) {
      const TileCoord tc = next_tc;
      const int64_t row0 = (int64_t)tc.mb * C::TILE_M + (int64_t)cta_rank * BM + quad * 32 + half * 16;   // 16 rows
      const int64_t col0 = (int64_t)tc.nb * BN;
      float* col_terms = terms_base + (it & 1) * 256;
      col_terms[gt] = epi.col_term(next_col[0]);
      col_terms[gt + 128] = epi.col_term(next_col[1]);
      const float rt0 = epi.row_term(next_row[0]), rt1 = epi.row_term(next_row[1]);
      fetch_norms(t + tile_step);
      if (half == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
      else asm volatile("bar.sync 2, 128;" ::: "memory");

      mbar_wait(bar_tfull + 8 * acc, acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32 + half * 16) << 16) + (uint32_t)(acc * BN);
      const bool interior = (row0 + 16 <= n) && (col0 + BN <= m);   // no bounds checks needed
#pragma unroll
      for (int p = 0; p < 2; ++p) {
        constexpr int W0 = 128;
        const int cl = p * W0;                                  // first tile-local column of this pass
        constexpr int WLAST = BN - W0;                          // 64, 96 (or 128 for BN = 256)
        const int W = p == 0 ? W0 : WLAST;
        uint32_t v[64];
        tmem_ld_16x256b_x8(taddr + cl, v);
        if (p == 0 || WLAST == 128) tmem_ld_16x256b_x8(taddr + cl + 64, v + 32);
        else if (WLAST == 96) tmem_ld_16x256b_x4(taddr + cl + 64, v + 32);
        tmem_wait_ld();
        if (p == 1) {
          // last TMEM read of this accumulator done -> release it before the math of this pass
          tc_fence_before();
          __syncwarp();
          if constexpr (kCG == 2) w::mbar_arrive_cluster(tempty_leader + 8 * acc);
          else w::mbar_arrive(bar_tempty + 8 * acc);
        }
        if ((dbg & 1) || row0 >= n || col0 + cl >= m) continue;   // warp-uniform
#pragma unroll
        for (int rs = 0; rs < 2; ++rs) {
          const float rterm = rs == 0 ? rt0 : rt1;
          const float2 nr2 = make_float2(rterm, rterm);
          const int G = (p == 0 ? W0 : WLAST) / 8;              // column-pair repetitions: 16, 12 or 8
          float2 res[16];
#pragma unroll
          for (int g = 0; g < 16; ++g) {
            if (g < G) {
              const float2 ct = *reinterpret_cast<const float2*>(&col_terms[cl + 8 * g + 2 * t0]);
              res[g] = epi.template pair<Kind::kF32Acc>(v[4 * g + 2 * rs], v[4 * g + 2 * rs + 1], nr2, ct);
            }
          }
          const bool full_rows = interior || (row0 + rs * 8 + 8 <= n && col0 + cl + W <= m);
          // the bulk copies of the previous box must have finished reading the staging buffer (bulk groups are
          // per thread: the eight issuing lanes wait for their own)
          if (lane < 8) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          __syncwarp();
#pragma unroll
          for (int g = 0; g < 16; ++g)
            if (g < G) *reinterpret_cast<float2*>(sb + t1 * PITCH + g * 32 + t0 * 8) = res[g];
          if (dbg & 4) { __syncwarp(); continue; }
          if (full_rows) {
            // one 1-D bulk copy per staged row: W*4 contiguous bytes, asynchronous (TMA engine), no tensor map
            fence_proxy_async_smem();
            __syncwarp();
            if (lane < 8) {
              const float* dst = out + (row0 + rs * 8 + lane) * ldo + col0 + cl;
              asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                           ::"l"(dst), "r"(smem_u32(sb) + (uint32_t)(lane * PITCH)), "r"((uint32_t)(W * 4)) : "memory");
              asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
          } else {
            // edge tiles: row-wise read-back with bounds checks
            __syncwarp();
            const int LPR = W / 4;
            if (lane < LPR) {
#pragma unroll
              for (int r = 0; r < 8; ++r) {
                const float4 val = *reinterpret_cast<const float4*>(sb + r * PITCH + lane * 16);
                const int64_t gr = row0 + rs * 8 + r;
                const int64_t gc = col0 + cl + lane * 4;
                float* dst = out + gr * ldo + gc;
                if (gr < n && gc + 4 <= m) {
                  __stcs(reinterpret_cast<float4*>(dst), val);
                } else if (gr < n) {
                  const float* pv = reinterpret_cast<const float*>(&val);
#pragma unroll
                  for (int e = 0; e < 4; ++e)
                    if (gc + e < m) dst[e] = pv[e];
                }
              }
            }
            __syncwarp();
          }
        }
      }
      if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
    }
    if (lane < 8) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // staging must outlive the last copies
    __syncwarp();
  } else if constexpr (kEW == 16) {
    // ================== epilogue, 16 warps: 16-column steps, 64-byte staging rows ==================
    // Same structure as the 8-warp packed epilogue below with half-size steps: four role groups of four warps
    // (one warp per TMEM lane quadrant); role r owns the 16-column chunks r, r+4, r+8, r+12 of the tile (4 or
    // 3 steps for BN = 224) and roles rotate every tile; each group shares its 64 column terms (and alphas)
    // through its own 128-thread named barrier; TMEM loads run one step ahead.  Per step: tcgen05.ld 32x32b.x16
    // -> 8 packed float2 results per lane -> 64 bytes of one staging row (swizzle-64B, conflict free) -> TMA
    // store of a 32 x 16 box.  90 registers, so 18 warps fit; measured no faster than 8 warps for the Gram
    // (memory bound) — see DESIGN.md.
    constexpr int CS = 16;
    constexpr int NCH = BN / CS;                  // 14 (mxf4) or 16 (i8) chunks per tile
    const int ew = warp - 2;                      // 0..15
    const int quad = warp & 3;                    // TMEM lane quadrant of this warp
    const int grp = ew >> 2;                      // role group 0..3
    const int gt = (ew & 3) * 32 + lane;          // thread index inside the group, 0..127
    uint8_t* sb = smem + C::OFF_EPI + ew * EPI_BUF_BYTES;
    const uint32_t sb_u32 = sbase + C::OFF_EPI + ew * EPI_BUF_BYTES;
    const uint32_t tempty_leader = (kCG == 2) ? mapa(bar_tempty, 0) : bar_tempty;
    // per group and tile parity: 64 column terms followed by 64 alphas
    float* terms_base = reinterpret_cast<float*>(smem + C::OFF_COL) + grp * 256;
    int acc = 0;
    uint32_t acc_phase = 0, it = 0;
    int next_row = 0;
    uint32_t next_val = 0;      // threads 0..63: column norm (int), threads 64..127: alpha (float bits)
    TileCoord next_tc = tile_coord(tile0 < total_tiles ? tile0 : 0, m_tiles, n_tiles, group_m);
    auto fetch_norms = [&](uint32_t t, uint32_t iter) {
      next_val = 0;
      next_row = 0;
      if (t < total_tiles) {
        next_tc = tile_coord(t, m_tiles, n_tiles, group_m);
        const int role = (grp + (int)iter) & 3;
        const int g = gt & 63;
        const int lc = ((g >> 4) * 4 + role) * CS + (g & 15);     // tile-local column of this thread
        const int64_t gc = (int64_t)next_tc.nb * BN + lc;
        if (lc < BN && gc < m) {
          if (gt < 64) next_val = (uint32_t)__ldg(nb + gc);
          else if constexpr (Epi::kRowdot) next_val = __float_as_uint(__ldg(epi.alpha + gc));
        }
        const int64_t gr = (int64_t)next_tc.mb * C::TILE_M + (int64_t)cta_rank * BM + quad * 32 + lane;
        if (gr < n) next_row = __ldg(na + gr);
      }
    };
    fetch_norms(tile0, 0);
    for (uint32_t t = tile0; t < total_tiles; t += tile_step, ++it) {
      const TileCoord tc = next_tc;
      const int role = (grp + (int)it) & 3;
      const int64_t row0 = (int64_t)tc.mb * C::TILE_M + (int64_t)cta_rank * BM + quad * 32;
      const int64_t col0 = (int64_t)tc.nb * BN;
      float* col_terms = terms_base + (it & 1) * 128;
      const float* alpha_s = col_terms + 64;
      if (gt < 64) col_terms[gt] = epi.col_term((int)next_val);
      else col_terms[gt] = __uint_as_float(next_val);
      const float rterm = epi.row_term(next_row);
      fetch_norms(t + tile_step, it + 1);
      // group barrier (ids 1..4): see the 8-warp epilogue for why one barrier per tile is enough
      switch (grp) {
        case 0: asm volatile("bar.sync 1, 128;" ::: "memory"); break;
        case 1: asm volatile("bar.sync 2, 128;" ::: "memory"); break;
        case 2: asm volatile("bar.sync 3, 128;" ::: "memory"); break;
        default: asm volatile("bar.sync 4, 128;" ::: "memory"); break;
      }

      const int STEPS = (NCH - role + 3) / 4;
      const bool rows_live = row0 < n;
      const float2 nr2 = make_float2(rterm, rterm);
      float rowacc = 0.f;   // fused screening: this lane's sum_j K[row, j] * alpha[j] over its columns

      mbar_wait(bar_tfull + 8 * acc, acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BN + role * CS);
      auto release_acc = [&]() {
        tc_fence_before();
        __syncwarp();
        if constexpr (kCG == 2) w::mbar_arrive_cluster(tempty_leader + 8 * acc);
        else w::mbar_arrive(bar_tempty + 8 * acc);
      };
      auto process = [&](const uint32_t (&v)[CS], int c) {
        const int cl = (4 * c + role) * CS;       // first tile-local column of this step
        if (!(rows_live && (col0 + cl < m)) || (dbg & 1)) return;   // warp-uniform
        float4 nc[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) nc[q] = *reinterpret_cast<const float4*>(&col_terms[c * CS + q * 4]);
        if constexpr (Epi::kRowdot) {
          float4 al[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) al[q] = *reinterpret_cast<const float4*>(&alpha_s[c * CS + q * 4]);
          float2 part = make_float2(0.f, 0.f);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float2 k0 = epi.template pair_fast<Kind::kF32Acc>(v[q * 4 + 0], v[q * 4 + 1], nr2, make_float2(nc[q].x, nc[q].y));
            const float2 k1 = epi.template pair_fast<Kind::kF32Acc>(v[q * 4 + 2], v[q * 4 + 3], nr2, make_float2(nc[q].z, nc[q].w));
            part = __ffma2_rn(k0, make_float2(al[q].x, al[q].y), part);
            part = __ffma2_rn(k1, make_float2(al[q].z, al[q].w), part);
          }
          rowacc += part.x + part.y;
        } else {
          float2 res[8];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            res[2 * q + 0] = epi.template pair<Kind::kF32Acc>(v[q * 4 + 0], v[q * 4 + 1], nr2, make_float2(nc[q].x, nc[q].y));
            res[2 * q + 1] = epi.template pair<Kind::kF32Acc>(v[q * 4 + 2], v[q * 4 + 3], nr2, make_float2(nc[q].z, nc[q].w));
          }
          // single staging buffer per warp: the previous TMA store must have finished reading it
          w::tma_store_wait_read<0>();
          __syncwarp();
#pragma unroll
          for (int q = 0; q < 4; ++q)
            *reinterpret_cast<float4*>(sb + lane * 64 + ((q ^ ((lane >> 1) & 3)) << 4)) =
                make_float4(res[2 * q].x, res[2 * q].y, res[2 * q + 1].x, res[2 * q + 1].y);
          fence_proxy_async_smem();     // generic-proxy smem writes -> visible to the TMA engine
          __syncwarp();
          if (!(dbg & 4)) w::tma_store_2d_commit(&map_out, sb_u32, (int)(col0 + cl), (int)row0);
        }
      };
      uint32_t va[CS], vb[CS];
      tmem_ld16(taddr, va);
#pragma unroll 1
      for (int c = 0; c < STEPS; c += 2) {
        tmem_wait_ld();                                        // va = step c
        if (c + 1 < STEPS) tmem_ld16(taddr + (c + 1) * 4 * CS, vb);
        else release_acc();                                    // last TMEM read of this accumulator done
        process(va, c);
        if (c + 1 < STEPS) {
          tmem_wait_ld();                                      // vb = step c + 1
          if (c + 2 < STEPS) tmem_ld16(taddr + (c + 2) * 4 * CS, va);
          else release_acc();
          process(vb, c + 1);
        }
      }
      if constexpr (Epi::kRowdot) {
        // four partial slabs per column tile, one per role: fixed column subsets, independent of the launch geometry
        const int64_t my_row = row0 + lane;
        if (my_row < n) epi.partials[(int64_t)(tc.nb * 4 + role) * epi.ldp + my_row] = rowacc;
      }
      if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
    }
    if constexpr (!Epi::kRowdot) {
      w::tma_store_wait_read<0>();   // staging smem must outlive the last store's reads
      __syncwarp();
    }
  } else if constexpr (kEW == 8 && kRB == 1 && Epi::kPacked) {
    // ============ epilogue, packed fp32 (Tanimoto Gram and fused screening), one row block ============
    // Two warps per TMEM lane quadrant; one takes the even 32-column chunks of the tile (4 steps), the other
    // the odd ones (3 steps for BN = 224) and the roles swap every tile, so each warp averages BN/64 steps.
    // The four warps with the same role share their column terms through a 128-thread named barrier —
    // they have identical work, so nobody waits for a slower role (the CTA-wide barrier of the generic
    // path cost 15 % of the epilogue, profiles/r1_ncu_source_stalls_mxf4x1.txt).  TMEM loads run one step
    // ahead of the float math (two register buffers): the accumulator is released after the last load,
    // i.e. one step earlier, and the load latency hides behind the previous step's math.
    constexpr int CS = 32;
    const int ew = warp - 2;                      // 0..7
    const int quad = warp & 3;                    // TMEM lane quadrant of this warp
    const int grp = ew >> 2;                      // role group: warps {2..5} / {6..9}
    const int gt = (ew & 3) * 32 + lane;          // thread index inside the group, 0..127
    uint8_t* sb = smem + C::OFF_EPI + ew * EPI_BUF_BYTES;
    const uint32_t sb_u32 = sbase + C::OFF_EPI + ew * EPI_BUF_BYTES;
    const uint32_t tempty_leader = (kCG == 2) ? mapa(bar_tempty, 0) : bar_tempty;
    // column terms: [group][tile parity][128] floats; alpha (fused screening) behind them
    float* terms_base = reinterpret_cast<float*>(smem + C::OFF_COL) + grp * 256;
    float* alpha_base = reinterpret_cast<float*>(smem + C::OFF_COL) + 512 + grp * 256;
    int acc = 0;
    uint32_t acc_phase = 0, it = 0;
    int next_col = 0, next_row = 0;
    float next_alpha = 0.f;
    TileCoord next_tc = tile_coord(tile0 < total_tiles ? tile0 : 0, m_tiles, n_tiles, group_m);
    auto fetch_norms = [&](uint32_t t, uint32_t iter) {
      next_col = 0;
      next_row = 0;
      next_alpha = 0.f;
      if (t < total_tiles) {
        next_tc = tile_coord(t, m_tiles, n_tiles, group_m);
        const int half = grp ^ (int)(iter & 1);
        const int lc = ((gt >> 5) * 2 + half) * CS + (gt & 31);   // tile-local column whose term this thread converts
        const int64_t gc = (int64_t)next_tc.nb * BN + lc;
        if (lc < BN && gc < m) {
          next_col = __ldg(nb + gc);
          if constexpr (Epi::kRowdot) next_alpha = __ldg(epi.alpha + gc);
        }
        const int64_t gr = (int64_t)next_tc.mb * C::TILE_M + (int64_t)cta_rank * BM + quad * 32 + lane;
        if (gr < n) next_row = __ldg(na + gr);
      }
    };
    fetch_norms(tile0, 0);
    for (uint32_t t = tile0; t < total_tiles; t += tile_step, ++it) {
      const TileCoord tc = next_tc;
      const int half = grp ^ (int)(it & 1);       // column half of this warp in this tile
      const int64_t row0 = (int64_t)tc.mb * C::TILE_M + (int64_t)cta_rank * BM + quad * 32;
      const int64_t col0 = (int64_t)tc.nb * BN;
      float* col_terms = terms_base + (it & 1) * 128;
      float* alpha_s = alpha_base + (it & 1) * 128;
      col_terms[gt] = epi.col_term(next_col);
      if constexpr (Epi::kRowdot) alpha_s[gt] = next_alpha;
      const float rterm = epi.row_term(next_row);
      fetch_norms(t + tile_step, it + 1);
      // group barrier (ids 1 and 2): terms of this tile visible; the buffer written now was last read two
      // tiles ago, before every warp of the group passed the previous barrier
      if (grp == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
      else asm volatile("bar.sync 2, 128;" ::: "memory");

      // role `half` owns the 32-column chunks half, half + 2, half + 4, ...: the two warps of a quadrant write
      // neighbouring 128-byte lines of the same rows at about the same time (256 contiguous bytes reach DRAM
      // together; tile-wise stores are bound by their 128-byte granularity, profiles/r1_store_stream_ceiling.log)
      const int STEPS = (BN / CS + 1 - half) / 2;                         // BN = 224: 4 and 3; BN = 256: 4 and 4
      const bool rows_live = row0 < n;
      const float2 nr2 = make_float2(rterm, rterm);
      float rowacc = 0.f;   // fused screening: this lane's sum_j K[row, j] * alpha[j] over its columns

      mbar_wait(bar_tfull + 8 * acc, acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BN + half * CS);
      auto release_acc = [&]() {
        tc_fence_before();
        __syncwarp();
        if constexpr (kCG == 2) w::mbar_arrive_cluster(tempty_leader + 8 * acc);
        else w::mbar_arrive(bar_tempty + 8 * acc);
      };
      auto process = [&](const uint32_t (&v)[CS], int c) {
        const int cl = (2 * c + half) * CS;       // first tile-local column of this step
        if (!(rows_live && (col0 + cl < m)) || (dbg & 1)) return;   // warp-uniform
        if (dbg & 2) {   // diagnostic: TMEM reads only
          if (v[0] == 0x7fffffffu && v[CS - 1] == 0x7ffffffeu) col_terms[0] = 0.f;
          return;
        }
        // column terms first, results in registers until the end: shared-memory loads are not reordered
        // across shared-memory stores, interleaving them serialises the 16 division chains of a step
        float4 nc[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) nc[q] = *reinterpret_cast<const float4*>(&col_terms[c * CS + q * 4]);
        if constexpr (Epi::kRowdot) {
          float4 al[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) al[q] = *reinterpret_cast<const float4*>(&alpha_s[c * CS + q * 4]);
          float2 part = make_float2(0.f, 0.f);
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float2 k0 = epi.template pair_fast<Kind::kF32Acc>(v[q * 4 + 0], v[q * 4 + 1], nr2, make_float2(nc[q].x, nc[q].y));
            const float2 k1 = epi.template pair_fast<Kind::kF32Acc>(v[q * 4 + 2], v[q * 4 + 3], nr2, make_float2(nc[q].z, nc[q].w));
            part = __ffma2_rn(k0, make_float2(al[q].x, al[q].y), part);
            part = __ffma2_rn(k1, make_float2(al[q].z, al[q].w), part);
          }
          rowacc += part.x + part.y;
        } else {
          float2 res[16];
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            res[2 * q + 0] = epi.template pair<Kind::kF32Acc>(v[q * 4 + 0], v[q * 4 + 1], nr2, make_float2(nc[q].x, nc[q].y));
            res[2 * q + 1] = epi.template pair<Kind::kF32Acc>(v[q * 4 + 2], v[q * 4 + 3], nr2, make_float2(nc[q].z, nc[q].w));
          }
          if (dbg & 64) {   // diagnostic: float math only (no staging, fence or store)
            uint32_t o = 0;
#pragma unroll
            for (int q = 0; q < 16; ++q) o |= __float_as_uint(res[q].x) | __float_as_uint(res[q].y);
            if (o == 0x7fffffffu) col_terms[0] = 0.f;
            return;
          }
          // single staging buffer per warp: the previous TMA store must have finished reading it.  (A second
          // buffer with two stores in flight per warp did not raise the store rate: the stream is bandwidth,
          // not latency bound — profiles/r1_epilogue_ablation.log.)
          w::tma_store_wait_read<0>();
          __syncwarp();
#pragma unroll
          for (int q = 0; q < 8; ++q)
            *reinterpret_cast<float4*>(sb + lane * 128 + ((q ^ (lane & 7)) << 4)) =
                make_float4(res[2 * q].x, res[2 * q].y, res[2 * q + 1].x, res[2 * q + 1].y);
          if (!(dbg & 128)) fence_proxy_async_smem();     // generic-proxy smem writes -> visible to the TMA engine
          __syncwarp();
          if (!(dbg & 4)) {   // same elected leader as the wait_read above (bulk groups are per thread)
            if (dbg & 16) w::tma_store_2d_hint_commit(&map_out, sb_u32, (int)(col0 + cl), (int)row0, l2_policy_evict_first());
            else w::tma_store_2d_commit(&map_out, sb_u32, (int)(col0 + cl), (int)row0);
          }
        }
      };
      uint32_t va[CS], vb[CS];
      tmem_ld32(taddr, va);
#pragma unroll 1
      for (int c = 0; c < STEPS; c += 2) {
        tmem_wait_ld();                                        // va = step c
        if (c + 1 < STEPS) tmem_ld32(taddr + (c + 1) * 2 * CS, vb);
        else release_acc();                                    // last TMEM read of this accumulator done
        process(va, c);
        if (c + 1 < STEPS) {
          tmem_wait_ld();                                      // vb = step c + 1
          if (c + 2 < STEPS) tmem_ld32(taddr + (c + 2) * 2 * CS, va);
          else release_acc();
          process(vb, c + 1);
        }
      }
      if constexpr (Epi::kRowdot) {
        const int64_t my_row = row0 + lane;
        if (my_row < n) epi.partials[(int64_t)(tc.nb * 2 + half) * epi.ldp + my_row] = rowacc;
      }
      if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
    }
    if constexpr (!Epi::kRowdot) {
      w::tma_store_wait_read<0>();   // staging smem must outlive the last store's reads
      __syncwarp();
    }
  } else if constexpr (kEW == 8 && kRB == 1 && !Epi::kPacked) {
    // ====== epilogue, element-wise functors (fp64 / raw-count / MinMax / exotic-eps outputs), one row block ======
    // Same structure as the packed fp32 epilogue above — two role groups with their own 128-thread barrier and
    // column terms, alternate CS-column chunks per role, roles swapped every tile, TMEM loads one step ahead —
    // around the generic `epi.one()` functor.  CS columns = one 128-byte staging row: 32 (4-byte outputs) or 16
    // (fp64: 14 chunks per 224-column tile, 7 per warp).
    constexpr int CS = 128 / (int)sizeof(OutT);
    constexpr int NCH = BN / CS;
    static_assert(CS == 32 || CS == 16, "4- or 8-byte outputs");
    static_assert((NCH + 1) / 2 * CS <= 128, "one column term per thread of a role group");
    const int ew = warp - 2;                      // 0..7
    const int quad = warp & 3;                    // TMEM lane quadrant of this warp
    const int grp = ew >> 2;                      // role group: warps {2..5} / {6..9}
    const int gt = (ew & 3) * 32 + lane;          // thread index inside the group, 0..127
    uint8_t* sb = smem + C::OFF_EPI + ew * EPI_BUF_BYTES;
    const uint32_t sb_u32 = sbase + C::OFF_EPI + ew * EPI_BUF_BYTES;
    const uint32_t tempty_leader = (kCG == 2) ? mapa(bar_tempty, 0) : bar_tempty;
    // column terms: [group][tile parity][128] TermT (<= 1 KB each, 4 KB in all)
    TermT* terms_base = reinterpret_cast<TermT*>(smem + C::OFF_COL + grp * 2048);
    int acc = 0;
    uint32_t acc_phase = 0, it = 0;
    int next_col = 0, next_row = 0;
    TileCoord next_tc = tile_coord(tile0 < total_tiles ? tile0 : 0, m_tiles, n_tiles, group_m);
    auto fetch_norms = [&](uint32_t t, uint32_t iter) {
      next_col = 0;
      next_row = 0;
      if (t < total_tiles) {
        next_tc = tile_coord(t, m_tiles, n_tiles, group_m);
        const int role = grp ^ (int)(iter & 1);
        const int lc = ((gt / CS) * 2 + role) * CS + (gt % CS);   // tile-local column whose term this thread converts
        const int64_t gc = (int64_t)next_tc.nb * BN + lc;
        if (lc < BN && gc < m) next_col = __ldg(nb + gc);
        const int64_t gr = (int64_t)next_tc.mb * C::TILE_M + (int64_t)cta_rank * BM + quad * 32 + lane;
        if (gr < n) next_row = __ldg(na + gr);
      }
    };
    fetch_norms(tile0, 0);
    for (uint32_t t = tile0; t < total_tiles; t += tile_step, ++it) {
      const TileCoord tc = next_tc;
      const int role = grp ^ (int)(it & 1);
      const int64_t row0 = (int64_t)tc.mb * C::TILE_M + (int64_t)cta_rank * BM + quad * 32;
      const int64_t col0 = (int64_t)tc.nb * BN;
      TermT* col_terms = terms_base + (it & 1) * 128;
      col_terms[gt] = epi.col_term(next_col);
      const TermT rterm = epi.row_term(next_row);
      fetch_norms(t + tile_step, it + 1);
      if (grp == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
      else asm volatile("bar.sync 2, 128;" ::: "memory");

      const int STEPS = (NCH + 1 - role) / 2;
      const bool rows_live = row0 < n;
      mbar_wait(bar_tfull + 8 * acc, acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BN + role * CS);
      auto release_acc = [&]() {
        tc_fence_before();
        __syncwarp();
        if constexpr (kCG == 2) w::mbar_arrive_cluster(tempty_leader + 8 * acc);
        else w::mbar_arrive(bar_tempty + 8 * acc);
      };
      auto load_step = [&](int c, uint32_t (&v)[CS]) {
        if constexpr (CS == 32) tmem_ld32(taddr + c * 2 * CS, v);
        else tmem_ld16(taddr + c * 2 * CS, v);
      };
      auto process = [&](const uint32_t (&v)[CS], int c) {
        const int cl = (2 * c + role) * CS;       // first tile-local column of this step
        if (!(rows_live && (col0 + cl < m)) || (dbg & 1)) return;   // warp-uniform
        TermT ct[CS];
#pragma unroll
        for (int j = 0; j < CS; ++j) ct[j] = col_terms[c * CS + j];
        OutT res[CS];
#pragma unroll
        for (int j = 0; j < CS; ++j) res[j] = epi.template one<Kind::kF32Acc>(v[j], rterm, ct[j]);
        w::tma_store_wait_read<0>();              // single staging buffer per warp
        __syncwarp();
#pragma unroll
        for (int q = 0; q < 8; ++q)
          *reinterpret_cast<uint4*>(sb + lane * 128 + ((q ^ (lane & 7)) << 4)) =
              *reinterpret_cast<const uint4*>(&res[q * (16 / (int)sizeof(OutT))]);
        fence_proxy_async_smem();
        __syncwarp();
        if (!(dbg & 4)) w::tma_store_2d_commit(&map_out, sb_u32, (int)(col0 + cl), (int)row0);
      };
      uint32_t va[CS], vb[CS];
      load_step(0, va);
#pragma unroll 1
      for (int c = 0; c < STEPS; c += 2) {
        tmem_wait_ld();                                        // va = step c
        if (c + 1 < STEPS) load_step(c + 1, vb);
        else release_acc();                                    // last TMEM read of this accumulator done
        process(va, c);
        if (c + 1 < STEPS) {
          tmem_wait_ld();                                      // vb = step c + 1
          if (c + 2 < STEPS) load_step(c + 2, va);
          else release_acc();
          process(vb, c + 1);
        }
      }
      if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
    }
    w::tma_store_wait_read<0>();   // staging smem must outlive the last store's reads
    __syncwarp();
  } else {
    // ================================== epilogue ==================================
    constexpr int EPC = 16 / (int)sizeof(OutT);   // elements per 16-byte chunk
    constexpr int CS = 8 * EPC;                   // columns per step: 128 bytes per row
    const int ew = warp - 2;                      // 0..7
    const int quad = warp & 3;                    // TMEM lane quadrant of this warp
    const int chalf = ew >> 2;                    // which 128-column half
    const int et = threadIdx.x - 64;
    uint8_t* sb = smem + C::OFF_EPI + ew * EPI_BUF_BYTES;
    const uint32_t sb_u32 = sbase + C::OFF_EPI + ew * EPI_BUF_BYTES;
    const uint32_t tempty_leader = (kCG == 2) ? mapa(bar_tempty, 0) : bar_tempty;
    int acc = 0;
    uint32_t acc_phase = 0, tile_par = 0;
    // integer norms of the NEXT tile are fetched one tile ahead so their global-load latency is
    // never exposed in the per-tile prologue
    float next_alpha = 0.f;
    auto fetch_norms = [&](uint32_t t, int& ncol, int (&nrow)[kRB]) {
      ncol = 0;
#pragma unroll
      for (int rb = 0; rb < kRB; ++rb) nrow[rb] = 0;
      next_alpha = 0.f;
      if (t < total_tiles) {
        const TileCoord tc = tile_coord(t, m_tiles, n_tiles, group_m);
        const int64_t gc = (int64_t)tc.nb * BN + et;
        if (et < BN && gc < m) {
          ncol = __ldg(nb + gc);
          if constexpr (Epi::kRowdot) next_alpha = __ldg(epi.alpha + gc);
        }
#pragma unroll
        for (int rb = 0; rb < kRB; ++rb) {
          const int64_t gr = (int64_t)tc.mb * C::TILE_M + (int64_t)cta_rank * BM * kRB + rb * BM + quad * 32 + lane;
          if (gr < n) nrow[rb] = __ldg(na + gr);
        }
      }
    };
    int next_col, next_row[kRB];
    fetch_norms(tile0, next_col, next_row);
    for (uint32_t t = tile0; t < total_tiles; t += tile_step) {
      const TileCoord tc = tile_coord(t, m_tiles, n_tiles, group_m);
      const int64_t row_base = (int64_t)tc.mb * C::TILE_M + (int64_t)cta_rank * BM * kRB + quad * 32;
      const int64_t col0 = (int64_t)tc.nb * BN;
      TermT* col_terms = reinterpret_cast<TermT*>(smem + C::OFF_COL + tile_par * COLTERM_BYTES);
      if (et < BN) {
        col_terms[et] = epi.col_term(next_col);
        if constexpr (Epi::kRowdot) reinterpret_cast<float*>(col_terms)[256 + et] = next_alpha;
      }
      TermT rterms[kRB];
#pragma unroll
      for (int rb = 0; rb < kRB; ++rb) rterms[rb] = epi.row_term(next_row[rb]);
      fetch_norms(t + tile_step, next_col, next_row);
      asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");

      mbar_wait(bar_tfull + 8 * acc, acc_phase);
      tc_fence_after();
      // the two warps of a quadrant take alternate CS-column chunks (chalf = 0: even, 1: odd): 7 + 7 chunks of
      // 16 columns for fp64 output on 224-column tiles instead of 8 + 6 with contiguous halves
      const int STEPS = (BN / CS + 1 - chalf) / 2;
#pragma unroll 1
      for (int rb = 0; rb < kRB; ++rb) {
      const int64_t row0 = row_base + rb * BM;
      const TermT rterm = rterms[rb];
      float rowacc = 0.f;   // fused screening: this lane's sum_j K[row, j] * alpha[j] over its columns
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BN * kRB + rb * BN);
      const bool rows_live = row0 < n;
#pragma unroll 1
      for (int c = 0; c < STEPS; ++c) {
        const int cl = (2 * c + chalf) * CS;         // first tile-local column of this step
        const bool live = rows_live && (col0 + cl < m) && !(dbg & 1);   // warp-uniform
        uint32_t v[CS];
        if (live) {
          if constexpr (CS == 32) tmem_ld32(taddr + cl, v); else tmem_ld16(taddr + cl, v);
          tmem_wait_ld();
        }
        if (c == STEPS - 1 && rb == kRB - 1) {
          // last TMEM read of this accumulator done -> release it before the math of this step
          tc_fence_before();
          __syncwarp();
          if constexpr (kCG == 2) w::mbar_arrive_cluster(tempty_leader + 8 * acc);
          else w::mbar_arrive(bar_tempty + 8 * acc);
        }
        if (!live) continue;
        // diagnostics: TMEM reads only — every warp (2), or only the warps that share a scheduler with the
        // MMA issuer (256: quadrant 1) / with the TMA producer (512: quadrant 0)
        if ((dbg & 2) || ((dbg & 256) && quad == 1) || ((dbg & 512) && quad == 0)) {
          if (v[0] == 0x7fffffffu && v[CS - 1] == 0x7ffffffeu) col_terms[0] = (TermT)0;
          continue;
        }
        if constexpr (Epi::kRowdot) {
          const float* alpha_s = reinterpret_cast<const float*>(col_terms) + 256;
          const float2 nr2 = make_float2(rterm, rterm);
          float4 nc[8], al[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            nc[q] = *reinterpret_cast<const float4*>(&col_terms[cl + q * 4]);
            al[q] = *reinterpret_cast<const float4*>(&alpha_s[cl + q * 4]);
          }
          float2 part = make_float2(0.f, 0.f);
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float2 k0 = epi.template pair_fast<Kind::kF32Acc>(v[q * 4 + 0], v[q * 4 + 1], nr2, make_float2(nc[q].x, nc[q].y));
            const float2 k1 = epi.template pair_fast<Kind::kF32Acc>(v[q * 4 + 2], v[q * 4 + 3], nr2, make_float2(nc[q].z, nc[q].w));
            part = __ffma2_rn(k0, make_float2(al[q].x, al[q].y), part);
            part = __ffma2_rn(k1, make_float2(al[q].z, al[q].w), part);
          }
          rowacc += part.x + part.y;
          continue;
        }
        // Column terms are loaded BEFORE any staging store and all results are kept in registers
        // until the end: shared-memory loads are not reordered across shared-memory stores, and
        // interleaving them (load, compute, store per 16-byte chunk) serialised the 16 independent
        // division chains of a step (profiles/r1_mxf4x1_epilogue_serialised.txt).
        if constexpr (Epi::kPacked) {
          const float2 nr2 = make_float2(rterm, rterm);
          float4 nc[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) nc[q] = *reinterpret_cast<const float4*>(&col_terms[cl + q * 4]);
          float2 res[16];
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            res[2 * q + 0] = epi.template pair<Kind::kF32Acc>(v[q * 4 + 0], v[q * 4 + 1], nr2, make_float2(nc[q].x, nc[q].y));
            res[2 * q + 1] = epi.template pair<Kind::kF32Acc>(v[q * 4 + 2], v[q * 4 + 3], nr2, make_float2(nc[q].z, nc[q].w));
          }
          if (dbg & 64) {   // diagnostic: float math only (no staging, fence or store)
            uint32_t o = 0;
#pragma unroll
            for (int q = 0; q < 16; ++q) o |= __float_as_uint(res[q].x) | __float_as_uint(res[q].y);
            if (o == 0x7fffffffu) col_terms[0] = (TermT)0;
            continue;
          }
          // single staging buffer per warp: the previous TMA store must have finished reading it;
          // the wait overlaps with the math above
          if (!lsu_store) {
            w::tma_store_wait_read<0>();
            __syncwarp();
          }
#pragma unroll
          for (int q = 0; q < 8; ++q)
            *reinterpret_cast<float4*>(sb + lane * 128 + ((q ^ (lane & 7)) << 4)) =
                make_float4(res[2 * q].x, res[2 * q].y, res[2 * q + 1].x, res[2 * q + 1].y);
        } else {
          TermT ct[CS];
#pragma unroll
          for (int j = 0; j < CS; ++j) ct[j] = col_terms[cl + j];
          OutT res[CS];
#pragma unroll
          for (int j = 0; j < CS; ++j) res[j] = epi.template one<Kind::kF32Acc>(v[j], rterm, ct[j]);
          if (!lsu_store) {
            w::tma_store_wait_read<0>();
            __syncwarp();
          }
#pragma unroll
          for (int q = 0; q < 8; ++q)
            *reinterpret_cast<uint4*>(sb + lane * 128 + ((q ^ (lane & 7)) << 4)) =
                *reinterpret_cast<const uint4*>(&res[q * EPC]);
        }
        if (lsu_store) {
          // LSU path: transposed read-back, 8 lanes cover one 128-byte row segment, streaming stores;
          // keeps the output stream out of the TMA engine that feeds the MMA operands
          __syncwarp();
          if (!(dbg & 4)) {
            const int64_t cbase = col0 + cl;
#pragma unroll
            for (int it = 0; it < 8; ++it) {
              const int r = it * 4 + (lane >> 3);
              const int q = lane & 7;
              const int64_t gr = row0 + r;
              const int64_t gc = cbase + q * EPC;
              if (gr < n && gc < m) {
                const uint4 val = *reinterpret_cast<const uint4*>(sb + r * 128 + ((q ^ (r & 7)) << 4));
                OutT* dst = out + gr * ldo + gc;
                if (gc + EPC <= m) {
                  __stcs(reinterpret_cast<uint4*>(dst), val);
                } else {
                  const OutT* pv = reinterpret_cast<const OutT*>(&val);
#pragma unroll
                  for (int e = 0; e < EPC; ++e)
                    if (gc + e < m) dst[e] = pv[e];
                }
              }
            }
          }
          __syncwarp();
        } else {
          if (!(dbg & 128)) fence_proxy_async_smem();     // generic-proxy smem writes -> visible to the TMA engine
          __syncwarp();
          if (!(dbg & 4)) {   // same elected leader as the wait_read above (bulk groups are per thread)
            if (dbg & 16) w::tma_store_2d_hint_commit(&map_out, sb_u32, (int)(col0 + cl), (int)row0, l2_policy_evict_first());
            else w::tma_store_2d_commit(&map_out, sb_u32, (int)(col0 + cl), (int)row0);
          }
        }
      }
      if constexpr (Epi::kRowdot) {
        const int64_t my_row = row0 + lane;
        if (my_row < n) epi.partials[(int64_t)(tc.nb * 2 + chalf) * epi.ldp + my_row] = rowacc;
      }
      }   // row blocks
      if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
      tile_par ^= 1;
    }
    w::tma_store_wait_read<0>();   // staging smem must outlive the last store's reads
    __syncwarp();
  }

  tc_fence_before();
  __syncthreads();
  if constexpr (kPair == 2) cluster_sync_all();   // no CTA of the pair exits while the other may signal it
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<kCG>(tmem_base, TMEM_COLS);
  }
}

// --------------------------------------------------------------------------------- host
// raster group: how many row blocks sweep the column tiles together (GAUCHE_B200_GROUP_M overrides)
static int group_m_setting() {
  static const int g = [] { const char* e = getenv("GAUCHE_B200_GROUP_M"); int v = e ? atoi(e) : GROUP_M; return v > 0 ? v : GROUP_M; }();
  return g;
}

// GAUCHE_B200_EW=8|16: epilogue warps of the packed-fp32 Tanimoto kernels (Gram and fused screening)
static int epi_warps_setting() {
  static const int w = [] { const char* e = getenv("GAUCHE_B200_EW"); return (e && atoi(e) == 16) ? 16 : 8; }();
  return w;
}

template <int kCG, typename Kind, int kRB = 1, bool kDirect = false, int kMC = 0, int kEW = 8, bool kRowStore = false,
          typename Epi = void>
static int launch(const uint8_t* a, int64_t lda, const int32_t* na, int64_t n, const uint8_t* b,
                  int64_t ldb, const int32_t* nb, int64_t m, int64_t k, void* out, int64_t ldo,
                  CUtensorMapDataType out_dt, Epi epi, cudaStream_t stream) {
  using C = Cfg<kCG, Kind, kRB, !(kDirect || Epi::kRowdot), kMC, kEW, kRowStore>;
  constexpr int BN = Kind::BN;
  using OutT = typename Epi::OutT;
  CUtensorMap ma, mb, mo;
  if (int e = make_map(&ma, CU_TENSOR_MAP_DATA_TYPE_UINT8, a, k, n, lda, BK, BM * kRB)) return e;
  if (int e = make_map(&mb, CU_TENSOR_MAP_DATA_TYPE_UINT8, b, k, m, ldb, BK, C::B_BOX_ROWS)) return e;
  if (int e = make_map(&mo, out_dt, out, m, n, ldo * (int64_t)sizeof(OutT), epi_row_bytes(kEW) / (int)sizeof(OutT), 32,
                       kEW == 16 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B)) return e;
  auto kern = tanimoto_umma2_kernel<kCG, Kind, Epi, kRB, kDirect, kMC, kEW, kRowStore>;
  static bool attr_done[64] = {false};
  int dev = 0;
  GK_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || !attr_done[dev]) {
    GK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_ALLOC));
    if (dev >= 0 && dev < 64) attr_done[dev] = true;
  }
  const int64_t tiles = ((n + C::TILE_M - 1) / C::TILE_M) * ((m + BN - 1) / BN);
  int64_t units = sm_count() / C::PAIR;       // persistent: one CTA (pair) per SM (pair)
  // GAUCHE_B200_MAX_CTAS: diagnostic cap on the persistent grid (is the limit per SM or chip wide?)
  static const int max_ctas = [] { const char* e = getenv("GAUCHE_B200_MAX_CTAS"); return e ? atoi(e) : 0; }();
  if (max_ctas > 0 && units > max_ctas / C::PAIR) units = max_ctas / C::PAIR;
  if (units > tiles) units = tiles;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(units * C::PAIR));
  cfg.blockDim = dim3(C::THREADS);
  cfg.dynamicSmemBytes = C::SMEM_ALLOC;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = C::PAIR;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  // GAUCHE_B200_DEBUG_MODE: timing diagnostics only (1 = no epilogue work, 2 = TMEM reads only,
  // 4 = no TMA stores, 8 = no operand loads, 16/32 = L2 hints, 64 = float math but no staging,
  // 128 = no proxy fence); any non-zero value produces garbage results.
  static const int dbg = [] { const char* e = getenv("GAUCHE_B200_DEBUG_MODE"); return e ? atoi(e) : 0; }();
  // GAUCHE_B200_STORE=lsu|tma selects the output path of the epilogue (default tma; both are exact and
  // equally fast: the limit is the SM<->L2 port, not the TMA engine — profiles/r1_store_path_lsu_vs_tma.log)
  static const int lsu_store = [] { const char* e = getenv("GAUCHE_B200_STORE"); return (e && e[0] == 'l') ? 1 : 0; }();
  GK_CUDA(cudaLaunchKernelEx(&cfg, kern, ma, mb, mo, na, nb, n, m, (int)(k / BK), dbg, (OutT*)out, ldo, lsu_store, group_m_setting(), epi));
  return 0;
}

// fused screening launch: `out`/`ldo` describe the partials buffer [2*n_tiles, ldp] (also used for the
// unused output tensor map)
template <int kCG, typename Kind, int kRB = 1, int kMC = 0, int kEW = 8>
static int launch_rowdot(const uint8_t* a, int64_t lda, const int32_t* na, int64_t n, const uint8_t* b,
                         int64_t ldb, const int32_t* nb, int64_t m, int64_t k, const float* alpha, double eps,
                         float* partials, int64_t ldp, cudaStream_t stream) {
  TanimotoRowdotF32x2 epi;
  epi.eps = (float)eps;
  epi.alpha = alpha;
  epi.partials = partials;
  epi.ldp = ldp;
  const int64_t slabs = (kEW / 4) * ((m + Kind::BN - 1) / Kind::BN);
  // the output tensor map is not used by this epilogue; describe the partials buffer so it is valid
  CUtensorMap ma, mb, mo;
  using C = Cfg<kCG, Kind, kRB, false, kMC, kEW>;   // fused epilogue never stages: the ring gets the staging bytes
  if (int e = make_map(&ma, CU_TENSOR_MAP_DATA_TYPE_UINT8, a, k, n, lda, BK, BM * kRB)) return e;
  if (int e = make_map(&mb, CU_TENSOR_MAP_DATA_TYPE_UINT8, b, k, m, ldb, BK, C::B_BOX_ROWS)) return e;
  if (int e = make_map(&mo, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, partials, ldp, slabs, ldp * 4, 32, 32)) return e;
  auto kern = tanimoto_umma2_kernel<kCG, Kind, TanimotoRowdotF32x2, kRB, false, kMC, kEW>;
  static bool attr_done[64] = {false};
  int dev = 0;
  GK_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || !attr_done[dev]) {
    GK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_ALLOC));
    if (dev >= 0 && dev < 64) attr_done[dev] = true;
  }
  const int64_t tiles = ((n + C::TILE_M - 1) / C::TILE_M) * ((m + Kind::BN - 1) / Kind::BN);
  int64_t units = sm_count() / C::PAIR;
  if (units > tiles) units = tiles;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(units * C::PAIR));
  cfg.blockDim = dim3(C::THREADS);
  cfg.dynamicSmemBytes = C::SMEM_ALLOC;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = C::PAIR;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  GK_CUDA(cudaLaunchKernelEx(&cfg, kern, ma, mb, mo, na, nb, n, m, (int)(k / BK), 0, (float*)partials, ldp, 0, group_m_setting(), epi));
  return 0;
}

template <int kCG, typename Kind, int kRB = 1>
static int dispatch(const uint8_t* a, int64_t lda, const int32_t* a_sq, int64_t n, const uint8_t* b,
                    int64_t ldb, const int32_t* b_sq, int64_t m, int64_t k, void* out, int64_t ldo,
                    int out_dtype, double eps, cudaStream_t s) {
  switch (out_dtype) {
    case GK_F32:
      if (eps >= 1e-12 && eps <= 1.0) {
        // GAUCHE_B200_EPILOGUE=direct selects register->global 8-byte stores (frees the staging smem for a
        // fifth ring stage).  Measured slower (737 vs 872 Gpairs/s): 32-byte sector writes cost the memory
        // system more than the TMA store's full 128-byte lines, so the staged TMA epilogue stays the default.
        static const bool direct = [] { const char* e = getenv("GAUCHE_B200_EPILOGUE"); return e && e[0] == 'd'; }();
        if constexpr (kRB == 1) {
          if (epi_warps_setting() == 16)
            return launch<kCG, Kind, 1, false, 0, 16>(a, lda, a_sq, n, b, ldb, b_sq, m, k, out, ldo,
                                                      CU_TENSOR_MAP_DATA_TYPE_FLOAT32, TanimotoF32x2{(float)eps}, s);
          if (direct)
            return launch<kCG, Kind, kRB, true>(a, lda, a_sq, n, b, ldb, b_sq, m, k, out, ldo,
                                                CU_TENSOR_MAP_DATA_TYPE_FLOAT32, TanimotoF32x2{(float)eps}, s);
        }
        return launch<kCG, Kind, kRB>(a, lda, a_sq, n, b, ldb, b_sq, m, k, out, ldo, CU_TENSOR_MAP_DATA_TYPE_FLOAT32,
                           TanimotoF32x2{(float)eps}, s);
      }
      return launch<kCG, Kind, kRB>(a, lda, a_sq, n, b, ldb, b_sq, m, k, out, ldo, CU_TENSOR_MAP_DATA_TYPE_FLOAT32,
                         Elementwise<TanimotoEpilogue<float>, float>{TanimotoEpilogue<float>{(float)eps}}, s);
    case GK_F64:
      if (eps >= 1e-12 && eps <= 1.0)   // normal operands, normal quotient: branch-free exact division
        return launch<kCG, Kind, kRB>(a, lda, a_sq, n, b, ldb, b_sq, m, k, out, ldo, CU_TENSOR_MAP_DATA_TYPE_FLOAT64,
                           Elementwise<TanimotoEpilogueInt<double>, double>{TanimotoEpilogueInt<double>{{eps}}}, s);
      return launch<kCG, Kind, kRB>(a, lda, a_sq, n, b, ldb, b_sq, m, k, out, ldo, CU_TENSOR_MAP_DATA_TYPE_FLOAT64,
                         Elementwise<TanimotoEpilogue<double>, double>{TanimotoEpilogue<double>{eps}}, s);
    case GK_I32:
      return launch<kCG, Kind, kRB>(a, lda, a_sq, n, b, ldb, b_sq, m, k, out, ldo, CU_TENSOR_MAP_DATA_TYPE_INT32,
                         Elementwise<RawEpilogue, int32_t>{RawEpilogue{}}, s);
    default:
      return set_err(GK_EINVAL, "gk_tanimoto_gram_umma2: unsupported out_dtype %d", out_dtype);
  }
}

}  // namespace umma2
}  // namespace gk

namespace gk {
namespace umma2 {
static int check_args(const char* name, const uint8_t* a, int64_t lda, const int32_t* a_sq, int64_t n,
                      const uint8_t* b, int64_t ldb, const int32_t* b_sq, int64_t m, int64_t k,
                      void* out, int64_t ldo, int out_dtype, int cta_group, bool check_out = true) {
  GK_CHECK_ARG(n >= 0 && m >= 0 && k > 0, GK_EINVAL, "%s: bad size", name);
  GK_CHECK_ARG(cta_group >= 1 && cta_group <= 6, GK_EINVAL,
               "%s: cta_group must be 1, 2 (CTA pair), 3 (single CTA, two accumulator row blocks), 4 (B multicast) "
               "or 5/6 (row-contiguous stores, 192/224-column tiles)", name);
  GK_CHECK_ARG(a && b && a_sq && b_sq && out, GK_EINVAL, "%s: null pointer", name);
  GK_CHECK_ARG(k % BK == 0, GK_EALIGN, "%s: k (%lld) must be a multiple of %d bytes", name, (long long)k, BK);
  GK_CHECK_ARG(lda >= k && ldb >= k && (!check_out || ldo >= m), GK_EINVAL, "%s: leading dimension too small", name);
  GK_CHECK_ARG(((uintptr_t)a) % 16 == 0 && ((uintptr_t)b) % 16 == 0 && lda % 16 == 0 && ldb % 16 == 0,
               GK_EALIGN, "%s: operands must be 16-byte aligned with lda/ldb multiples of 16", name);
  const int64_t osz = (out_dtype == GK_F64) ? 8 : 4;
  GK_CHECK_ARG(!check_out || (((uintptr_t)out) % 16 == 0 && (ldo * osz) % 16 == 0), GK_EALIGN,
               "%s: out must be 16-byte aligned with a row pitch that is a multiple of 16 bytes "
               "(TMA store); use gk_tanimoto_gram_umma for unaligned outputs", name);
  GK_CHECK_ARG(n < (1LL << 31) && m < (1LL << 31) && k < (1LL << 31), GK_ERANGE, "%s: size too large", name);
  GK_CHECK_ARG(((n + BM - 1) / BM) * ((m + 191) / 192) < (1LL << 31), GK_ERANGE,
               "%s: more than 2^31 output tiles; split the row range", name);
  int64_t cc = 0;
  if (int e = gk_query(GK_Q_CC, &cc)) return e;
  GK_CHECK_ARG(cc / 10 == 10, GK_EDEVICE, "%s: requires an sm_100 device (found sm_%lld)", name, (long long)cc);
  return 0;
}
}  // namespace umma2
}  // namespace gk

extern "C" int gk_tanimoto_gram_umma2(const uint8_t* a, int64_t lda, const int32_t* a_sq, int64_t n,
                                      const uint8_t* b, int64_t ldb, const int32_t* b_sq, int64_t m,
                                      int64_t k, void* out, int64_t ldo, int out_dtype, double eps,
                                      int cta_group, void* stream) {
  using namespace gk::umma2;
  if (n == 0 || m == 0) return 0;
  if (int e = check_args("gk_tanimoto_gram_umma2", a, lda, a_sq, n, b, ldb, b_sq, m, k, out, ldo, out_dtype, cta_group))
    return e;
  cudaStream_t s = (cudaStream_t)stream;
  if (cta_group == 2) return dispatch<2, KindI8>(a, lda, a_sq, n, b, ldb, b_sq, m, k, out, ldo, out_dtype, eps, s);
  if (cta_group == 3) return dispatch<1, KindI8, 2>(a, lda, a_sq, n, b, ldb, b_sq, m, k, out, ldo, out_dtype, eps, s);
  return dispatch<1, KindI8>(a, lda, a_sq, n, b, ldb, b_sq, m, k, out, ldo, out_dtype, eps, s);
}

extern "C" int gkl_tanimoto_gram_mxf4(const uint8_t* a4, int64_t lda, const int32_t* a_n, int64_t n,
                                     const uint8_t* b4, int64_t ldb, const int32_t* b_n, int64_t m,
                                     int64_t k_bytes, void* out, int64_t ldo, int out_dtype, double eps,
                                     int cta_group, void* stream) {
  using namespace gk::umma2;
  if (n == 0 || m == 0) return 0;
  if (int e = check_args("gkl_tanimoto_gram_mxf4", a4, lda, a_n, n, b4, ldb, b_n, m, k_bytes, out, ldo, out_dtype, cta_group))
    return e;
  // f32 accumulation of 0/1 products is exact while every count stays below 2^24
  if (k_bytes * 2 >= (1LL << 24))
    return gk::set_err(GK_ERANGE, "gk_tanimoto_gram_mxf4: %lld bits per fingerprint exceed exact f32 accumulation",
                       (long long)(k_bytes * 2));
  cudaStream_t s = (cudaStream_t)stream;
  if (cta_group == 2) return dispatch<2, KindMxf4>(a4, lda, a_n, n, b4, ldb, b_n, m, k_bytes, out, ldo, out_dtype, eps, s);
  if (cta_group == 3) return dispatch<1, KindMxf4, 2>(a4, lda, a_n, n, b4, ldb, b_n, m, k_bytes, out, ldo, out_dtype, eps, s);
  if (cta_group == 5 || cta_group == 6) {   // row-contiguous LSU stores; 5: 128 x 192 tiles, 6: 128 x 224 tiles
    if (out_dtype != GK_F32 || !(eps >= 1e-12 && eps <= 1.0))
      return gk::set_err(GK_EINVAL, "gk_tanimoto_gram_mxf4: the row-store variants serve fp32 output with the default eps range");
    if (cta_group == 5)
      return launch<1, KindMxf4N192, 1, false, 0, 8, true>(a4, lda, a_n, n, b4, ldb, b_n, m, k_bytes, out, ldo,
                                                           CU_TENSOR_MAP_DATA_TYPE_FLOAT32, TanimotoF32x2{(float)eps}, s);
    return launch<1, KindMxf4, 1, false, 0, 8, true>(a4, lda, a_n, n, b4, ldb, b_n, m, k_bytes, out, ldo,
                                                     CU_TENSOR_MAP_DATA_TYPE_FLOAT32, TanimotoF32x2{(float)eps}, s);
  }
  if (cta_group == 4) {   // clusters of two single-CTA tiles sharing B through TMA multicast
    if (out_dtype != GK_F32 || !(eps >= 1e-12 && eps <= 1.0))
      return gk::set_err(GK_EINVAL, "gk_tanimoto_gram_mxf4: the multicast variant serves fp32 output with the default eps range");
    return launch<1, KindMxf4, 1, false, 1>(a4, lda, a_n, n, b4, ldb, b_n, m, k_bytes, out, ldo,
                                            CU_TENSOR_MAP_DATA_TYPE_FLOAT32, TanimotoF32x2{(float)eps}, s);
  }
  return dispatch<1, KindMxf4>(a4, lda, a_n, n, b4, ldb, b_n, m, k_bytes, out, ldo, out_dtype, eps, s);
}

namespace gk {
namespace umma2 {
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
}  // namespace umma2
}  // namespace gk

namespace gk {
namespace umma2 {
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
}  // namespace umma2
}  // namespace gk

extern "C" int64_t gk_block_occupancy_words(int64_t rows, int64_t k_bytes, int tile_rows) {
  if (rows < 0 || k_bytes <= 0 || tile_rows <= 0) return -1;
  const int64_t tiles = (rows + tile_rows - 1) / tile_rows;
  const int64_t words = (k_bytes / gk::umma2::BK + 63) / 64;
  return tiles * words;
}

extern "C" int gk_block_occupancy(const uint8_t* q, int64_t ld, int64_t rows, int64_t k_bytes, int tile_rows,
                                  uint64_t* occ, void* stream) {
  using namespace gk;
  using namespace gk::umma2;
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

extern "C" int gkl_minmax_gram_mxf4_sparse(const uint8_t* a4t, int64_t lda, const int32_t* a_l1, int64_t n,
                                          const uint8_t* b4t, int64_t ldb, const int32_t* b_l1, int64_t m,
                                          int64_t k_bytes, const uint64_t* occ_a, const uint64_t* occ_b, void* out,
                                          int64_t ldo, int out_dtype, double eps, void* stream) {
  using namespace gk;
  using namespace gk::umma2;
  const char* name = "gkl_minmax_gram_mxf4_sparse";
  if (n == 0 || m == 0) return 0;
  if (int e = check_args(name, a4t, lda, a_l1, n, b4t, ldb, b_l1, m, k_bytes, out, ldo, out_dtype, 1)) return e;
  GK_CHECK_ARG(occ_a && occ_b, GK_EINVAL, "%s: null occupancy map", name);
  GK_CHECK_ARG(k_bytes * 2 < (1LL << 24), GK_ERANGE, "%s: %lld thermometer bits exceed exact f32 accumulation", name,
               (long long)(k_bytes * 2));
  GK_CHECK_ARG(eps >= 1e-12 && eps <= 1.0, GK_ERANGE, "%s: eps outside [1e-12, 1] (use gk_minmax_gram_mxf4)", name);
  cudaStream_t s = (cudaStream_t)stream;
  const int words = (int)((k_bytes / BK + 63) / 64);
  switch (out_dtype) {
    case GK_F32: {
      ElementwiseSparse<MinMaxFromMinEpilogueFast<float>, float> epi;
      epi.base = MinMaxFromMinEpilogueFast<float>{{(float)eps}};
      epi.occ_a = occ_a; epi.occ_b = occ_b; epi.occ_words = words;
      return launch<1, KindMxf4>(a4t, lda, a_l1, n, b4t, ldb, b_l1, m, k_bytes, out, ldo, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, epi, s);
    }
    case GK_F64: {
      ElementwiseSparse<MinMaxFromMinEpilogueFast<double>, double> epi;
      epi.base = MinMaxFromMinEpilogueFast<double>{{eps}};
      epi.occ_a = occ_a; epi.occ_b = occ_b; epi.occ_words = words;
      return launch<1, KindMxf4>(a4t, lda, a_l1, n, b4t, ldb, b_l1, m, k_bytes, out, ldo, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, epi, s);
    }
    case GK_I32: {
      ElementwiseSparse<RawL1FromMinEpilogue, int32_t> epi;
      epi.base = RawL1FromMinEpilogue{};
      epi.occ_a = occ_a; epi.occ_b = occ_b; epi.occ_words = words;
      return launch<1, KindMxf4>(a4t, lda, a_l1, n, b4t, ldb, b_l1, m, k_bytes, out, ldo, CU_TENSOR_MAP_DATA_TYPE_INT32, epi, s);
    }
    default:
      return set_err(GK_EINVAL, "%s: unsupported out_dtype %d", name, out_dtype);
  }
}

extern "C" int gk_minmax_gram_mxf4(const uint8_t* a4t, int64_t lda, const int32_t* a_l1, int64_t n,
                                   const uint8_t* b4t, int64_t ldb, const int32_t* b_l1, int64_t m,
                                   int64_t k_bytes, void* out, int64_t ldo, int out_dtype, double eps,
                                   void* stream) {
  using namespace gk;
  using namespace gk::umma2;
  const char* name = "gk_minmax_gram_mxf4";
  if (n == 0 || m == 0) return 0;
  if (int e = check_args(name, a4t, lda, a_l1, n, b4t, ldb, b_l1, m, k_bytes, out, ldo, out_dtype, 1)) return e;
  GK_CHECK_ARG(k_bytes * 2 < (1LL << 24), GK_ERANGE, "%s: %lld thermometer bits exceed exact f32 accumulation", name,
               (long long)(k_bytes * 2));
  cudaStream_t s = (cudaStream_t)stream;
  const bool fast_div = eps >= 1e-12 && eps <= 1.0;   // branch-free exact division (common.cuh)
  switch (out_dtype) {
    case GK_F32:
      if (fast_div)
        return launch<1, KindMxf4>(a4t, lda, a_l1, n, b4t, ldb, b_l1, m, k_bytes, out, ldo, CU_TENSOR_MAP_DATA_TYPE_FLOAT32,
                                   Elementwise<MinMaxFromMinEpilogueFast<float>, float>{MinMaxFromMinEpilogueFast<float>{{(float)eps}}}, s);
      return launch<1, KindMxf4>(a4t, lda, a_l1, n, b4t, ldb, b_l1, m, k_bytes, out, ldo, CU_TENSOR_MAP_DATA_TYPE_FLOAT32,
                                 Elementwise<MinMaxFromMinEpilogue<float>, float>{MinMaxFromMinEpilogue<float>{(float)eps}}, s);
    case GK_F64:
      if (fast_div)
        return launch<1, KindMxf4>(a4t, lda, a_l1, n, b4t, ldb, b_l1, m, k_bytes, out, ldo, CU_TENSOR_MAP_DATA_TYPE_FLOAT64,
                                   Elementwise<MinMaxFromMinEpilogueFast<double>, double>{MinMaxFromMinEpilogueFast<double>{{eps}}}, s);
      return launch<1, KindMxf4>(a4t, lda, a_l1, n, b4t, ldb, b_l1, m, k_bytes, out, ldo, CU_TENSOR_MAP_DATA_TYPE_FLOAT64,
                                 Elementwise<MinMaxFromMinEpilogue<double>, double>{MinMaxFromMinEpilogue<double>{eps}}, s);
    case GK_I32:
      return launch<1, KindMxf4>(a4t, lda, a_l1, n, b4t, ldb, b_l1, m, k_bytes, out, ldo, CU_TENSOR_MAP_DATA_TYPE_INT32,
                                 Elementwise<RawL1FromMinEpilogue, int32_t>{RawL1FromMinEpilogue{}}, s);
    default:
      return set_err(GK_EINVAL, "%s: unsupported out_dtype %d", name, out_dtype);
  }
}

extern "C" int gk_reduce_partials(const float* partials, int64_t slabs, int64_t n, int64_t ldp, float bias,
                                  float* scores, void* stream) {
  using namespace gk;
  GK_CHECK_ARG(slabs >= 0 && n >= 0, GK_EINVAL, "gk_reduce_partials: negative size");
  if (n == 0) return 0;
  GK_CHECK_ARG(partials && scores && ldp >= n, GK_EINVAL, "gk_reduce_partials: bad arguments");
  int64_t blocks = (n + 31) / 32;
  const int64_t cap = (int64_t)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  gk::umma2::reduce_partials_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(partials, slabs, n, ldp,
                                                                                       bias, scores);
  return cuda_err(cudaGetLastError(), "reduce_partials_kernel launch");
}

extern "C" int64_t gk_tanimoto_rowdot_slabs(int64_t m, int operand_kind) {
  const int bn = operand_kind == 1 ? gk::umma2::KindMxf4::BN : gk::umma2::KindI8::BN;
  return (gk::umma2::epi_warps_setting() / 4) * ((m + bn - 1) / bn);   // one slab per epilogue warp of a quadrant
}

extern "C" int gkl_tanimoto_rowdot(const uint8_t* a, int64_t lda, const int32_t* a_sq, int64_t n,
                                  const uint8_t* b, int64_t ldb, const int32_t* b_sq, int64_t m,
                                  int64_t k_bytes, int operand_kind, const float* alpha, double eps, float bias,
                                  float* partials, int64_t ldp, float* scores, void* stream) {
  using namespace gk;
  using namespace gk::umma2;
  const char* name = "gk_tanimoto_rowdot";
  GK_CHECK_ARG(operand_kind == 0 || operand_kind == 1, GK_EINVAL, "%s: operand_kind must be 0 (u8) or 1 (e2m1)", name);
  GK_CHECK_ARG(n >= 0 && m > 0, GK_EINVAL, "%s: bad size", name);
  if (n == 0) return 0;
  GK_CHECK_ARG(alpha && partials && scores, GK_EINVAL, "%s: null pointer", name);
  GK_CHECK_ARG(ldp >= n && ldp % 4 == 0 && ((uintptr_t)partials) % 16 == 0, GK_EALIGN,
               "%s: partials must be 16-byte aligned with ldp >= n and a multiple of 4", name);
  GK_CHECK_ARG(eps >= 1e-12 && eps <= 1.0, GK_ERANGE, "%s: eps outside [1e-12, 1]", name);
  if (int e = check_args(name, a, lda, a_sq, n, b, ldb, b_sq, m, k_bytes, partials, ldp, GK_F32, 1, false)) return e;
  cudaStream_t s = (cudaStream_t)stream;
  int rc;
  // GAUCHE_B200_RB=1|2: accumulator row blocks per CTA of the fused FP4 kernel (experiment switch)
  static const int rb = [] { const char* e = getenv("GAUCHE_B200_RB"); return (e && e[0] == '2') ? 2 : 1; }();
  static const bool mc = [] { const char* e = getenv("GAUCHE_B200_MC"); return e && e[0] == '1'; }();
  const bool ew16 = epi_warps_setting() == 16;
  if (operand_kind == 1 && ew16 && !mc && rb != 2)
    rc = launch_rowdot<1, KindMxf4, 1, 0, 16>(a, lda, a_sq, n, b, ldb, b_sq, m, k_bytes, alpha, eps, partials, ldp, s);
  else if (operand_kind == 0 && ew16)
    rc = launch_rowdot<2, KindI8, 1, 0, 16>(a, lda, a_sq, n, b, ldb, b_sq, m, k_bytes, alpha, eps, partials, ldp, s);
  else if (operand_kind == 1 && mc)
    rc = launch_rowdot<1, KindMxf4, 1, 1>(a, lda, a_sq, n, b, ldb, b_sq, m, k_bytes, alpha, eps, partials, ldp, s);
  else if (operand_kind == 1 && rb == 2)
    rc = launch_rowdot<1, KindMxf4, 2>(a, lda, a_sq, n, b, ldb, b_sq, m, k_bytes, alpha, eps, partials, ldp, s);
  else if (operand_kind == 1)
    rc = launch_rowdot<1, KindMxf4>(a, lda, a_sq, n, b, ldb, b_sq, m, k_bytes, alpha, eps, partials, ldp, s);
  else
    rc = launch_rowdot<2, KindI8>(a, lda, a_sq, n, b, ldb, b_sq, m, k_bytes, alpha, eps, partials, ldp, s);
  if (rc) return rc;
  // slabs actually written by the variant that ran (the workspace query returns the largest)
  const int bn = operand_kind == 1 ? KindMxf4::BN : KindI8::BN;
  const bool ran16 = ew16 && !(operand_kind == 1 && (mc || rb == 2));
  const int64_t slabs = (ran16 ? 4 : 2) * ((m + bn - 1) / bn);
  int64_t blocks = (n + 31) / 32;
  const int64_t cap = (int64_t)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  reduce_partials_kernel<<<(unsigned)blocks, 256, 0, s>>>(partials, slabs, n, ldp, bias, scores);
  return cuda_err(cudaGetLastError(), "reduce_partials_kernel launch");
}
