// gram_tc.cuh — the tcgen05 Gram kernel of the path: Tanimoto Gram, fused K·alpha screening, MinMax on
// thermometer planes, sibling-kernel epilogues.  Contract: gauche/kernels/fingerprint_kernels/
// tanimoto_kernel.py:36-46 (and minmax_kernel.py:33-44 through the thermometer identity); results are
// bit-identical to the oracle.
//
// One persistent, warp-specialised kernel template, one CTA (or CTA pair) per SM:
//   warp 0        TMA producer
//   warp 1        MMA issuer: tcgen05.mma, two accumulator stages in TMEM
//   warps 2..9    epilogue: tcgen05.ld -> reference-order float math -> swizzled staging -> TMA store (or LSU
//                 stores for outputs whose pitch is not 16-byte aligned); or the K·alpha row sums of screening
//   kBX != 0: 20 warps in five warpgroups — {producer, MMA, 2 idle}, epilogue warps 4..11, bit expanders 12..19 — so that
//   setmaxnreg can move registers from the lean roles to the epilogue (40 / 152 / 64 registers per thread)
// Template parameters
//   kCG   1 single-CTA tiles | 2 cta_group::2 pairs (256-row tiles, each CTA stages half of B)
//   Kind  KindI8: u8 operands, kind::i8, s32 accumulators, 128 x 256 tiles (counts)
//         KindMxf4T<N>: packed e2m1 nibbles, kind::mxf4.block_scale, exact f32 accumulators, 128 x N tiles (N = 224; 192 hybrid)
//   Epi   TanimotoF32x2 (packed fp32 Gram) | TanimotoRowdotF32x2 (fused screening) | Elementwise<...> | ElementwiseSparse[Sym]<...>
//   kBX   the mainloop:
//     0   TMA-FED: the producer TMA-loads ready-made K-major operand tiles (u8 / e2m1, 128B swizzle) into the ring, SS-form MMAs.
//     3   HYBRID (KindMxf4T<192>, kCG = 1; default for fp32 Tanimoto Grams of bit fingerprints).  The A tile stays packed bits
//         all the way into the SM (32 B per row and k-block instead of 128 B): the producer TMA-loads bit tiles into a small
//         ring, expander warps turn every 32-bit word w into the four e2m1 words
//             w & 0x11111111 (nibbles 0.5), w & 0x22222222 (1.0), w & 0x44444444 (2.0), (w>>1) & 0x44444444 (2.0)
//         — one LOP3 each, one shift per input word — and store them into TENSOR MEMORY (tcgen05.st); the MMAs read A from
//         there (TS form).  The four words go to the four 32-byte K segments of the k-block row, i.e. to the four MMAs of the
//         k-block, whose A-side UE8M0 scale factors (2^+1, 2^0, 2^-1, 2^-1) turn every product back into exactly 1.  The bit
//         words arrive pre-permuted (gk_permute_bits_kblock) so that this order equals the natural element order of the B
//         tile, which is TMA-fed ready-made e2m1 as in mainloop 0.  A k-block costs 24 KB of shared memory instead of 44, so
//         seven of them (and six more of A bits) are in flight: the operand pipeline is latency bound and the ring depth
//         is what the kernel lives on (4 / 5 / 6 / 7 stages -> 857 / 954 / 1012 / 1071 Gpairs/s).
//     1   BIT-EXPANDING (KindMxf4, kCG = 1; engine "bx", for libraries that must not keep an e2m1 copy in HBM): BOTH tiles
//         arrive as bits and are expanded into the 128B-swizzled operand ring by the expander warps (generic-proxy stores +
//         fence.proxy.async + mbarrier; scale factors 2 / 1 / 0.5 / 0.5 on both operands, the element order inside a k-block
//         is the same fixed permutation on both sides).  L2 -> SM traffic drops 4x but the shared-memory data pipe saturates.
#pragma once
#include <type_traits>
#include "tc_common.cuh"

namespace gk {
namespace gtc {
using namespace gk::tc;

constexpr int TMEM_COLS = 512;            // whole TMEM: 2 accumulator stages (+ scale factors for mxf4)
constexpr int EPI_WARPS = 8;              // two per TMEM lane quadrant
constexpr int EPI_THREADS = EPI_WARPS * 32;
constexpr int EPI_BUF_BYTES = 32 * 128;   // one 32-row x 128-byte staging buffer per epilogue warp
constexpr int COLTERM_BYTES = 256 * 8;
#ifndef GK_GROUP_M
#define GK_GROUP_M 32
#endif
constexpr int GROUP_M = GK_GROUP_M;       // raster group, measured best of {4..128} (profiles/r1_raster_group_m.log)
#ifndef GK_CONV_WARPS
#define GK_CONV_WARPS 8
#endif
constexpr int CONV_WARPS = GK_CONV_WARPS;  // bit expanders (kBX): teams of four warps (one per TMEM lane quadrant / scheduler)
static_assert(CONV_WARPS == 4 || CONV_WARPS == 8, "one or two teams of four expander warps");
constexpr int CONV_TEAMS = CONV_WARPS / 4;
// TMEM columns of the mxf4 scale factors: 16 columns each, every byte the same UE8M0 value
constexpr int SF_COL_0 = 448, SF_COL_P1 = 464, SF_COL_M1 = 480;
// Bit expansion flavour.  true: nibbles 0.5 / 1.0 / 2.0 / 2.0 per K segment, undone by per-MMA scale factors (5 ALU ops per
// input word).  false (-DGK_BX_UNIT_NIBBLES): every nibble 1.0, unit scale factors (7 ALU ops per input word).
#ifdef GK_BX_UNIT_NIBBLES
constexpr bool kSfTrick = false;
#else
constexpr bool kSfTrick = true;
#endif

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
  static_assert(kBN % 32 == 0 && kBN <= 224, "two accumulator stages + scale factors must fit 512 TMEM columns");
  static constexpr int BN = kBN;          // 224: widest tile that leaves room for the scale factors
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
using KindMxf4 = KindMxf4T<224>;
using KindMxf4N192 = KindMxf4T<192>;      // leaves 64 TMEM columns for the A operand of the hybrid mainloop (kBX = 3)
constexpr int TA_COL = 384;               // kBX = 3: TMEM columns [384, 448) hold two expanded A k-blocks (32 columns each)

// kEpi: 0 fused screening (no output staging), 1 packed fp32 Gram epilogue, 2 element-wise epilogues
template <typename Epi>
struct epi_code { static constexpr int value = Epi::kRowdot ? 0 : (Epi::kPacked ? 1 : 2); };
template <int kCG, typename Kind, int kEpi, int kBX>
struct Cfg {
  static constexpr bool kStaging = kEpi != 0;
  static_assert(!kBX || (kCG == 1 && Kind::kBlockScaled), "bit-expanding mainloop: single-CTA kind::mxf4 tiles");
  static constexpr int BN = Kind::BN;
  static constexpr int EPI_WARP0 = kBX ? 4 : 2;              // first epilogue warp (kBX: roles are warpgroup aligned)
  static constexpr int THREADS = EPI_WARP0 * 32 + EPI_THREADS + (kBX ? CONV_WARPS * 32 : 0);   // 320 | 640
  static constexpr int ACC_STAGES = 2;
  static constexpr int A_BYTES = BM * BK;                   // 16 KB
  static constexpr int B_ROWS = (kCG == 2) ? BN / 2 : BN;   // B rows staged by this CTA
  static constexpr int B_BYTES = B_ROWS * BK;
  static_assert(kBX == 0 || kBX == 1 || kBX == 3, "mainloops: 0 TMA-fed, 1 bit-expanding, 3 hybrid");
  static constexpr bool kHY = (kBX == 3);                   // hybrid: A = packed bits -> TMEM, B = ready-made e2m1 tiles by TMA
  static constexpr bool kTA = kHY;                          // A operand expanded into TMEM (no shared-memory copy of it)
  static_assert(!kTA || 2 * BN <= TA_COL, "kBX = 3: accumulators + A buffers + scale factors share the 512 TMEM columns");
  static constexpr int STAGE_A_BYTES = kTA ? 0 : A_BYTES;
  static constexpr int STAGE_BYTES = STAGE_A_BYTES + B_BYTES;     // one k-block of MMA operands in shared memory
  static constexpr int EPI_BYTES = kStaging ? EPI_WARPS * EPI_BUF_BYTES : 0;
  // column terms: [role group][tile parity][128] x 4 bytes (packed epilogue; + as much again for alpha in the fused one) or
  // x 8 bytes (element-wise epilogues).  kHY spends every spare byte on the B ring: no alpha block for the Gram epilogue, and
  // no 1 KB alignment slack (the dynamic shared-memory array is declared 1024-byte aligned; the kernel traps if it is not).
  static constexpr int COL_BYTES = (kHY && kEpi == 1) ? COLTERM_BYTES : 2 * COLTERM_BYTES;
  static constexpr int ALIGN_SLACK = kHY ? 0 : 1024;
  static constexpr int FIXED_BYTES = EPI_BYTES + COL_BYTES + 256 + ALIGN_SLACK;
  // bit ring (kBX): KBS k-blocks of packed bits per stage = IB bytes per row, TMA swizzle = IB bytes
  static constexpr int KBS = 2;
  static constexpr int IB = 32 * KBS;
  static constexpr int BITS_STAGE_BYTES = (kHY ? BM : BM + B_ROWS) * IB;
  // The TMA pipeline is LATENCY bound (profiles/r1_ring_depth_experiment.log): the ring takes every byte the
  // epilogue does not need.  kBX: two expanded stages (one expander team each) decouple expanders and MMAs, the
  // rest holds bit tiles in flight (each byte of bits stands for four operand bytes).
  // kHY: the B ring is TMA-fed again and takes what is left after HY_BSTAGES bit stages of A (8 KB each = two k-blocks):
  // seven 24 KB B stages + six k-blocks of A bits in flight instead of four 44 KB stages.  The ring depth is what the
  // kernel lives on: 4 / 5 / 6 B stages run the cfg4 Gram at 857 / 954 / 1012 Gpairs/s (profiles/r2_bxh_hybrid_kernel.log).
  static constexpr int XSTAGES = 2;
#ifndef GK_HY_BSTAGES
#define GK_HY_BSTAGES 3
#endif
  static constexpr int HY_BSTAGES = GK_HY_BSTAGES;
  static constexpr int ASLOTS = 2;                          // kTA: expanded A k-blocks in TMEM
#ifdef GK_HY_STAGES          // lab: shallower B ring
  static constexpr int HY_STAGES = GK_HY_STAGES;
#else
  static constexpr int HY_STAGES = (227 * 1024 - FIXED_BYTES - HY_BSTAGES * BITS_STAGE_BYTES) / STAGE_BYTES;
#endif
  static constexpr int STAGES = kHY ? HY_STAGES : kBX ? XSTAGES : (227 * 1024 - FIXED_BYTES) / STAGE_BYTES;
  static constexpr int BSTAGES_FIT = kBX ? (227 * 1024 - FIXED_BYTES - STAGES * STAGE_BYTES) / BITS_STAGE_BYTES : 0;
  static constexpr int BSTAGES = kHY ? HY_BSTAGES : (BSTAGES_FIT > 6 ? 6 : BSTAGES_FIT);
  static_assert(!kBX || BSTAGES >= 2, "bit ring too shallow");
  static constexpr int OFF_RING = 0;
  static constexpr int OFF_BITS = OFF_RING + STAGES * STAGE_BYTES;
  static constexpr int OFF_EPI = OFF_BITS + BSTAGES * BITS_STAGE_BYTES;
  static constexpr int OFF_COL = OFF_EPI + EPI_BYTES;
  static constexpr int OFF_BAR = OFF_COL + COL_BYTES;
  static constexpr int NUM_BARS = 2 * STAGES + 2 * ACC_STAGES + 2 * BSTAGES + (kHY ? 2 * ASLOTS : 0);
  static_assert(NUM_BARS * 8 + 16 <= 256, "barrier block");
  static constexpr int OFF_TMEMPTR = OFF_BAR + NUM_BARS * 8;
  static constexpr int SMEM_BYTES = OFF_TMEMPTR + 16;
  static constexpr int SMEM_ALLOC = SMEM_BYTES + ALIGN_SLACK;
  static constexpr int TILE_M = BM * kCG;                   // rows per (pair-)tile
  static_assert(SMEM_ALLOC <= 227 * 1024, "shared memory budget exceeded");
  static_assert(STAGE_BYTES % 1024 == 0 && BITS_STAGE_BYTES % 1024 == 0 && (BM * IB) % 1024 == 0, "swizzle atoms");
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

// Symmetric Grams (Epi::kSym: x1 is x2 and K[i,j] == K[j,i] bit for bit, i.e. MinMax): tiles that lie entirely below the
// diagonal are skipped by every role; the epilogue writes each computed chunk twice, as it is and mirrored.
template <typename E, typename = void>
struct epi_sym : std::false_type {};
template <typename E>
struct epi_sym<E, std::void_t<decltype(E::kSym)>> : std::bool_constant<E::kSym> {};
template <int TILE_M, int BN>
__device__ __forceinline__ bool tile_below_diagonal(const TileCoord& c) {
  return (int64_t)(c.nb + 1) * BN <= (int64_t)c.mb * TILE_M;     // every column index < every row index
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
// The eight 16-byte chunks of one expanded operand row, stored back to back from 32 DISTINCT registers.  Written as one
// asm block on purpose: with separate stores ptxas recycles the same four data registers for every STS.128, and each
// LOP3 that refills them waits (short scoreboard) until the previous store has read them — ~35 clk per store, ~1000 clk
// per k-block and team (ncu source page of the first bit-expanding build: the expander warps sat on those LOP3s).
__device__ __forceinline__ void sts128x8(const uint32_t (&a)[8], const uint4 (&o)[8]) {
  asm volatile(
      "st.shared.v4.b32 [%0], {%8, %9, %10, %11};\n\t"
      "st.shared.v4.b32 [%1], {%12, %13, %14, %15};\n\t"
      "st.shared.v4.b32 [%2], {%16, %17, %18, %19};\n\t"
      "st.shared.v4.b32 [%3], {%20, %21, %22, %23};\n\t"
      "st.shared.v4.b32 [%4], {%24, %25, %26, %27};\n\t"
      "st.shared.v4.b32 [%5], {%28, %29, %30, %31};\n\t"
      "st.shared.v4.b32 [%6], {%32, %33, %34, %35};\n\t"
      "st.shared.v4.b32 [%7], {%36, %37, %38, %39};"
      ::"r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]),
        "r"(o[0].x), "r"(o[0].y), "r"(o[0].z), "r"(o[0].w), "r"(o[1].x), "r"(o[1].y), "r"(o[1].z), "r"(o[1].w),
        "r"(o[2].x), "r"(o[2].y), "r"(o[2].z), "r"(o[2].w), "r"(o[3].x), "r"(o[3].y), "r"(o[3].z), "r"(o[3].w),
        "r"(o[4].x), "r"(o[4].y), "r"(o[4].z), "r"(o[4].w), "r"(o[5].x), "r"(o[5].y), "r"(o[5].z), "r"(o[5].w),
        "r"(o[6].x), "r"(o[6].y), "r"(o[6].z), "r"(o[6].w), "r"(o[7].x), "r"(o[7].y), "r"(o[7].z), "r"(o[7].w)
      : "memory");
}
// One operand row of a k-block: 256 bits (lo = words 0-3, hi = words 4-7) -> eight 16-byte chunks of e2m1 nibbles.
// K segment s (32 bytes = one MMA) <- bits 4i+s of every word; chunk c = 2s + h holds the words of half h.
__device__ __forceinline__ void expand_row(const uint4& lo, const uint4& hi, uint4 (&o)[8]);
// fill 16 TMEM columns of this warp's lane quadrant with one 32-bit pattern
__device__ __forceinline__ void tmem_fill16(uint32_t taddr, uint32_t pat) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};"
      ::"r"(taddr), "r"(pat)
      : "memory");
}

// Everything the three warp roles share.
template <typename C>
struct Ctx {
  uint8_t* smem;
  uint32_t sbase, tmem_base, cta_rank;
  uint32_t bar_full, bar_empty, bar_tfull, bar_tempty, bar_bfull, bar_bempty, bar_afull, bar_aempty;
  int m_tiles, n_tiles, group_m, k_blocks;
  uint32_t total_tiles, tile0, tile_step;
  int64_t n, m;
  int warp, lane;
};

// Tile index -> coordinates.  Symmetric launches enumerate, raster group by raster group, only the column tiles from the
// group's first live one on (the columns left of it lie below the diagonal for every row tile of the group); the tiles
// of the band where the diagonal crosses the group are still visited and skipped by predicate.  Without this the static
// schedule (tile t, t + stride, ...) leaves the CTAs with very different numbers of live tiles.
template <bool kSym, int TILE_M, int BN>
__device__ __forceinline__ TileCoord tile_coord_t(uint32_t t, int m_tiles, int n_tiles, int group_m) {
  if constexpr (!kSym) {
    return tile_coord(t, m_tiles, n_tiles, group_m);
  } else {
    int first_m = 0;
    for (;;) {
      const int gsz = min(group_m, m_tiles - first_m);
      const int nb_a = (int)(((int64_t)first_m * TILE_M) / BN);       // first column tile that is live for row tile first_m
      const uint32_t cnt = (uint32_t)gsz * (uint32_t)(n_tiles - nb_a);
      if (t < cnt || first_m + group_m >= m_tiles) {
        TileCoord c;
        c.mb = first_m + (int)(t % (uint32_t)gsz);
        c.nb = nb_a + (int)(t / (uint32_t)gsz);
        return c;
      }
      t -= cnt;
      first_m += group_m;
    }
  }
}
template <bool kSym, int TILE_M, int BN>
__device__ __forceinline__ uint32_t total_tiles_t(int m_tiles, int n_tiles, int group_m) {
  if constexpr (!kSym) {
    return (uint32_t)m_tiles * (uint32_t)n_tiles;
  } else {
    uint32_t total = 0;
    for (int first_m = 0; first_m < m_tiles; first_m += group_m) {
      const int gsz = min(group_m, m_tiles - first_m);
      total += (uint32_t)gsz * (uint32_t)(n_tiles - (int)(((int64_t)first_m * TILE_M) / BN));
    }
    return total;
  }
}
// first tile index >= t (in steps of the CTA's stride) that this launch computes
template <bool kSym, typename C, int BN>
__device__ __forceinline__ uint32_t next_live_tile(const Ctx<C>& cx, uint32_t t) {
  if constexpr (kSym) {
    while (t < cx.total_tiles &&
           tile_below_diagonal<C::TILE_M, BN>(tile_coord_t<true, C::TILE_M, BN>(t, cx.m_tiles, cx.n_tiles, cx.group_m)))
      t += cx.tile_step;
  }
  return t;
}

// ------------------------------------------------------------------------------------------------ producer
template <int kCG, typename Kind, typename Epi, int kBX, typename C>
__device__ __forceinline__ void producer_role(const Ctx<C>& cx, const CUtensorMap* map_a, const CUtensorMap* map_b,
                                              const Epi& epi) {
  constexpr int BN = Kind::BN;
  const bool is_leader = (kCG == 1) || cx.cta_rank == 0;
  int stage = 0;
  uint32_t phase = 0;
  if constexpr (C::kHY) {
    // hybrid: A as packed bits (one 8 KB stage = KBS k-blocks of the 128 A rows) for the expanders, B as ready-made e2m1
    // k-block tiles straight into the operand ring.  The bit ring runs further ahead (in k-blocks) than the B ring.
    const int k_stages = cx.k_blocks / C::KBS;
    int bstage = 0;
    uint32_t bphase = 0;
    for (uint32_t t = cx.tile0; t < cx.total_tiles; t += cx.tile_step) {
      const TileCoord tc = tile_coord(t, cx.m_tiles, cx.n_tiles, cx.group_m);
      for (int ks = 0; ks < k_stages; ++ks) {
        mbar_wait(cx.bar_bempty + 8 * bstage, bphase ^ 1);
        w::mbar_expect_tx(cx.bar_bfull + 8 * bstage, C::BITS_STAGE_BYTES);
        w::tma_load_2d<1>(cx.sbase + C::OFF_BITS + bstage * C::BITS_STAGE_BYTES, map_a, cx.bar_bfull + 8 * bstage, ks * C::IB, tc.mb * BM);
        if (++bstage == C::BSTAGES) { bstage = 0; bphase ^= 1; }
#pragma unroll
        for (int j = 0; j < C::KBS; ++j) {
          mbar_wait(cx.bar_empty + 8 * stage, phase ^ 1);
          w::mbar_expect_tx(cx.bar_full + 8 * stage, C::B_BYTES);
          w::tma_load_2d<1>(cx.sbase + C::OFF_RING + stage * C::STAGE_BYTES, map_b, cx.bar_full + 8 * stage,
                            (ks * C::KBS + j) * BK, tc.nb * BN);
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if constexpr (kBX != 0) {
    // packed-bit tiles: one stage = KBS k-blocks = IB bytes of every A and B row of the tile
    const int k_stages = cx.k_blocks / C::KBS;
    for (uint32_t t = cx.tile0; t < cx.total_tiles; t += cx.tile_step) {
      const TileCoord tc = tile_coord(t, cx.m_tiles, cx.n_tiles, cx.group_m);
      for (int ks = 0; ks < k_stages; ++ks) {
        mbar_wait(cx.bar_bempty + 8 * stage, phase ^ 1);
        const uint32_t dst = cx.sbase + C::OFF_BITS + stage * C::BITS_STAGE_BYTES;
        w::mbar_expect_tx(cx.bar_bfull + 8 * stage, C::BITS_STAGE_BYTES);
        w::tma_load_2d<1>(dst, map_a, cx.bar_bfull + 8 * stage, ks * C::IB, tc.mb * BM);
        w::tma_load_2d<1>(dst + BM * C::IB, map_b, cx.bar_bfull + 8 * stage, ks * C::IB, tc.nb * BN);
        if (++stage == C::BSTAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    // transaction bytes of BOTH CTAs of a pair land on the leader's full barrier
    const uint32_t full_base = (kCG == 2) ? mapa(cx.bar_full, 0) : cx.bar_full;
    for (uint32_t t = next_live_tile<epi_sym<Epi>::value, C, BN>(cx, cx.tile0); t < cx.total_tiles;
         t = next_live_tile<epi_sym<Epi>::value, C, BN>(cx, t + cx.tile_step)) {
      const TileCoord tc = tile_coord_t<epi_sym<Epi>::value, C::TILE_M, BN>(t, cx.m_tiles, cx.n_tiles, cx.group_m);
      const int a_row = tc.mb * C::TILE_M + (int)cx.cta_rank * BM;
      const int b_row = tc.nb * BN + (int)cx.cta_rank * C::B_ROWS * (kCG - 1);
      auto load_block = [&](int kb) {
        mbar_wait(cx.bar_empty + 8 * stage, phase ^ 1);
        // all 32 lanes stay converged; tc::w:: wrappers elect the issuing lane inside the asm block
        const uint32_t sa = cx.sbase + C::OFF_RING + stage * C::STAGE_BYTES;
#if defined(GK_TMA_ABLATE)      // lab: fetch the A (bit 0) / B (bit 1) tile of only every fourth k-block (timing of 4x less operand traffic)
        const bool skip_a = (GK_TMA_ABLATE & 1) && (kb & 3), skip_b = (GK_TMA_ABLATE & 2) && (kb & 3);
        if (is_leader) w::mbar_expect_tx(cx.bar_full + 8 * stage, ((skip_a ? 0 : C::A_BYTES) + (skip_b ? 0 : C::B_BYTES)) * kCG);
        if (!skip_a) w::tma_load_2d<kCG>(sa, map_a, full_base + 8 * stage, kb * BK, a_row);
        if (!skip_b) w::tma_load_2d<kCG>(sa + C::A_BYTES, map_b, full_base + 8 * stage, kb * BK, b_row);
#else
        if (is_leader) w::mbar_expect_tx(cx.bar_full + 8 * stage, C::STAGE_BYTES * kCG);
        w::tma_load_2d<kCG>(sa, map_a, full_base + 8 * stage, kb * BK, a_row);
        w::tma_load_2d<kCG>(sa + C::A_BYTES, map_b, full_base + 8 * stage, kb * BK, b_row);
#endif
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
        for (int kb = 0; kb < cx.k_blocks; ++kb) load_block(kb);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ MMA issuer
template <int kCG, typename Kind, typename Epi, int kBX, typename C>
__device__ __forceinline__ void mma_role(const Ctx<C>& cx, const Epi& epi) {
  constexpr int BN = Kind::BN;
  int stage = 0;
  uint32_t phase = 0;
  int acc = 0;
  uint32_t acc_phase = 0;
  int aslot = 0;                 // kHY: TMEM slot of the expanded A k-block (k-blocks alternate between the two slots)
  uint32_t aphase = 0;
  constexpr uint32_t idesc = Kind::template idesc<kCG>();
  for (uint32_t t = next_live_tile<epi_sym<Epi>::value, C, BN>(cx, cx.tile0); t < cx.total_tiles;
       t = next_live_tile<epi_sym<Epi>::value, C, BN>(cx, t + cx.tile_step)) {
    mbar_wait(cx.bar_tempty + 8 * acc, acc_phase ^ 1);   // both CTAs drained this accumulator
    tc_fence_after();
    const uint32_t d_tmem = cx.tmem_base + (uint32_t)(acc * BN);
    auto mma_block = [&](int kb) {
      mbar_wait(cx.bar_full + 8 * stage, phase);          // operands landed (TMA) / expanded (kBX)
      if constexpr (C::kHY) mbar_wait(cx.bar_afull + 8 * aslot, aphase);     // A k-block expanded into TMEM
      tc_fence_after();
      const uint32_t sa = cx.sbase + C::OFF_RING + stage * C::STAGE_BYTES;
      const uint64_t da = make_smem_desc(sa);
      const uint64_t db = make_smem_desc(sa + C::STAGE_A_BYTES);
      const uint32_t ta = cx.tmem_base + (uint32_t)(TA_COL + 32 * aslot);   // kBX = 3: this k-block's A in TMEM
#if defined(GK_BX_ABLATE) && GK_BX_ABLATE == 5       // lab 5: one MMA per k-block instead of four (tensor-core smem reads / 4)
      if (kBX) {
        const uint32_t sf0 = cx.tmem_base + (uint32_t)SF_COL_0;
        if constexpr (Kind::kBlockScaled) w::mma_mxf4<kCG>(d_tmem, da, db, idesc, (uint32_t)(kb != 0), sf0, sf0);
      } else
#endif
#pragma unroll
      for (int kk = 0; kk < BK / UK; ++kk) {
        // k-block 0 is always the first one issued for a tile (also in sparse mode): it overwrites
        if constexpr (Kind::kBlockScaled) {
          // kBX: K segment kk of the row holds nibbles 0.5 / 1.0 / 2.0 / 2.0 -> scales 2 / 1 / 0.5 / 0.5 per operand
          const uint32_t sf = cx.tmem_base + (uint32_t)(!(kBX && kSfTrick) ? SF_COL_0 : (kk == 0 ? SF_COL_P1 : (kk == 1 ? SF_COL_0 : SF_COL_M1)));
          if constexpr (C::kTA)     // B is plain e2m1 1.0 -> unit scale factors on the B side
            w::mma_mxf4_ta(d_tmem, ta + (uint32_t)(kk * 8), db + (uint64_t)(kk * (UK >> 4)), idesc, (uint32_t)((kb | kk) != 0), sf,
                           cx.tmem_base + (uint32_t)SF_COL_0);
          else
            w::mma_mxf4<kCG>(d_tmem, da + (uint64_t)(kk * (UK >> 4)), db + (uint64_t)(kk * (UK >> 4)), idesc,
                             (uint32_t)((kb | kk) != 0), sf, sf);
        } else {
          w::mma_i8<kCG>(d_tmem, da + (uint64_t)(kk * (UK >> 4)), db + (uint64_t)(kk * (UK >> 4)), idesc,
                         (uint32_t)((kb | kk) != 0));
        }
      }
      w::mma_commit<kCG>(cx.bar_empty + 8 * stage);       // frees the ring slot once these MMAs retire
      if constexpr (C::kHY) {
        w::mma_commit<kCG>(cx.bar_aempty + 8 * aslot);    // ... and the A slot in TMEM
        if (++aslot == C::ASLOTS) { aslot = 0; aphase ^= 1; }
      }
      if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
    };
    if constexpr (Epi::kSparse) {
      // same k-block list as the producer: both A and B blocks non-empty, block 0 always
      const TileCoord tc = tile_coord_t<epi_sym<Epi>::value, C::TILE_M, BN>(t, cx.m_tiles, cx.n_tiles, cx.group_m);
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
      for (int kb = 0; kb < cx.k_blocks; ++kb) mma_block(kb);
    }
    w::mma_commit<kCG>(cx.bar_tfull + 8 * acc);   // accumulator complete once every MMA issued so far retires
    if (++acc == C::ACC_STAGES) { acc = 0; acc_phase ^= 1; }
  }
}

__device__ __forceinline__ void expand_row(const uint4& lo, const uint4& hi, uint4 (&o)[8]) {
  auto cls = [](const uint4& v, uint32_t mask, int rsh, int lsh) {
    return make_uint4(((v.x >> rsh) << lsh) & mask, ((v.y >> rsh) << lsh) & mask, ((v.z >> rsh) << lsh) & mask,
                      ((v.w >> rsh) << lsh) & mask);
  };
  if constexpr (kSfTrick) {
    o[0] = cls(lo, 0x11111111u, 0, 0); o[1] = cls(hi, 0x11111111u, 0, 0);     // 0.5
    o[2] = cls(lo, 0x22222222u, 0, 0); o[3] = cls(hi, 0x22222222u, 0, 0);     // 1.0
    o[4] = cls(lo, 0x44444444u, 0, 0); o[5] = cls(hi, 0x44444444u, 0, 0);     // 2.0
    o[6] = cls(lo, 0x44444444u, 1, 0); o[7] = cls(hi, 0x44444444u, 1, 0);     // 2.0
  } else {
    o[0] = cls(lo, 0x22222222u, 0, 1); o[1] = cls(hi, 0x22222222u, 0, 1);     // all 1.0
    o[2] = cls(lo, 0x22222222u, 0, 0); o[3] = cls(hi, 0x22222222u, 0, 0);
    o[4] = cls(lo, 0x22222222u, 1, 0); o[5] = cls(hi, 0x22222222u, 1, 0);
    o[6] = cls(lo, 0x22222222u, 2, 0); o[7] = cls(hi, 0x22222222u, 2, 0);
  }
}
// store the eight chunks of row r (128B-swizzled K-major stage at x_base)
__device__ __forceinline__ void store_row(uint32_t x_base, uint32_t r, const uint4 (&o)[8]) {
  const uint32_t xrow = x_base + r * 128, sw = (r & 7u) << 4;
  uint32_t a[8];
#pragma unroll
  for (uint32_t c = 0; c < 8; ++c) a[c] = xrow + ((c << 4) ^ sw);
  sts128x8(a, o);
}

// ------------------------------------------------------------------------------------------------ bit expanders
// One row and k-block per thread and step: two conflict-free LDS.128 of the row's 32 bytes of bits, 32 LOP3 + 8 SHF,
// eight conflict-free STS.128 into the 128B-swizzled operand stage (lanes = consecutive rows, so the swizzle puts the
// eight lanes of a store phase into eight different 16-byte bank groups).
// The conversion of one k-block is a LATENCY chain (barrier wait -> LDS -> ALU -> STS -> proxy fence -> arrive, ~500 clk
// for a warp that converts 3 of the 11 (= (128 + 224) / 32) row groups), longer than the 448 clk its four MMAs take.  So
// the eight warps form one TEAM PER OPERAND STAGE: team s converts the k-blocks that go to stage s (every STAGES-th
// one), the teams run concurrently on consecutive k-blocks and together deliver one k-block per ~250 clk.  Within a team
// the row groups are dealt round-robin, rotating with the team's turn; the bit words are fetched before the wait for
// the operand stage, so that wait overlaps the MMAs still reading it.
template <typename C>
__device__ __forceinline__ void expander_role(const Ctx<C>& cx) {
  constexpr int GROUPS = (BM + C::B_ROWS) / 32;
  static_assert(C::STAGES == 2, "two operand stages");
  constexpr int WPT = 4;                                   // warps per team
  constexpr int MAXG = (GROUPS + WPT - 1) / WPT;           // row groups per warp and k-block (3 for 11 groups on 4 warps)
  constexpr uint32_t SWZ_MASK = C::IB / 16 - 1;
  const int cw = cx.warp - (C::EPI_WARP0 + EPI_WARPS);
  const int team = cw / WPT, tw = cw % WPT;
  const int k_stages = cx.k_blocks / C::KBS;
  int bstage = 0;
  uint32_t bphase = 0, xphase = 0, kbc = 0, turn = 0;      // xphase: bit s = phase of operand stage s
  auto fetch = [&](uint32_t bits_base, uint32_t r, int j, uint4& lo, uint4& hi) {
    uint32_t o0 = r * C::IB + (uint32_t)(2 * j) * 16, o1 = o0 + 16;
    o0 ^= ((o0 >> 7) & SWZ_MASK) << 4;                     // TMA swizzle of the bit tile
    o1 ^= ((o1 >> 7) & SWZ_MASK) << 4;
    lo = lds128(bits_base + o0);
    hi = lds128(bits_base + o1);
  };
  for (uint32_t t = cx.tile0; t < cx.total_tiles; t += cx.tile_step) {
    for (int ks = 0; ks < k_stages; ++ks) {
      mbar_wait(cx.bar_bfull + 8 * bstage, bphase);
      const uint32_t bits_base = cx.sbase + C::OFF_BITS + bstage * C::BITS_STAGE_BYTES;
#pragma unroll
      for (int j = 0; j < C::KBS; ++j, ++kbc) {
        if ((int)(kbc % CONV_TEAMS) != team) continue;      // another team's k-block (warp-uniform)
        const uint32_t st = kbc % C::STAGES;                // operand stage of this k-block
        const uint32_t x_base = cx.sbase + C::OFF_RING + st * C::STAGE_BYTES;
        const uint32_t bar_xfull = cx.bar_full + 8 * st, bar_xempty = cx.bar_empty + 8 * st;
        const uint32_t g0 = (uint32_t)(tw + turn) % WPT;    // this warp's row groups: g0, g0 + WPT, ...
        uint4 lo, hi;
        fetch(bits_base, g0 * 32 + cx.lane, j, lo, hi);     // first group's bit words: in flight during the wait below
        mbar_wait(bar_xempty, ((xphase >> st) & 1u) ^ 1u);  // the MMAs that read this operand stage have retired
#pragma unroll
        for (int i = 0; i < MAXG; ++i) {
          const uint32_t g = g0 + i * WPT;
          if (g < GROUPS) {                                  // warp-uniform
            uint4 o[8];
            expand_row(lo, hi, o);
            if (g + WPT < GROUPS) fetch(bits_base, (g + WPT) * 32 + cx.lane, j, lo, hi);   // next group's words behind the stores
            store_row(x_base, g * 32 + cx.lane, o);
          }
        }
#if !(defined(GK_BX_ABLATE) && GK_BX_ABLATE == 4)     // lab 4: no proxy fence
        fence_proxy_async_smem();     // generic-proxy stores -> visible to the tensor core's async-proxy reads
#endif
        __syncwarp();
        w::mbar_arrive(bar_xfull);
        xphase ^= 1u << st;
        ++turn;
      }
      __syncwarp();                   // every lane has consumed its bit words
      w::mbar_arrive(cx.bar_bempty + 8 * bstage);
      if (++bstage == C::BSTAGES) { bstage = 0; bphase ^= 1; }
    }
  }
}

// kBX = 3 (hybrid): only the A tile is expanded in the SM — packed bits -> TENSOR MEMORY (tcgen05.st, 32 columns per k-block:
// lane = row, column c = bytes [4c, 4c + 4) of the K-major k-block row; the MMAs read A from there in the TS form) — while the
// B tile arrives as ready-made e2m1 by TMA.  A warp can only touch the TMEM lane quadrant warp % 4.  The A bits are stored PRE-PERMUTED in HBM (gk_permute_bits_kblock) so that the
// five-op expansion yields the natural element order of the e2m1 B operand; the nibble values 0.5 / 1 / 2 / 2 of the four K
// segments are undone by the A-side scale factors alone.  One 32-row group per warp and k-block: the warp of lane quadrant
// q converts A rows [32q, 32q + 32).  k-block kbc goes to TMEM slot kbc % 2; with two teams each team owns one slot.
template <typename C>
__device__ __forceinline__ void expander_role_hy(const Ctx<C>& cx) {
  constexpr int WPT = 4;
  constexpr uint32_t SWZ_MASK = C::IB / 16 - 1;
  const int cw = cx.warp - (C::EPI_WARP0 + EPI_WARPS);
  const int team = cw / WPT, quad = cx.warp & 3;
  const int k_stages = cx.k_blocks / C::KBS;
  int bstage = 0;
  uint32_t bphase = 0, xphase = 0, kbc = 0;      // xphase: bit s = phase of TMEM slot s
  for (uint32_t t = cx.tile0; t < cx.total_tiles; t += cx.tile_step) {
    for (int ks = 0; ks < k_stages; ++ks) {
      mbar_wait(cx.bar_bfull + 8 * bstage, bphase);
      const uint32_t bits_base = cx.sbase + C::OFF_BITS + bstage * C::BITS_STAGE_BYTES;
#pragma unroll
      for (int j = 0; j < C::KBS; ++j, ++kbc) {
        if ((int)(kbc % CONV_TEAMS) != team) continue;      // another team's k-block (warp-uniform)
        const uint32_t st = kbc % C::ASLOTS;
        const uint32_t ta = cx.tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)TA_COL + 32u * st;
        uint32_t o0 = (uint32_t)(quad * 32 + cx.lane) * C::IB + (uint32_t)(2 * j) * 16, o1 = o0 + 16;
        o0 ^= ((o0 >> 7) & SWZ_MASK) << 4;                  // TMA swizzle of the bit tile
        o1 ^= ((o1 >> 7) & SWZ_MASK) << 4;
        const uint4 lo = lds128(bits_base + o0), hi = lds128(bits_base + o1);
        uint4 o[8];
        expand_row(lo, hi, o);
        mbar_wait(cx.bar_aempty + 8 * st, ((xphase >> st) & 1u) ^ 1u);   // the MMAs that read this TMEM slot have retired
        tc_fence_after();
        asm volatile(
            "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
            "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
            "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
            ::"r"(ta), "r"(o[0].x), "r"(o[0].y), "r"(o[0].z), "r"(o[0].w), "r"(o[1].x), "r"(o[1].y), "r"(o[1].z), "r"(o[1].w),
              "r"(o[2].x), "r"(o[2].y), "r"(o[2].z), "r"(o[2].w), "r"(o[3].x), "r"(o[3].y), "r"(o[3].z), "r"(o[3].w),
              "r"(o[4].x), "r"(o[4].y), "r"(o[4].z), "r"(o[4].w), "r"(o[5].x), "r"(o[5].y), "r"(o[5].z), "r"(o[5].w),
              "r"(o[6].x), "r"(o[6].y), "r"(o[6].z), "r"(o[6].w), "r"(o[7].x), "r"(o[7].y), "r"(o[7].z), "r"(o[7].w)
            : "memory");
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        w::mbar_arrive(cx.bar_afull + 8 * st);
        xphase ^= 1u << st;
      }
      __syncwarp();                    // every lane has consumed its bit words
      w::mbar_arrive(cx.bar_bempty + 8 * bstage);
      if (++bstage == C::BSTAGES) { bstage = 0; bphase ^= 1; }
    }
  }
}

// ------------------------------------------------------------------------------------------------ epilogues
// LSU write-back of one staged 32-row x 128-byte box for outputs the TMA cannot address (row pitch or base not
// 16-byte aligned): transposed read-back, 8 lanes per 128-byte row segment, element-wise bounded stores.
template <typename OutT>
__device__ __forceinline__ void store_box_ragged(const uint8_t* sb, OutT* __restrict__ out, int64_t ldo, int64_t row0,
                                                 int64_t cbase, int64_t n, int64_t m, int lane) {
  constexpr int EPC = 16 / (int)sizeof(OutT);
  __syncwarp();
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int r = it * 4 + (lane >> 3), q = lane & 7;
    const int64_t gr = row0 + r, gc = cbase + q * EPC;
    if (gr < n && gc < m) {
      const uint4 val = *reinterpret_cast<const uint4*>(sb + r * 128 + ((q ^ (r & 7)) << 4));
      const OutT* pv = reinterpret_cast<const OutT*>(&val);
      OutT* dst = out + gr * ldo + gc;
#pragma unroll
      for (int e = 0; e < EPC; ++e)
        if (gc + e < m) dst[e] = pv[e];
    }
  }
  __syncwarp();
}

// Packed fp32 epilogue (Tanimoto Gram and fused screening).  Two warps per TMEM lane quadrant; one takes the even
// 32-column chunks of the tile (4 steps), the other the odd ones (3 steps for BN = 224) and the roles swap every tile.
// The four warps with the same role share their column terms through a 128-thread named barrier — identical work, so
// nobody waits for a slower role.  TMEM loads run one step ahead of the float math (two register buffers): the
// accumulator is released after the last load and the load latency hides behind the previous step's math.
template <int kCG, typename Kind, typename Epi, typename C>
__device__ __forceinline__ void epilogue_packed(const Ctx<C>& cx, const CUtensorMap* map_out, const int32_t* __restrict__ na,
                                                const int32_t* __restrict__ nb, typename Epi::OutT* __restrict__ out,
                                                int64_t ldo, int ragged, const Epi& epi) {
  constexpr int BN = Kind::BN;
  constexpr int CS = 32;
  const int lane = cx.lane;
  const int ew = cx.warp - C::EPI_WARP0;        // 0..7
  const int quad = cx.warp & 3;                 // TMEM lane quadrant of this warp
  const int grp = ew >> 2;                      // role group: first / second four epilogue warps
  const int gt = (ew & 3) * 32 + lane;          // thread index inside the group, 0..127
  uint8_t* sb = cx.smem + C::OFF_EPI + ew * EPI_BUF_BYTES;
  const uint32_t sb_u32 = cx.sbase + C::OFF_EPI + ew * EPI_BUF_BYTES;
  const uint32_t tempty_leader = (kCG == 2) ? mapa(cx.bar_tempty, 0) : cx.bar_tempty;
  // column terms: [group][tile parity][128] floats; alpha (fused screening) behind them
  float* terms_base = reinterpret_cast<float*>(cx.smem + C::OFF_COL) + grp * 256;
  float* alpha_base = reinterpret_cast<float*>(cx.smem + C::OFF_COL) + 512 + grp * 256;
  int acc = 0;
  uint32_t acc_phase = 0, it = 0;
  int next_col = 0, next_row = 0;
  float next_alpha = 0.f;
  constexpr bool kSym = epi_sym<Epi>::value;
  TileCoord next_tc = tile_coord(0, cx.m_tiles, cx.n_tiles, cx.group_m);
  auto fetch_norms = [&](uint32_t t, uint32_t iter) {
    next_col = 0;
    next_row = 0;
    next_alpha = 0.f;
    if (t < cx.total_tiles) {
      next_tc = tile_coord_t<kSym, C::TILE_M, BN>(t, cx.m_tiles, cx.n_tiles, cx.group_m);
      const int half = grp ^ (int)(iter & 1);
      const int lc = ((gt >> 5) * 2 + half) * CS + (gt & 31);   // tile-local column whose term this thread converts
      const int64_t gc = (int64_t)next_tc.nb * BN + lc;
      if (lc < BN && gc < cx.m) {
        next_col = __ldg(nb + gc);
        if constexpr (Epi::kRowdot) next_alpha = __ldg(epi.alpha + gc);
      }
      const int64_t gr = (int64_t)next_tc.mb * C::TILE_M + (int64_t)cx.cta_rank * BM + quad * 32 + lane;
      if (gr < cx.n) next_row = __ldg(na + gr);
    }
  };
  uint32_t t = next_live_tile<kSym, C, BN>(cx, cx.tile0);
  fetch_norms(t, 0);
  for (; t < cx.total_tiles; ++it) {
    const uint32_t t_next = next_live_tile<kSym, C, BN>(cx, t + cx.tile_step);
    const TileCoord tc = next_tc;
    const int half = grp ^ (int)(it & 1);       // column half of this warp in this tile
    const int64_t row0 = (int64_t)tc.mb * C::TILE_M + (int64_t)cx.cta_rank * BM + quad * 32;
    const int64_t col0 = (int64_t)tc.nb * BN;
    float* col_terms = terms_base + (it & 1) * 128;
    float* alpha_s = alpha_base + (it & 1) * 128;
    col_terms[gt] = epi.col_term(next_col);
    if constexpr (Epi::kRowdot) alpha_s[gt] = next_alpha;
    const float rterm = epi.row_term(next_row);
    fetch_norms(t_next, it + 1);
    t = t_next;
    // group barrier (ids 1 and 2): terms of this tile visible; the buffer written now was last read two
    // tiles ago, before every warp of the group passed the previous barrier
    if (grp == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
    else asm volatile("bar.sync 2, 128;" ::: "memory");

    // role `half` owns the 32-column chunks half, half + 2, half + 4, ...: the two warps of a quadrant write
    // neighbouring 128-byte lines of the same rows at about the same time
    const int STEPS = (BN / CS + 1 - half) / 2;                         // BN = 224: 4 and 3; BN = 256: 4 and 4
    const bool rows_live = row0 < cx.n;
    const float2 nr2 = make_float2(rterm, rterm);
    float rowacc = 0.f;   // fused screening: this lane's sum_j K[row, j] * alpha[j] over its columns

    mbar_wait(cx.bar_tfull + 8 * acc, acc_phase);
    tc_fence_after();
    const uint32_t taddr = cx.tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BN + half * CS);
    auto release_acc = [&]() {
      tc_fence_before();
      __syncwarp();
      if constexpr (kCG == 2) w::mbar_arrive_cluster(tempty_leader + 8 * acc);
      else w::mbar_arrive(cx.bar_tempty + 8 * acc);
    };
    auto process = [&](const uint32_t (&v)[CS], int c) {
      const int cl = (2 * c + half) * CS;       // first tile-local column of this step
      if (!(rows_live && (col0 + cl < cx.m))) return;   // warp-uniform
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
        // single staging buffer per warp: the previous TMA store must have finished reading it
        w::tma_store_wait_read<0>();
        __syncwarp();
#pragma unroll
        for (int q = 0; q < 8; ++q)
          *reinterpret_cast<float4*>(sb + lane * 128 + ((q ^ (lane & 7)) << 4)) =
              make_float4(res[2 * q].x, res[2 * q].y, res[2 * q + 1].x, res[2 * q + 1].y);
        if (ragged) {
          store_box_ragged<typename Epi::OutT>(sb, out, ldo, row0, col0 + cl, cx.n, cx.m, lane);
        } else {
          fence_proxy_async_smem();     // generic-proxy smem writes -> visible to the TMA engine
          __syncwarp();
          w::tma_store_2d_commit(map_out, sb_u32, (int)(col0 + cl), (int)row0);
        }
        if constexpr (kSym) {
          // the mirrored chunk out[col0 + cl + j][row0 + lane] = K[row0 + lane][col0 + cl + j]: lane = column of the mirrored
          // row, so every column j of the chunk is ONE coalesced 128-byte store straight from the registers
          const int64_t mc = row0 + lane;
          float* dst = out + (col0 + cl) * ldo + mc;
          if (mc < cx.m) {
#pragma unroll
            for (int q = 0; q < 16; ++q) {
              if (col0 + cl + 2 * q < cx.n) __stcs(dst + (int64_t)(2 * q) * ldo, res[q].x);
              if (col0 + cl + 2 * q + 1 < cx.n) __stcs(dst + (int64_t)(2 * q + 1) * ldo, res[q].y);
            }
          }
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
      if (my_row < cx.n) epi.partials[(int64_t)(tc.nb * 2 + half) * epi.ldp + my_row] = rowacc;
    }
    if (++acc == C::ACC_STAGES) { acc = 0; acc_phase ^= 1; }
  }
  if constexpr (!Epi::kRowdot) {
    w::tma_store_wait_read<0>();   // staging smem must outlive the last store's reads
    __syncwarp();
  }
}

// Element-wise epilogue (fp64 / raw-count / MinMax / sibling / exotic-eps outputs): same structure as the packed one —
// two role groups with their own 128-thread barrier and column terms, alternate CS-column chunks per role, roles swapped
// every tile, TMEM loads one step ahead — around the generic `epi.one()` functor.  CS columns = one 128-byte staging
// row: 32 (4-byte outputs) or 16 (fp64: 14 chunks per 224-column tile, 7 per warp).
template <int kCG, typename Kind, typename Epi, typename C>
__device__ __forceinline__ void epilogue_elementwise(const Ctx<C>& cx, const CUtensorMap* map_out,
                                                     const int32_t* __restrict__ na, const int32_t* __restrict__ nb,
                                                     typename Epi::OutT* __restrict__ out, int64_t ldo, int ragged,
                                                     const Epi& epi) {
  using OutT = typename Epi::OutT;
  using TermT = typename Epi::TermT;     // column term
  using RowT = typename Epi::RowT;       // row term (may combine two per-row integers: Epi::kRow2)
  constexpr int BN = Kind::BN;
  constexpr int CS = 128 / (int)sizeof(OutT);
  constexpr int NCH = BN / CS;
  static_assert(CS == 32 || CS == 16, "4- or 8-byte outputs");
  static_assert((NCH + 1) / 2 * CS <= 128, "one column term per thread of a role group");
  static_assert(sizeof(TermT) <= 8, "column terms: 2 groups x 2 parities x 128 x 8 bytes");
  const int lane = cx.lane;
  const int ew = cx.warp - C::EPI_WARP0;
  const int quad = cx.warp & 3;
  const int grp = ew >> 2;
  const int gt = (ew & 3) * 32 + lane;
  uint8_t* sb = cx.smem + C::OFF_EPI + ew * EPI_BUF_BYTES;
  const uint32_t sb_u32 = cx.sbase + C::OFF_EPI + ew * EPI_BUF_BYTES;
  const uint32_t tempty_leader = (kCG == 2) ? mapa(cx.bar_tempty, 0) : cx.bar_tempty;
  TermT* terms_base = reinterpret_cast<TermT*>(cx.smem + C::OFF_COL + grp * 2048);
  int acc = 0;
  uint32_t acc_phase = 0, it = 0;
  int next_col = 0, next_row = 0, next_row2 = 0;
  TileCoord next_tc = tile_coord(0, cx.m_tiles, cx.n_tiles, cx.group_m);
  auto fetch_norms = [&](uint32_t t, uint32_t iter) {
    next_col = 0;
    next_row = 0;
    next_row2 = 0;
    if (t < cx.total_tiles) {
      next_tc = tile_coord_t<epi_sym<Epi>::value, C::TILE_M, BN>(t, cx.m_tiles, cx.n_tiles, cx.group_m);
      const int role = grp ^ (int)(iter & 1);
      const int lc = ((gt / CS) * 2 + role) * CS + (gt % CS);   // tile-local column whose term this thread converts
      const int64_t gc = (int64_t)next_tc.nb * BN + lc;
      if (lc < BN && gc < cx.m) next_col = __ldg(nb + gc);
      const int64_t gr = (int64_t)next_tc.mb * C::TILE_M + (int64_t)cx.cta_rank * BM + quad * 32 + lane;
      if (gr < cx.n) {
        next_row = __ldg(na + gr);
        if constexpr (Epi::kRow2) next_row2 = epi.row2 ? __ldg(epi.row2 + gr) : 0;
      }
    }
  };
  constexpr bool kSym = epi_sym<Epi>::value;
  static_assert(!kSym || CS == 32, "mirrored stores: 4-byte outputs (a transposed chunk is again 32 rows x 128 bytes)");
  uint32_t t = next_live_tile<kSym, C, BN>(cx, cx.tile0);
  fetch_norms(t, 0);
  for (; t < cx.total_tiles; ++it) {
    const uint32_t t_next = next_live_tile<kSym, C, BN>(cx, t + cx.tile_step);
    const TileCoord tc = next_tc;
    const int role = grp ^ (int)(it & 1);
    const int64_t row0 = (int64_t)tc.mb * C::TILE_M + (int64_t)cx.cta_rank * BM + quad * 32;
    const int64_t col0 = (int64_t)tc.nb * BN;
    TermT* col_terms = terms_base + (it & 1) * 128;
    col_terms[gt] = epi.col_term(next_col);
    const RowT rterm = epi.row_term2(next_row, next_row2);
    fetch_norms(t_next, it + 1);
    t = t_next;
    if (grp == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
    else asm volatile("bar.sync 2, 128;" ::: "memory");

    const int STEPS = (NCH + 1 - role) / 2;
    const bool rows_live = row0 < cx.n;
    mbar_wait(cx.bar_tfull + 8 * acc, acc_phase);
    tc_fence_after();
    const uint32_t taddr = cx.tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BN + role * CS);
    auto release_acc = [&]() {
      tc_fence_before();
      __syncwarp();
      if constexpr (kCG == 2) w::mbar_arrive_cluster(tempty_leader + 8 * acc);
      else w::mbar_arrive(cx.bar_tempty + 8 * acc);
    };
    auto load_step = [&](int c, uint32_t (&v)[CS]) {
      if constexpr (CS == 32) tmem_ld32(taddr + c * 2 * CS, v);
      else tmem_ld16(taddr + c * 2 * CS, v);
    };
    auto process = [&](const uint32_t (&v)[CS], int c) {
      const int cl = (2 * c + role) * CS;       // first tile-local column of this step
      if (!(rows_live && (col0 + cl < cx.m))) return;   // warp-uniform
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
      if (ragged) {
        store_box_ragged<OutT>(sb, out, ldo, row0, col0 + cl, cx.n, cx.m, lane);
      } else {
        fence_proxy_async_smem();
        __syncwarp();
        w::tma_store_2d_commit(map_out, sb_u32, (int)(col0 + cl), (int)row0);
      }
      if constexpr (kSym) {
        // the mirrored chunk out[col0 + cl + j][row0 + lane] = res[j]: lane = column of the mirrored row, so every column j
        // of the chunk is ONE coalesced 128-byte store straight from the registers — no second staging pass, no alignment
        // requirement on the output pitch
        const int64_t mc = row0 + lane;
        OutT* dst = out + (col0 + cl) * ldo + mc;
        if (mc < cx.m) {
#pragma unroll
          for (int j = 0; j < CS; ++j)
            if (col0 + cl + j < cx.n) __stcs(dst + (int64_t)j * ldo, res[j]);
        }
      }
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
    if (++acc == C::ACC_STAGES) { acc = 0; acc_phase ^= 1; }
  }
  w::tma_store_wait_read<0>();   // staging smem must outlive the last store's reads
  __syncwarp();
}

// ------------------------------------------------------------------------------------------------ kernel
template <int kCG, typename Kind, typename Epi, int kBX>
__global__ void __launch_bounds__((Cfg<kCG, Kind, epi_code<Epi>::value, kBX>::THREADS), 1)
gram_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
               const __grid_constant__ CUtensorMap map_out, const int32_t* __restrict__ na,
               const int32_t* __restrict__ nb, int64_t n, int64_t m, int k_blocks,
               typename Epi::OutT* __restrict__ out, int64_t ldo, int ragged, int group_m, Epi epi) {
  static_assert(!Epi::kSparse || !kBX, "block-sparse k loop: TMA-fed tiles (occupancy maps per 128 * kCG row tile / BN column tile)");
  using C = Cfg<kCG, Kind, epi_code<Epi>::value, kBX>;
  constexpr int BN = Kind::BN;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  Ctx<C> cx;
  cx.smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  if constexpr (C::ALIGN_SLACK == 0) {
    if (cx.smem != smem_raw) __trap();     // the swizzled stages need the 1024-byte alignment the declaration asks for
  }
  cx.sbase = smem_u32(cx.smem);
  cx.warp = threadIdx.x >> 5;
  cx.lane = threadIdx.x & 31;
  cx.cta_rank = (kCG == 2) ? cluster_ctarank() : 0u;
  cx.bar_full = cx.sbase + C::OFF_BAR;                          // [STAGES]  operand stage ready for the MMAs
  cx.bar_empty = cx.bar_full + C::STAGES * 8;                   // [STAGES]  operand stage consumed
  cx.bar_tfull = cx.bar_empty + C::STAGES * 8;                  // [ACC_STAGES]
  cx.bar_tempty = cx.bar_tfull + C::ACC_STAGES * 8;             // [ACC_STAGES]
  cx.bar_bfull = cx.bar_tempty + C::ACC_STAGES * 8;             // [BSTAGES] bit tile landed (kBX)
  cx.bar_bempty = cx.bar_bfull + C::BSTAGES * 8;                // [BSTAGES] bit tile expanded
  cx.bar_afull = cx.bar_bempty + C::BSTAGES * 8;                // [ASLOTS]  kHY: A k-block expanded into TMEM
  cx.bar_aempty = cx.bar_afull + C::ASLOTS * 8;                 // [ASLOTS]  kHY: TMEM slot consumed
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(cx.smem + C::OFF_TMEMPTR);
  cx.n = n;
  cx.m = m;
  cx.k_blocks = k_blocks;
  cx.group_m = group_m;
  cx.m_tiles = (int)((n + C::TILE_M - 1) / C::TILE_M);
  cx.n_tiles = (int)((m + BN - 1) / BN);
  cx.total_tiles = total_tiles_t<epi_sym<Epi>::value, C::TILE_M, BN>(cx.m_tiles, cx.n_tiles, group_m);
  cx.tile0 = blockIdx.x / kCG;
  cx.tile_step = gridDim.x / kCG;
  const int warp = cx.warp;

  if (warp == 0 && cx.lane == 0) {
    prefetch_tmap(&map_a);
    prefetch_tmap(&map_b);
    if (!Epi::kRowdot && !ragged) prefetch_tmap(&map_out);
    for (int s = 0; s < C::STAGES; ++s) {
      // TMA-fed: the leader's producer arrive(+expect_tx) carries the bytes of both CTAs; kBX: one arrive per warp of the stage's expander team
      mbar_init(cx.bar_full + 8 * s, (kBX && !C::kHY) ? 4 : 1);
      mbar_init(cx.bar_empty + 8 * s, 1);                       // one tcgen05.commit
    }
    for (int a = 0; a < C::ACC_STAGES; ++a) {
      mbar_init(cx.bar_tfull + 8 * a, 1);                       // one (multicast) tcgen05.commit
      mbar_init(cx.bar_tempty + 8 * a, EPI_WARPS * kCG);        // every epilogue warp of the pair
    }
    for (int s = 0; s < C::BSTAGES; ++s) {
      mbar_init(cx.bar_bfull + 8 * s, 1);
      mbar_init(cx.bar_bempty + 8 * s, CONV_WARPS);
    }
    if constexpr (C::kHY) {
      for (int s = 0; s < C::ASLOTS; ++s) {
        mbar_init(cx.bar_afull + 8 * s, 4);                     // the four warps (lane quadrants) of the slot's expander team
        mbar_init(cx.bar_aempty + 8 * s, 1);                    // one tcgen05.commit
      }
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<kCG>(smem_u32(tmem_ptr_smem), TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  cx.tmem_base = *tmem_ptr_smem;
  if constexpr (Kind::kBlockScaled) {
    // constant scale factors: every byte the MMA can read as SFA/SFB is the same UE8M0 value, whatever the
    // scale-factor layout; one epilogue warp per lane quadrant fills columns [448,512)
    if (warp >= C::EPI_WARP0 && warp < C::EPI_WARP0 + 4) {
      const uint32_t lane_base = cx.tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
      tmem_fill16(lane_base + SF_COL_0, 0x7F7F7F7Fu);                          // 2^0
      tmem_fill16(lane_base + SF_COL_P1, (kBX && kSfTrick) ? 0x80808080u : 0x7F7F7F7Fu);     // 2^+1 (bit-expanding mainloop only)
      tmem_fill16(lane_base + SF_COL_M1, (kBX && kSfTrick) ? 0x7E7E7E7Eu : 0x7F7F7F7Fu);     // 2^-1
      tmem_fill16(lane_base + SF_COL_M1 + 16, 0x7F7F7F7Fu);
      tmem_wait_st();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }
  if constexpr (kCG == 2) cluster_sync_all();   // peer barriers + scale factors ready before any remote signal

  // kBX: 640 threads x 96 registers at launch (the CTA's register pool); the epilogue wants ~160, producer / MMA /
  // expanders far fewer, so each warpgroup resizes its share first (setmaxnreg is warpgroup-wide: one instruction per
  // group).  The shares must fit the pool or setmaxnreg.inc waits forever.
#define GK_BX_REGS_LEAN 40
#if GK_CONV_WARPS == 8
#define GK_BX_REGS_EPI 152
#define GK_BX_REGS_EXP 64
#else
#define GK_BX_REGS_EPI 168
#define GK_BX_REGS_EXP 128
#endif
#define GK_STR2(x) #x
#define GK_STR(x) GK_STR2(x)
  static_assert(!kBX || (C::EPI_WARP0 * 32 * GK_BX_REGS_LEAN + EPI_THREADS * GK_BX_REGS_EPI + CONV_WARPS * 32 * GK_BX_REGS_EXP <=
                         C::THREADS * ((65536 / C::THREADS) & ~7)), "setmaxnreg shares exceed the CTA's register pool");
  if (warp < C::EPI_WARP0) {
    if constexpr (kBX) asm volatile("setmaxnreg.dec.sync.aligned.u32 " GK_STR(GK_BX_REGS_LEAN) ";");
    if (warp == 0) {
      producer_role<kCG, Kind, Epi, kBX, C>(cx, &map_a, &map_b, epi);
    } else if (warp == 1) {
      if (kCG == 1 || cx.cta_rank == 0) mma_role<kCG, Kind, Epi, kBX, C>(cx, epi);   // only the leader CTA of a pair issues
    }
  } else if (warp < C::EPI_WARP0 + EPI_WARPS) {
    if constexpr (kBX) asm volatile("setmaxnreg.inc.sync.aligned.u32 " GK_STR(GK_BX_REGS_EPI) ";");
    if constexpr (Epi::kPacked) epilogue_packed<kCG, Kind, Epi, C>(cx, &map_out, na, nb, out, ldo, ragged, epi);
    else epilogue_elementwise<kCG, Kind, Epi, C>(cx, &map_out, na, nb, out, ldo, ragged, epi);
  } else {
    if constexpr (kBX) {
      asm volatile("setmaxnreg.dec.sync.aligned.u32 " GK_STR(GK_BX_REGS_EXP) ";");
      if constexpr (C::kHY) expander_role_hy<C>(cx);
      else expander_role<C>(cx);
    }
  }

  tc_fence_before();
  __syncthreads();
  if constexpr (kCG == 2) cluster_sync_all();   // no CTA of the pair exits while the other may signal it
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<kCG>(cx.tmem_base, TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------ host
// Operand descriptions.  kBX = false: a/b are K-major byte matrices (u8 or packed e2m1), k = row length in bytes
// (multiple of 128).  kBX = true: a/b are packed-bit matrices, k = row length in bytes (multiple of 4).
template <int kCG, typename Kind, int kBX, typename Epi>
static int launch(const void* a, int64_t lda, const int32_t* na, int64_t n, const void* b, int64_t ldb,
                  const int32_t* nb, int64_t m, int64_t k, void* out, int64_t ldo, CUtensorMapDataType out_dt,
                  const Epi& epi, cudaStream_t stream, int64_t k_b = 0) {
  using C = Cfg<kCG, Kind, epi_code<Epi>::value, kBX>;
  constexpr int BN = Kind::BN;
  using OutT = typename Epi::OutT;
  CUtensorMap ma, mb, mo;
  int k_blocks;
  if constexpr (kBX == 3) {
    // hybrid: a = pre-permuted packed bits (k bytes per row), b = packed e2m1 (k_b bytes per row)
    if (int e = make_map(&ma, CU_TENSOR_MAP_DATA_TYPE_UINT8, a, k, n, lda, C::IB, BM, CU_TENSOR_MAP_SWIZZLE_64B)) return e;
    if (int e = make_map(&mb, CU_TENSOR_MAP_DATA_TYPE_UINT8, b, k_b, m, ldb, BK, C::B_ROWS)) return e;
    k_blocks = (int)((k + C::IB - 1) / C::IB) * C::KBS;     // k-blocks past the end of either row are zero-filled by the TMA
  } else if constexpr (kBX != 0) {
    if (int e = make_map(&ma, CU_TENSOR_MAP_DATA_TYPE_UINT8, a, k, n, lda, C::IB, BM, CU_TENSOR_MAP_SWIZZLE_64B)) return e;
    if (int e = make_map(&mb, CU_TENSOR_MAP_DATA_TYPE_UINT8, b, k, m, ldb, C::IB, C::B_ROWS, CU_TENSOR_MAP_SWIZZLE_64B)) return e;
    static_assert(C::IB == 64, "bit-tile swizzle mode follows IB");
    k_blocks = (int)((k + C::IB - 1) / C::IB) * C::KBS;     // whole bit stages; bytes past the row end are zero-filled
  } else {
    if (int e = make_map(&ma, CU_TENSOR_MAP_DATA_TYPE_UINT8, a, k, n, lda, BK, BM)) return e;
    if (int e = make_map(&mb, CU_TENSOR_MAP_DATA_TYPE_UINT8, b, k, m, ldb, BK, C::B_ROWS)) return e;
    k_blocks = (int)(k / BK);
  }
  // outputs the TMA cannot address take the LSU write-back of the epilogue
  const int ragged = !Epi::kRowdot && ((((uintptr_t)out) % 16 != 0) || ((ldo * (int64_t)sizeof(OutT)) % 16 != 0));
  if (Epi::kRowdot || ragged) mo = ma;    // never dereferenced
  else if (int e = make_map(&mo, out_dt, out, m, n, ldo * (int64_t)sizeof(OutT), 128 / (int)sizeof(OutT), 32)) return e;
  auto kern = gram_tc_kernel<kCG, Kind, Epi, kBX>;
  static bool attr_done[64] = {false};
  int dev = 0;
  GK_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || !attr_done[dev]) {
    GK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_ALLOC));
    if (dev >= 0 && dev < 64) attr_done[dev] = true;
  }
  const int64_t tiles = ((n + C::TILE_M - 1) / C::TILE_M) * ((m + BN - 1) / BN);
  int64_t units = sm_count() / kCG;       // persistent: one CTA (pair) per SM (pair)
  if (units > tiles) units = tiles;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(units * kCG));
  cfg.blockDim = dim3(C::THREADS);
  cfg.dynamicSmemBytes = C::SMEM_ALLOC;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kCG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  // raster: GROUP_M row tiles share one column tile before the next column tile starts.  The hybrid mainloop fetches 4 KB
  // of A per k-block instead of 16 KB, so it prefers wider groups (more CTAs on the same B tile at the same time):
  // 32 / 64 / 128 -> 1071 / 1110 / 1108 Gpairs/s on the cfg4 block; the TMA-fed kernels are best at 32.
  const int group_m = (kBX == 3) ? 2 * GROUP_M : GROUP_M;
  GK_CUDA(cudaLaunchKernelEx(&cfg, kern, ma, mb, mo, na, nb, n, m, k_blocks, (OutT*)out, ldo, ragged, group_m, epi));
  return 0;
}

// Shared argument checks of the tensor-core entry points.  `k` = operand row length in bytes.
static inline int check_args(const char* name, const void* a, int64_t lda, const int32_t* a_n, int64_t n, const void* b,
                             int64_t ldb, const int32_t* b_n, int64_t m, int64_t k, int64_t k_align, const void* out,
                             int64_t ldo, int out_dtype) {
  GK_CHECK_ARG(n >= 0 && m >= 0 && k > 0, GK_EINVAL, "%s: bad size", name);
  GK_CHECK_ARG(a && b && a_n && b_n && out, GK_EINVAL, "%s: null pointer", name);
  GK_CHECK_ARG(k % k_align == 0, GK_EALIGN, "%s: k (%lld) must be a multiple of %lld bytes", name, (long long)k,
               (long long)k_align);
  GK_CHECK_ARG(lda >= k && ldb >= k && ldo >= m, GK_EINVAL, "%s: leading dimension too small", name);
  GK_CHECK_ARG(((uintptr_t)a) % 16 == 0 && ((uintptr_t)b) % 16 == 0 && lda % 16 == 0 && ldb % 16 == 0, GK_EALIGN,
               "%s: operands must be 16-byte aligned with row pitches that are multiples of 16 bytes", name);
  const int64_t osz = (out_dtype == GK_F64) ? 8 : 4;
  GK_CHECK_ARG(((uintptr_t)out) % osz == 0, GK_EALIGN, "%s: out must be aligned to its element size", name);
  GK_CHECK_ARG(n < (1LL << 31) && m < (1LL << 31) && k < (1LL << 31), GK_ERANGE, "%s: size too large", name);
  GK_CHECK_ARG(((n + BM - 1) / BM) * ((m + 223) / 224) < (1LL << 31), GK_ERANGE,
               "%s: more than 2^31 output tiles; split the row range", name);
  int64_t cc = 0;
  if (int e = gk_query(GK_Q_CC, &cc)) return e;
  GK_CHECK_ARG(cc / 10 == 10, GK_EDEVICE, "%s: requires an sm_100 device (found sm_%lld)", name, (long long)cc);
  return 0;
}

}  // namespace gtc
}  // namespace gk
