// tc_common.cuh — shared pieces of the tcgen05 Gram kernels (umma2.cu, umma3.cu): inline-PTX wrappers
// for mbarrier / TMA / TMEM / tcgen05.mma, the UMMA shared-memory descriptor, the reference-order
// epilogue functors, and the host-side tensor-map encoder.
#pragma once
#include <cuda.h>
#include <stdlib.h>

#include <mutex>

#include "common.cuh"

namespace gk {
namespace tc {

constexpr int BM = 128;                  // accumulator rows per CTA (TMEM lanes)
constexpr int BK = 128;                  // K bytes per stage = one 128B swizzle atom
constexpr int UK = 32;                   // K bytes per tcgen05.mma (32 x i8 or 64 x e2m1)
constexpr int SF_COL_A = 448, SF_COL_B = 480; // TMEM columns of the (constant) mxf4 scale factors

// --------------------------------------------------------------------------------- PTX
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t mapa(uint32_t local_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// arrive on a barrier that may live in the peer CTA (address from mapa)
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// try_wait that lets the hardware park the thread for up to `ns` nanoseconds while the phase is incomplete
__device__ __forceinline__ bool mbar_try_wait_hint(uint32_t bar, uint32_t parity, uint32_t ns) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity), "r"(ns)
      : "memory");
  return ok != 0;
}
// 0: plain try_wait turns (shipped).  A suspend hint (try_wait + NANOSLEEP.SYNCS) made no measurable difference, burst or
// over 2 s at the power cap (profiles/r2_bxh_hybrid_kernel.log), so the simpler loop stays.
#ifndef GK_MBAR_HINT_NS
#define GK_MBAR_HINT_NS 0
#endif
// Bounded wait: a protocol bug must surface as a trapped launch, never as a hung GPU.  The spin loop is kept to the
// try_wait itself: with twenty warps per CTA, a loop that also read the clock and compared 64-bit times on every turn was
// 30 % of all instructions the hybrid kernel issued (ncu source page) — issue slots and power taken from the epilogue.
// The clock is looked at every 4096 turns only.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  uint32_t turns = 0;
  long long t0 = 0;
#if GK_MBAR_HINT_NS > 0
  while (!mbar_try_wait_hint(bar, parity, GK_MBAR_HINT_NS)) {
#else
  while (!mbar_try_wait(bar, parity)) {
#endif
    if ((++turns & 0xFFFu) == 0) {
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      if (now - t0 > 4000000000LL) {  // ~2 s at 2 GHz
        printf("gauche_b200: mbarrier timeout (block %d thread %d bar 0x%x parity %u)\n",
               (int)blockIdx.x, (int)threadIdx.x, bar, parity);
        __trap();
      }
    }
  }
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
template <int kCG>
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar,
                                            int c0, int c1) {
  if constexpr (kCG == 2) {
    // bar is a shared::cluster address (the leader CTA's full barrier)
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
  } else {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
  }
}
// TMA load multicast to every CTA of the cluster named in `mask`: the tile lands at the same smem offset and
// signals the mbarrier at the same offset in each destination CTA
__device__ __forceinline__ void tma_load_2d_mcast(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1,
                                                  uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}
// single-CTA MMAs, completion signalled on the barrier at the same offset in every CTA of `mask`
__device__ __forceinline__ void mma_commit_mcast(uint32_t bar, uint16_t mask) {
  asm volatile(
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(bar), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(map), "r"(src), "r"(c0), "r"(c1)
               : "memory");
}
// L2 eviction-priority policies for the two streams of the Gram kernel
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void tma_store_2d_hint(const CUtensorMap* map, uint32_t src, int c0, int c1, uint64_t policy) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%2, %3}], [%1], %4;"
               ::"l"(map), "r"(src), "r"(c0), "r"(c1), "l"(policy)
               : "memory");
}
__device__ __forceinline__ void tma_load_2d_hint(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1,
                                                 uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "l"(policy)
      : "memory");
}
__device__ __forceinline__ void tma_store_commit() {
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
template <int N>
__device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
template <int kCG>
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t cols) {
  if constexpr (kCG == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  } else {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
}
template <int kCG>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
  if constexpr (kCG == 2)
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
  else
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
template <int kCG>
__device__ __forceinline__ void mma_i8(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                       uint32_t idesc, uint32_t accumulate) {
  if constexpr (kCG == 2) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}
// D[tmem] (+)= A[smem] * B[smem], e2m1 x e2m1 -> f32, per-32-element UE8M0 scale factors in TMEM
template <int kCG>
__device__ __forceinline__ void mma_mxf4(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                         uint32_t accumulate, uint32_t sfa_tmem, uint32_t sfb_tmem) {
  if constexpr (kCG == 2) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::mxf4.block_scale.block32 [%0], %1, %2, %3, [%5], [%6], p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(sfa_tmem), "r"(sfb_tmem)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::mxf4.block_scale.block32 [%0], %1, %2, %3, [%5], [%6], p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(sfa_tmem), "r"(sfb_tmem)
        : "memory");
  }
}
// fill 32 TMEM columns of this warp's lane quadrant with one 32-bit pattern
__device__ __forceinline__ void tmem_fill32(uint32_t taddr, uint32_t pat) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, "
      "%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};"
      ::"r"(taddr), "r"(pat)
      : "memory");
}
__device__ __forceinline__ void tmem_wait_st() {
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
// arrive (in every CTA of the pair for kCG=2) once all previously issued MMAs have retired
template <int kCG>
__device__ __forceinline__ void mma_commit(uint32_t bar) {
  if constexpr (kCG == 2) {
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
        ::"r"(bar), "h"((uint16_t)3)
        : "memory");
  } else {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
                 : "memory");
  }
}
__device__ __forceinline__ void tmem_wait_ld() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
        "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
// 16 lanes x 256 bit, 4 repetitions = 16 rows x 32 columns.  Thread t = t0 + 4*t1 receives, for repetition g:
// v[4g+0..1] = row t1,   columns 8g + 2*t0 + {0,1};   v[4g+2..3] = row t1 + 8, same columns
// (four neighbouring lanes hold 32 contiguous bytes of one row -> sector-aligned direct global stores).
__device__ __forceinline__ void tmem_ld_16x256b_x4(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.16x256b.x4.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
// 16 lanes x 256 bit, 8 repetitions = 16 rows x 64 columns: repetition g puts row t1 (cols 8g + 2*t0 + {0,1}) in
// v[4g+0..1] and row t1 + 8 in v[4g+2..3]
__device__ __forceinline__ void tmem_ld_16x256b_x8(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.16x256b.x8.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
        "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}

// ------------------------------------------------------------------------------------------------
// Warp-collective issue wrappers: called by ALL 32 lanes of a converged warp; one lane, elected INSIDE the asm
// block, issues the instruction.  ptxas recognises `elect.sync` + `@e <uniform-datapath op>` and emits a bare
// UTMALDG / UTCQMMA / UTCOMMA / UTCBAR / UTMASTG with its operands in uniform registers.  Issuing the same
// instructions under `if (lane == 0)` (or under a separately computed elect predicate) makes ptxas wrap EVERY
// one in an ELECT / PLOP3 / BRA.U.ANY "for each active thread" loop: ~110 SASS instructions and ~550 clk per
// k-block on the MMA warp (profiles/r1_ncu_source_stalls_mxf4x1.txt).  The leader is deterministic for a
// given member mask, so per-thread state (tcgen05.commit tracking, bulk async-groups) stays with one lane.
namespace w {
#define GK_ELECT "{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\t"
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile(GK_ELECT "@e mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n\t}" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile(GK_ELECT "@e mbarrier.arrive.shared::cta.b64 _, [%0];\n\t}" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile(GK_ELECT "@e mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];\n\t}" ::"r"(cluster_addr) : "memory");
}
template <int kCG>
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  if constexpr (kCG == 2) {
    asm volatile(GK_ELECT
        "@e cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4}], [%2];\n\t}"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
  } else {
    asm volatile(GK_ELECT
        "@e cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4}], [%2];\n\t}"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
  }
}
__device__ __forceinline__ void tma_load_2d_mcast(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1,
                                                  uint16_t mask) {
  asm volatile(GK_ELECT
      "@e cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%3, %4}], [%2], %5;\n\t}"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "h"(mask) : "memory");
}
__device__ __forceinline__ void tma_load_2d_hint(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1,
                                                 uint64_t policy) {
  asm volatile(GK_ELECT
      "@e cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5;\n\t}"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "l"(policy) : "memory");
}
template <int kCG>
__device__ __forceinline__ void mma_i8(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                       uint32_t accumulate) {
  if constexpr (kCG == 2) {
    asm volatile(GK_ELECT ".reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "@e tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
  } else {
    asm volatile(GK_ELECT ".reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
  }
}
template <int kCG>
__device__ __forceinline__ void mma_mxf4(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                         uint32_t accumulate, uint32_t sfa_tmem, uint32_t sfb_tmem) {
  if constexpr (kCG == 2) {
    asm volatile(GK_ELECT ".reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "@e tcgen05.mma.cta_group::2.kind::mxf4.block_scale.block32 [%0], %1, %2, %3, [%5], [%6], p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(sfa_tmem), "r"(sfb_tmem) : "memory");
  } else {
    asm volatile(GK_ELECT ".reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::mxf4.block_scale.block32 [%0], %1, %2, %3, [%5], [%6], p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(sfa_tmem), "r"(sfb_tmem) : "memory");
  }
}
// A operand in tensor memory (TS form), B in shared memory; single-CTA
__device__ __forceinline__ void mma_mxf4_ta(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate, uint32_t sfa_tmem, uint32_t sfb_tmem) {
  asm volatile(GK_ELECT ".reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::mxf4.block_scale.block32 [%0], [%1], %2, %3, [%5], [%6], p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(sfa_tmem), "r"(sfb_tmem) : "memory");
}
template <int kCG>
__device__ __forceinline__ void mma_commit(uint32_t bar) {
  if constexpr (kCG == 2) {
    asm volatile(GK_ELECT
        "@e tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n\t}"
        ::"r"(bar), "h"((uint16_t)3) : "memory");
  } else {
    asm volatile(GK_ELECT "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}"
                 ::"r"(bar) : "memory");
  }
}
__device__ __forceinline__ void mma_commit_mcast(uint32_t bar, uint16_t mask) {
  asm volatile(GK_ELECT
      "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n\t}"
      ::"r"(bar), "h"(mask) : "memory");
}
// TMA store of one staging box + commit of its bulk group (one leader, one asm block)
__device__ __forceinline__ void tma_store_2d_commit(const CUtensorMap* map, uint32_t src, int c0, int c1) {
  asm volatile(GK_ELECT
      "@e cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];\n\t"
      "@e cp.async.bulk.commit_group;\n\t}"
      ::"l"(map), "r"(src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_2d_hint_commit(const CUtensorMap* map, uint32_t src, int c0, int c1,
                                                         uint64_t policy) {
  asm volatile(GK_ELECT
      "@e cp.async.bulk.tensor.2d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%2, %3}], [%1], %4;\n\t"
      "@e cp.async.bulk.commit_group;\n\t}"
      ::"l"(map), "r"(src), "r"(c0), "r"(c1), "l"(policy) : "memory");
}
template <int N>
__device__ __forceinline__ void tma_store_wait_read() {
  asm volatile(GK_ELECT "@e cp.async.bulk.wait_group.read %0;\n\t}" ::"n"(N) : "memory");
}
#undef GK_ELECT
}  // namespace w

// K-major 128B-swizzle shared-memory matrix descriptor (see umma.cu for the field map)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// --------------------------------------------------------------------------------- epilogues
// Packed fp32 Tanimoto epilogue.  Terms are kept NEGATED (-(eps+n1), -n2) so that the chain
//   nrc = (-r) + (-c) = -fl(r+c);  nden = nrc + dot = -fl(fl(r+c) - dot)
// needs no negation instruction; every step is the exact negative/identical value of the
// reference-order scalar sequence in common.cuh (TanimotoEpilogueInt + fdiv_rn_normal).
struct TanimotoF32x2 {
  using TermT = float;
  using OutT = float;
  static constexpr bool kPacked = true;
  static constexpr bool kRowdot = false;
  static constexpr bool kSparse = false;
  float eps;
  __device__ __forceinline__ float row_term(int n1) const { return -(eps + (float)n1); }
  __device__ __forceinline__ float col_term(int n2) const { return -(float)n2; }
  template <bool kF32Acc>
  __device__ __forceinline__ float2 pair(uint32_t a0, uint32_t a1, float2 nr, float2 nc) const {
    // accumulators are exact integers: s32 (kind::i8) or integer-valued f32 (kind::mxf4)
    const float2 dot = kF32Acc ? make_float2(__uint_as_float(a0), __uint_as_float(a1))
                               : make_float2((float)(int)a0, (float)(int)a1);
    const float2 nrc = __fadd2_rn(nr, nc);
    const float2 nden = __fadd2_rn(nrc, dot);
    const float2 num = __fadd2_rn(dot, make_float2(eps, eps));
    float2 y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y.x) : "f"(-nden.x));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y.y) : "f"(-nden.y));
    const float2 e = __ffma2_rn(nden, y, make_float2(1.0f, 1.0f));
    y = __ffma2_rn(y, e, y);
    const float2 q = __fmul2_rn(num, y);
    const float2 r = __ffma2_rn(nden, q, num);
    return __ffma2_rn(y, r, q);
  }
};

// Packed fp32 MinMax epilogue over the thermometer-plane accumulator acc = sum_c min(a_c, b_c) (MinMaxFromMinEpilogueFast
// two results at a time).  With every integer below 2^24 (checked by the caller: thermometer bits per row < 2^23) the
// reference's float ops collapse exactly: s = fl(l1_i + l1_j) and d = s - 2 acc are exact integers, so
//   num = fl(fl(s - d) + eps) = fma(acc, 2, eps)          den = fl(fl(s + d) + eps) = fma(s - acc, 2, eps)
// (s + d = 2 (s - acc) is even and < 2^25, hence representable) — four packed instructions before the exact division.
template <bool kSymT>
struct MinMaxSparseF32x2 {
  using TermT = float;
  using OutT = float;
  static constexpr bool kPacked = true;
  static constexpr bool kRowdot = false;
  static constexpr bool kSparse = true;
  static constexpr bool kSym = kSymT;
  float eps;
  const uint64_t* occ_a;
  const uint64_t* occ_b;
  int occ_words;
  __device__ __forceinline__ float row_term(int l1) const { return (float)l1; }
  __device__ __forceinline__ float col_term(int l1) const { return (float)l1; }
  template <bool kF32Acc>
  __device__ __forceinline__ float2 pair(uint32_t a0, uint32_t a1, float2 r, float2 c) const {
    const float2 acc = kF32Acc ? make_float2(__uint_as_float(a0), __uint_as_float(a1))
                               : make_float2((float)(int)a0, (float)(int)a1);
    const float2 two = make_float2(2.0f, 2.0f), e2 = make_float2(eps, eps);
    const float2 s = __fadd2_rn(r, c);
    const float2 sd = __ffma2_rn(acc, make_float2(-1.0f, -1.0f), s);      // s - acc (exact)
    const float2 num = __ffma2_rn(acc, two, e2);
    const float2 den = __ffma2_rn(sd, two, e2);
    float2 y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y.x) : "f"(den.x));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y.y) : "f"(den.y));
    const float2 nden = make_float2(-den.x, -den.y);
    const float2 e = __ffma2_rn(nden, y, make_float2(1.0f, 1.0f));
    y = __ffma2_rn(y, e, y);
    const float2 q = __fmul2_rn(num, y);
    const float2 rr = __ffma2_rn(nden, q, num);
    return __ffma2_rn(y, rr, q);
  }
};

// Fused screening epilogue: instead of writing the Gram tile, every lane accumulates
// sum_j K[i,j] * alpha[j] over the columns it converts and the warp writes one partial per row and
// (n-tile, column-half) slab: the 4 B/pair Gram write and the second pass over it disappear
// (caller side of the path: gauche/gp.py:169-226, K(X*,X) @ alpha).
struct TanimotoRowdotF32x2 : TanimotoF32x2 {
  static constexpr bool kRowdot = true;
  // K[i,j] for the fused K·alpha sum: numerator * rcp.approx(denominator), <= 2 ulp from the correctly rounded
  // quotient of the Gram path.  The sum over ~1e5 columns carries ~sqrt(m) * 2^-24 of rounding anyway, so the
  // Newton step and the residual correction (4 of the 9 packed instructions per column pair) buy nothing here.
  template <bool kF32Acc>
  __device__ __forceinline__ float2 pair_fast(uint32_t a0, uint32_t a1, float2 nr, float2 nc) const {
    const float2 dot = kF32Acc ? make_float2(__uint_as_float(a0), __uint_as_float(a1))
                               : make_float2((float)(int)a0, (float)(int)a1);
    const float2 nrc = __fadd2_rn(nr, nc);
    const float2 nden = __fadd2_rn(nrc, dot);
    const float2 num = __fadd2_rn(dot, make_float2(eps, eps));
    float2 y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y.x) : "f"(-nden.x));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y.y) : "f"(-nden.y));
    return __fmul2_rn(num, y);
  }
  const float* alpha;    // [m]
  float* partials;       // [2 * n_tiles, ldp]
  int64_t ldp;
};

// Element-wise adaptor around the common.cuh functors (fp64 output, raw counts, exotic eps).
template <typename Base, typename Out>
struct Elementwise {
  using TermT = decltype(Base{}.row_term(0));
  using RowT = TermT;
  using OutT = Out;
  static constexpr bool kPacked = false;
  static constexpr bool kRowdot = false;
  static constexpr bool kSparse = false;
  static constexpr bool kRow2 = false;       // no second per-row integer array
  Base base;
  __device__ __forceinline__ TermT row_term(int v) const { return base.row_term(v); }
  __device__ __forceinline__ RowT row_term2(int v, int) const { return base.row_term(v); }
  __device__ __forceinline__ TermT col_term(int v) const { return base.col_term(v); }
  template <bool kF32Acc>
  __device__ __forceinline__ Out one(uint32_t a, TermT r, TermT c) const {
    const int acc = kF32Acc ? __float2int_rn(__uint_as_float(a)) : (int)a;
    return (Out)base(acc, r, c);
  }
};

// Element-wise epilogue + block-occupancy masks of both operands: bit kb of occ_a[mb * occ_words + kb/64] says
// whether the 128-row x 128-byte operand block (row tile mb, k-block kb) holds any non-zero byte (occ_b: BN-row
// tiles).  The producer and the MMA issuer skip every k-block whose A or B block is all zero — it contributes
// nothing to the accumulators.  Thermometer planes of count fingerprints are the use: planes t >= 5 of ECFP counts
// are > 90 % empty blocks (MinMax, gk_minmax_gram_mxf4_sparse).
template <typename Base, typename Out>
struct ElementwiseSparse : Elementwise<Base, Out> {
  static constexpr bool kSparse = true;
  const uint64_t* occ_a;
  const uint64_t* occ_b;
  int occ_words;
};

// ... of a SYMMETRIC problem (x1 is x2, K[i,j] == K[j,i] bit for bit — MinMax: fl(n_i + n_j) commutes and the L1 distance
// is an exact integer): tiles strictly below the diagonal are never computed, every computed chunk is also stored mirrored.
template <typename Base, typename Out>
struct ElementwiseSparseSym : ElementwiseSparse<Base, Out> {
  static constexpr bool kSym = true;
};

// --------------------------------------------------------------------------------- host
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  });
  return fn;
}

static int make_map(CUtensorMap* map, CUtensorMapDataType dt, const void* base, int64_t inner,
                    int64_t rows, int64_t row_stride_bytes, int box_inner, int box_rows,
                    CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return set_err(GK_EDEVICE, "cuTensorMapEncodeTiled driver entry point not available");
  cuuint64_t gdim[2] = {(cuuint64_t)inner, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)row_stride_bytes};
  cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_rows};
  cuuint32_t estride[2] = {1, 1};
  CUresult r = enc(map, dt, 2, (void*)base, gdim, gstride, box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_err(GK_EINVAL, "cuTensorMapEncodeTiled failed (CUresult %d)", (int)r);
  return 0;
}


}  // namespace tc
}  // namespace gk
