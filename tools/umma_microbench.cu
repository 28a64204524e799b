// umma_microbench.cu — raw tcgen05.mma issue rate (kind::i8, kind::f16, kind::mxf4.block_scale) on B200 (operands resident in smem, no TMA,
// no epilogue).  Answers: how many SM clocks does one SS-mode UMMA of shape (M=128*cg, N, K=32) take
// when issued back to back?  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_microbench umma_microbench.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ bool try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}

template <int CG, int KIND>  // KIND 0 = i8, 1 = f16(bf16), 2 = mxf4.block_scale (e2m1, UE8M0 scale factors in TMEM), 3 = 2 in the TS form (A operand in TMEM columns [384, 448))
__global__ void __launch_bounds__(128, 1) bench(int n_mma, int N, int stages, long long* out_cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_ptr;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t rank = 0;
  if (CG == 2) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  // deterministic small operand values
  for (int i = threadIdx.x; i < stages * 48 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0x01000100u;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    if (CG == 2) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_ptr)), "r"(512) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_ptr)), "r"(512) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (CG == 2) {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_ptr;
  const int M = 128 * CG;
  if (KIND >= 2) {   // unit scale factors in TMEM columns [448, 512) (KIND 3: and an A operand in columns [384, 448))
    const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
    for (int c = (KIND == 3 ? 384 : 448); c < 512; c += 16)
      asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};"
                   ::"r"(lane_base + c), "r"(c < 448 ? 0x22222222u : 0x7F7F7F7Fu) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }
  uint32_t idesc;
  if (KIND == 0) idesc = (2u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);                      // s32 acc, u8 x u8
  else if (KIND >= 2) idesc = (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | (1u << 23) | ((uint32_t)(M >> 4) << 24);   // e2m1 x e2m1, UE8M0
  else idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);       // f32 acc, bf16 x bf16
  long long t0 = 0, t1 = 0;
  if (warp == 0 && rank == 0) {
    if (lane == 0) {
      t0 = clock64();
      for (int i = 0; i < n_mma; ++i) {
        const int st = (i >> 2) % stages;
        const uint32_t sa = smem_u32(smem) + st * 48 * 1024;
        const uint64_t da = make_desc(sa) + (uint64_t)((i & 3) * 2);
        const uint64_t db = make_desc(sa + 16 * 1024) + (uint64_t)((i & 3) * 2);
        const uint32_t acc = (i != 0);
        if (KIND == 3) {
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::mxf4.block_scale.block32 [%0], [%1], %2, %3, [%5], [%6], p;\n\t}" ::"r"(tmem), "r"(tmem + 384 + (uint32_t)(((i >> 2) & 1) * 32 + (i & 3) * 8)), "l"(db), "r"(idesc), "r"(acc), "r"(tmem + 448), "r"(tmem + 480) : "memory");
        } else if (KIND == 2) {
          if (CG == 2)
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::mxf4.block_scale.block32 [%0], %1, %2, %3, [%5], [%6], p;\n\t}" ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc), "r"(tmem + 448), "r"(tmem + 480) : "memory");
          else
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::mxf4.block_scale.block32 [%0], %1, %2, %3, [%5], [%6], p;\n\t}" ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc), "r"(tmem + 448), "r"(tmem + 480) : "memory");
        } else if (CG == 2) {
          if (KIND == 0)
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
          else
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
        } else {
          if (KIND == 0)
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
          else
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
        }
      }
      if (CG == 2)
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "h"((uint16_t)1) : "memory");
      else
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
      while (!try_wait(smem_u32(&bar), 0)) {}
      t1 = clock64();
      if (blockIdx.x == 0) out_cycles[0] = t1 - t0;
    }
    __syncwarp();
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (CG == 2) {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (CG == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
  }
}

template <int CG, int KIND>
void run(const char* name, int N, int grid, int n_mma = 4096, int reps = 2) {
  const int stages = 4;
  const int smem = stages * 48 * 1024;
  auto k = bench<CG, KIND>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  long long* d;
  cudaMalloc(&d, 8);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(128);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CG; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms = 0.f;
  for (int rep = 0; rep < reps; ++rep) {
    if (rep == 1) cudaEventRecord(e0);     // rep 0 warms up; reps 1.. are timed back to back (wall clock, real SM clocks)
    cudaError_t e = cudaLaunchKernelEx(&cfg, k, n_mma, N, stages, d);
    if (e != cudaSuccess) { printf("%s: error %s\n", name, cudaGetErrorString(e)); return; }
  }
  cudaEventRecord(e1);
  cudaError_t e2 = cudaDeviceSynchronize();
  if (e2 != cudaSuccess) { printf("%s: error %s\n", name, cudaGetErrorString(e2)); return; }
  cudaEventElapsedTime(&ms, e0, e1);
  long long c;
  cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
  const double per = (double)c / n_mma;
  const double macs = (double)128 * CG * N * (KIND == 0 ? 32 : (KIND >= 2 ? 64 : 16));
  const double tops = 2.0 * macs * n_mma * (grid / CG) * (reps - 1) / (ms * 1e-3) / 1e12;
  printf("%-30s grid=%3d N=%3d n_mma=%7d: %.1f clk/MMA -> %.0f MAC/clk/SM ; %d launch(es) %.2f ms -> %.0f TOP/s chip-wide\n", name, grid,
         N, n_mma, per, macs / per / CG, reps - 1, ms, tops);
  cudaFree(d);
}

int main() {
  // kind::mxf4.block_scale (the Tanimoto bit engine): issue rate, then burst (~20 ms) and sustained (~3 s) chip-wide rate
  for (int grid : {1, 148}) {
    run<1, 2>("mxf4 cta_group::1 M=128", 256, grid);
    run<1, 2>("mxf4 cta_group::1 M=128", 224, grid);
    run<1, 2>("mxf4 cta_group::1 M=128", 192, grid);
    run<1, 2>("mxf4 cta_group::1 M=128", 128, grid);
    run<1, 3>("mxf4 TS form (A in TMEM) M=128", 192, grid);
    run<1, 3>("mxf4 TS form (A in TMEM) M=128", 128, grid);
  }
  run<2, 2>("mxf4 cta_group::2 M=256", 256, 148);
  run<2, 2>("mxf4 cta_group::2 M=256", 224, 148);
  run<1, 2>("mxf4 burst  M=128", 224, 148, 400000, 2);
  run<1, 2>("mxf4 sustained M=128", 224, 148, 400000, 101);
  run<1, 0>("i8   burst  M=128", 256, 148, 400000, 2);
  run<1, 0>("i8   sustained M=128", 256, 148, 400000, 101);
  for (int grid : {1, 148}) {
    run<1, 0>("i8  cta_group::1 M=128", 256, grid);
    run<1, 0>("i8  cta_group::1 M=128", 240, grid);
    run<1, 0>("i8  cta_group::1 M=128", 224, grid);
    run<1, 0>("i8  cta_group::1 M=128", 192, grid);
    run<1, 0>("i8  cta_group::1 M=128", 160, grid);
    run<1, 0>("i8  cta_group::1 M=128", 128, grid);
    run<1, 0>("i8  cta_group::1 M=128", 64, grid);
    run<1, 1>("bf16 cta_group::1 M=128", 256, grid);
    run<1, 1>("bf16 cta_group::1 M=128", 128, grid);
  }
  for (int grid : {2, 148}) {
    run<2, 0>("i8  cta_group::2 M=256", 256, grid);
    run<2, 0>("i8  cta_group::2 M=256", 128, grid);
    run<2, 1>("bf16 cta_group::2 M=256", 256, grid);
  }
  return 0;
}
