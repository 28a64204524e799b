// store_microbench.cu — how fast can 148 persistent CTAs (8 storing warps each) write an fp32 matrix with a
// 400 KB row pitch, by store method?  No loads, no math.  Each warp owns a 32-row strip of a 128 x 224 tile
// and writes it in 32-column steps, like the Gram epilogue.
//   0: TMA store of a 32 x 128 B swizzled box from a per-warp 4 KB buffer (the kernel's path)
//   1: STG.128, 8 lanes per 128-byte row segment, 4 rows per instruction (the "lsu" path)
//   2: STG.128, the warp covers 512 contiguous bytes of ONE row per instruction (128 columns x 32 rows per 32 instr)
//   3: STG.128, fully linear fill of the same number of bytes
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o store_microbench store_microbench.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int MODE>
__global__ void __launch_bounds__(256, 1) k(const __grid_constant__ CUtensorMap map, float* out, int64_t ldo, int m_tiles,
                                            int n_tiles, int bn) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int quad = warp & 3, half = warp >> 2;
  uint8_t* sb = smem + warp * 4096;
  for (int i = lane; i < 1024; i += 32) reinterpret_cast<float*>(sb)[i] = (float)(i + warp);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
  const float4 val = make_float4(1.f, 2.f, 3.f, (float)lane);
  const long total = (long)m_tiles * n_tiles;
  for (long t = blockIdx.x; t < total; t += gridDim.x) {
    const int mb = (int)(t % m_tiles), nb = (int)(t / m_tiles);
    const int64_t row0 = (int64_t)mb * 128 + quad * 32;
    const int64_t col0 = (int64_t)nb * bn;
    if (MODE == 4 || MODE == 5) {
      uint64_t pol;
      if (MODE == 4) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
      else asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
      const int c0 = half * 128, steps = half == 0 ? 4 : (bn - 128) / 32;
      for (int c = 0; c < steps; ++c) {
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncwarp();
        *reinterpret_cast<float4*>(sb + lane * 128 + ((c ^ (lane & 7)) << 4)) = val;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
          asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%2, %3}], [%1], %4;"
                       ::"l"(&map), "r"(smem_u32(sb)), "r"((int)(col0 + c0 + c * 32)), "r"((int)row0), "l"(pol) : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
    } else if (MODE == 6) {
      // the whole 32-row strip of the tile (7 boxes) issued back to back by ONE warp per quadrant from the same
      // buffer (data is a token): lines of a row segment reach L2 within a few hundred clocks
      if (half == 0) {
        if (lane == 0) {
          asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          for (int c = 0; c < bn / 32; ++c)
            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                         ::"l"(&map), "r"(smem_u32(sb)), "r"((int)(col0 + c * 32)), "r"((int)row0) : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        __syncwarp();
      }
    } else if (MODE == 0) {
      const int c0 = half * 128, steps = half == 0 ? 4 : (bn - 128) / 32;
      for (int c = 0; c < steps; ++c) {
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncwarp();
        *reinterpret_cast<float4*>(sb + lane * 128 + ((c ^ (lane & 7)) << 4)) = val;   // token smem write
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
          asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                       ::"l"(&map), "r"(smem_u32(sb)), "r"((int)(col0 + c0 + c * 32)), "r"((int)row0) : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
    } else if (MODE == 1) {
      const int c0 = half * 128, steps = half == 0 ? 4 : (bn - 128) / 32;
      for (int c = 0; c < steps; ++c)
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int r = it * 4 + (lane >> 3), q = lane & 7;
          __stcs(reinterpret_cast<float4*>(out + (row0 + r) * ldo + col0 + c0 + c * 32 + q * 4), val);
        }
    } else if (MODE == 2) {
      // the two warps of a quadrant split the 32 rows; each instruction writes 512 contiguous bytes of one row
      // (columns [0,128)), then the remaining bn-128 columns with the lanes that fit
      for (int r = half * 16; r < half * 16 + 16; ++r) {
        float* p = out + (row0 + r) * ldo + col0;
        __stcs(reinterpret_cast<float4*>(p + lane * 4), val);
        if (128 + lane * 4 < bn) __stcs(reinterpret_cast<float4*>(p + 128 + lane * 4), val);
      }
    } else {
      // linear: same bytes per tile (128 x bn x 4), contiguous
      float* p = out + t * (long)(128 * bn);
      for (int i = threadIdx.x * 4; i < 128 * bn; i += 256 * 4) __stcs(reinterpret_cast<float4*>(p + i), val);
    }
  }
  if (MODE == 0 || MODE >= 4) { if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int MODE>
void run(const char* name, const CUtensorMap& map, float* out, int64_t n, int64_t m, int bn, int grid) {
  auto kern = k<MODE>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 4096 + 1024);
  const int m_tiles = (int)(n / 128), n_tiles = (int)(m / bn);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e9f;
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(e0);
    kern<<<grid, 256, 8 * 4096 + 1024>>>(map, out, m, m_tiles, n_tiles, bn);
    cudaEventRecord(e1);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && ms < best) best = ms;
  }
  const double bytes = (double)m_tiles * n_tiles * 128 * bn * 4;
  printf("%-44s grid %3d: %.3f ms  %.0f GB/s  %.1f B/clk/SM @1.965GHz  (= %.0f Gpairs/s of fp32 Gram)\n", name, grid, best,
         bytes / best / 1e6, bytes / best / 1e6 / grid / 1.965, bytes / 4 / best / 1e6);
}

int main() {
  const int64_t n = 16384, m = 99904;   // 446 tiles of 224 columns
  float* out; cudaMalloc(&out, n * m * 4);
  void* fn = nullptr; cudaDriverEntryPointQueryResult qres;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  CUtensorMap map;
  cuuint64_t gdim[2] = {(cuuint64_t)m, (cuuint64_t)n}; cuuint64_t gstride[1] = {(cuuint64_t)m * 4};
  cuuint32_t box[2] = {32, 32}; cuuint32_t es[2] = {1, 1};
  CUresult r = ((EncodeFn)fn)(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, out, gdim, gstride, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
  for (int grid : {148, 74}) {
    run<0>("0 TMA store 32x128B boxes", map, out, n, m, 224, grid);
    run<1>("1 STG.128, 128B row segments x4 rows", map, out, n, m, 224, grid);
    run<2>("2 STG.128, 512B contiguous per instruction", map, out, n, m, 224, grid);
    run<3>("3 STG.128 linear", map, out, n, m, 224, grid);
    run<4>("4 TMA boxes, L2 evict_last hint", map, out, n, m, 224, grid);
    run<5>("5 TMA boxes, L2 evict_first hint", map, out, n, m, 224, grid);
    run<6>("6 TMA boxes, a strip's 7 boxes back to back", map, out, n, m, 224, grid);
  }
  return 0;
}
