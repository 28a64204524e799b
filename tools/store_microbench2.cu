// store_microbench2.cu — store-pattern ceiling for the hybrid Gram kernel's tiles: 148 persistent CTAs x 8 storing warps
// write a 16384 x 100000 fp32 matrix (true 400 000-byte pitch) in 128 x 192 tiles, no loads, no math.  Which shape of
// store reaches what share of the linear-fill rate?  (Companion of store_microbench.cu, which used 224-column tiles on a
// 99 904-column matrix.)
//   A  32-row x 128-byte swizzled TMA boxes, the two warps of a lane quadrant take alternate chunks (the kernel's path)
//   B  32 x 256 B unswizzled boxes          C  32 x 384 B unswizzled boxes (columns [96h, 96h + 96))
//   D  16 x 384 B unswizzled boxes          E  8 x 768 B unswizzled boxes (rows 16h .. 16h + 15 of the quadrant)
//   F  16 x 768 B unswizzled boxes
//   G  per-lane 1-D bulk copies of 256 B    H  per-lane 1-D bulk copies of 384 B     I  16 lanes x 768 B 1-D bulk copies
//   J  three 32 x 128 B boxes (alternate chunks) back to back     K  three ADJACENT 32 x 128 B boxes back to back
//   L  linear fill (STG.128)        M  as A with two staging buffers per warp
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o store_microbench2 store_microbench2.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void box_store(const CUtensorMap* m, uint32_t src, int col, int row) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(m), "r"(src), "r"(col), "r"(row) : "memory");
}
__device__ __forceinline__ void commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk1d(float* dst, uint32_t src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}

constexpr int BN = 192;
constexpr int STAGE = 13312;       // per-warp staging bytes (32 rows x (384 + 32))

template <char MODE>
__global__ void __launch_bounds__(256, 1) k(const __grid_constant__ CUtensorMap map, float* out, int64_t ldo, int m_tiles, int n_tiles,
                                            int group_m) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int quad = warp & 3, half = warp >> 2;
  uint8_t* sb = smem + warp * STAGE;
  const uint32_t sbu = smem_u32(sb);
  for (int i = lane; i < STAGE / 4; i += 32) reinterpret_cast<float*>(sb)[i] = (float)(i + warp);
  fence_async();
  __syncwarp();
  const float4 val = make_float4(1.f, 2.f, 3.f, (float)lane);
  const long total = (long)m_tiles * n_tiles;
  long it = 0;
  for (long t = blockIdx.x; t < total; t += gridDim.x, ++it) {
    // the Gram kernel's raster: group_m row tiles share a column tile
    const long per_group = (long)group_m * n_tiles;
    const long g = t / per_group;
    const int first_m = (int)g * group_m;
    const int gsz = min(group_m, m_tiles - first_m);
    const long r = t - g * per_group;
    const int mb = first_m + (int)(r % gsz), nb = (int)(r / gsz);
    const int64_t row0 = (int64_t)mb * 128 + quad * 32;
    const int col0 = nb * BN;
    const int role = half ^ (int)(it & 1);
    auto token = [&](int c) {
      if (lane == 0) wait_read0();
      __syncwarp();
      *reinterpret_cast<float4*>(sb + lane * 128 + ((c & 7) << 4)) = val;
      fence_async();
      __syncwarp();
    };
    if (MODE == 'A') {
      for (int c = role; c < 6; c += 2) {
        token(c);
        if (lane == 0) { box_store(&map, sbu, col0 + c * 32, (int)row0); commit(); }
      }
    } else if (MODE == 'M') {
      // A with TWO staging buffers per warp: the next chunk is staged while the previous store still reads its buffer
      int nbuf = 0;
      for (int c = role; c < 6; c += 2, nbuf ^= 1) {
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        __syncwarp();
        *reinterpret_cast<float4*>(sb + nbuf * 4096 + lane * 128 + ((c & 7) << 4)) = val;
        fence_async();
        __syncwarp();
        if (lane == 0) { box_store(&map, sbu + nbuf * 4096, col0 + c * 32, (int)row0); commit(); }
      }
    } else if (MODE == 'B') {
      for (int p = role; p < 3; p += 2) {     // role 0: column pairs 0 and 2, role 1: pair 1 (roles swap every tile)
        token(p);
        if (lane == 0) { box_store(&map, sbu, col0 + p * 64, (int)row0); commit(); }
      }
    } else if (MODE == 'C') {
      token(0);
      if (lane == 0) { box_store(&map, sbu, col0 + half * 96, (int)row0); commit(); }
    } else if (MODE == 'D') {
      for (int s = 0; s < 2; ++s) {
        token(s);
        if (lane == 0) { box_store(&map, sbu, col0 + half * 96, (int)row0 + s * 16); commit(); }
      }
    } else if (MODE == 'E') {
      for (int s = 0; s < 2; ++s) {
        token(s);
        if (lane == 0) { box_store(&map, sbu, col0, (int)row0 + half * 16 + s * 8); commit(); }
      }
    } else if (MODE == 'F') {
      token(0);
      if (lane == 0) { box_store(&map, sbu, col0, (int)row0 + half * 16); commit(); }
    } else if (MODE == 'G') {
      for (int p = role; p < 3; p += 2) {
        wait_read0();
        __syncwarp();
        *reinterpret_cast<float4*>(sb + lane * 272 + p * 16) = val;
        fence_async();
        __syncwarp();
        bulk1d(out + (row0 + lane) * ldo + col0 + p * 64, sbu + lane * 272, 256);
        commit();
      }
    } else if (MODE == 'H') {
      wait_read0();
      __syncwarp();
      *reinterpret_cast<float4*>(sb + lane * 400) = val;
      fence_async();
      __syncwarp();
      bulk1d(out + (row0 + lane) * ldo + col0 + half * 96, sbu + lane * 400, 384);
      commit();
    } else if (MODE == 'I') {
      if (lane < 16) wait_read0();
      __syncwarp();
      *reinterpret_cast<float4*>(sb + (lane & 15) * 784 + (lane >> 4) * 16) = val;
      fence_async();
      __syncwarp();
      if (lane < 16) {
        bulk1d(out + (row0 + half * 16 + lane) * ldo + col0, sbu + lane * 784, 768);
        commit();
      }
    } else if (MODE == 'J' || MODE == 'K') {
      token(0);
      if (lane == 0) {
        for (int j = 0; j < 3; ++j) {
          const int c = (MODE == 'J') ? role + 2 * j : half * 3 + j;
          box_store(&map, sbu + j * 4096, col0 + c * 32, (int)row0);
        }
        commit();
      }
    } else {
      float* p = out + t * (long)(128 * BN);
      for (int i = threadIdx.x * 4; i < 128 * BN; i += 256 * 4) __stcs(reinterpret_cast<float4*>(p + i), val);
    }
  }
  if (MODE != 'L') wait_read0();
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeFn g_encode;

template <char MODE>
void run(const char* name, float* out, int64_t n, int64_t m, int box_cols, int box_rows, bool swz, int grid, int group_m) {
  CUtensorMap map;
  cuuint64_t gdim[2] = {(cuuint64_t)m, (cuuint64_t)n};
  cuuint64_t gstride[1] = {(cuuint64_t)m * 4};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows}, es[2] = {1, 1};
  CUresult r = g_encode(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, out, gdim, gstride, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        swz ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("%s: encode failed %d\n", name, (int)r); return; }
  auto kern = k<MODE>;
  const int smem = 8 * STAGE + 1024;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int m_tiles = (int)(n / 128), n_tiles = (int)(m / BN);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 1e9f;
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(e0);
    kern<<<grid, 256, smem>>>(map, out, m, m_tiles, n_tiles, group_m);
    cudaEventRecord(e1);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && ms < best) best = ms;
  }
  const double bytes = (double)m_tiles * n_tiles * 128 * BN * 4;
  printf("%-52s grid %3d group_m %3d: %.3f ms  %5.0f GB/s  (= %4.0f Gpairs/s of fp32 Gram)\n", name, grid, group_m, best,
         bytes / best / 1e6, bytes / 4 / best / 1e6);
}

int main() {
  const int64_t n = 16384, m = 100000;
  float* out;
  cudaMalloc(&out, n * m * 4);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  g_encode = (EncodeFn)fn;
  for (int gm : {64, 128, 8}) {
    run<'A'>("A 32 x 128 B swizzled boxes, alternate chunks", out, n, m, 32, 32, true, 148, gm);
    run<'M'>("M as A, two staging buffers per warp", out, n, m, 32, 32, true, 148, gm);
    run<'J'>("J three alternate 32 x 128 B boxes back to back", out, n, m, 32, 32, true, 148, gm);
    run<'K'>("K three adjacent 32 x 128 B boxes back to back", out, n, m, 32, 32, true, 148, gm);
    run<'B'>("B 32 x 256 B unswizzled boxes", out, n, m, 64, 32, false, 148, gm);
    run<'C'>("C 32 x 384 B unswizzled boxes", out, n, m, 96, 32, false, 148, gm);
    run<'D'>("D 16 x 384 B unswizzled boxes", out, n, m, 96, 16, false, 148, gm);
    run<'E'>("E 8 x 768 B unswizzled boxes", out, n, m, 192, 8, false, 148, gm);
    run<'F'>("F 16 x 768 B unswizzled boxes", out, n, m, 192, 16, false, 148, gm);
    run<'G'>("G per-lane 1-D bulk copies, 256 B", out, n, m, 32, 32, true, 148, gm);
    run<'H'>("H per-lane 1-D bulk copies, 384 B", out, n, m, 32, 32, true, 148, gm);
    run<'I'>("I 16 lanes x 768 B 1-D bulk copies", out, n, m, 32, 32, true, 148, gm);
    run<'L'>("L linear fill", out, n, m, 32, 32, true, 148, gm);
  }
  return 0;
}
