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
__global__ void __launch_bounds__(256, 1) k(const __grid_constant__ CUtensorMap map, const __grid_constant__ CUtensorMap map2,
  const __grid_constant__ CUtensorMap m3a, const __grid_constant__ CUtensorMap m3b, const __grid_constant__ CUtensorMap m3c, const __grid_constant__ CUtensorMap m3d, float* out, int64_t ldo, int m_tiles,
                                            int n_tiles, int bn, int run_len) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int quad = warp & 3, half = warp >> 2;
  uint8_t* sb = smem + warp * 4608;
  for (int i = lane; i < 1024; i += 32) reinterpret_cast<float*>(sb)[i] = (float)(i + warp);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
  const float4 val = make_float4(1.f, 2.f, 3.f, (float)lane);
  const long total = (long)m_tiles * n_tiles;
  const long total_r = ((long)(n_tiles + run_len - 1) / run_len) * run_len * m_tiles;
  for (long t0 = (long)blockIdx.x * run_len; t0 < total_r; t0 += (long)gridDim.x * run_len)
  for (long t = t0; t < t0 + run_len; ++t) {
    // run_len consecutive column tiles of the same row block are visited back to back by one CTA
    const long grp = t / run_len; const int within = (int)(t % run_len);
    const int mb = (int)(grp % m_tiles), nb = (int)(grp / m_tiles) * run_len + within;
    if (nb >= n_tiles) continue;
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
    } else if (MODE == 8 || MODE == 9) {
      // 3-D tensor maps: each warp owns 16 rows of the tile; four stores per tile: 8 rows x {4 chunks, 3 chunks} of 128 B
      // MODE 8: smem/traversal order [chunk][row][128 B]   MODE 9: [row][chunk][128 B] (512 contiguous bytes back to back)
      const int chunk0 = (int)(col0 / 32);
      for (int c = 0; c < 4; ++c) {
        const int rows_at = (int)row0 + half * 16 + (c & 1) * 8;
        const int ch = chunk0 + (c >> 1) * 4;
        const CUtensorMap* mp = (MODE == 8) ? ((c >> 1) ? &m3b : &m3a) : ((c >> 1) ? &m3d : &m3c);
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncwarp();
        *reinterpret_cast<float4*>(sb + lane * 128 + c * 16) = val;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
          if (MODE == 8)
            asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                         ::"l"(mp), "r"(smem_u32(sb)), "r"(0), "r"(rows_at), "r"(ch) : "memory");
          else
            asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                         ::"l"(mp), "r"(smem_u32(sb)), "r"(0), "r"(ch), "r"(rows_at) : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
    } else if (MODE == 10) {
      // 1-D bulk stores: each warp owns 16 rows; per box 8 lanes issue one cp.async.bulk of 512 (or bn*4-512) contiguous
      // bytes each from a padded (544 B pitch) staging buffer
      for (int c = 0; c < 4; ++c) {
        const int64_t r = row0 + half * 16 + (c & 1) * 8 + lane;
        const int cl = (c >> 1) * 128;
        const uint32_t bytes = (c >> 1) ? (uint32_t)(bn - 128) * 4u : 512u;
        if (lane < 8) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncwarp();
        *reinterpret_cast<float2*>(sb + (lane >> 2) * 544 + (lane & 3) * 8 + c * 32) = make_float2(val.x, val.y);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane < 8) {
          asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                       ::"l"(out + r * ldo + col0 + cl), "r"(smem_u32(sb) + lane * 544), "r"(bytes) : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
    } else if (MODE == 7) {
      // 8 rows x 512 B unswizzled boxes: each warp covers its 32-row strip's column half [half*128, +128) in 4 boxes
      for (int c = 0; c < 4; ++c) {
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncwarp();
        *reinterpret_cast<float4*>(sb + lane * 128 + c * 16) = val;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
          asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                       ::"l"(&map2), "r"(smem_u32(sb)), "r"((int)(col0 + half * 128)), "r"((int)(row0 + c * 8)) : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
    } else if (MODE == 0) {
      const int nst = bn / 32, s0 = (nst + 1) / 2;
      const int c0 = half * s0 * 32, steps = half == 0 ? s0 : nst - s0;
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
      const int nst = bn / 32, s0 = (nst + 1) / 2;
      const int c0 = half * s0 * 32, steps = half == 0 ? s0 : nst - s0;
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
        if (lane * 4 < bn) __stcs(reinterpret_cast<float4*>(p + lane * 4), val);
        if (128 + lane * 4 < bn) __stcs(reinterpret_cast<float4*>(p + 128 + lane * 4), val);
      }
    } else {
      // linear: same bytes per tile (128 x bn x 4), contiguous
      float* p = out + t * (long)(128 * bn);
      for (int i = threadIdx.x * 4; i < 128 * bn; i += 256 * 4) __stcs(reinterpret_cast<float4*>(p + i), val);
    }
  }
  if (MODE == 10) { if (lane < 8) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
  else if (MODE == 0 || MODE >= 4) { if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static CUtensorMap g3[4];
template <int MODE>
void run(const char* name, const CUtensorMap& map, const CUtensorMap& map2, float* out, int64_t n, int64_t m, int bn, int grid, int run_len = 1) {
  auto kern = k<MODE>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 4608 + 1024);
  const int m_tiles = (int)(n / 128), n_tiles = (int)(m / bn);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e9f;
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(e0);
    kern<<<grid, 256, 8 * 4608 + 1024>>>(map, map2, g3[0], g3[1], g3[2], g3[3], out, m, m_tiles, n_tiles, bn, run_len);
    cudaEventRecord(e1);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && ms < best) best = ms;
  }
  const double bytes = (double)m_tiles * n_tiles * 128 * bn * 4;
  printf("%-44s run %2d grid %3d: %.3f ms  %.0f GB/s  %.1f B/clk/SM @1.965GHz  (= %.0f Gpairs/s of fp32 Gram)\n", name, run_len, grid, best,
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
  CUtensorMap map2;
  cuuint32_t box2[2] = {128, 8};
  r = ((EncodeFn)fn)(&map2, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, out, gdim, gstride, box2, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode2 failed %d\n", (int)r); return 1; }
  {
    // 3-D views of the same matrix: (32 floats, rows, chunks of 128 B) and (32 floats, chunks, rows)
    cuuint64_t d_rc[3] = {32, (cuuint64_t)n, (cuuint64_t)(m / 32)}; cuuint64_t s_rc[2] = {(cuuint64_t)m * 4, 128};
    cuuint64_t d_cr[3] = {32, (cuuint64_t)(m / 32), (cuuint64_t)n}; cuuint64_t s_cr[2] = {128, (cuuint64_t)m * 4};
    cuuint32_t es3[3] = {1, 1, 1};
    cuuint32_t b0[3] = {32, 8, 4}, b1[3] = {32, 8, 3}, b2[3] = {32, 4, 8}, b3[3] = {32, 3, 8};
    CUresult r0 = ((EncodeFn)fn)(&g3[0], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, out, d_rc, s_rc, b0, es3, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CUresult r1 = ((EncodeFn)fn)(&g3[1], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, out, d_rc, s_rc, b1, es3, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CUresult r2 = ((EncodeFn)fn)(&g3[2], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, out, d_cr, s_cr, b2, es3, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CUresult r3 = ((EncodeFn)fn)(&g3[3], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, out, d_cr, s_cr, b3, es3, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r0 || r1 || r2 || r3) { printf("3-D encode failed %d %d %d %d\n", (int)r0, (int)r1, (int)r2, (int)r3); return 1; }
  }
  for (int bnv : {128, 160, 192, 224, 256}) {
    const int64_t mm = (m / bnv) * bnv;
    char nm[96];
    snprintf(nm, sizeof nm, "0 TMA 32x128B boxes, bn=%d", bnv); run<0>(nm, map, map2, out, n, mm, bnv, 148);
    snprintf(nm, sizeof nm, "1 STG 128B segments, bn=%d", bnv); run<1>(nm, map, map2, out, n, mm, bnv, 148);
    snprintf(nm, sizeof nm, "2 STG row-contiguous, bn=%d", bnv); run<2>(nm, map, map2, out, n, mm, bnv, 148);
  }
  run<0>("0 TMA 32 rows x 128 B boxes, bn=224", map, map2, out, n, m, 224, 148);
  run<8>("8 TMA 3-D boxes [chunk][8 rows][128B], bn=224", map, map2, out, n, m, 224, 148);
  run<9>("9 TMA 3-D boxes [8 rows][chunk][128B], bn=224", map, map2, out, n, m, 224, 148);
  run<2>("2 STG.128 512B contiguous, bn=224", map, map2, out, n, m, 224, 148);
  run<10>("10 1-D bulk stores 512+384 B per row, bn=224", map, map2, out, n, m, 224, 148);
  run<10>("10 1-D bulk stores 512+512 B per row, bn=256", map, map2, out, n, 390 * 256, 256, 148);
  // tiles of 128 x 256 for the box-shape comparison (m = 99904 is not a multiple of 256: 390 tiles)
  run<0>("0 TMA 32 rows x 128 B boxes, bn=256", map, map2, out, n, 390 * 256, 256, 148);
  run<7>("7 TMA 8 rows x 512 B boxes (no swizzle), bn=256", map, map2, out, n, 390 * 256, 256, 148);
  run<2>("2 STG.128 512B contiguous, bn=256", map, map2, out, n, 390 * 256, 256, 148);
  for (int rl : {2, 4, 8, 16}) {
    run<0>("0 TMA store 32x128B boxes", map, map2, out, n, m, 224, 148, rl);
    run<2>("2 STG.128, 512B contiguous per instruction", map, map2, out, n, m, 224, 148, rl);
  }
  for (int grid : {148, 74}) {
    run<0>("0 TMA store 32x128B boxes", map, map2, out, n, m, 224, grid);
    run<1>("1 STG.128, 128B row segments x4 rows", map, map2, out, n, m, 224, grid);
    run<2>("2 STG.128, 512B contiguous per instruction", map, map2, out, n, m, 224, grid);
    run<3>("3 STG.128 linear", map, map2, out, n, m, 224, grid);
    run<4>("4 TMA boxes, L2 evict_last hint", map, map2, out, n, m, 224, grid);
    run<5>("5 TMA boxes, L2 evict_first hint", map, map2, out, n, m, 224, grid);
    run<6>("6 TMA boxes, a strip's 7 boxes back to back", map, map2, out, n, m, 224, grid);
  }
  return 0;
}
