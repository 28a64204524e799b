// fp32x2_microbench.cu — issue rates of the packed f32x2 instructions the Gram epilogue uses (FADD2 / FMUL2 /
// FFMA2) against their scalar forms, plus MUFU.RCP, at full occupancy.  Registers only; 8 independent chains.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp32x2_microbench fp32x2_microbench.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

template <int OP>
__global__ void __launch_bounds__(1024) k(float seed, int iters, float* out, long long* cyc) {
  float2 a[8];
  for (int i = 0; i < 8; ++i) a[i] = make_float2(seed * (threadIdx.x + 1) + i, seed + i * 0.5f);
  const float2 b = make_float2(seed * 0.999f, seed * 1.001f);
  const float2 c = make_float2(1e-3f, -1e-3f);
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (OP == 0) a[i] = __ffma2_rn(a[i], b, c);
      if (OP == 1) { a[i].x = __fmaf_rn(a[i].x, b.x, c.x); a[i].y = __fmaf_rn(a[i].y, b.y, c.y); }
      if (OP == 2) a[i] = __fadd2_rn(a[i], c);
      if (OP == 3) { a[i].x = __fadd_rn(a[i].x, c.x); a[i].y = __fadd_rn(a[i].y, c.y); }
      if (OP == 4) a[i] = __fmul2_rn(a[i], b);
      if (OP == 5) { a[i].x = __fmul_rn(a[i].x, b.x); a[i].y = __fmul_rn(a[i].y, b.y); }
      if (OP == 6) { asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a[i].x)); asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a[i].y)); }
    }
  }
  const long long t1 = clock64();
  float s = 0;
  for (int i = 0; i < 8; ++i) s += a[i].x + a[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

template <int OP>
void run(const char* name, int threads) {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 2 * 1024 * 4); cudaMalloc(&cyc, 8);
  const int iters = 4096;
  for (int r = 0; r < 2; ++r) k<OP><<<148, threads>>>(1.0001f, iters, out, cyc);
  cudaDeviceSynchronize();
  long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  const double flane = (double)threads * 8 * 2 * iters / (double)c;   // scalar results per clk per SM
  printf("%-26s %4d thr/SM: %.1f results/clk/SM\n", name, threads, flane);
  cudaFree(out); cudaFree(cyc);
}

int main() {
  for (int threads : {1024, 320, 128}) {
    run<0>("FFMA2 (packed)", threads);
    run<1>("FFMA x2 (scalar)", threads);
    run<2>("FADD2 (packed)", threads);
    run<3>("FADD x2 (scalar)", threads);
    run<4>("FMUL2 (packed)", threads);
    run<5>("FMUL x2 (scalar)", threads);
    run<6>("MUFU.RCP x2", threads);
  }
  return 0;
}
