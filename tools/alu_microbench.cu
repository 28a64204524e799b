// alu_microbench.cu — issue rates (lanes/clk/SM) of the integer ops behind the CUDA-core Gram kernels:
// VABSDIFF4.U8.ACC (__vsadu4 + add), IDP.4A (__dp4a), POPC (+LOP3+IADD).  Registers only.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

template <int OP>
__global__ void __launch_bounds__(1024) k(uint32_t seed, int iters, uint32_t* out, long long* cyc) {
  uint32_t a[8], acc[8];
  for (int i = 0; i < 8; ++i) { a[i] = seed * (threadIdx.x + 1) + i * 0x01010101u; acc[i] = i; }
  uint32_t b = seed ^ 0x5a5a5a5au;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (OP == 0) acc[i] += __vsadu4(a[i], b);
      if (OP == 1) acc[i] = __dp4a(a[i], b, acc[i]);
      if (OP == 2) acc[i] += __popc(a[i] & b);
    }
    b += 0x01020304u;
  }
  const long long t1 = clock64();
  uint32_t s = 0;
  for (int i = 0; i < 8; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

template <int OP>
void run(const char* name) {
  uint32_t* out; long long* cyc;
  cudaMalloc(&out, 148 * 2 * 1024 * 4); cudaMalloc(&cyc, 8);
  const int iters = 4096;
  for (int r = 0; r < 2; ++r) k<OP><<<148 * 2, 1024>>>(12345u, iters, out, cyc);
  cudaDeviceSynchronize();
  long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  // 2 CTAs x 1024 threads per SM
  const double lanes = 2.0 * 1024 * 8 * iters / (double)c;
  printf("%-22s %.1f lane-ops/clk/SM (%.2f warp-instr/clk/SM)\n", name, lanes, lanes / 32);
  cudaFree(out); cudaFree(cyc);
}

int main() {
  run<0>("VABSDIFF4.U8.ACC");
  run<1>("IDP.4A.U8.U8");
  run<2>("LOP3+POPC+IADD");
  return 0;
}
