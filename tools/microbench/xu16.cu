// Microbenchmark: is MUFU.EX2 on fp16 operands (ex2.approx.f16 / f16x2) any faster than on fp32?
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o xu16 xu16.cu && ./xu16
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float* out, long long* cyc, int iters) {
  float a[8];
  uint32_t h[8];
  for (int i = 0; i < 8; ++i) { a[i] = threadIdx.x * 1e-3f + i * 0.01f - 1.f; h[i] = 0xb800b800u + i + threadIdx.x; }
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (MODE == 1) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h[i]));
      if (MODE == 2) { uint16_t x = (uint16_t)h[i]; asm volatile("ex2.approx.f16 %0, %0;" : "+h"(x)); h[i] = x; }
      if (MODE == 3) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(a[i]));
      if (MODE == 4) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
    }
  }
  long long t1 = clock64();
  float s = 0;
  for (int i = 0; i < 8; ++i) s += a[i] + __uint_as_float(h[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int MODE>
void run(const char* name, float* out, long long* cyc, int results_per_instr) {
  const int iters = 2000, threads = 512;
  k<MODE><<<148, threads>>>(out, cyc, iters);
  cudaDeviceSynchronize();
  long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  const double per = double(c) / (iters * 8.0 * (threads / 32 / 4));
  printf("%-24s %.2f cycles per warp instruction per SMSP = %.1f results/clk/SM\n", name, per, 4 * 32.0 * results_per_instr / per);
}
int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
  run<0>("ex2.approx.ftz.f32", out, cyc, 1); run<1>("ex2.approx.f16x2", out, cyc, 2); run<2>("ex2.approx.f16", out, cyc, 1);
  run<3>("tanh.approx.f32", out, cyc, 1); run<4>("rcp.approx.ftz.f32", out, cyc, 1);
  return 0;
}
