// Microbenchmark: issue cost (cycles per warp instruction per SM sub-partition) of the FMA-pipe instructions the
// epilogues and norms are made of: FFMA, FFMA2 (f32x2), FADD2, HFMA2, F2FP (cvt.rn.f16x2.f32), FMNMX, LOP3.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fma fma.cu && ./fma
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
template <int MODE>
__global__ void k(float* out, long long* cyc, int iters) {
  float a[8], b[8];
  uint64_t p[8];
  uint32_t h[8];
  for (int i = 0; i < 8; ++i) { a[i] = threadIdx.x * 1e-3f + i; b[i] = 1.0f + i * 1e-3f; h[i] = 0x3c003c00u + i; }
  for (int i = 0; i < 8; ++i) asm volatile("mov.b64 %0, {%1, %2};" : "=l"(p[i]) : "f"(a[i]), "f"(b[i]));
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b[i]), "f"(b[(i + 1) & 7]));
      if (MODE == 1) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(p[(i + 1) & 7]), "l"(p[(i + 2) & 7]));
      if (MODE == 2) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(p[(i + 1) & 7]));
      if (MODE == 3) asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(h[i]) : "r"(h[(i + 1) & 7]), "r"(h[(i + 2) & 7]));
      if (MODE == 4) asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h[i]) : "f"(a[i]), "f"(b[i]));
      if (MODE == 5) asm volatile("max.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b[i]));
      if (MODE == 6) asm volatile("xor.b32 %0, %0, %1;" : "+r"(h[i]) : "r"(h[(i + 1) & 7]));
      if (MODE == 7) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b[i]));
    }
  }
  long long t1 = clock64();
  float s = 0;
  for (int i = 0; i < 8; ++i) { float x, y; asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(p[i])); s += a[i] + x + y + __uint_as_float(h[i]); }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int MODE>
void run(const char* name, float* out, long long* cyc) {
  const int iters = 4000, threads = 512;
  k<MODE><<<148, threads>>>(out, cyc, iters);
  cudaDeviceSynchronize();
  long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-22s %.2f cycles per warp instruction per SMSP\n", name, double(c) / (iters * 8.0 * (threads / 32 / 4)));
}
int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
  run<0>("FFMA", out, cyc); run<7>("FADD", out, cyc); run<1>("FFMA2 (f32x2)", out, cyc); run<2>("FADD2 (f32x2)", out, cyc);
  run<3>("HFMA2", out, cyc); run<4>("F2FP (cvt f16x2.f32)", out, cyc); run<5>("FMNMX", out, cyc); run<6>("LOP3", out, cyc);
  return 0;
}
