// Microbenchmark: MUFU.EX2 throughput per SM sub-partition on B200, alone and mixed with FMA-pipe work.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o xu xu.cu && ./xu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
template <int FMA_PER_MUFU>
__global__ void k(float* out, long long* cyc, int iters) {
  float a[8];
  for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-3f + i;
  float f[8];
  for (int i = 0; i < 8; ++i) f[i] = i * 0.5f;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      a[i] = ex2(a[i]);
#pragma unroll
      for (int q = 0; q < FMA_PER_MUFU; ++q) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[(i + q) & 7]) : "f"(1.0001f), "f"(0.5f));
    }
  }
  long long t1 = clock64();
  float s = 0;
  for (int i = 0; i < 8; ++i) s += a[i] + f[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int F>
void run(int threads, float* out, long long* cyc) {
  const int iters = 2000;
  k<F><<<148, threads>>>(out, cyc, iters);
  cudaDeviceSynchronize();
  long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  double warps_per_smsp = threads / 32 / 4.0;
  double mufu_per_smsp = warps_per_smsp * iters * 8;
  printf("threads=%4d fma/mufu=%d cycles=%lld  cycles per warp-MUFU per SMSP=%.2f\n", threads, F, c, c / mufu_per_smsp);
}
int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
  for (int t : {128, 256, 512, 1024}) { run<0>(t, out, cyc); run<1>(t, out, cyc); run<2>(t, out, cyc); run<4>(t, out, cyc); run<8>(t, out, cyc); }
  return 0;
}
