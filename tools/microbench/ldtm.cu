// Microbenchmark: tcgen05.ld (TMEM -> registers) throughput per SM, 4 / 8 / 16 warps, 32x32b.x32 and .x16.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../diff-mining_b200/csrc -o ldtm ldtm.cu && ./ldtm
#include <cstdio>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace dm;
template <int X32>
__global__ void k(float* out, long long* cyc, int iters) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t base = slot + (static_cast<uint32_t>((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      uint32_t r[32];
      if (X32) tmem_ld_x32(base + ((c * 32 + it) & 255), r); else { tmem_ld_x16(base + ((c * 16 + it) & 255), r); }
      tmem_wait_ld();
      acc ^= r[0] ^ r[X32 ? 31 : 15];
    }
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = __uint_as_float(acc);
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(slot, 512); }
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int X32>
void run(int threads, float* out, long long* cyc) {
  const int iters = 1000;
  k<X32><<<148, threads>>>(out, cyc, iters);
  cudaError_t e = cudaDeviceSynchronize();
  long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  const double bytes = double(threads) * (X32 ? 32 : 16) * 4 * 8.0 * iters;
  printf("%s threads=%d: %s  %.1f bytes/clk/SM (%.1f cycles per warp instruction)\n", X32 ? "ld.32x32b.x32" : "ld.32x32b.x16", threads,
         cudaGetErrorString(e), bytes / c, double(c) / (8.0 * iters));
}
int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
  for (int t : {128, 256, 512}) { run<1>(t, out, cyc); run<0>(t, out, cyc); }
  return 0;
}
