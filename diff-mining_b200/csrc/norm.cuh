// GroupNorm(32 groups)(+SiLU) and LayerNorm over NHWC fp16 activations.  HBM-bound passes:
// 16-byte vector loads/stores, fp32 statistics, one rounding to fp16 at the end -- the same rounding
// point torch.autocast gives the reference (group_norm / layer_norm run in fp32 and their fp32 output is
// cast to fp16 by the next conv / Linear).  GroupNorm reads up to two source tensors so the up-blocks'
// cat(hidden, skip) (reference: /root/reference/diffmining/typicality/dift.py:141-165) is never materialised
// un-normalised; groups may straddle the concat boundary.
#pragma once
#include "ptx.cuh"

namespace dm {

struct NormSrc {
  const __half* ptr;
  int C;                 // channels taken from this source
  long long pix_stride;  // elements between consecutive pixels
};

// partial[n][split][g][2] = (sum, sumsq) over this split's pixels.  Fully deterministic: per-thread register
// accumulation over a fixed pixel sequence, then a fixed-order fold through shared memory (no atomics), then a
// fixed-order sum over splits in gn_apply_kernel -- results are bit-identical run to run and rank to rank.
__global__ void __launch_bounds__(256) gn_stats_kernel(NormSrc s0, NormSrc s1, int HW, int cpg, int splits,
                                                       float* __restrict__ partial) {
  __shared__ float sm_a[2048], sm_q[2048];  // [r][local channel] for one pass over <= 256 vector columns
  const int n = blockIdx.y, split = blockIdx.x;
  const int C = s0.C + s1.C;
  const int vcols = C >> 3;
  const int VT = vcols < (int)blockDim.x ? vcols : (int)blockDim.x;  // vector columns handled concurrently
  const int R = blockDim.x / VT;                                      // pixel rows in flight
  const int r = threadIdx.x / VT, vt = threadIdx.x % VT;
  const int per = (HW + splits - 1) / splits;
  const int p0 = split * per, p1 = min(HW, p0 + per);
  float gs = 0.f, gq = 0.f;  // thread g < 32 owns group g
  for (int vbase = 0; vbase < vcols; vbase += VT) {
    const int v = vbase + vt;
    float a[8], q[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = 0.f; q[i] = 0.f; }
    if (r < R && v < vcols) {
      const int c = v << 3;
      const bool first = c < s0.C;
      const __half* base = first ? s0.ptr + c : s1.ptr + (c - s0.C);
      const long long ps = first ? s0.pix_stride : s1.pix_stride;
      base += static_cast<long long>(n) * HW * ps;
      for (int px = p0 + r; px < p1; px += R) {
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(base + px * ps));
        const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 f = __half22float2(h[i]);
          a[2 * i] += f.x; q[2 * i] += f.x * f.x;
          a[2 * i + 1] += f.y; q[2 * i + 1] += f.y * f.y;
        }
      }
    }
    __syncthreads();  // previous pass fully consumed
    if (r < R) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        sm_a[(r * VT + vt) * 8 + i] = a[i];
        sm_q[(r * VT + vt) * 8 + i] = q[i];
      }
    }
    __syncthreads();
    if (threadIdx.x < 32) {
      // channels of group g that fall inside this pass: [max(g*cpg, cb), min((g+1)*cpg, ce))
      const int cb = vbase << 3, ce = min(C, (vbase + VT) << 3);
      const int lo = max((int)threadIdx.x * cpg, cb), hi = min(((int)threadIdx.x + 1) * cpg, ce);
      for (int c = lo; c < hi; ++c)
        for (int rr = 0; rr < R; ++rr) {
          gs += sm_a[rr * VT * 8 + (c - cb)];
          gq += sm_q[rr * VT * 8 + (c - cb)];
        }
    }
  }
  if (threadIdx.x < 32) {
    float* o = partial + ((static_cast<long long>(n) * splits + split) * 32 + threadIdx.x) * 2;
    o[0] = gs;
    o[1] = gq;
  }
}

// y = (x - mean) * rstd * gamma + beta (+SiLU) -> dense NHWC fp16 [Nimg, HW, C]
__global__ void __launch_bounds__(256) gn_apply_kernel(NormSrc s0, NormSrc s1, int HW, int cpg, int splits,
                                                       const float* __restrict__ partial,
                                                       const float* __restrict__ gamma,
                                                       const float* __restrict__ beta, float eps, int silu,
                                                       int chunks, __half* __restrict__ out) {
  extern __shared__ float2 ab[];  // [C] (scale, shift)
  __shared__ float mean_s[32], rstd_s[32];
  const int n = blockIdx.y;
  const int C = s0.C + s1.C;
  if (threadIdx.x < 32) {
    float s = 0.f, q = 0.f;
    const float* pp = partial + (static_cast<long long>(n) * splits * 32 + threadIdx.x) * 2;
    for (int i = 0; i < splits; ++i) { s += pp[i * 64]; q += pp[i * 64 + 1]; }
    const float cnt = static_cast<float>(HW) * cpg;
    const float mean = s / cnt;
    const float var = fmaxf(q / cnt - mean * mean, 0.f);
    mean_s[threadIdx.x] = mean;
    rstd_s[threadIdx.x] = rsqrtf(var + eps);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const int g = c / cpg;
    const float a = gamma[c] * rstd_s[g];
    ab[c] = make_float2(a, beta[c] - mean_s[g] * a);
  }
  __syncthreads();
  const int vcols = C >> 3;
  const int per = (HW + chunks - 1) / chunks;
  const int p0 = blockIdx.x * per, p1 = min(HW, p0 + per);
  const long long total = static_cast<long long>(p1 - p0) * vcols;
  for (long long idx = threadIdx.x; idx < total; idx += blockDim.x) {
    const int px = p0 + static_cast<int>(idx / vcols);
    const int c = static_cast<int>(idx % vcols) << 3;
    const bool first = c < s0.C;
    const __half* src = first ? s0.ptr + (static_cast<long long>(n) * HW + px) * s0.pix_stride + c
                              : s1.ptr + (static_cast<long long>(n) * HW + px) * s1.pix_stride + (c - s0.C);
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(src));
    const __half2* h = reinterpret_cast<const __half2*>(&u);
    uint32_t pk[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 f = __half22float2(h[i]);
      const float2 ab0 = ab[c + 2 * i], ab1 = ab[c + 2 * i + 1];
      float y0 = f.x * ab0.x + ab0.y, y1 = f.y * ab1.x + ab1.y;
      if (silu) { y0 = silu_f(y0); y1 = silu_f(y1); }
      pk[i] = pack_h2(y0, y1);
    }
    *reinterpret_cast<uint4*>(out + (static_cast<long long>(n) * HW + px) * C + c) =
        make_uint4(pk[0], pk[1], pk[2], pk[3]);
  }
}

// LayerNorm over the last dim (C <= 1280, multiple of 8); one warp per token.
__global__ void __launch_bounds__(256) layernorm_kernel(const __half* __restrict__ x, long long ld_x,
                                                        const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, float eps, long long rows,
                                                        int C, __half* __restrict__ out, long long ld_out) {
  const long long row = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const int vcols = C >> 3;
  constexpr int MAXV = 5;
  float v[MAXV][8];
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < MAXV; ++k) {
    const int vc = lane + 32 * k;
    if (vc < vcols) {
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(x + row * ld_x + (vc << 3)));
      const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 f = __half22float2(h[i]);
        v[k][2 * i] = f.x; v[k][2 * i + 1] = f.y;
        s += f.x + f.y;
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / C;
  float q = 0.f;
#pragma unroll
  for (int k = 0; k < MAXV; ++k) {
    if (lane + 32 * k < vcols) {
#pragma unroll
      for (int i = 0; i < 8; ++i) { const float d = v[k][i] - mean; q += d * d; }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  const float rstd = rsqrtf(q / C + eps);
#pragma unroll
  for (int k = 0; k < MAXV; ++k) {
    const int vc = lane + 32 * k;
    if (vc < vcols) {
      const int c = vc << 3;
      uint32_t pk[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float y0 = (v[k][2 * i] - mean) * rstd * gamma[c + 2 * i] + beta[c + 2 * i];
        const float y1 = (v[k][2 * i + 1] - mean) * rstd * gamma[c + 2 * i + 1] + beta[c + 2 * i + 1];
        pk[i] = pack_h2(y0, y1);
      }
      *reinterpret_cast<uint4*>(out + row * ld_out + c) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    }
  }
}

}  // namespace dm
