// GroupNorm(32 groups)(+SiLU) and LayerNorm over NHWC fp16 activations.  HBM-bound streaming passes:
// every thread owns one fixed 16-byte channel vector and walks pixel rows with four independent loads in
// flight; fp32 statistics, one rounding to fp16 at the end -- the same rounding point torch.autocast gives
// the reference (group_norm / layer_norm run in fp32 and their fp32 output is cast to fp16 by the next
// conv / Linear).  GroupNorm reads up to two source tensors so the up-blocks' cat(hidden, skip)
// (reference: /root/reference/diffmining/typicality/dift.py:141-165) is never materialised un-normalised;
// groups may straddle the concat boundary.
#pragma once
#include <cooperative_groups.h>

#include "ptx.cuh"

namespace dm {

struct NormSrc {
  const __half* ptr;
  int C;                 // channels taken from this source
  long long pix_stride;  // elements between consecutive pixels
};

__device__ __forceinline__ float silu_fast(float x) { return __fdividef(x, 1.f + __expf(-x)); }

__device__ __forceinline__ const __half* norm_src_ptr(const NormSrc& s0, const NormSrc& s1, int c, long long& ps) {
  const bool first = c < s0.C;
  ps = first ? s0.pix_stride : s1.pix_stride;
  return first ? s0.ptr + c : s1.ptr + (c - s0.C);
}

// ---- numerically stable statistics.  var = E[x^2] - E[x]^2 cancels when |mean| >> std (real SD-1.5 activations have
// large-mean channels), so every sum is taken about a per-(image, channel) shift k_c = x[n, pixel 0, c]:
//   S_c = sum (x - k_c),  Q_c = sum (x - k_c)^2            (per thread -> per CTA -> per image, fixed order)
//   mean_g = sum_{c in g} (S_c + n k_c) / (n cpg)
//   M2_g   = sum_{c in g} [ Q_c - 2 d_c S_c + n d_c^2 ],  d_c = mean_g - k_c      (= sum (x - mean_g)^2 exactly)
// which is as stable as the Welford update torch's group_norm uses.  All threads / CTAs / splits of an image use the
// same k_c, so the combination stays a plain fixed-order sum: deterministic and batch-invariant as before.
__device__ __forceinline__ void gn_unpack8(const uint4& u, float (&f)[8]) {
  const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 v = __half22float2(h[i]);
    f[2 * i] = v.x;
    f[2 * i + 1] = v.y;
  }
}

// per-thread shifted sums over pixels p0+r, p0+r+R, ... < p1 of one channel vector (UNR independent loads in flight)
template <int UNR>
__device__ __forceinline__ void gn_accumulate(const __half* base, long long ps, int p0, int p1, int r, int R,
                                              const float (&k)[8], float (&a)[8], float (&q)[8]) {
  auto acc = [&](const uint4& u) {
    float f[8];
    gn_unpack8(u, f);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float d = f[i] - k[i];
      a[i] += d;
      q[i] = fmaf(d, d, q[i]);
    }
  };
  int px = p0 + r;
  for (; px + (UNR - 1) * R < p1; px += UNR * R) {
    uint4 u[UNR];
#pragma unroll
    for (int j = 0; j < UNR; ++j) u[j] = __ldg(reinterpret_cast<const uint4*>(base + (px + j * R) * ps));
#pragma unroll
    for (int j = 0; j < UNR; ++j) acc(u[j]);
  }
  for (; px < p1; px += R) acc(__ldg(reinterpret_cast<const uint4*>(base + px * ps)));
}

// shared memory layout of the GroupNorm kernels: [R*C] S | [R*C] Q | [C] shifts   (floats)
__host__ __device__ inline size_t gn_smem_floats(int VT, int R) { return static_cast<size_t>(VT) * 8 * (2 * R + 1); }

// fold the per-thread sums of a CTA into per-channel totals (left in row 0 of sm_a / sm_q).  Caller syncs after.
__device__ __forceinline__ void gn_fold_channels(float* sm_a, float* sm_q, int C, int R) {
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.f, qq = 0.f;
    for (int rr = 0; rr < R; ++rr) {
      s += sm_a[rr * C + c];
      qq += sm_q[rr * C + c];
    }
    sm_a[c] = s;
    sm_q[c] = qq;
  }
}
// group g = threadIdx.x >> 3 is folded by its 8 lanes (fixed xor tree); valid in lane k8 == 0 of the first 256 threads
__device__ __forceinline__ float gn_group_sum(const float* sm_a, const float* sm_k, float cnt, int cpg) {
  const int g = threadIdx.x >> 3, k8 = threadIdx.x & 7;
  float gs = 0.f;
  if (g < 32)
    for (int cc = k8; cc < cpg; cc += 8) gs += sm_a[g * cpg + cc] + cnt * sm_k[g * cpg + cc];
  if (threadIdx.x < 256) {
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) gs += __shfl_xor_sync(0xffffffffu, gs, o);
  }
  return gs;
}
__device__ __forceinline__ float gn_group_m2(const float* sm_a, const float* sm_q, const float* sm_k, const float* mean,
                                             float cnt, int cpg) {
  const int g = threadIdx.x >> 3, k8 = threadIdx.x & 7;
  float m2 = 0.f;
  if (g < 32)
    for (int cc = k8; cc < cpg; cc += 8) {
      const int c = g * cpg + cc;
      const float d = mean[g] - sm_k[c];
      m2 += sm_q[c] - 2.f * d * sm_a[c] + cnt * d * d;
    }
  if (threadIdx.x < 256) {
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) m2 += __shfl_xor_sync(0xffffffffu, m2, o);
  }
  return m2;
}

// Pass 1 of the two-kernel path.  grid (splits, Nimg), block >= max(256, VT*R) threads: thread (r, vt), r < R, owns
// channel vector vt and pixels p0+r, p0+r+R, ... of this split.  partial[n][split][2][C] = per-channel (S_c, Q_c) of the
// split; the LAST block of an image (atomic ticket) folds the splits in index order and writes the per-(image, channel)
// affine ab[n][c] = (gamma*rstd, beta - mean*gamma*rstd).  Deterministic: fixed per-thread pixel sequence, fixed
// shared-memory fold, fixed split order -- no floating-point atomics; the ticket only elects who folds.
// The split geometry depends on (HW, C) alone, so an image's statistics do not depend on the batch it is in.
__global__ void __launch_bounds__(384) gn_stats_kernel(NormSrc s0, NormSrc s1, int HW, int cpg, int splits, int px_per,
                                                        int VT, int R, float* __restrict__ partial,
                                                        const float* __restrict__ gamma, const float* __restrict__ beta,
                                                        float eps, float2* __restrict__ ab, unsigned* __restrict__ tickets) {
  extern __shared__ float sm_gn[];
  __shared__ int is_last;
  __shared__ float stat[64];  // mean[32], rstd[32]
  const int n = blockIdx.y, split = blockIdx.x;
  const int C = s0.C + s1.C;
  const int nthr = VT * R;
  float* sm_a = sm_gn;
  float* sm_q = sm_gn + nthr * 8;
  float* sm_k = sm_gn + 2 * nthr * 8;
  const int r = threadIdx.x / VT, vt = threadIdx.x % VT;
  const int p0 = split * px_per, p1 = min(HW, p0 + px_per);
  const bool active = threadIdx.x < nthr;
  if (active) {
    long long ps;
    const __half* base = norm_src_ptr(s0, s1, vt << 3, ps);
    base += static_cast<long long>(n) * HW * ps;
    float a[8], q[8], k[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = 0.f; q[i] = 0.f; }
    gn_unpack8(__ldg(reinterpret_cast<const uint4*>(base)), k);
    gn_accumulate<4>(base, ps, p0, p1, r, R, k, a, q);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      sm_a[threadIdx.x * 8 + i] = a[i];  // index = (r*VT + vt)*8 + i = r*C + channel
      sm_q[threadIdx.x * 8 + i] = q[i];
      if (r == 0) sm_k[vt * 8 + i] = k[i];
    }
  }
  __syncthreads();
  gn_fold_channels(sm_a, sm_q, C, R);
  __syncthreads();
  {
    float* o = partial + (static_cast<long long>(n) * splits + split) * 2 * C;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      o[c] = sm_a[c];
      o[C + c] = sm_q[c];
    }
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) is_last = (atomicAdd(&tickets[n], 1u) == static_cast<unsigned>(splits - 1));
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  {
    const float* pp = partial + static_cast<long long>(n) * splits * 2 * C;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      float s = 0.f, qq = 0.f;
      for (int i = 0; i < splits; ++i) {
        s += __ldcg(pp + static_cast<long long>(i) * 2 * C + c);
        qq += __ldcg(pp + static_cast<long long>(i) * 2 * C + C + c);
      }
      sm_a[c] = s;
      sm_q[c] = qq;
    }
  }
  __syncthreads();
  const float cnt = static_cast<float>(HW);
  const float gs = gn_group_sum(sm_a, sm_k, cnt, cpg);
  if (threadIdx.x < 256 && (threadIdx.x & 7) == 0) stat[threadIdx.x >> 3] = gs / (cnt * cpg);
  __syncthreads();
  const float m2 = gn_group_m2(sm_a, sm_q, sm_k, stat, cnt, cpg);
  if (threadIdx.x < 256 && (threadIdx.x & 7) == 0) stat[32 + (threadIdx.x >> 3)] = rsqrtf(fmaxf(m2 / (cnt * cpg), 0.f) + eps);
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const int g = c / cpg;
    const float sc = gamma[c] * stat[32 + g];
    ab[static_cast<long long>(n) * C + c] = make_float2(sc, beta[c] - stat[g] * sc);
  }
  if (threadIdx.x == 0) tickets[n] = 0;  // self-cleaning for the next GroupNorm on this stream
}

// Pass 2.  y = x * a + b (+SiLU) -> dense NHWC fp16 [Nimg, HW, C]; grid (pixel blocks, Nimg), same thread map.
__global__ void __launch_bounds__(384) gn_apply_kernel(NormSrc s0, NormSrc s1, int HW, int px_per, int VT, int R,
                                                        const float2* __restrict__ ab, int silu,
                                                        __half* __restrict__ out) {
  const int n = blockIdx.y;
  const int C = s0.C + s1.C;
  if (threadIdx.x >= VT * R) return;
  const int r = threadIdx.x / VT, vt = threadIdx.x % VT;
  const int c = vt << 3;
  const int p0 = blockIdx.x * px_per, p1 = min(HW, p0 + px_per);
  float sa[8], sb[8];
  {
    const float4* p = reinterpret_cast<const float4*>(ab + static_cast<long long>(n) * C + c);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float4 f = p[i];
      sa[2 * i] = f.x; sb[2 * i] = f.y; sa[2 * i + 1] = f.z; sb[2 * i + 1] = f.w;
    }
  }
  long long ps;
  const __half* base = norm_src_ptr(s0, s1, c, ps);
  base += static_cast<long long>(n) * HW * ps;
  __half* obase = out + static_cast<long long>(n) * HW * C + c;
  auto apply = [&](const uint4& u, int px) {
    const __half2* h = reinterpret_cast<const __half2*>(&u);
    uint32_t pk[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 f = __half22float2(h[i]);
      float y0 = f.x * sa[2 * i] + sb[2 * i], y1 = f.y * sa[2 * i + 1] + sb[2 * i + 1];
      if (silu) { y0 = silu_fast(y0); y1 = silu_fast(y1); }
      pk[i] = pack_h2(y0, y1);
    }
    *reinterpret_cast<uint4*>(obase + static_cast<long long>(px) * C) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
  };
  int px = p0 + r;
  for (; px + 3 * R < p1; px += 4 * R) {
    const uint4 u0 = __ldg(reinterpret_cast<const uint4*>(base + px * ps));
    const uint4 u1 = __ldg(reinterpret_cast<const uint4*>(base + (px + R) * ps));
    const uint4 u2 = __ldg(reinterpret_cast<const uint4*>(base + (px + 2 * R) * ps));
    const uint4 u3 = __ldg(reinterpret_cast<const uint4*>(base + (px + 3 * R) * ps));
    apply(u0, px); apply(u1, px + R); apply(u2, px + 2 * R); apply(u3, px + 3 * R);
  }
  for (; px < p1; px += R) apply(__ldg(reinterpret_cast<const uint4*>(base + px * ps)), px);
}

// Fused GroupNorm for images that fit in L2 (every U-Net activation): ONE kernel, one thread-block cluster per image.
// grid (CL, Nimg), cluster (CL, 1, 1): CTA `rank` owns pixels [rank*px_per, (rank+1)*px_per) of image blockIdx.y.
//   pass 1  per-thread shifted channel sums over the CTA's pixel slice (same thread map as gn_stats_kernel), folded to
//           per-channel totals and 32 group sums in shared memory
//   cluster.sync, every CTA reads the CL group sums of its image through distributed shared memory IN RANK ORDER
//           (deterministic: fixed slices, fixed fold, fixed order -- and the geometry depends on (HW, C) only, so an
//           image's statistics do not depend on the batch it is in) -> group means; a second exchange of the
//           per-CTA sums of squared deviations about those means -> variances
//   pass 2  the slice is read again -- an L2 hit, it was streamed a few microseconds ago -- normalised (+SiLU) and
//           written.  HBM traffic = 1 read + 1 write of the activation (the two-kernel path reads it twice).
// MINB = co-resident CTAs per SM the register budget is held to, UNR = independent 16-byte loads in flight per thread
template <int MINB, int UNR>
__global__ void __launch_bounds__(384, MINB) gn_fused_kernel(NormSrc s0, NormSrc s1, int HW, int cpg, int px_per, int VT, int R,
                                                       const float* __restrict__ gamma, const float* __restrict__ beta,
                                                       float eps, int silu, __half* __restrict__ out) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  extern __shared__ float sm_gn[];
  __shared__ float part[32];   // this CTA's raw sum per group
  __shared__ float part2[32];  // this CTA's sum of squared deviations per group
  __shared__ float stat[64];   // mean[32], rstd[32]
  const int n = blockIdx.y;
  const int rank = static_cast<int>(cluster.block_rank());
  const int CL = static_cast<int>(cluster.num_blocks());
  const int C = s0.C + s1.C;
  const int nthr = VT * R;
  float* sm_a = sm_gn;
  float* sm_q = sm_gn + nthr * 8;
  float* sm_k = sm_gn + 2 * nthr * 8;
  const int r = threadIdx.x / VT, vt = threadIdx.x % VT;
  const int p0 = min(HW, rank * px_per), p1 = min(HW, p0 + px_per);
  const bool active = threadIdx.x < nthr;
  long long ps = 0;
  const __half* base = nullptr;
  if (active) {
    base = norm_src_ptr(s0, s1, vt << 3, ps);
    base += static_cast<long long>(n) * HW * ps;
    float a[8], q[8], k[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = 0.f; q[i] = 0.f; }
    gn_unpack8(__ldg(reinterpret_cast<const uint4*>(base)), k);
    gn_accumulate<UNR>(base, ps, p0, p1, r, R, k, a, q);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      sm_a[threadIdx.x * 8 + i] = a[i];  // index = r*C + channel
      sm_q[threadIdx.x * 8 + i] = q[i];
      if (r == 0) sm_k[vt * 8 + i] = k[i];
    }
  }
  __syncthreads();
  gn_fold_channels(sm_a, sm_q, C, R);
  __syncthreads();
  const float cnt = static_cast<float>(p1 - p0);
  {
    const float gs = gn_group_sum(sm_a, sm_k, cnt, cpg);
    if (threadIdx.x < 256 && (threadIdx.x & 7) == 0) part[threadIdx.x >> 3] = gs;
  }
  cluster.sync();
  const float tot = static_cast<float>(HW) * cpg;
  if (threadIdx.x < 32) {
    float s = 0.f;
    for (int c = 0; c < CL; ++c) s += cluster.map_shared_rank(part, c)[threadIdx.x];
    stat[threadIdx.x] = s / tot;
  }
  __syncthreads();
  {
    const float m2 = gn_group_m2(sm_a, sm_q, sm_k, stat, cnt, cpg);
    if (threadIdx.x < 256 && (threadIdx.x & 7) == 0) part2[threadIdx.x >> 3] = m2;
  }
  cluster.sync();
  if (threadIdx.x < 32) {
    float m2 = 0.f;
    for (int c = 0; c < CL; ++c) m2 += cluster.map_shared_rank(part2, c)[threadIdx.x];
    stat[32 + threadIdx.x] = rsqrtf(fmaxf(m2 / tot, 0.f) + eps);
  }
  cluster.sync();  // also keeps every CTA's partials alive until all remote reads are done
  if (!active) return;
  const int c = vt << 3;
  float sa[8], sb[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int g = (c + i) / cpg;
    const float sc = gamma[c + i] * stat[32 + g];
    sa[i] = sc;
    sb[i] = beta[c + i] - stat[g] * sc;
  }
  __half* obase = out + static_cast<long long>(n) * HW * C + c;
  auto apply = [&](const uint4& u, int px) {
    const __half2* h = reinterpret_cast<const __half2*>(&u);
    uint32_t pk[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 f = __half22float2(h[i]);
      float y0 = f.x * sa[2 * i] + sb[2 * i], y1 = f.y * sa[2 * i + 1] + sb[2 * i + 1];
      if (silu) { y0 = silu_fast(y0); y1 = silu_fast(y1); }
      pk[i] = pack_h2(y0, y1);
    }
    *reinterpret_cast<uint4*>(obase + static_cast<long long>(px) * C) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
  };
  int px = p0 + r;
  for (; px + (UNR - 1) * R < p1; px += UNR * R) {
    uint4 u[UNR];
#pragma unroll
    for (int k = 0; k < UNR; ++k) u[k] = __ldcg(reinterpret_cast<const uint4*>(base + (px + k * R) * ps));
#pragma unroll
    for (int k = 0; k < UNR; ++k) apply(u[k], px + k * R);
  }
  for (; px < p1; px += R) apply(__ldcg(reinterpret_cast<const uint4*>(base + px * ps)), px);
}

// GroupNorm whose statistics were formed by the epilogues of the convs that produced its input(s) (IgGn, igemm.cuh): fold +
// apply in ONE kernel, one cluster of CL (4, 8 or 16) CTAs per image; like the kernels above it normalises cat(src0, src1)
// without materialising it.  grid (CL, Nimg), cluster (CL, 1, 1).
//   first   every thread issues the first UNR 16-byte loads of its apply pass (they do not depend on the statistics), so
//           the fold below runs in the shadow of their latency
//   fold    CTA `rank` owns 32/CL groups (C/CL channels): thread (slice, channel) sums every NS-th entry of the source's
//           record [E][C_src][2] in index order, the slices are added in index order, each group is reduced by one warp
//           with a fixed xor tree (same shifted-sum algebra as above) -> (a, b) per channel, which the CTA writes into the
//           shared memory of all CTAs of the image (distributed shared memory)
//   cluster.sync (the only one)
//   apply   CTA `rank` normalises (+SiLU) pixels [rank*px_per, (rank+1)*px_per): 1 read + 1 write of the activation.
// Large images (128x128 latents: few images per micro-batch) are split over gridDim.z clusters, each folding the (small)
// records again and applying to its own quarter of the pixels, so the pass still fills the machine.
// Deterministic and batch-invariant: the entry order, slice count (a function of C), split count (a function of HW) and
// trees are fixed.
struct GnStatSrc {
  const __half* x;   // dense [Nimg][HW][C]
  const float* rec;  // per image [E][C][2] partial (S, Q), then [C] shifts
  int C;
};
constexpr int GN_FOLD_MAXC = 2560;
template <int THREADS, int MINB, int UNR>
__global__ void __launch_bounds__(THREADS, MINB) gn_fold_apply_kernel(GnStatSrc s0, GnStatSrc s1, int HW, int cpg, int px_per,
                                                                       int VT, int R, int E, const float* __restrict__ gamma,
                                                                       const float* __restrict__ beta, float eps, int silu,
                                                                       __half* __restrict__ out) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  __shared__ float2 ab_s[GN_FOLD_MAXC];  // (a, b) of every channel of the image, filled by the CTAs of the cluster
  __shared__ float sl_s[THREADS], sl_q[THREADS];  // [slice][channel of this CTA]
  __shared__ float ch_s[THREADS], ch_q[THREADS], ch_k[THREADS];
  const int n = blockIdx.y;
  const int rank = static_cast<int>(cluster.block_rank());
  const int CL = static_cast<int>(cluster.num_blocks());
  const int tid = threadIdx.x;
  const int C = s0.C + s1.C;
  // ---- apply-pass thread map and its first batch of loads
  const bool active = tid < VT * R;
  const int r = tid / VT, vt = tid % VT;
  const int c = vt << 3;
  const int p0 = min(HW, (static_cast<int>(blockIdx.z) * CL + rank) * px_per), p1 = min(HW, p0 + px_per);
  const bool in0 = c < s0.C;
  const long long ps = in0 ? s0.C : s1.C;  // pixel stride of this thread's source
  const __half* base = (in0 ? s0.x + c : s1.x + (c - s0.C)) + static_cast<long long>(n) * HW * ps;
  uint4 u[UNR];
  int px = p0 + r;
  const bool first_full = active && px + (UNR - 1) * R < p1;
  if (first_full) {
#pragma unroll
    for (int k = 0; k < UNR; ++k) u[k] = __ldg(reinterpret_cast<const uint4*>(base + (px + k * R) * ps));
  }
  // ---- fold
  const int nch = C / CL;        // channels folded by this CTA
  const int c0 = rank * nch;
  const int NS = THREADS / nch;  // entry slices (>= 1: the launcher keeps C / CL <= THREADS)
  if (tid < NS * nch) {
    const int sl = tid / nch, ch = tid % nch;
    const int gc = c0 + ch;
    const bool f0 = gc < s0.C;
    const int Cs = f0 ? s0.C : s1.C, lc = f0 ? gc : gc - s0.C;
    const float2* pp = reinterpret_cast<const float2*>((f0 ? s0.rec : s1.rec) + static_cast<long long>(n) * (2 * E + 1) * Cs) + lc;
    float s = 0.f, q = 0.f;
    int e = sl;
    for (; e + 7 * NS < E; e += 8 * NS) {  // eight independent loads in flight, added in index order
      float2 v[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = __ldcg(pp + static_cast<long long>(e + k * NS) * Cs);
#pragma unroll
      for (int k = 0; k < 8; ++k) { s += v[k].x; q += v[k].y; }
    }
    for (; e < E; e += NS) {
      const float2 v = __ldcg(pp + static_cast<long long>(e) * Cs);
      s += v.x; q += v.y;
    }
    sl_s[tid] = s;
    sl_q[tid] = q;
  }
  __syncthreads();
  if (tid < nch) {
    float s = 0.f, q = 0.f;
    for (int i = 0; i < NS; ++i) { s += sl_s[i * nch + tid]; q += sl_q[i * nch + tid]; }
    ch_s[tid] = s;
    ch_q[tid] = q;
    const int gc = c0 + tid;
    const bool f0 = gc < s0.C;
    const int Cs = f0 ? s0.C : s1.C, lc = f0 ? gc : gc - s0.C;
    ch_k[tid] = __ldcg((f0 ? s0.rec : s1.rec) + static_cast<long long>(n) * (2 * E + 1) * Cs + static_cast<long long>(2 * E) * Cs + lc);
  }
  __syncthreads();
  if (tid < 32 * (32 / CL)) {  // warp g: group (32/CL)*rank + g
    const int g = tid >> 5, lane = tid & 31;
    const float cnt = static_cast<float>(HW);
    float gs = 0.f;
    for (int cc = lane; cc < cpg; cc += 32) gs += ch_s[g * cpg + cc] + cnt * ch_k[g * cpg + cc];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) gs += __shfl_xor_sync(0xffffffffu, gs, o);
    const float mean = gs / (cnt * cpg);
    float m2 = 0.f;
    for (int cc = lane; cc < cpg; cc += 32) {
      const float d = mean - ch_k[g * cpg + cc];
      m2 += ch_q[g * cpg + cc] - 2.f * d * ch_s[g * cpg + cc] + cnt * d * d;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m2 += __shfl_xor_sync(0xffffffffu, m2, o);
    const float rstd = rsqrtf(fmaxf(m2 / (cnt * cpg), 0.f) + eps);
    for (int cc = lane; cc < cpg; cc += 32) {
      const int ch = c0 + g * cpg + cc;
      const float sc = __ldg(gamma + ch) * rstd;
      const float2 v = make_float2(sc, __ldg(beta + ch) - mean * sc);
      for (int peer = 0; peer < CL; ++peer) cluster.map_shared_rank(ab_s, peer)[ch] = v;
    }
  }
  cluster.sync();
  if (!active) return;
  // ---- apply
  float sa[8], sb[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { sa[i] = ab_s[c + i].x; sb[i] = ab_s[c + i].y; }
  __half* obase = out + static_cast<long long>(n) * HW * C + c;
  auto apply = [&](const uint4& v, int pxl) {
    const __half2* h = reinterpret_cast<const __half2*>(&v);
    uint32_t pk[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 f = __half22float2(h[i]);
      float y0 = f.x * sa[2 * i] + sb[2 * i], y1 = f.y * sa[2 * i + 1] + sb[2 * i + 1];
      if (silu) { y0 = silu_fast(y0); y1 = silu_fast(y1); }
      pk[i] = pack_h2(y0, y1);
    }
    *reinterpret_cast<uint4*>(obase + static_cast<long long>(pxl) * C) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
  };
  // (register double buffering of the batches was measured: fewer resident CTAs cost more than the overlap gains --
  // 2.96-3.84 ms against 2.81 for the 44 layers of a micro-batch, profiles/r02_experiments/norm_sweep4_prefetch.txt)
  if (first_full) {
#pragma unroll
    for (int k = 0; k < UNR; ++k) apply(u[k], px + k * R);
    px += UNR * R;
  }
  for (; px + (UNR - 1) * R < p1; px += UNR * R) {
#pragma unroll
    for (int k = 0; k < UNR; ++k) u[k] = __ldg(reinterpret_cast<const uint4*>(base + (px + k * R) * ps));
#pragma unroll
    for (int k = 0; k < UNR; ++k) apply(u[k], px + k * R);
  }
  for (; px < p1; px += R) apply(__ldg(reinterpret_cast<const uint4*>(base + px * ps)), px);
}

// LayerNorm over the last dim (C <= 32*8*MAXV, multiple of 8).  Persistent warps: a warp walks rows
// w, w+W, ... holding one row in registers while the next row's loads are already in flight.  The kernel is
// FMA-pipe-bound, not HBM-bound (plain fp32 FMA-pipe instructions issue every other cycle), so all per-element math
// runs on packed f32x2 instructions and gamma / beta are read as float pairs from shared memory (LSU pipe).
__device__ __forceinline__ uint64_t ln_pack(float a, float b) {
  uint64_t d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(a), "f"(b));
  return d;
}
__device__ __forceinline__ void ln_unpack(uint64_t v, float& a, float& b) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ uint64_t ln_fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t ln_add2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t ln_mul2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

template <int MAXV>
__global__ void __launch_bounds__(256) layernorm_kernel(const __half* __restrict__ x, long long ld_x,
                                                        const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, float eps, long long rows,
                                                        int C, __half* __restrict__ out, long long ld_out) {
  extern __shared__ float sm_ln[];  // gamma[C] | beta[C]
  float* sg = sm_ln;
  float* sb = sm_ln + C;
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    sg[i] = gamma[i];
    sb[i] = beta[i];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const long long nw = static_cast<long long>(gridDim.x) * (blockDim.x >> 5);
  long long row = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int vcols = C >> 3;
  uint4 cur[MAXV], nxt[MAXV];
  auto load = [&](uint4* dst, long long rw) {
#pragma unroll
    for (int k = 0; k < MAXV; ++k) {
      const int vc = lane + 32 * k;
      dst[k] = make_uint4(0, 0, 0, 0);
      if (vc < vcols) dst[k] = __ldg(reinterpret_cast<const uint4*>(x + rw * ld_x + (vc << 3)));
    }
  };
  if (row < rows) load(cur, row);
  const float invC = 1.f / static_cast<float>(C);
  while (row < rows) {
    const long long nrow = row + nw;
    if (nrow < rows) load(nxt, nrow);
    uint64_t v[MAXV][4];  // the row slice as packed fp32 pairs
    uint64_t s2 = ln_pack(0.f, 0.f);
#pragma unroll
    for (int k = 0; k < MAXV; ++k) {
      const __half2* h = reinterpret_cast<const __half2*>(&cur[k]);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 f = __half22float2(h[i]);
        v[k][i] = ln_pack(f.x, f.y);
        s2 = ln_add2(s2, v[k][i]);  // lanes past vcols hold zeros
      }
    }
    float sa, sb2;
    ln_unpack(s2, sa, sb2);
    float s = sa + sb2;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s * invC;
    const uint64_t nm2 = ln_pack(-mean, -mean);
    uint64_t q2 = ln_pack(0.f, 0.f);
#pragma unroll
    for (int k = 0; k < MAXV; ++k) {
      if (lane + 32 * k < vcols) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          v[k][i] = ln_add2(v[k][i], nm2);  // d = x - mean
          q2 = ln_fma2(v[k][i], v[k][i], q2);
        }
      }
    }
    float qa, qb;
    ln_unpack(q2, qa, qb);
    float q = qa + qb;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = rsqrtf(q * invC + eps);
    const uint64_t r2 = ln_pack(rstd, rstd);
#pragma unroll
    for (int k = 0; k < MAXV; ++k) {
      const int vc = lane + 32 * k;
      if (vc < vcols) {
        const float4* g4 = reinterpret_cast<const float4*>(sg + (vc << 3));
        const float4* b4 = reinterpret_cast<const float4*>(sb + (vc << 3));
        uint32_t pk[4];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const float4 g = g4[i], b = b4[i];
          float y0, y1, y2, y3;
          ln_unpack(ln_fma2(ln_mul2(v[k][2 * i], r2), ln_pack(g.x, g.y), ln_pack(b.x, b.y)), y0, y1);
          ln_unpack(ln_fma2(ln_mul2(v[k][2 * i + 1], r2), ln_pack(g.z, g.w), ln_pack(b.z, b.w)), y2, y3);
          pk[2 * i] = pack_h2(y0, y1);
          pk[2 * i + 1] = pack_h2(y2, y3);
        }
        *reinterpret_cast<uint4*>(out + row * ld_out + (vc << 3)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      }
    }
#pragma unroll
    for (int k = 0; k < MAXV; ++k) cur[k] = nxt[k];
    row = nrow;
  }
}

// LayerNorm for C = 40 * LPR (320 / 640 / 1280: every Transformer2DModel width): LPR lanes per row, five 16-byte vectors per
// lane, 32 / LPR rows per warp -- every lane loads, computes and stores (the generic kernel above leaves 37.5 % of a warp
// idle at C = 320).  Persistent warps with the next rows' loads in flight, two-pass statistics in registers, xor-tree
// reductions over the LPR lanes of a row (fixed order: deterministic).
template <int LPR>
__global__ void __launch_bounds__(256) layernorm5_kernel(const __half* __restrict__ x, long long ld_x,
                                                         const float* __restrict__ gamma, const float* __restrict__ beta,
                                                         float eps, long long rows, __half* __restrict__ out,
                                                         long long ld_out) {
  constexpr int C = 40 * LPR, RPW = 32 / LPR;
  extern __shared__ float sm_ln[];  // gamma[C] | beta[C]
  float* sg = sm_ln;
  float* sb = sm_ln + C;
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    sg[i] = gamma[i];
    sb[i] = beta[i];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int sub = lane / LPR, j = lane % LPR;
  const long long stride = static_cast<long long>(gridDim.x) * (blockDim.x >> 5) * RPW;
  long long row = (static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5)) * RPW + sub;
  uint4 cur[5], nxt[5];
  auto load = [&](uint4* dst, long long rw) {
    const __half* px = x + rw * ld_x + (j << 3);
#pragma unroll
    for (int k = 0; k < 5; ++k) dst[k] = __ldg(reinterpret_cast<const uint4*>(px + k * LPR * 8));
  };
  if (row < rows) load(cur, row);
  constexpr float invC = 1.f / static_cast<float>(C);
  // rows of a warp run out together except in the last sweep: the shuffles below stay inside a row's own LPR lanes, and
  // lanes whose row is past the end keep executing them on stale registers without storing
  while (__any_sync(0xffffffffu, row < rows)) {
    const bool live = row < rows;
    const long long nrow = row + stride;
    if (nrow < rows) load(nxt, nrow);
    float v[5][8];
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      const __half2* h = reinterpret_cast<const __half2*>(&cur[k]);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 f = __half22float2(h[i]);
        v[k][2 * i] = f.x;
        v[k][2 * i + 1] = f.y;
        s += f.x + f.y;
      }
    }
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s * invC;
    float q = 0.f;
#pragma unroll
    for (int k = 0; k < 5; ++k)
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        v[k][i] -= mean;
        q = fmaf(v[k][i], v[k][i], q);
      }
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = rsqrtf(q * invC + eps);
    if (live) {
      __half* po = out + row * ld_out + (j << 3);
#pragma unroll
      for (int k = 0; k < 5; ++k) {
        const float4* g4 = reinterpret_cast<const float4*>(sg + ((j + k * LPR) << 3));
        const float4* b4 = reinterpret_cast<const float4*>(sb + ((j + k * LPR) << 3));
        const float4 g0 = g4[0], g1 = g4[1], b0 = b4[0], b1 = b4[1];
        const uint32_t p0 = pack_h2(fmaf(v[k][0] * rstd, g0.x, b0.x), fmaf(v[k][1] * rstd, g0.y, b0.y));
        const uint32_t p1 = pack_h2(fmaf(v[k][2] * rstd, g0.z, b0.z), fmaf(v[k][3] * rstd, g0.w, b0.w));
        const uint32_t p2 = pack_h2(fmaf(v[k][4] * rstd, g1.x, b1.x), fmaf(v[k][5] * rstd, g1.y, b1.y));
        const uint32_t p3 = pack_h2(fmaf(v[k][6] * rstd, g1.z, b1.z), fmaf(v[k][7] * rstd, g1.w, b1.w));
        *reinterpret_cast<uint4*>(po + k * LPR * 8) = make_uint4(p0, p1, p2, p3);
      }
    }
#pragma unroll
    for (int k = 0; k < 5; ++k) cur[k] = nxt[k];
    row = nrow;
  }
}

// Same row map without the register prefetch: the row stays packed (20 registers), the three passes convert on the fly, and
// the overlap comes from occupancy (4 CTAs = 32 warps per SM) instead of from a second row buffer.  gamma / beta sit in
// shared memory as fp16 (they ARE fp16 weights in the engine; the fp32 copies the launcher receives convert back exactly),
// gamma and beta of one 8-channel vector side by side: 32 bytes of shared-memory reads per 16-byte vector instead of 64 --
// the kernel was L1/LSU-pipe-bound (84 % l1tex throughput, profiles/r02_norm_kernels.txt), not HBM-bound.
template <int LPR, int MINB>
__global__ void __launch_bounds__(256, MINB) layernorm5b_kernel(const __half* __restrict__ x, long long ld_x,
                                                                const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                float eps, long long rows, __half* __restrict__ out,
                                                                long long ld_out) {
  constexpr int C = 40 * LPR, RPW = 32 / LPR;
  __shared__ __align__(16) __half sgb[2 * C];  // per 8-channel vector: gamma[8] | beta[8]
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    sgb[(i >> 3) * 16 + (i & 7)] = __float2half_rn(gamma[i]);
    sgb[(i >> 3) * 16 + 8 + (i & 7)] = __float2half_rn(beta[i]);
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int sub = lane / LPR, j = lane % LPR;
  const long long stride = static_cast<long long>(gridDim.x) * (blockDim.x >> 5) * RPW;
  constexpr float invC = 1.f / static_cast<float>(C);
  for (long long row0 = (static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5)) * RPW; row0 < rows;
       row0 += stride) {
    const long long row = row0 + sub;
    const bool live = row < rows;
    uint4 cur[5];
    if (live) {
      const __half* px = x + row * ld_x + (j << 3);
#pragma unroll
      for (int k = 0; k < 5; ++k) cur[k] = __ldg(reinterpret_cast<const uint4*>(px + k * LPR * 8));
    } else {
#pragma unroll
      for (int k = 0; k < 5; ++k) cur[k] = make_uint4(0, 0, 0, 0);
    }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      const __half2* h = reinterpret_cast<const __half2*>(&cur[k]);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 f = __half22float2(h[i]);
        s += f.x + f.y;
      }
    }
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s * invC;
    float q = 0.f;
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      const __half2* h = reinterpret_cast<const __half2*>(&cur[k]);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 f = __half22float2(h[i]);
        const float d0 = f.x - mean, d1 = f.y - mean;
        q = fmaf(d0, d0, q);
        q = fmaf(d1, d1, q);
      }
    }
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = rsqrtf(q * invC + eps);
    const float nmr = -mean * rstd;
    if (live) {
      __half* po = out + row * ld_out + (j << 3);
#pragma unroll
      for (int k = 0; k < 5; ++k) {
        const __half2* h = reinterpret_cast<const __half2*>(&cur[k]);
        const uint4 gv = *reinterpret_cast<const uint4*>(sgb + (j + k * LPR) * 16);
        const uint4 bv = *reinterpret_cast<const uint4*>(sgb + (j + k * LPR) * 16 + 8);
        const __half2* gh = reinterpret_cast<const __half2*>(&gv);
        const __half2* bh = reinterpret_cast<const __half2*>(&bv);
        uint32_t pk[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 f = __half22float2(h[i]), g = __half22float2(gh[i]), b = __half22float2(bh[i]);
          // (x - mean) * rstd = x * rstd + (-mean * rstd), then * gamma + beta
          pk[i] = pack_h2(fmaf(fmaf(f.x, rstd, nmr), g.x, b.x), fmaf(fmaf(f.y, rstd, nmr), g.y, b.y));
        }
        *reinterpret_cast<uint4*>(po + k * LPR * 8) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      }
    }
  }
}

}  // namespace dm
