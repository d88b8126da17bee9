// Small HBM-bound helpers around the tensor-core ops: boundary layout changes (NCHW fp32 <-> NHWC fp16),
// add_noise + 3x3 patch gather for the 3/4-channel input convs, resampling, the timestep sinusoid,
// the VAE posterior sample, the row softmax of the VAE's single-head attention, the per-element loss
// and its Monte-Carlo reduction into T(x|c).
#pragma once
#include "ptx.cuh"

namespace dm {

// A[m, tap*Cin + c] = src(img(b), c, y+dy, x+dx) (zero outside), zero-padded to 64 columns; fp16.
// With noise != null:  src = sqrt_acp[t]*x0 + sqrt_1m_acp[t]*noise computed in fp32 then rounded to fp16
// (reference: scheduler.add_noise at /root/reference/diffmining/typicality/compute.py:99, fp32 operands,
// cast to fp16 by autocast at conv_in).
__global__ void __launch_bounds__(256) patch3x3_kernel(const float* __restrict__ x0, const int* __restrict__ x_index,
                                                       const float* __restrict__ noise,
                                                       const int* __restrict__ noise_index,
                                                       const long long* __restrict__ t, const float* __restrict__ ca,
                                                       const float* __restrict__ cb, int sched_n, int* err_flag,
                                                       int Bf, int Cin, int H, int W, int index_stride,
                                                       __half* __restrict__ out) {
  const long long total = static_cast<long long>(Bf) * H * W * 8;  // 8 x 16-byte vectors per row
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int v = static_cast<int>(idx & 7);
    const long long m = idx >> 3;
    const int x = static_cast<int>(m % W), y = static_cast<int>((m / W) % H), b = static_cast<int>(m / (static_cast<long long>(W) * H));
    // index_stride > 1: row b stands for the group of `index_stride` consecutive forwards that share (x_t, t)
    const int xi = x_index ? x_index[b * index_stride] : b;
    const int ni = noise_index ? noise_index[b * index_stride] : b;
    float a = 1.f, bb = 0.f;
    if (noise) {
      long long tt = t[ni];
      if (tt < 0 || tt >= sched_n) {  // never index the schedule tables out of bounds; the host reports the sticky flag
        if (err_flag) *err_flag = 1;
        tt = tt < 0 ? 0 : sched_n - 1;
      }
      a = ca[tt];
      bb = cb[tt];
    }
    float vals[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int k = v * 8 + i;
      float r = 0.f;
      if (k < 9 * Cin) {
        const int tap = k / Cin, c = k % Cin;
        const int yy = y + tap / 3 - 1, xx = x + tap % 3 - 1;
        if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
          const long long o = (static_cast<long long>(c) * H + yy) * W + xx;
          r = x0[static_cast<long long>(xi) * Cin * H * W + o];
          if (noise) r = a * r + bb * noise[static_cast<long long>(ni) * Cin * H * W + o];
        }
      }
      vals[i] = r;
    }
    *reinterpret_cast<uint4*>(out + m * 64 + v * 8) = make_uint4(pack_h2(vals[0], vals[1]), pack_h2(vals[2], vals[3]),
                                                                 pack_h2(vals[4], vals[5]), pack_h2(vals[6], vals[7]));
  }
}

// sinusoidal timestep embedding, flip_sin_to_cos, freq_shift 0: [cos(t f_k) | sin(t f_k)], k < 160, fp16 out
// (reference: time_proj + cast to model dtype, /root/reference/diffmining/typicality/dift.py:84-89).
__global__ void timestep_embed_kernel(const long long* __restrict__ t, const int* __restrict__ t_index, int Bf,
                                      __half* __restrict__ out) {
  const int b = blockIdx.x, k = threadIdx.x;  // 160 threads
  const float tt = static_cast<float>(t[t_index ? t_index[b] : b]);
  const float f = expf(-9.210340371976184f * static_cast<float>(k) / 160.f);
  const float e = tt * f;
  out[b * 320 + k] = __float2half_rn(cosf(e));
  out[b * 320 + 160 + k] = __float2half_rn(sinf(e));
}

// out[(u*G + g), :] = in[u, :], g < G: replicates the activations computed once per (x_t, t) group to every
// condition row of the group (cond/uncond prefix sharing); `row_vecs` = 16-byte vectors per row
__global__ void __launch_bounds__(256) repeat_rows_kernel(const uint4* __restrict__ in, long long rows, long long row_vecs,
                                                          int G, uint4* __restrict__ out) {
  const long long total = rows * row_vecs;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long u = idx / row_vecs, v = idx % row_vecs;
    const uint4 val = __ldg(in + idx);
    for (int g = 0; g < G; ++g) out[(u * G + g) * row_vecs + v] = val;
  }
}

// nearest resize NHWC -> NHWC (torch 'nearest': src = min(floor(dst * in/out), in-1))
__global__ void __launch_bounds__(256) upsample_nearest_kernel(const __half* __restrict__ in, int N, int H, int W,
                                                               int C, int Ho, int Wo, __half* __restrict__ out) {
  const int vc = C >> 3;
  const long long total = static_cast<long long>(N) * Ho * Wo * vc;
  const float sy = static_cast<float>(H) / Ho, sx = static_cast<float>(W) / Wo;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int v = static_cast<int>(idx % vc);
    const long long pix = idx / vc;
    const int xo = static_cast<int>(pix % Wo), yo = static_cast<int>((pix / Wo) % Ho), n = static_cast<int>(pix / (static_cast<long long>(Wo) * Ho));
    const int yi = min(static_cast<int>(floorf(yo * sy)), H - 1), xi = min(static_cast<int>(floorf(xo * sx)), W - 1);
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(in + ((static_cast<long long>(n) * H + yi) * W + xi) * C + v * 8));
    *reinterpret_cast<uint4*>(out + pix * C + v * 8) = u;
  }
}

// [N,H,W,C] -> four parity planes [(py*2+px)*N + n, H2, W2, C] (zero where the source pixel does not exist),
// so a stride-2 3x3 conv becomes 9 unit-stride shifted box loads.
__global__ void __launch_bounds__(256) space_to_planes_kernel(const __half* __restrict__ in, int N, int H, int W, int C,
                                                              int H2, int W2, __half* __restrict__ out) {
  const int vc = C >> 3;
  const long long total = 4ll * N * H2 * W2 * vc;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int v = static_cast<int>(idx % vc);
    const long long pix = idx / vc;
    const int x2 = static_cast<int>(pix % W2), y2 = static_cast<int>((pix / W2) % H2);
    const long long pn = pix / (static_cast<long long>(W2) * H2);
    const int n = static_cast<int>(pn % N), plane = static_cast<int>(pn / N);
    const int y = 2 * y2 + (plane >> 1), x = 2 * x2 + (plane & 1);
    uint4 u = make_uint4(0, 0, 0, 0);
    if (y < H && x < W) u = __ldg(reinterpret_cast<const uint4*>(in + ((static_cast<long long>(n) * H + y) * W + x) * C + v * 8));
    *reinterpret_cast<uint4*>(out + pix * C + v * 8) = u;
  }
}

// Per-element loss of the typicality path (reference: F.mse_loss(noise_pred.float(), noise, 'none'),
// /root/reference/diffmining/typicality/compute.py:101): pred is the fp16 conv_out result (NHWC, 16-col padded),
// loss = (float(pred) - eps)^2 in fp32.  Writes any of: fp32 NCHW loss rows, the fp16 raw grid laid out
// [img][sample][cond][4][h][w] (compute.py:155-160), and accumulates T(x|c) terms.
struct LossMap {
  const int* noise_index;  // [Bf] row -> noise row
  const int* grid_row;     // [Bf] row -> (img*N + sample)*n_cond + cond  (raw-grid row), or null
  float* loss_f32;         // [Bf,4,h,w] or null
  __half* grid_f16;        // [rows,4,h,w] or null
  float* eps_f32;          // [Bf,4,h,w] raw prediction out or null
};
__global__ void __launch_bounds__(256) loss_kernel(const __half* __restrict__ pred, int ld_pred,
                                                   const float* __restrict__ noise, LossMap mp, int Bf, int HW) {
  const long long total = static_cast<long long>(Bf) * HW;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int b = static_cast<int>(idx / HW), px = static_cast<int>(idx % HW);
    const uint2 u = *reinterpret_cast<const uint2*>(pred + idx * ld_pred);
    const __half2* h = reinterpret_cast<const __half2*>(&u);
    const float2 p01 = __half22float2(h[0]), p23 = __half22float2(h[1]);
    const float pr[4] = {p01.x, p01.y, p23.x, p23.y};
    const int ni = mp.noise_index ? mp.noise_index[b] : b;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const long long o = (static_cast<long long>(b) * 4 + c) * HW + px;
      if (mp.eps_f32) mp.eps_f32[o] = pr[c];
      if (noise) {
        const float d = pr[c] - noise[(static_cast<long long>(ni) * 4 + c) * HW + px];
        const float l = d * d;
        if (mp.loss_f32) mp.loss_f32[o] = l;
        if (mp.grid_f16) {
          const int gr = mp.grid_row ? mp.grid_row[b] : b;
          mp.grid_f16[(static_cast<long long>(gr) * 4 + c) * HW + px] = __float2half_rn(l);
        }
      }
    }
  }
}

// T(x|c_k)[img, k, px] = mean_s [ mean_ch grid[img,s,n_cond-1,ch,px] - mean_ch grid[img,s,k,ch,px] ], k < n_cond-1,
// from the fp16 raw grid (the values the reference consumers read back from the .npy:
// /root/reference/diffmining/typicality/cluster.py:112-123).  Fixed summation order -> bit-identical on any rank.
__global__ void __launch_bounds__(256) tmap_kernel(const __half* __restrict__ grid, int Bi, int N, int n_cond, int HW,
                                                   float* __restrict__ T) {
  const long long total = static_cast<long long>(Bi) * (n_cond - 1) * HW;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int px = static_cast<int>(idx % HW);
    const int k = static_cast<int>((idx / HW) % (n_cond - 1));
    const int img = static_cast<int>(idx / (static_cast<long long>(HW) * (n_cond - 1)));
    float acc = 0.f;
    for (int s = 0; s < N; ++s) {
      const __half* gu = grid + ((static_cast<long long>(img) * N + s) * n_cond + (n_cond - 1)) * 4 * HW + px;
      const __half* gc = grid + ((static_cast<long long>(img) * N + s) * n_cond + k) * 4 * HW + px;
      float mu = 0.f, mc = 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        mu += __half2float(gu[c * HW]);
        mc += __half2float(gc[c * HW]);
      }
      acc += mu * 0.25f - mc * 0.25f;
    }
    T[idx] = acc / N;
  }
}

// VAE tail: moments = quant_conv(conv_out) (1x1, 8->8), mean/logvar split, logvar clamp [-30, 20],
// z = (mean + exp(0.5*logvar) * eps) * scaling.  fp16 roundings follow autocast: conv_out and quant_conv
// outputs are fp16, exp runs in fp32, the sample is fp32.
__global__ void __launch_bounds__(256) vae_sample_kernel(const __half* __restrict__ h16, int ld_h,
                                                         const __half* __restrict__ wq, const float* __restrict__ bq,
                                                         const float* __restrict__ eps, float scaling, int B, int HW,
                                                         float* __restrict__ z, float* __restrict__ mean_out,
                                                         float* __restrict__ logvar_out) {
  const long long total = static_cast<long long>(B) * HW;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int b = static_cast<int>(idx / HW), px = static_cast<int>(idx % HW);
    const uint4 u = *reinterpret_cast<const uint4*>(h16 + idx * ld_h);
    const __half* hv = reinterpret_cast<const __half*>(&u);
    float in[8], mom[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) in[i] = __half2float(hv[i]);
#pragma unroll
    for (int o = 0; o < 8; ++o) {
      float a = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) a += __half2float(wq[o * 8 + i]) * in[i];
      mom[o] = round_h(a + bq[o]);
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const float mean = mom[c];
      const float logvar = fminf(fmaxf(mom[4 + c], -30.f), 20.f);
      const long long o = (static_cast<long long>(b) * 4 + c) * HW + px;
      const float stdv = expf(round_h(0.5f * logvar));
      if (z) z[o] = (mean + stdv * (eps ? eps[o] : 0.f)) * scaling;
      if (mean_out) mean_out[o] = mean;
      if (logvar_out) logvar_out[o] = logvar;
    }
  }
}

// row softmax: P[row, :] = softmax(scale * S[row, :]) fp32 in -> fp16 out (the VAE's 1-head attention).
__global__ void __launch_bounds__(256) softmax_rows_kernel(const float* __restrict__ S, long long ld_s, int cols,
                                                           float scale, __half* __restrict__ P, long long ld_p) {
  __shared__ float red[32];
  const float* s = S + blockIdx.x * ld_s;
  __half* p = P + blockIdx.x * ld_p;
  float mx = -INFINITY;
  for (int i = threadIdx.x; i < cols; i += blockDim.x) mx = fmaxf(mx, s[i]);
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  mx = red[0];
  for (int i = 1; i < (int)(blockDim.x >> 5); ++i) mx = fmaxf(mx, red[i]);
  __syncthreads();
  float sum = 0.f;
  for (int i = threadIdx.x; i < cols; i += blockDim.x) sum += __expf((s[i] - mx) * scale);
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sum;
  __syncthreads();
  sum = 0.f;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) sum += red[i];
  const float inv = 1.f / sum;
  for (int i = threadIdx.x; i < cols; i += blockDim.x) p[i] = __float2half_rn(__expf((s[i] - mx) * scale) * inv);
}

// NHWC fp16 [B*E, HW, C] -> NCHW fp32 [B, C, HW], averaged over the E ensemble members
// (reference: unet_ft.mean(0, keepdim=True), /root/reference/diffmining/typicality/dift.py:231).
__global__ void __launch_bounds__(256) nhwc_to_nchw_mean_kernel(const __half* __restrict__ in, int B, int E, int HW, int C,
                                                                float* __restrict__ out) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z, c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int r = ty; r < 32; r += 8) {
    const int px = p0 + r, c = c0 + tx;
    float a = 0.f;
    if (px < HW && c < C)
      for (int e = 0; e < E; ++e) a += __half2float(in[((static_cast<long long>(b) * E + e) * HW + px) * C + c]);
    tile[r][tx] = a / E;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int c = c0 + r, px = p0 + tx;
    if (px < HW && c < C) out[(static_cast<long long>(b) * C + c) * HW + px] = tile[tx][r];
  }
}

}  // namespace dm
