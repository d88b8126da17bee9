// T-map consumer: what the reference does on the host right after the hot path --
//   load_typicality (/root/reference/diffmining/typicality/cluster.py:125-137): channel mean -> bilinear resize to
//   the image size -> AvgPool2d((kx, ky), stride 1) of cond / uncond -> difference -> mean over the N draws,
//   df_D + get_non_overlapping (cluster.py:183-205, utils.py:74-102): every window position scored, sorted, and the
//   k best mutually non-overlapping windows kept (touching windows count as overlapping: inclusive comparisons)
// -- as three small kernels on the T map the engine already holds in HBM.  All steps are linear, so the engine's
// latent-resolution T (mean over draws of the channel-mean difference) is resized and pooled once instead of per
// draw and per condition.  Output per image: k boxes (x_start, y_start, x_end, y_end) in the reference's
// (row, column) convention with x_end = x_start + kx, and their scores D.
#include "../../include/dm_abi.h"
#include "abi_util.h"

namespace dm {

// torch.nn.functional.interpolate(mode="bilinear", align_corners=False): src = (dst + 0.5) * in / out - 0.5, clamped at 0
__global__ void __launch_bounds__(256) tmap_resize_kernel(const float* __restrict__ T, int B, int h, int w, int H, int W,
                                                          float* __restrict__ out) {
  const long long total = static_cast<long long>(B) * H * W;
  const float sy = static_cast<float>(h) / H, sx = static_cast<float>(w) / W;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int X = static_cast<int>(idx % W), Y = static_cast<int>((idx / W) % H), b = static_cast<int>(idx / (static_cast<long long>(W) * H));
    const float fy = fmaxf((Y + 0.5f) * sy - 0.5f, 0.f), fx = fmaxf((X + 0.5f) * sx - 0.5f, 0.f);
    const int y0 = static_cast<int>(fy), x0 = static_cast<int>(fx);
    const int y1 = min(y0 + 1, h - 1), x1 = min(x0 + 1, w - 1);
    const float wy = fy - y0, wx = fx - x0;
    const float* t = T + static_cast<long long>(b) * h * w;
    const float top = t[y0 * w + x0] * (1.f - wx) + t[y0 * w + x1] * wx;
    const float bot = t[y1 * w + x0] * (1.f - wx) + t[y1 * w + x1] * wx;
    out[idx] = top * (1.f - wy) + bot * wy;
  }
}

// separable box sums: window `win` along the fastest (stride 1, axis 1) or the row (axis 0) dimension, "valid" extent
__global__ void __launch_bounds__(256) box_sum_kernel(const float* __restrict__ in, int B, int H, int W, int win, int axis,
                                                      float scale, float* __restrict__ out) {
  const int Ho = axis == 0 ? H - win + 1 : H, Wo = axis == 1 ? W - win + 1 : W;
  const long long total = static_cast<long long>(B) * Ho * Wo;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int X = static_cast<int>(idx % Wo), Y = static_cast<int>((idx / Wo) % Ho), b = static_cast<int>(idx / (static_cast<long long>(Wo) * Ho));
    const float* p = in + (static_cast<long long>(b) * H + Y) * W + X;
    const long long step = axis == 0 ? W : 1;
    float s = 0.f;
    for (int d = 0; d < win; ++d) s += p[d * step];
    out[idx] = s * scale;
  }
}

// greedy selection of the k best mutually non-overlapping windows of one image per CTA (k passes of a block-wide
// arg-max; ties go to the lowest linear index).  A window (i, j) is dropped once |i - i*| <= kx and |j - j*| <= ky
// for an already chosen (i*, j*) -- the reference's inclusive overlap test.
__global__ void __launch_bounds__(1024) topk_nms_kernel(const float* __restrict__ S, int Ho, int Wo, int kx, int ky, int k,
                                                        int descending, int* __restrict__ boxes,
                                                        float* __restrict__ scores, int* __restrict__ count) {
  __shared__ float sv[32];
  __shared__ int si[32];
  __shared__ int chosen[64][2];
  __shared__ int nchosen;
  const int b = blockIdx.x;
  const float* s = S + static_cast<long long>(b) * Ho * Wo;
  const int n = Ho * Wo;
  if (threadIdx.x == 0) nchosen = 0;
  __syncthreads();
  for (int pass = 0; pass < k; ++pass) {
    float best = -INFINITY;
    int besti = 0x7fffffff;
    const int nc = nchosen;
    for (int idx = threadIdx.x; idx < n; idx += blockDim.x) {
      const int i = idx / Wo, j = idx % Wo;
      bool ok = true;
      for (int c = 0; c < nc; ++c) ok = ok && !(abs(i - chosen[c][0]) <= kx && abs(j - chosen[c][1]) <= ky);
      if (!ok) continue;
      const float v = descending ? s[idx] : -s[idx];
      if (v > best || (v == best && idx < besti)) { best = v; besti = idx; }
    }
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
      if (ov > best || (ov == best && oi < besti)) { best = ov; besti = oi; }
    }
    if ((threadIdx.x & 31) == 0) { sv[threadIdx.x >> 5] = best; si[threadIdx.x >> 5] = besti; }
    __syncthreads();
    if (threadIdx.x < 32) {
      best = threadIdx.x < (blockDim.x >> 5) ? sv[threadIdx.x] : -INFINITY;
      besti = threadIdx.x < (blockDim.x >> 5) ? si[threadIdx.x] : 0x7fffffff;
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
        if (ov > best || (ov == best && oi < besti)) { best = ov; besti = oi; }
      }
      if (threadIdx.x == 0) {
        if (besti != 0x7fffffff) {
          const int i = besti / Wo, j = besti % Wo;
          chosen[nc][0] = i; chosen[nc][1] = j;
          int* bx = boxes + (static_cast<long long>(b) * k + nc) * 4;
          bx[0] = i; bx[1] = j; bx[2] = i + kx; bx[3] = j + ky;
          scores[static_cast<long long>(b) * k + nc] = s[besti];
          nchosen = nc + 1;
        } else {
          nchosen = -(nc + 1);  // nothing left: stop (get_non_overlapping breaks when the frame is empty)
        }
      }
    }
    __syncthreads();
    if (nchosen < 0) break;
  }
  if (threadIdx.x == 0) count[b] = nchosen < 0 ? -nchosen - 1 : nchosen;
}

}  // namespace dm

using namespace dm;

static int grid_for(long long total) {
  return static_cast<int>(std::max<long long>(1, std::min<long long>((total + 255) / 256, 148 * 16)));
}

extern "C" int dm_patch_topk(const float* T, int B, int h, int w, int H, int W, int kx, int ky, int k, int descending,
                             float* work, int32_t* boxes, float* scores, int32_t* count, void* stream) {
  return abi_guard([&] {
    DM_CHECK(T && work && boxes && scores && count, "dm_patch_topk: null argument");
    DM_CHECK(B > 0 && h > 0 && w > 0 && H >= kx && W >= ky && kx > 0 && ky > 0, "dm_patch_topk: bad sizes");
    DM_CHECK(k > 0 && k <= 64, "dm_patch_topk: k must be in 1..64");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    // work: [B,H,W] resized map | [B,H,W-ky+1] row sums ; the pooled scores overwrite the first buffer
    float* up = work;
    float* rows = work + static_cast<size_t>(B) * H * W;
    const int Ho = H - kx + 1, Wo = W - ky + 1;
    tmap_resize_kernel<<<grid_for(static_cast<long long>(B) * H * W), 256, 0, s>>>(T, B, h, w, H, W, up);
    DM_CUDA(cudaGetLastError());
    box_sum_kernel<<<grid_for(static_cast<long long>(B) * H * Wo), 256, 0, s>>>(up, B, H, W, ky, 1, 1.f, rows);
    DM_CUDA(cudaGetLastError());
    box_sum_kernel<<<grid_for(static_cast<long long>(B) * Ho * Wo), 256, 0, s>>>(rows, B, H, Wo, kx, 0,
                                                                                 1.f / (static_cast<float>(kx) * ky), up);
    DM_CUDA(cudaGetLastError());
    topk_nms_kernel<<<B, 1024, 0, s>>>(up, Ho, Wo, kx, ky, k, descending, boxes, scores, count);
    DM_CUDA(cudaGetLastError());
  });
}

extern "C" int64_t dm_patch_topk_work_floats(int B, int H, int W, int ky) {
  return static_cast<int64_t>(B) * H * W + static_cast<int64_t>(B) * H * (W - ky + 1);
}
