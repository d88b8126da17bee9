// Host-side operator layer: prepares (tensor maps, tile shapes) and launches the CUDA kernels.
// Ops are prepared once per plan (buffers are static inside the engine's arena) and replayed.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <stdexcept>
#include <string>

#include "attention3.cuh"
#include "igemm.cuh"

namespace dm {

struct DmError : std::runtime_error {
  explicit DmError(const std::string& s) : std::runtime_error(s) {}
};
#define DM_CHECK(cond, msg)                                                                         \
  do {                                                                                              \
    if (!(cond)) throw ::dm::DmError(std::string(__FILE__) + ":" + std::to_string(__LINE__) + ": " + (msg)); \
  } while (0)
#define DM_CUDA(expr)                                                                      \
  do {                                                                                     \
    cudaError_t e__ = (expr);                                                              \
    if (e__ != cudaSuccess)                                                                \
      throw ::dm::DmError(std::string(__FILE__) + ":" + std::to_string(__LINE__) + ": " + #expr + ": " + \
                          cudaGetErrorString(e__));                                        \
  } while (0)

// ---- NHWC fp16 tensor view handed to TMA: dims as the conv sees them
struct ActView {
  const __half* ptr = nullptr;
  int N = 1, H = 1, W = 1, C = 0;  // addressable extent
  long long pix_stride = 0;        // elements between pixels (>= C)
};

struct IgemmDesc {
  int Nimg = 1, H = 1, W = 1;  // output pixel grid
  int nsrc = 1;
  ActView src[2];
  int nseg = 0;
  IgSeg seg[IG_MAX_SEG];
  const __half* Wt = nullptr;  // [N, K] K-major
  int N = 0, K = 0;
  long long w_ld = 0;  // weight row stride in elements (0 = K)
  int k_ragged = 0;    // allow K % 64 != 0: the tail chunk is zero-filled by TMA on both operands
  const float* bias = nullptr;
  const __half* rowbias = nullptr;
  int ld_rowbias = 0;
  const __half* residual = nullptr;
  long long ld_res = 0;
  void* out = nullptr;
  long long ld_out = 0;
  int out_f32 = 0, geglu = 0, act_silu = 0;
  const IgLossArgs* loss = nullptr;  // device-resident typicality epilogue arguments (conv_out, direct epilogue only)
  // GroupNorm statistics of the output formed in the epilogue (igemm.cuh: IgGn): rec = per-image records
  // [(2 * 4 * tiles_img + 1) * N floats]; tiles_img = 128-pixel tiles per image (0 = tiles_x * tiles_y, the conv layout;
  // a flattened Linear passes H * W / 128)
  struct {
    float* rec = nullptr;  // null: off
    int tiles_img = 0;
  } gn;
  int bn = 0;  // 0 = choose
  int cg = 0;  // CTAs per tile: 0 = choose, 1 = single CTA, 2 = CTA pair (cta_group::2)
};

struct IgemmOp {
  IgMaps maps;
  IgParams p;
  int bn = 0, grid = 0;
  int direct = 0;  // 1 = plain global-store epilogue (fp32 output or N-tile < 32 columns)
  int cg = 1;      // CTAs per tile (2 = CTA pair)
  int ng = 2;      // epilogue warpgroups (4 for short-K, epilogue-bound layers)
  int ws = 0;      // 1 = weight-stationary CTA pairs (short-K Linears with many M-tiles)
  double flops = 0;
};

void set_variant(const std::string& name, int value);  // "igemm_pair" / "gn_fused": 0, 1, or -1 = default
IgemmOp igemm_prepare(const IgemmDesc& d, int num_sms);
// true when a conv / Linear writing fp16 [Nimg, H*W, N] through the staged epilogue can form GroupNorm statistics there:
// whole 128-pixel tiles per image and whole N-tiles (holds for the conv layout and for the flattened Linear layout)
bool igemm_gn_fusable(int H, int W, int N);
size_t gn_record_floats(int Nimg, int HW, int C);  // size of the statistics record of an [Nimg, HW, C] tensor
int gn_epilogue_mode();  // variant "gn_epilogue" (DM_GN_EPILOGUE) bit mask: 1 = 3x3 convs, 2 = 1x1 convs / Linears form statistics
void igemm_launch(const IgemmOp& op, cudaStream_t s);
// convenience segment builders
void seg_conv3x3(IgemmDesc& d, int Cin_total, int C0);           // 9 taps over src0 (C0 ch) [+ src1]
void seg_conv3x3_s2(IgemmDesc& d, int C, int Nimg, bool vae_pad);  // 9 taps over the 4 parity planes
void seg_1x1(IgemmDesc& d, int C0, int C1);                       // plain GEMM over src0 [+ src1]

struct AttnDesc {
  int B = 0, heads = 8, D = 0, Tq = 0, Tk = 0;
  const __half *q = nullptr, *k = nullptr, *v = nullptr;
  long long ld_q = 0, ld_k = 0, ld_v = 0;  // token strides (elements)
  long long bs_q = 0, bs_k = 0, bs_v = 0;  // batch strides (elements)
  int kv_batches = 0;                      // extent of the K/V batch dim (context slots); 0 = B
  const int* kv_index = nullptr;
  __half* out = nullptr;
  long long ld_out = 0;
};
struct AttnOp {
  AttnMaps maps;
  AttnParams p;
  int D = 0;
  int v2 = 0;     // 1 = warp-specialised two-Q-tile kernel (attention2.cuh)
  int v3 = 0;     // 1 = its persistent successor (attention3.cuh): grid = min(#work items, #SMs)
  int xattn = 0;  // 1 = short-key-set kernel (<= 80 keys, P in tensor memory)
  int vattn = 0;  // 1 = single-head 512-wide kernel of the VAE mid block (two CTAs per query tile)
  dim3 grid;
  double flops = 0;
};
AttnOp attn_prepare(const AttnDesc& d);
void attn_launch(const AttnOp& op, cudaStream_t s);

// ---- norm / misc launchers
struct GnDesc {
  const __half* src0 = nullptr; int C0 = 0; long long ps0 = 0;
  const __half* src1 = nullptr; int C1 = 0; long long ps1 = 0;
  int Nimg = 0, HW = 0;
  const float *gamma = nullptr, *beta = nullptr;
  float eps = 1e-5f;
  int silu = 0;
  float* partial = nullptr;  // scratch [Nimg * splits * 2 * C] (two-kernel path: per-split per-channel shifted sums)
  float* ab = nullptr;       // scratch [Nimg * C * 2]: per-(image, channel) scale / shift
  unsigned* tickets = nullptr;  // scratch [Nimg], zero on entry, left zero
  __half* out = nullptr;     // dense [Nimg, HW, C0+C1]
};
int gn_splits(int Nimg, int HW);
int gn_launch_count(int HW, int C);  // kernels gn_launch() issues for this shape (1 = cluster-fused, 2 = stats + apply)
void gn_launch(const GnDesc& d, cudaStream_t s);
// GroupNorm over one or two dense sources whose statistics records the producing convs' epilogues left in rec0 / rec1
// (igemm.cuh: IgGn): fold + apply in one cluster kernel.  d.src0/C0 [, d.src1/C1], d.Nimg, d.HW, gamma / beta / eps / silu /
// out are read.  Requires HW % 128 == 0, C0 + C1 <= 2560, C0 and C1 multiples of 8, (C0 + C1) % 32 == 0.
bool gn_fold_apply_supported(int HW, int C0, int C1);
void gn_fold_apply_launch(const GnDesc& d, const float* rec0, const float* rec1, cudaStream_t s);
void layernorm_launch(const __half* x, long long ld_x, const float* gamma, const float* beta, float eps, long long rows,
                      int C, __half* out, long long ld_out, cudaStream_t s);
void patch3x3_launch(const float* x0, const int* x_index, const float* noise, const int* noise_index, const long long* t,
                     const float* ca, const float* cb, int Bf, int Cin, int H, int W, __half* out, cudaStream_t s,
                     int index_stride = 1, int sched_n = 0, int* err_flag = nullptr);
void repeat_rows_launch(const __half* in, long long rows, long long row_elems, int G, __half* out, cudaStream_t s);
int variant_prefix_share();  // 1 = run the (x_t, t)-only prefix of the U-Net once per draw in dm_typicality
void timestep_embed_launch(const long long* t, const int* t_index, int Bf, __half* out, cudaStream_t s);
void upsample_nearest_launch(const __half* in, int N, int H, int W, int C, int Ho, int Wo, __half* out, cudaStream_t s);
void space_to_planes_launch(const __half* in, int N, int H, int W, int C, int H2, int W2, __half* out, cudaStream_t s);
void loss_launch(const __half* pred, int ld_pred, const float* noise, const int* noise_index, const int* grid_row,
                 float* loss_f32, __half* grid_f16, float* eps_f32, int Bf, int HW, cudaStream_t s);
void tmap_launch(const __half* grid, int Bi, int N, int n_cond, int HW, float* T, cudaStream_t s);
void vae_sample_launch(const __half* h16, int ld_h, const __half* wq, const float* bq, const float* eps, float scaling,
                       int B, int HW, float* z, float* mean_out, float* logvar_out, cudaStream_t s);
void softmax_rows_launch(const float* S, long long ld_s, int rows, int cols, float scale, __half* P, long long ld_p,
                         cudaStream_t s);
void nhwc_to_nchw_mean_launch(const __half* in, int B, int E, int HW, int C, float* out, cudaStream_t s);

}  // namespace dm
