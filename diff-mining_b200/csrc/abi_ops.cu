// Operator-level C-ABI entry points (include/dm_abi.h, "operator-level" section): thin wrappers that prepare and
// launch exactly the kernels the engine uses, for unit tests against torch ops on the GPU box.
#include "../../include/dm_abi.h"
#include "abi_util.h"

#include <cstdio>
#include <cstdlib>
#include <vector>

namespace dm {
std::string& last_error_ref() {
  static thread_local std::string s;
  return s;
}
int device_sm_count() {
  int dev = 0, n = 0;
  DM_CUDA(cudaGetDevice(&dev));
  DM_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
  return n;
}
}  // namespace dm

using namespace dm;

extern "C" const char* dm_last_error(void) { return last_error_ref().c_str(); }
extern "C" int dm_abi_version(void) { return 1; }

extern "C" int dm_op_conv(const void* x, const void* x2, int N, int H, int W, int C0, int C1, const void* w, int Cout,
                          int ks, int stride, int vae_pad, const float* bias, const void* rowbias, const void* residual,
                          void* out, int out_f32, int geglu, int act_silu, int bn, void* stream) {
  return abi_guard([&] {
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    DM_CHECK(ks == 1 || ks == 3, "conv: kernel size must be 1 or 3");
    DM_CHECK(stride == 1 || (stride == 2 && ks == 3 && !x2), "conv: stride 2 needs a single-source 3x3");
    DM_CHECK(C0 % 64 == 0 && C1 % 64 == 0, "conv: channels must be multiples of 64");
    IgemmDesc d;
    d.src[0] = ActView{static_cast<const __half*>(x), N, H, W, C0, C0};
    d.nsrc = 1;
    if (x2) {
      d.src[1] = ActView{static_cast<const __half*>(x2), N, H, W, C1, C1};
      d.nsrc = 2;
    }
    __half* planes = nullptr;
    int Ho = H, Wo = W;
    if (stride == 2) {
      const int H2 = (H + 1) / 2, W2 = (W + 1) / 2;
      DM_CUDA(cudaMalloc(&planes, 4ull * N * H2 * W2 * C0 * sizeof(__half)));
      space_to_planes_launch(static_cast<const __half*>(x), N, H, W, C0, H2, W2, planes, s);
      d.src[0] = ActView{planes, 4 * N, H2, W2, C0, C0};
      Ho = vae_pad ? H / 2 : H2;
      Wo = vae_pad ? W / 2 : W2;
      seg_conv3x3_s2(d, C0, N, vae_pad != 0);
    } else if (ks == 3) {
      seg_conv3x3(d, C0 + C1, C0);
    } else {
      seg_1x1(d, C0, C1);
    }
    d.Nimg = N; d.H = Ho; d.W = Wo;
    d.Wt = static_cast<const __half*>(w);
    d.N = Cout;
    d.K = ks * ks * (C0 + C1);
    d.bias = bias;
    d.rowbias = static_cast<const __half*>(rowbias);
    d.ld_rowbias = Cout;
    d.residual = static_cast<const __half*>(residual);
    d.ld_res = Cout;
    d.out = out;
    d.ld_out = geglu ? Cout / 2 : Cout;
    d.out_f32 = out_f32; d.geglu = geglu; d.act_silu = act_silu;
    d.bn = bn;
    IgemmOp op = igemm_prepare(d, device_sm_count());
    igemm_launch(op, s);
    if (planes) {
      DM_CUDA(cudaStreamSynchronize(s));
      DM_CUDA(cudaFree(planes));
    }
  });
}

extern "C" int dm_op_conv_gn(const void* x, int N, int H, int W, int C0, const void* w, int Cout, int ks, const float* bias,
                             const void* rowbias, const void* residual, const float* gamma, const float* beta, float eps,
                             int silu, void* out, void* gn_out, void* stream) {
  return abi_guard([&] {
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    DM_CHECK(ks == 1 || ks == 3, "conv_gn: kernel size must be 1 or 3");
    DM_CHECK(C0 % 64 == 0 && Cout % 64 == 0 && Cout <= 1280, "conv_gn: channels must be multiples of 64, Cout <= 1280");
    DM_CHECK(igemm_gn_fusable(H, W, Cout) && gn_fold_apply_supported(H * W, Cout, 0),
             "conv_gn: this shape cannot form GroupNorm statistics in the conv epilogue");
    IgemmDesc d;
    d.src[0] = ActView{static_cast<const __half*>(x), N, H, W, C0, C0};
    d.nsrc = 1;
    if (ks == 3) seg_conv3x3(d, C0, C0);
    else seg_1x1(d, C0, 0);
    d.Nimg = N; d.H = H; d.W = W;
    d.Wt = static_cast<const __half*>(w);
    d.N = Cout;
    d.K = ks * ks * C0;
    d.bias = bias;
    d.rowbias = static_cast<const __half*>(rowbias);
    d.ld_rowbias = Cout;
    d.residual = static_cast<const __half*>(residual);
    d.ld_res = Cout;
    d.out = out;
    d.ld_out = Cout;
    float* rec = nullptr;
    DM_CUDA(cudaMalloc(&rec, sizeof(float) * gn_record_floats(N, H * W, Cout)));
    float* scratch = rec;
    d.gn.rec = rec;
    IgemmOp op = igemm_prepare(d, device_sm_count());
    igemm_launch(op, s);
    GnDesc g;
    g.src0 = static_cast<const __half*>(out); g.C0 = Cout; g.ps0 = Cout;
    g.Nimg = N; g.HW = H * W; g.gamma = gamma; g.beta = beta; g.eps = eps; g.silu = silu;
    g.out = static_cast<__half*>(gn_out);
    gn_fold_apply_launch(g, rec, nullptr, s);
    DM_CUDA(cudaStreamSynchronize(s));
    DM_CUDA(cudaFree(scratch));
  });
}

extern "C" int dm_op_attention(const void* q, const void* k, const void* v, int64_t ld_q, int64_t ld_k, int64_t ld_v,
                               int64_t bs_q, int64_t bs_k, int64_t bs_v, int B, int heads, int D, int Tq, int Tk,
                               int kv_batches, const int32_t* kv_index_dev, void* out, int64_t ld_out, void* stream) {
  return abi_guard([&] {
    AttnDesc d;
    d.B = B; d.heads = heads; d.D = D; d.Tq = Tq; d.Tk = Tk;
    d.q = static_cast<const __half*>(q); d.k = static_cast<const __half*>(k); d.v = static_cast<const __half*>(v);
    d.ld_q = ld_q; d.ld_k = ld_k; d.ld_v = ld_v;
    d.bs_q = bs_q; d.bs_k = bs_k; d.bs_v = bs_v;
    d.kv_batches = kv_batches;
    d.kv_index = kv_index_dev;
    d.out = static_cast<__half*>(out);
    d.ld_out = ld_out;
    AttnOp op = attn_prepare(d);
    attn_launch(op, static_cast<cudaStream_t>(stream));
  });
}

extern "C" int dm_op_groupnorm(const void* x, const void* x2, int N, int HW, int C0, int C1, const float* gamma,
                               const float* beta, float eps, int silu, void* out, void* stream) {
  return abi_guard([&] {
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    GnDesc d;
    d.src0 = static_cast<const __half*>(x); d.C0 = C0; d.ps0 = C0;
    d.src1 = static_cast<const __half*>(x2); d.C1 = x2 ? C1 : 0; d.ps1 = C1;
    d.Nimg = N; d.HW = HW; d.gamma = gamma; d.beta = beta; d.eps = eps; d.silu = silu;
    d.out = static_cast<__half*>(out);
    float* partial = nullptr;
    const size_t n_part = 2ull * (C0 + d.C1) * N * gn_splits(N, HW), n_ab = 2ull * N * (C0 + d.C1);
    DM_CUDA(cudaMalloc(&partial, sizeof(float) * (n_part + n_ab) + sizeof(unsigned) * N));
    d.partial = partial;
    d.ab = partial + n_part;
    d.tickets = reinterpret_cast<unsigned*>(partial + n_part + n_ab);
    DM_CUDA(cudaMemsetAsync(d.tickets, 0, sizeof(unsigned) * N, s));
    gn_launch(d, s);
    DM_CUDA(cudaStreamSynchronize(s));
    DM_CUDA(cudaFree(partial));
  });
}

extern "C" int dm_op_set_variant(const char* name, int value) {
  return abi_guard([&] {
    DM_CHECK(name != nullptr, "dm_op_set_variant: null name");
    set_variant(name, value);
  });
}

extern "C" int dm_op_layernorm(const void* x, int64_t rows, int C, const float* gamma, const float* beta, float eps,
                               void* out, void* stream) {
  return abi_guard([&] {
    layernorm_launch(static_cast<const __half*>(x), C, gamma, beta, eps, rows, C, static_cast<__half*>(out), C,
                     static_cast<cudaStream_t>(stream));
  });
}
