// Model-level C-ABI entry points (include/dm_abi.h): lifetime, weights, context, VAE encode, U-Net epsilon /
// loss rows, the Monte-Carlo typicality driver, DIFT features, introspection.
#include <algorithm>
#include <cstring>
#include <vector>

#include "../../include/dm_abi.h"
#include "abi_util.h"
#include "engine.h"

using namespace dm;

struct dm_engine {
  Engine eng;
  // device scratch for per-call index arrays / timesteps / the raw loss grid
  int* idx_dev = nullptr;
  size_t idx_cap = 0;
  long long* t_dev = nullptr;
  size_t t_cap = 0;
  __half* grid_dev = nullptr;
  size_t grid_cap = 0;
};

namespace {

// Every engine entry point runs on the engine's device (function attributes, arenas and plans are per device) and
// restores the caller's current device on exit.
struct DeviceScope {
  int prev = -1;
  explicit DeviceScope(int dev) {
    DM_CUDA(cudaGetDevice(&prev));
    if (prev != dev) DM_CUDA(cudaSetDevice(dev));
    else prev = -1;
  }
  ~DeviceScope() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

int* ensure_idx(dm_engine* h, size_t n) {
  if (h->idx_cap < n) {
    if (h->idx_dev) cudaFree(h->idx_dev);
    h->idx_cap = std::max<size_t>(n, 4096);
    DM_CUDA(cudaMalloc(reinterpret_cast<void**>(&h->idx_dev), h->idx_cap * sizeof(int)));
  }
  return h->idx_dev;
}
long long* ensure_t(dm_engine* h, size_t n) {
  if (h->t_cap < n) {
    if (h->t_dev) cudaFree(h->t_dev);
    h->t_cap = std::max<size_t>(n, 1024);
    DM_CUDA(cudaMalloc(reinterpret_cast<void**>(&h->t_dev), h->t_cap * sizeof(long long)));
  }
  return h->t_dev;
}
__half* ensure_grid(dm_engine* h, size_t n) {
  if (h->grid_cap < n) {
    if (h->grid_dev) cudaFree(h->grid_dev);
    h->grid_cap = n;
    DM_CUDA(cudaMalloc(reinterpret_cast<void**>(&h->grid_dev), h->grid_cap * sizeof(__half)));
  }
  return h->grid_dev;
}

// One U-Net micro-batch: rows b < Bf read latent x[x_index[b]], noise[noise_index[b]], t[t_index[b]],
// context slot ctx_idx[b] (all index arrays on the DEVICE).
Plan* unet_microbatch(Engine& e, int kind, int aux, const float* x, const int* x_index, const float* noise,
                      const int* noise_index, const long long* t, const int* t_index, const int* ctx_idx, int Bf, int h,
                      int w, cudaStream_t s, __half* fused_grid = nullptr) {
  Plan* p = e.get_plan(PlanKey{kind, Bf, h, w, aux});
  // kPlanUnet with aux = G > 1: the input patch matrix is built once per group of G forwards sharing (x_t, t)
  const int G = (kind == kPlanUnet && aux > 1) ? aux : 1;
  patch3x3_launch(x, x_index, noise, noise_index, t, e.sched_a, e.sched_b, Bf / G, 4, h, w, p->a_in, s, G, e.sched_n,
                  e.err_dev);
  timestep_embed_launch(t, t_index, Bf, p->temb_sin, s);
  DM_CUDA(cudaMemcpyAsync(p->ctx_idx, ctx_idx, Bf * sizeof(int), cudaMemcpyDeviceToDevice, s));
  if (kind == kPlanUnet && p->loss_args) {
    // conv_out's fused typicality epilogue: grid rows of this micro-batch start at `fused_grid` (null = epilogue off)
    const IgLossArgs la{noise, noise_index, nullptr, fused_grid};
    DM_CUDA(cudaMemcpyAsync(p->loss_args, &la, sizeof(la), cudaMemcpyHostToDevice, s));
  }
  e.launch_count += 2;
  e.run_plan(p, s);
  if (kind == kPlanUnet) e.last_unet_plan = p;
  return p;
}

void check_slots(const int32_t* slots, int n) {
  for (int i = 0; i < n; ++i)
    DM_CHECK(slots[i] >= 0 && slots[i] < kMaxCtxSlots, "context slot out of range");
}

// default micro-batch cap: 56 forwards at 64x64 latents (fills the 148 SMs at every U-Net level),
// scaled down with the latent area so the plan arena stays a few GB at 128x128
inline int default_max_forwards(int h, int w) {
  const long long cap = 56ll * 4096 / std::max(1ll, static_cast<long long>(h) * w);
  return static_cast<int>(std::max(8ll, std::min(56ll, cap)));
}

}  // namespace

extern "C" int dm_create(int device, dm_engine** out) {
  return abi_guard([&] {
    DM_CHECK(out != nullptr, "dm_create: null output");
    int ndev = 0;
    DM_CUDA(cudaGetDeviceCount(&ndev));
    DM_CHECK(device >= 0 && device < ndev, "dm_create: no such CUDA device");
    DM_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    DM_CUDA(cudaGetDeviceProperties(&prop, device));
    DM_CHECK(prop.major == 10, std::string("this engine is built for sm_100a (B200) only; device is sm_") +
                                   std::to_string(prop.major) + std::to_string(prop.minor));
    dm_engine* h = new dm_engine();
    h->eng.device = device;
    h->eng.num_sms = prop.multiProcessorCount;
    const char* g = getenv("DM_GRAPH");
    if (g && g[0] == '0') h->eng.use_graph = false;
    *out = h;
  });
}

extern "C" int dm_destroy(dm_engine* h) {
  return abi_guard([&] {
    if (!h) return;
    cudaDeviceSynchronize();
    if (h->idx_dev) cudaFree(h->idx_dev);
    if (h->t_dev) cudaFree(h->t_dev);
    if (h->grid_dev) cudaFree(h->grid_dev);
    delete h;
  });
}

extern "C" int dm_load_tensor(dm_engine* h, const char* key, const void* host_ptr, int dtype, int ndim,
                              const int64_t* shape) {
  return abi_guard([&] {
    DM_CHECK(h && key, "dm_load_tensor: null argument");
    h->eng.load_tensor(key, host_ptr, dtype, ndim, shape);
  });
}
extern "C" int dm_finalize_weights(dm_engine* h) {
  return abi_guard([&] {
    DM_CHECK(h, "null engine");
    DeviceScope dev_scope(h->eng.device);
    h->eng.finalize();
  });
}
extern "C" int dm_save_packed(dm_engine* h, const char* path) {
  return abi_guard([&] {
    DM_CHECK(h && path, "dm_save_packed: null argument");
    DeviceScope dev_scope(h->eng.device);
    h->eng.save_packed(path);
  });
}
extern "C" int dm_load_packed(dm_engine* h, const char* path) {
  return abi_guard([&] {
    DM_CHECK(h && path, "dm_load_packed: null argument");
    DeviceScope dev_scope(h->eng.device);
    h->eng.load_packed(path);
  });
}
extern "C" int dm_set_schedule(dm_engine* h, const float* a, const float* b, int n) {
  return abi_guard([&] {
    DM_CHECK(h, "null engine");
    DeviceScope dev_scope(h->eng.device);
    h->eng.set_schedule(a, b, n);
  });
}
extern "C" int dm_set_context(dm_engine* h, int slot, const float* ctx, void* stream) {
  return abi_guard([&] {
    DM_CHECK(h && ctx, "dm_set_context: null argument");
    DeviceScope dev_scope(h->eng.device);
    h->eng.set_context(slot, ctx, static_cast<cudaStream_t>(stream));
  });
}

extern "C" int dm_vae_encode(dm_engine* h, const float* img, const float* eps, int B, int H, int W, float* z, float* mean,
                             float* logvar, void* stream) {
  return abi_guard([&] {
    DM_CHECK(h && img, "dm_vae_encode: null argument");
    DeviceScope dev_scope(h->eng.device);
    h->eng.check_async_error();
    DM_CHECK(B > 0 && H >= 8 && W >= 8, "dm_vae_encode: empty batch or image smaller than 8x8");
    Engine& e = h->eng;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int ho = H / 2 / 2 / 2, wo = W / 2 / 2 / 2;
    const long long img_px = static_cast<long long>(H) * W;
    // bound the arena: ~1.3 KB of live activations per input pixel
    int mb = static_cast<int>(std::max<long long>(1, std::min<long long>(B, (8ll << 20) / img_px)));
    for (int b0 = 0; b0 < B; b0 += mb) {
      const int nb = std::min(mb, B - b0);
      Plan* p = e.get_plan(PlanKey{kPlanVae, nb, H, W, 0});
      patch3x3_launch(img + static_cast<size_t>(b0) * 3 * img_px, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nb, 3,
                      H, W, p->a_in, s);
      e.launch_count += 1;
      e.run_plan(p, s);
      const size_t o = static_cast<size_t>(b0) * 4 * ho * wo;
      vae_sample_launch(p->out, 16, e.H("vae.quant_conv.weight"), e.F("vae.quant_conv.bias"), eps ? eps + o : nullptr,
                        0.18215f, nb, ho * wo, z ? z + o : nullptr, mean ? mean + o : nullptr,
                        logvar ? logvar + o : nullptr, s);
      e.launch_count += 1;
    }
  });
}

extern "C" int dm_unet_rows(dm_engine* h, const float* x, const float* noise, const int64_t* t, const int32_t* x_index,
                            const int32_t* noise_index, const int32_t* ctx_slots, int M, int hh, int ww,
                            float* loss_out, float* eps_out, int max_forwards, void* stream) {
  return abi_guard([&] {
    DM_CHECK(h && x && t && ctx_slots, "dm_unet_rows: null argument");
    DeviceScope dev_scope(h->eng.device);
    h->eng.check_async_error();
    DM_CHECK(M > 0 && hh > 0 && ww > 0, "dm_unet_rows: empty problem");
    DM_CHECK(noise || !loss_out, "dm_unet_rows: a loss needs the noise");
    Engine& e = h->eng;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    check_slots(ctx_slots, M);
    std::vector<int> host(3 * static_cast<size_t>(M));
    for (int i = 0; i < M; ++i) {
      host[i] = x_index ? x_index[i] : i;
      host[M + i] = noise_index ? noise_index[i] : i;
      host[2 * M + i] = ctx_slots[i];
    }
    int* dev = ensure_idx(h, host.size());
    DM_CUDA(cudaMemcpyAsync(dev, host.data(), host.size() * sizeof(int), cudaMemcpyHostToDevice, s));
    const int Bf = std::min(M, max_forwards > 0 ? max_forwards : default_max_forwards(hh, ww));
    const int HW = hh * ww;
    for (int m0 = 0; m0 < M; m0 += Bf) {
      const int nb = std::min(Bf, M - m0);
      // when there is no noise, x rows are already noisy and t is indexed like x
      const int* tix = noise ? dev + M + m0 : dev + m0;
      Plan* p = unet_microbatch(e, kPlanUnet, 0, x, dev + m0, noise, dev + M + m0, reinterpret_cast<const long long*>(t),
                                tix, dev + 2 * M + m0, nb, hh, ww, s);
      loss_launch(p->out, 16, noise, dev + M + m0, nullptr, loss_out ? loss_out + static_cast<size_t>(m0) * 4 * HW : nullptr,
                  nullptr, eps_out ? eps_out + static_cast<size_t>(m0) * 4 * HW : nullptr, nb, HW, s);
      e.launch_count += 1;
    }
  });
}

extern "C" int dm_unet_eps(dm_engine* h, const float* x_noisy, const int64_t* t, const int32_t* ctx_slots, int Bf, int hh,
                           int ww, float* eps_out, void* stream) {
  return dm_unet_rows(h, x_noisy, nullptr, t, nullptr, nullptr, ctx_slots, Bf, hh, ww, nullptr, eps_out, 0, stream);
}

extern "C" int dm_compute_loss(dm_engine* h, const float* x0, const float* noise, const int64_t* t,
                               const int32_t* ctx_slots, int S, int n_cond, int hh, int ww, float* loss_out,
                               void* stream) {
  std::vector<int32_t> xi(static_cast<size_t>(S) * n_cond, 0), ni(xi.size()), cs(xi.size());
  for (int c = 0; c < n_cond; ++c)
    for (int s = 0; s < S; ++s) {
      ni[static_cast<size_t>(c) * S + s] = s;
      cs[static_cast<size_t>(c) * S + s] = ctx_slots ? ctx_slots[c] : -1;
    }
  return dm_unet_rows(h, x0, noise, t, xi.data(), ni.data(), cs.data(), S * n_cond, hh, ww, loss_out, nullptr, 0, stream);
}

extern "C" int dm_typicality(dm_engine* h, const float* x0, const float* noise, const int64_t* t,
                             const int32_t* ctx_slots, int Bi, int N, int n_cond, int hh, int ww, void* grid_out,
                             float* T_out, int max_forwards, void* stream) {
  return abi_guard([&] {
    DM_CHECK(h && x0 && noise && t && ctx_slots, "dm_typicality: null argument");
    DeviceScope dev_scope(h->eng.device);
    h->eng.check_async_error();
    DM_CHECK(Bi > 0 && N > 0 && n_cond > 0 && hh > 0 && ww > 0, "dm_typicality: empty problem");
    DM_CHECK(grid_out || T_out, "dm_typicality: no output requested");
    DM_CHECK(!T_out || n_cond >= 2, "dm_typicality: T(x|c) needs a condition and the unconditional slot");
    Engine& e = h->eng;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    check_slots(ctx_slots, n_cond);
    const long long F = static_cast<long long>(Bi) * N * n_cond;
    DM_CHECK(F < (1ll << 30), "dm_typicality: too many forwards in one call");
    // forward f = (img*N + sample)*n_cond + cond  == row of the raw grid [Bi][N][n_cond]
    std::vector<int> host(3 * static_cast<size_t>(F));
    for (long long f = 0; f < F; ++f) {
      const int c = static_cast<int>(f % n_cond);
      const int sm = static_cast<int>((f / n_cond) % N);
      const int img = static_cast<int>(f / (static_cast<long long>(n_cond) * N));
      host[f] = img;
      host[F + f] = sm;
      host[2 * F + f] = ctx_slots[c];
    }
    int* dev = ensure_idx(h, host.size());
    DM_CUDA(cudaMemcpyAsync(dev, host.data(), host.size() * sizeof(int), cudaMemcpyHostToDevice, s));
    const int HW = hh * ww;
    __half* grid = grid_out ? static_cast<__half*>(grid_out) : ensure_grid(h, static_cast<size_t>(F) * 4 * HW);
    // balanced micro-batches: as few as the cap allows, equal sizes (whole (eps,t) draws: multiples of n_cond), so no
    // small remainder batch under-fills the 148 SMs
    const long long want = max_forwards > 0 ? max_forwards : default_max_forwards(hh, ww);
    const long long cap = std::max<long long>(n_cond, want / n_cond * n_cond);  // whole draws only
    const long long n_mb = (F + cap - 1) / cap;
    const int Bf = static_cast<int>(((F + n_mb - 1) / n_mb + n_cond - 1) / n_cond * n_cond);
    for (long long f0 = 0; f0 < F; f0 += Bf) {
      const int nb = static_cast<int>(std::min<long long>(Bf, F - f0));
      // the n_cond forwards of one (eps, t) draw are consecutive rows: share their context-free prefix
      const int share = (n_cond > 1 && variant_prefix_share()) ? n_cond : 0;
      // (pred - eps)^2 -> fp16 grid rows [f0, f0 + nb) is formed by conv_out's epilogue: no separate loss pass
      unet_microbatch(e, kPlanUnet, share, x0, dev + f0, noise, dev + F + f0, reinterpret_cast<const long long*>(t), dev + F + f0,
                      dev + 2 * F + f0, nb, hh, ww, s, grid + static_cast<size_t>(f0) * 4 * HW);
    }
    if (T_out) {
      tmap_launch(grid, Bi, N, n_cond, HW, T_out, s);
      e.launch_count += 1;
    }
  });
}

extern "C" int dm_dift_shape(int hh, int ww, int up_ft_index, int* C, int* ho, int* wo) {
  return abi_guard([&] {
    DM_CHECK(up_ft_index >= 0 && up_ft_index <= 2, "dm_dift: up_ft_index must be 0, 1 or 2 (3 is the full U-Net)");
    int sh[4] = {hh, 0, 0, 0}, sw[4] = {ww, 0, 0, 0};
    for (int i = 1; i < 4; ++i) { sh[i] = (sh[i - 1] + 1) / 2; sw[i] = (sw[i - 1] + 1) / 2; }
    static const int ch[3] = {1280, 1280, 640};
    // after up block i (with its upsampler) the map has the resolution of down level 2-i
    if (C) *C = ch[up_ft_index];
    if (ho) *ho = sh[2 - up_ft_index];
    if (wo) *wo = sw[2 - up_ft_index];
  });
}

extern "C" int dm_dift(dm_engine* h, const float* latents, const float* noise, int64_t t, int ctx_slot, int B, int E,
                       int hh, int ww, int up_ft_index, float* feat_out, void* stream) {
  return abi_guard([&] {
    DM_CHECK(h && latents && feat_out, "dm_dift: null argument");
    DeviceScope dev_scope(h->eng.device);
    h->eng.check_async_error();
    DM_CHECK(B > 0 && E > 0, "dm_dift: empty batch");
    DM_CHECK(up_ft_index >= 0 && up_ft_index <= 2, "dm_dift: up_ft_index must be 0, 1 or 2");
    DM_CHECK(t >= 0 && t < h->eng.sched_n, "dm_dift: timestep out of range");
    Engine& e = h->eng;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    int32_t slot = ctx_slot;
    check_slots(&slot, 1);
    const int imgs_per_mb = std::max(1, 64 / E);
    const int max_rows = imgs_per_mb * E;
    std::vector<int> host(2 * static_cast<size_t>(max_rows));
    for (int i = 0; i < max_rows; ++i) { host[i] = i; host[max_rows + i] = ctx_slot; }
    int* dev = ensure_idx(h, host.size());
    DM_CUDA(cudaMemcpyAsync(dev, host.data(), host.size() * sizeof(int), cudaMemcpyHostToDevice, s));
    std::vector<long long> th(max_rows, t);
    long long* tdev = ensure_t(h, max_rows);
    DM_CUDA(cudaMemcpyAsync(tdev, th.data(), th.size() * sizeof(long long), cudaMemcpyHostToDevice, s));
    const size_t lat = static_cast<size_t>(4) * hh * ww;
    for (int b0 = 0; b0 < B; b0 += imgs_per_mb) {
      const int nb = std::min(imgs_per_mb, B - b0);
      const int rows = nb * E;
      Plan* p = unet_microbatch(e, kPlanDift, up_ft_index, latents + static_cast<size_t>(b0) * E * lat, dev,
                                noise ? noise + static_cast<size_t>(b0) * E * lat : nullptr, dev, tdev, dev, dev + max_rows,
                                rows, hh, ww, s);
      nhwc_to_nchw_mean_launch(p->out, nb, E, p->out_H * p->out_W, p->out_C,
                               feat_out + static_cast<size_t>(b0) * p->out_C * p->out_H * p->out_W, s);
      e.launch_count += 1;
    }
  });
}

extern "C" int64_t dm_launch_count(dm_engine* h) { return h ? h->eng.launch_count : -1; }
extern "C" double dm_flop_count(dm_engine* h) { return h ? h->eng.flop_count : -1.0; }

extern "C" int dm_debug_keep(dm_engine* h, int on) {
  return abi_guard([&] {
    DM_CHECK(h, "null engine");
    if (h->eng.debug_keep != (on != 0)) {
      h->eng.last_unet_plan = nullptr;
      h->eng.plans.clear();
    }
    h->eng.debug_keep = on != 0;
  });
}

extern "C" int64_t dm_debug_fetch(dm_engine* h, const char* name, float* out_dev, int64_t capacity, int* dims4,
                                  void* stream) {
  int64_t n = -1;
  int rc = abi_guard([&] {
    DM_CHECK(h && name, "null argument");
    DeviceScope dev_scope(h->eng.device);
    Plan* p = h->eng.last_unet_plan;
    DM_CHECK(p != nullptr, "no U-Net forward has run with dm_debug_keep(1)");
    auto it = p->taps.find(name);
    if (it == p->taps.end()) it = p->taps.find(std::string("unet.") + name);
    DM_CHECK(it != p->taps.end(), std::string("no such tap: ") + name);
    const Act& a = it->second;
    n = static_cast<int64_t>(a.pixels()) * a.C;
    if (dims4) { dims4[0] = a.N; dims4[1] = a.C; dims4[2] = a.H; dims4[3] = a.W; }
    if (out_dev) {
      DM_CHECK(capacity >= n, "dm_debug_fetch: output buffer too small");
      nhwc_to_nchw_mean_launch(reinterpret_cast<const __half*>(p->arena + a.off), a.N, 1, a.H * a.W, a.C, out_dev,
                               static_cast<cudaStream_t>(stream));
    }
  });
  return rc == 0 ? n : -1;
}

namespace {
void profile_plan(Engine& e, PlanKey key, int iters, double* ms_igemm, double* ms_attn, double* ms_other,
                  double* flops_igemm, double* flops_attn) {
  Plan* p = e.get_plan(key);
  cudaStream_t s = e.cap_stream;
  const size_t n = p->steps.size();
  std::vector<cudaEvent_t> ev(n + 1);
  for (auto& x : ev) DM_CUDA(cudaEventCreate(&x));
  for (auto& st : p->steps) st.run(s);  // warm-up
  DM_CUDA(cudaStreamSynchronize(s));
  double acc[3] = {0, 0, 0};
  const bool verbose = getenv("DM_PROFILE_VERBOSE") != nullptr;
  std::vector<double> per(n, 0.0);
  for (int it = 0; it < iters; ++it) {
    DM_CUDA(cudaEventRecord(ev[0], s));
    for (size_t i = 0; i < n; ++i) {
      p->steps[i].run(s);
      DM_CUDA(cudaEventRecord(ev[i + 1], s));
    }
    DM_CUDA(cudaStreamSynchronize(s));
    for (size_t i = 0; i < n; ++i) {
      float ms = 0;
      DM_CUDA(cudaEventElapsedTime(&ms, ev[i], ev[i + 1]));
      acc[p->steps[i].cls] += ms;
      per[i] += ms;
    }
  }
  if (verbose)
    for (size_t i = 0; i < n; ++i)
      printf("DMPROF %-70s cls=%d ms=%.4f gflop=%.2f\n", p->steps[i].name.c_str(), p->steps[i].cls, per[i] / iters,
             p->steps[i].flops * 1e-9);
  for (auto& x : ev) cudaEventDestroy(x);
  if (ms_igemm) *ms_igemm = acc[0] / iters;
  if (ms_attn) *ms_attn = acc[1] / iters;
  if (ms_other) *ms_other = acc[2] / iters;
  if (flops_igemm) *flops_igemm = p->flops_igemm;
  if (flops_attn) *flops_attn = p->flops_attn;
}
}  // namespace

extern "C" int dm_profile_unet(dm_engine* h, int Bf, int hh, int ww, int iters, double* ms_igemm, double* ms_attn,
                               double* ms_other, double* flops_igemm, double* flops_attn) {
  return abi_guard([&] {
    DM_CHECK(h && iters > 0, "dm_profile_unet: bad arguments");
    DeviceScope dev_scope(h->eng.device);
    // DM_PROFILE_KIND=vae profiles the VAE-encoder plan instead (hh, ww = image size): development aid
    const char* kind_env = getenv("DM_PROFILE_KIND");
    const int kind = (kind_env && std::string(kind_env) == "vae") ? kPlanVae : kPlanUnet;
    profile_plan(h->eng, PlanKey{kind, Bf, hh, ww, 0}, iters, ms_igemm, ms_attn, ms_other, flops_igemm, flops_attn);
  });
}

extern "C" int dm_profile_plan(dm_engine* h, int kind, int Bf, int hh, int ww, int aux, int iters, double* ms_by_class,
                               double* flops_by_class) {
  return abi_guard([&] {
    DM_CHECK(h && iters > 0 && ms_by_class && flops_by_class, "dm_profile_plan: bad arguments");
    DM_CHECK(kind == kPlanUnet || kind == kPlanDift || kind == kPlanVae, "dm_profile_plan: kind must be 0 (U-Net), 1 (DIFT) or 2 (VAE)");
    DeviceScope dev_scope(h->eng.device);
    profile_plan(h->eng, PlanKey{kind, Bf, hh, ww, aux}, iters, &ms_by_class[0], &ms_by_class[1], &ms_by_class[2],
                 &flops_by_class[0], &flops_by_class[1]);
  });
}
