// Engine implementation: weight intake + packing, context K/V caches, plan cache, plan replay (eager / CUDA graph).
#include "engine.h"

#include <algorithm>
#include <cmath>
#include <cstring>

#include "plan_builder.h"

namespace dm {

// ------------------------------------------------------------------ arena planner
size_t ArenaPlanner::alloc(size_t bytes) {
  if (bytes == 0) bytes = 1024;
  for (auto it = free_.begin(); it != free_.end(); ++it) {
    if (it->second >= bytes) {
      const size_t off = it->first, sz = it->second;
      free_.erase(it);
      if (sz > bytes) free_[off + bytes] = sz - bytes;
      live_[off] = bytes;
      return off;
    }
  }
  // extend the top (merging with a free block that touches the top)
  size_t off = top_;
  if (!free_.empty()) {
    auto last = std::prev(free_.end());
    if (last->first + last->second == top_) {
      off = last->first;
      free_.erase(last);
    }
  }
  top_ = off + bytes;
  high_ = std::max(high_, top_);
  live_[off] = bytes;
  return off;
}
void ArenaPlanner::free(size_t off) {
  auto it = live_.find(off);
  if (it == live_.end()) return;
  size_t sz = it->second;
  live_.erase(it);
  auto nx = free_.lower_bound(off);
  if (nx != free_.end() && off + sz == nx->first) {
    sz += nx->second;
    nx = free_.erase(nx);
  }
  if (nx != free_.begin()) {
    auto pv = std::prev(nx);
    if (pv->first + pv->second == off) {
      pv->second += sz;
      return;
    }
  }
  free_[off] = sz;
}

Plan::~Plan() {
  if (graph) cudaGraphExecDestroy(graph);
  if (arena) cudaFree(arena);
}

Engine::~Engine() {
  plans.clear();
  for (void* p : owned) cudaFree(p);
  if (err_host) cudaFreeHost(err_host);
  if (cap_stream) cudaStreamDestroy(cap_stream);
}

void* Engine::dmalloc(size_t bytes) {
  void* p = nullptr;
  DM_CUDA(cudaMalloc(&p, std::max<size_t>(bytes, 256)));
  owned.push_back(p);
  return p;
}
const __half* Engine::H(const std::string& k) const {
  auto it = wh.find(k);
  DM_CHECK(it != wh.end(), "missing packed weight '" + k + "' (was the model loaded and finalized?)");
  return it->second;
}
const float* Engine::F(const std::string& k) const {
  auto it = wf.find(k);
  DM_CHECK(it != wf.end(), "missing packed vector '" + k + "' (was the model loaded and finalized?)");
  return it->second;
}

// ------------------------------------------------------------------ weight intake
static inline __half bf16_to_half(uint16_t b) {
  uint32_t u = static_cast<uint32_t>(b) << 16;
  float f;
  std::memcpy(&f, &u, 4);
  return __float2half_rn(f);
}

void Engine::load_tensor(const std::string& key, const void* host, int dtype, int ndim, const int64_t* shape) {
  DM_CHECK(!finalized, "load_tensor after finalize_weights");
  DM_CHECK(host != nullptr && ndim >= 1 && ndim <= 4, "load_tensor: bad arguments for " + key);
  HostTensor t;
  t.shape.assign(shape, shape + ndim);
  const int64_t n = t.numel();
  t.data.resize(n);
  if (dtype == 0) {
    const float* f = static_cast<const float*>(host);
    for (int64_t i = 0; i < n; ++i) t.data[i] = __float2half_rn(f[i]);
  } else if (dtype == 1) {
    std::memcpy(t.data.data(), host, n * sizeof(__half));
  } else if (dtype == 2) {
    const uint16_t* b = static_cast<const uint16_t*>(host);
    for (int64_t i = 0; i < n; ++i) t.data[i] = bf16_to_half(b[i]);
  } else {
    DM_CHECK(false, "load_tensor: unknown dtype tag");
  }
  staging[key] = std::move(t);
}

namespace {

void alloc_kv_cache(Engine& e, const std::string& a2, int C) {
  __half* kv = static_cast<__half*>(e.dmalloc(static_cast<size_t>(kMaxCtxSlots) * kCtxTokens * 2 * C * sizeof(__half)));
  DM_CUDA(cudaMemset(kv, 0, static_cast<size_t>(kMaxCtxSlots) * kCtxTokens * 2 * C * sizeof(__half)));
  e.kv_cache[a2] = kv;
  e.kv_C[a2] = C;
  e.xattn_layers.push_back(a2);
}

struct Packer {
  Engine& e;
  const HostTensor& get(const std::string& k) {
    auto it = e.staging.find(k);
    DM_CHECK(it != e.staging.end(), "state dict is missing '" + k + "'");
    return it->second;
  }
  void expect(const HostTensor& t, const std::string& k, std::initializer_list<int64_t> shp) {
    DM_CHECK(t.shape == std::vector<int64_t>(shp), "tensor '" + k + "' has an unexpected shape");
  }
  __half* up_h(const std::string& name, const std::vector<__half>& v) {
    __half* d = static_cast<__half*>(e.dmalloc(v.size() * sizeof(__half)));
    DM_CUDA(cudaMemcpy(d, v.data(), v.size() * sizeof(__half), cudaMemcpyHostToDevice));
    e.wh[name] = d;
    e.wh_n[name] = v.size();
    return d;
  }
  float* up_f(const std::string& name, const std::vector<float>& v) {
    float* d = static_cast<float*>(e.dmalloc(v.size() * sizeof(float)));
    DM_CUDA(cudaMemcpy(d, v.data(), v.size() * sizeof(float), cudaMemcpyHostToDevice));
    e.wf[name] = d;
    e.wf_n[name] = v.size();
    return d;
  }
  void vec(const std::string& k, int pad_to = 0) {
    const HostTensor& t = get(k);
    std::vector<float> v(std::max<int64_t>(t.numel(), pad_to), 0.f);
    for (int64_t i = 0; i < t.numel(); ++i) v[i] = __half2float(t.data[i]);
    up_f(k, v);
  }
  void norm(const std::string& k) {
    vec(k + ".weight");
    vec(k + ".bias");
  }
  // [O, I, 3, 3] -> [max(O, opad), 9*I], k = (r*3+s)*I + i
  void conv3(const std::string& k, int opad = 0) {
    const HostTensor& t = get(k + ".weight");
    DM_CHECK(t.shape.size() == 4 && t.shape[2] == 3 && t.shape[3] == 3, "'" + k + ".weight' is not a 3x3 conv");
    const int64_t O = t.shape[0], I = t.shape[1];
    const int64_t Op = std::max<int64_t>(O, opad);
    std::vector<__half> v(Op * 9 * I, __float2half_rn(0.f));
    for (int64_t o = 0; o < O; ++o)
      for (int64_t i = 0; i < I; ++i)
        for (int tap = 0; tap < 9; ++tap) v[(o * 9 + tap) * I + i] = t.data[(o * I + i) * 9 + tap];
    up_h(k + ".weight", v);
    vec(k + ".bias", static_cast<int>(Op));
  }
  // [O, Cin(3|4), 3, 3] -> [O, 64], k = tap*Cin + c (zero padded)
  void conv_in(const std::string& k) {
    const HostTensor& t = get(k + ".weight");
    const int64_t O = t.shape[0], I = t.shape[1];
    DM_CHECK(t.shape.size() == 4 && 9 * I <= 64, "'" + k + ".weight' is not a small-Cin 3x3 conv");
    std::vector<__half> v(O * 64, __float2half_rn(0.f));
    for (int64_t o = 0; o < O; ++o)
      for (int64_t i = 0; i < I; ++i)
        for (int tap = 0; tap < 9; ++tap) v[o * 64 + tap * I + i] = t.data[(o * I + i) * 9 + tap];
    up_h(k + ".weight", v);
    vec(k + ".bias");
  }
  // Linear [O, I] or 1x1 conv [O, I, 1, 1]
  void mat(const std::string& k, bool bias) {
    const HostTensor& t = get(k + ".weight");
    DM_CHECK(t.shape.size() == 2 || (t.shape.size() == 4 && t.shape[2] == 1 && t.shape[3] == 1),
             "'" + k + ".weight' is not a matrix");
    up_h(k + ".weight", t.data);
    if (bias) vec(k + ".bias");
  }
  void resnet(const std::string& k, int Cin, int Cout, bool temb) {
    norm(k + ".norm1");
    conv3(k + ".conv1");
    norm(k + ".norm2");
    conv3(k + ".conv2");
    if (Cin != Cout) mat(k + ".conv_shortcut", true);
    (void)temb;
  }
  void transformer(const std::string& k, int C) {
    norm(k + ".norm");
    mat(k + ".proj_in", true);
    mat(k + ".proj_out", true);
    const std::string t = k + ".transformer_blocks.0";
    norm(t + ".norm1"); norm(t + ".norm2"); norm(t + ".norm3");
    {  // fused q | k | v of the self-attention
      std::vector<__half> v;
      for (const char* n : {".attn1.to_q", ".attn1.to_k", ".attn1.to_v"}) {
        const HostTensor& w = get(t + n + ".weight");
        expect(w, t + n, {C, C});
        v.insert(v.end(), w.data.begin(), w.data.end());
      }
      up_h(t + ".attn1.qkv.weight", v);
    }
    mat(t + ".attn1.to_out.0", true);
    mat(t + ".attn2.to_q", false);
    {  // fused k | v projection of the text context (used by set_context)
      std::vector<__half> v;
      for (const char* n : {".attn2.to_k", ".attn2.to_v"}) {
        const HostTensor& w = get(t + n + ".weight");
        expect(w, t + n, {C, kCtxDim});
        v.insert(v.end(), w.data.begin(), w.data.end());
      }
      up_h(t + ".attn2.kv.weight", v);
    }
    mat(t + ".attn2.to_out.0", true);
    {  // GEGLU: interleave value / gate rows so each accumulator column pair is (value_j, gate_j)
      const HostTensor& w = get(t + ".ff.net.0.proj.weight");
      const HostTensor& b = get(t + ".ff.net.0.proj.bias");
      expect(w, t + ".ff.net.0.proj", {8 * C, C});
      const int64_t half = 4 * C;
      std::vector<__half> v(w.data.size());
      std::vector<float> bv(8 * C);
      for (int64_t j = 0; j < half; ++j) {
        std::memcpy(&v[(2 * j) * C], &w.data[j * C], C * sizeof(__half));
        std::memcpy(&v[(2 * j + 1) * C], &w.data[(half + j) * C], C * sizeof(__half));
        bv[2 * j] = __half2float(b.data[j]);
        bv[2 * j + 1] = __half2float(b.data[half + j]);
      }
      up_h(t + ".ff.net.0.proj.weight", v);
      up_f(t + ".ff.net.0.proj.bias", bv);
    }
    mat(t + ".ff.net.2", true);
    // context K/V cache for this cross-attention layer
    alloc_kv_cache(e, t + ".attn2", C);
  }
};

}  // namespace

void Engine::finalize() {
  DM_CHECK(!finalized, "finalize_weights called twice");
  DM_CUDA(cudaSetDevice(device));
  Packer pk{*this};
  has_unet = staging.count("unet.conv_in.weight") != 0;
  has_vae = staging.count("vae.encoder.conv_in.weight") != 0;
  DM_CHECK(has_unet || has_vae, "no 'unet.*' or 'vae.*' tensors were loaded");
  static const int ch[4] = {320, 640, 1280, 1280};
  if (has_unet) {
    const std::string U = "unet.";
    std::vector<std::pair<std::string, int>> resnets;  // (key, Cout) in execution order, for the stacked temb proj
    pk.conv_in(U + "conv_in");
    pk.mat(U + "time_embedding.linear_1", true);
    pk.mat(U + "time_embedding.linear_2", true);
    int cin = 320;
    for (int i = 0; i < 4; ++i) {
      for (int j = 0; j < 2; ++j) {
        const std::string rk = U + "down_blocks." + std::to_string(i) + ".resnets." + std::to_string(j);
        pk.resnet(rk, j == 0 ? cin : ch[i], ch[i], true);
        resnets.push_back({rk, ch[i]});
        if (i < 3) pk.transformer(U + "down_blocks." + std::to_string(i) + ".attentions." + std::to_string(j), ch[i]);
      }
      if (i < 3) pk.conv3(U + "down_blocks." + std::to_string(i) + ".downsamplers.0.conv");
      cin = ch[i];
    }
    pk.resnet(U + "mid_block.resnets.0", 1280, 1280, true);
    resnets.push_back({U + "mid_block.resnets.0", 1280});
    pk.transformer(U + "mid_block.attentions.0", 1280);
    pk.resnet(U + "mid_block.resnets.1", 1280, 1280, true);
    resnets.push_back({U + "mid_block.resnets.1", 1280});
    static const int rev[4] = {1280, 1280, 640, 320};
    int out_c = rev[0];
    for (int i = 0; i < 4; ++i) {
      const int prev = out_c;
      out_c = rev[i];
      const int in_c = rev[std::min(i + 1, 3)];
      for (int j = 0; j < 3; ++j) {
        const int skip = j == 2 ? in_c : out_c;
        const int rin = j == 0 ? prev : out_c;
        const std::string rk = U + "up_blocks." + std::to_string(i) + ".resnets." + std::to_string(j);
        pk.resnet(rk, rin + skip, out_c, true);
        resnets.push_back({rk, out_c});
        if (i > 0) pk.transformer(U + "up_blocks." + std::to_string(i) + ".attentions." + std::to_string(j), out_c);
      }
      if (i < 3) pk.conv3(U + "up_blocks." + std::to_string(i) + ".upsamplers.0.conv");
    }
    pk.norm(U + "conv_norm_out");
    pk.conv3(U + "conv_out", 16);
    // all 22 time_emb_proj Linears stacked into one [sum Cout, 1280] GEMM
    tproj_total = 0;
    for (auto& r : resnets) { tproj_off[r.first] = tproj_total; tproj_total += r.second; }
    std::vector<__half> w(static_cast<size_t>(tproj_total) * kTimeDim);
    std::vector<float> b(tproj_total);
    for (auto& r : resnets) {
      const HostTensor& tw = pk.get(r.first + ".time_emb_proj.weight");
      const HostTensor& tb = pk.get(r.first + ".time_emb_proj.bias");
      pk.expect(tw, r.first + ".time_emb_proj", {r.second, kTimeDim});
      std::memcpy(&w[static_cast<size_t>(tproj_off[r.first]) * kTimeDim], tw.data.data(), tw.data.size() * sizeof(__half));
      for (int i = 0; i < r.second; ++i) b[tproj_off[r.first] + i] = __half2float(tb.data[i]);
    }
    pk.up_h(U + "time_emb_proj_all.weight", w);
    pk.up_f(U + "time_emb_proj_all.bias", b);
  }
  if (has_vae) {
    const std::string V = "vae.encoder.";
    static const int vch[4] = {128, 256, 512, 512};
    pk.conv_in(V + "conv_in");
    int cin = 128;
    for (int i = 0; i < 4; ++i) {
      for (int j = 0; j < 2; ++j)
        pk.resnet(V + "down_blocks." + std::to_string(i) + ".resnets." + std::to_string(j), j == 0 ? cin : vch[i], vch[i], false);
      if (i < 3) pk.conv3(V + "down_blocks." + std::to_string(i) + ".downsamplers.0.conv");
      cin = vch[i];
    }
    pk.resnet(V + "mid_block.resnets.0", 512, 512, false);
    pk.resnet(V + "mid_block.resnets.1", 512, 512, false);
    const std::string a = V + "mid_block.attentions.0";
    pk.norm(a + ".group_norm");
    {  // fused q | k | v projection (weights [3*512, 512], biases [3*512])
      std::vector<__half> v;
      std::vector<float> b;
      for (const char* n : {".to_q", ".to_k", ".to_v"}) {
        const HostTensor& w = pk.get(a + n + ".weight");
        const HostTensor& bb = pk.get(a + n + ".bias");
        pk.expect(w, a + n, {512, 512});
        v.insert(v.end(), w.data.begin(), w.data.end());
        for (auto h : bb.data) b.push_back(__half2float(h));
      }
      pk.up_h(a + ".qkv.weight", v);
      pk.up_f(a + ".qkv.bias", b);
    }
    pk.mat(a + ".to_out.0", true);
    pk.norm(V + "conv_norm_out");
    pk.conv3(V + "conv_out", 16);
    {  // quant_conv 1x1 8->8 stays fp16 [8,8] + fp32 bias, applied inside vae_sample_kernel
      const HostTensor& w = pk.get("vae.quant_conv.weight");
      DM_CHECK(w.numel() == 64, "'vae.quant_conv.weight' must be [8,8,1,1]");
      pk.up_h("vae.quant_conv.weight", w.data);
      pk.vec("vae.quant_conv.bias");
    }
  }
  staging.clear();
  finish_setup();
}

void Engine::finish_setup() {
  gn_partial_floats = 64ull * 64 * 8192;  // 64 floats x <=64 splits x <=8192 images
  gn_partial = static_cast<float*>(dmalloc(gn_partial_floats * sizeof(float)));
  gn_ab_floats = 8ull << 20;  // 2 floats x (images x channels) <= 4 Mi entries
  gn_ab = static_cast<float*>(dmalloc(gn_ab_floats * sizeof(float)));
  gn_ticket_count = 8192;
  gn_tickets = static_cast<unsigned*>(dmalloc(gn_ticket_count * sizeof(unsigned)));
  DM_CUDA(cudaMemset(gn_tickets, 0, gn_ticket_count * sizeof(unsigned)));
  if (!sched_a) {
    // default SD-1.5 schedule (scaled_linear 0.00085 -> 0.012, 1000 steps) until dm_set_schedule overrides it
    std::vector<float> a(1000), b(1000);
    const float s0 = std::sqrt(0.00085f), s1 = std::sqrt(0.012f);
    float acpf = 1.f;
    for (int i = 0; i < 1000; ++i) {
      const float lin = s0 + (s1 - s0) * (static_cast<float>(i) / 999.f);
      const float beta = lin * lin;
      acpf *= (1.f - beta);
      a[i] = std::sqrt(acpf);
      b[i] = std::sqrt(1.f - acpf);
    }
    set_schedule(a.data(), b.data(), 1000);
  }
  DM_CUDA(cudaStreamCreateWithFlags(&cap_stream, cudaStreamNonBlocking));
  DM_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&err_host), sizeof(int), cudaHostAllocMapped));
  *err_host = 0;
  DM_CUDA(cudaHostGetDevicePointer(reinterpret_cast<void**>(&err_dev), err_host, 0));
  DM_CUDA(cudaDeviceSynchronize());
  finalized = true;
}

void Engine::check_async_error() {
  if (err_host && *reinterpret_cast<volatile int*>(err_host) != 0) {
    *err_host = 0;
    DM_CHECK(false, "an earlier call on this engine used timesteps outside [0, " + std::to_string(sched_n) +
                        ") (they were clamped on the device; its results are invalid)");
  }
}

void Engine::set_schedule(const float* a, const float* b, int n) {
  DM_CHECK(n > 0 && a && b, "set_schedule: bad arguments");
  if (!sched_a || n != sched_n) {
    sched_a = static_cast<float*>(dmalloc(n * sizeof(float)));
    sched_b = static_cast<float*>(dmalloc(n * sizeof(float)));
    sched_n = n;
  }
  DM_CUDA(cudaMemcpy(sched_a, a, n * sizeof(float), cudaMemcpyHostToDevice));
  DM_CUDA(cudaMemcpy(sched_b, b, n * sizeof(float), cudaMemcpyHostToDevice));
}

// ------------------------------------------------------------------ packed-weight cache
// One file holding the engine's packed device buffers exactly as finalize() leaves them (repacked conv weights, fused
// q|k|v, interleaved GEGLU rows, stacked time_emb_proj, fp32 bias / norm vectors), so a later process skips the diffusers
// state-dict intake and the host-side repacking.  Layout: "DMPK0002", flags, then named records.
namespace {
constexpr char kPackMagic[8] = {'D', 'M', 'P', 'K', '0', '0', '0', '2'};
struct FileW {
  FILE* f;
  void raw(const void* p, size_t n) { DM_CHECK(fwrite(p, 1, n, f) == n, "packed-weight cache: short write"); }
  template <class T> void pod(const T& v) { raw(&v, sizeof(T)); }
  void str(const std::string& s) { pod<uint32_t>(static_cast<uint32_t>(s.size())); raw(s.data(), s.size()); }
};
struct FileR {
  FILE* f;
  void raw(void* p, size_t n) { DM_CHECK(fread(p, 1, n, f) == n, "packed-weight cache: truncated file"); }
  template <class T> T pod() { T v; raw(&v, sizeof(T)); return v; }
  std::string str() {
    const uint32_t n = pod<uint32_t>();
    DM_CHECK(n < 4096, "packed-weight cache: corrupt name");
    std::string s(n, '\0');
    raw(&s[0], n);
    return s;
  }
};
}  // namespace

void Engine::save_packed(const std::string& path) const {
  DM_CHECK(finalized, "save_packed needs finalized weights");
  DM_CUDA(cudaSetDevice(device));
  FILE* f = fopen(path.c_str(), "wb");
  DM_CHECK(f != nullptr, "cannot open '" + path + "' for writing");
  try {
    FileW w{f};
    w.raw(kPackMagic, 8);
    w.pod<uint8_t>(has_unet ? 1 : 0);
    w.pod<uint8_t>(has_vae ? 1 : 0);
    w.pod<int32_t>(tproj_total);
    w.pod<uint32_t>(static_cast<uint32_t>(tproj_off.size()));
    for (auto& kv : tproj_off) { w.str(kv.first); w.pod<int32_t>(kv.second); }
    w.pod<uint32_t>(static_cast<uint32_t>(xattn_layers.size()));
    for (auto& a2 : xattn_layers) { w.str(a2); w.pod<int32_t>(kv_C.at(a2)); }
    std::vector<char> host;
    w.pod<uint32_t>(static_cast<uint32_t>(wh.size()));
    for (auto& kv : wh) {
      const size_t n = wh_n.at(kv.first);
      host.resize(n * sizeof(__half));
      DM_CUDA(cudaMemcpy(host.data(), kv.second, host.size(), cudaMemcpyDeviceToHost));
      w.str(kv.first); w.pod<uint64_t>(n); w.raw(host.data(), host.size());
    }
    w.pod<uint32_t>(static_cast<uint32_t>(wf.size()));
    for (auto& kv : wf) {
      const size_t n = wf_n.at(kv.first);
      host.resize(n * sizeof(float));
      DM_CUDA(cudaMemcpy(host.data(), kv.second, host.size(), cudaMemcpyDeviceToHost));
      w.str(kv.first); w.pod<uint64_t>(n); w.raw(host.data(), host.size());
    }
    w.raw(kPackMagic, 8);  // trailer: a truncated file never validates
  } catch (...) {
    fclose(f);
    remove(path.c_str());
    throw;
  }
  fclose(f);
}

void Engine::load_packed(const std::string& path) {
  DM_CHECK(!finalized && staging.empty(), "load_packed needs a fresh engine");
  DM_CUDA(cudaSetDevice(device));
  FILE* f = fopen(path.c_str(), "rb");
  DM_CHECK(f != nullptr, "cannot open '" + path + "'");
  try {
    FileR r{f};
    char magic[8];
    r.raw(magic, 8);
    DM_CHECK(std::memcmp(magic, kPackMagic, 8) == 0, "'" + path + "' is not a packed-weight cache of this engine version");
    has_unet = r.pod<uint8_t>() != 0;
    has_vae = r.pod<uint8_t>() != 0;
    tproj_total = r.pod<int32_t>();
    for (uint32_t i = 0, n = r.pod<uint32_t>(); i < n; ++i) { std::string k = r.str(); tproj_off[k] = r.pod<int32_t>(); }
    std::vector<std::pair<std::string, int>> xl;
    for (uint32_t i = 0, n = r.pod<uint32_t>(); i < n; ++i) { std::string k = r.str(); xl.push_back({k, r.pod<int32_t>()}); }
    std::vector<char> host;
    for (uint32_t i = 0, n = r.pod<uint32_t>(); i < n; ++i) {
      std::string k = r.str();
      const uint64_t ne = r.pod<uint64_t>();
      DM_CHECK(ne < (1ull << 32), "packed-weight cache: corrupt record");
      host.resize(ne * sizeof(__half));
      r.raw(host.data(), host.size());
      __half* d = static_cast<__half*>(dmalloc(host.size()));
      DM_CUDA(cudaMemcpy(d, host.data(), host.size(), cudaMemcpyHostToDevice));
      wh[k] = d; wh_n[k] = ne;
    }
    for (uint32_t i = 0, n = r.pod<uint32_t>(); i < n; ++i) {
      std::string k = r.str();
      const uint64_t ne = r.pod<uint64_t>();
      DM_CHECK(ne < (1ull << 32), "packed-weight cache: corrupt record");
      host.resize(ne * sizeof(float));
      r.raw(host.data(), host.size());
      float* d = static_cast<float*>(dmalloc(host.size()));
      DM_CUDA(cudaMemcpy(d, host.data(), host.size(), cudaMemcpyHostToDevice));
      wf[k] = d; wf_n[k] = ne;
    }
    r.raw(magic, 8);
    DM_CHECK(std::memcmp(magic, kPackMagic, 8) == 0, "'" + path + "' is truncated");
    for (auto& x : xl) alloc_kv_cache(*this, x.first, x.second);
  } catch (...) {
    fclose(f);
    throw;
  }
  fclose(f);
  finish_setup();
}

void Engine::set_context(int slot, const float* ctx, cudaStream_t s) {
  DM_CHECK(finalized && has_unet, "set_context needs a finalized U-Net");
  DM_CHECK(slot >= 0 && slot < kMaxCtxSlots, "context slot out of range (0.." + std::to_string(kMaxCtxSlots - 1) + ")");
  std::vector<__half> h(kCtxTokens * kCtxDim);
  for (size_t i = 0; i < h.size(); ++i) h[i] = __float2half_rn(ctx[i]);
  __half* tmp = nullptr;
  DM_CUDA(cudaMalloc(&tmp, h.size() * sizeof(__half)));
  DM_CUDA(cudaMemcpyAsync(tmp, h.data(), h.size() * sizeof(__half), cudaMemcpyHostToDevice, s));
  for (const std::string& a2 : xattn_layers) {
    const int C = kv_C.at(a2);
    IgemmDesc d;
    d.Nimg = 1; d.H = 1; d.W = kCtxTokens;
    d.nsrc = 1;
    d.src[0] = ActView{tmp, 1, 1, kCtxTokens, kCtxDim, kCtxDim};
    seg_1x1(d, kCtxDim, 0);
    d.Wt = H(a2 + ".kv.weight");
    d.N = 2 * C; d.K = kCtxDim;
    d.out = kv_cache.at(a2) + static_cast<size_t>(slot) * kCtxTokens * 2 * C;
    d.ld_out = 2 * C;
    IgemmOp op = igemm_prepare(d, num_sms);
    igemm_launch(op, s);
    launch_count += 1;
    flop_count += op.flops;
  }
  DM_CUDA(cudaStreamSynchronize(s));
  DM_CUDA(cudaFree(tmp));
}

// ------------------------------------------------------------------ plans
Plan* Engine::get_plan(PlanKey key) {
  DM_CHECK(finalized, "weights are not finalized");
  auto it = plans.find(key);
  if (it != plans.end()) return it->second.get();
  DM_CHECK(key.B > 0 && key.h > 0 && key.w > 0, "empty batch or image");
  if (key.kind == kPlanVae) DM_CHECK(has_vae, "VAE encoder weights were not loaded");
  else DM_CHECK(has_unet, "U-Net weights were not loaded");
  if (plans.size() >= 12) {  // bound device memory held by cached arenas
    for (auto pit = plans.begin(); pit != plans.end();) {
      if (pit->second.get() != last_unet_plan) pit = plans.erase(pit);
      else ++pit;
    }
  }
  auto build = [&](Plan& p, bool dry, ArenaPlanner& ar) {
    if (key.kind == kPlanVae) build_vae_plan(*this, p, dry, ar);
    else build_unet_plan(*this, p, dry, ar);
  };
  size_t need = 0;
  {
    Plan dryp;
    dryp.key = key;
    ArenaPlanner ar;
    build(dryp, true, ar);
    need = ar.high_water();
  }
  std::unique_ptr<Plan> p(new Plan());
  p->key = key;
  p->arena_bytes = need + 4096;
  DM_CUDA(cudaMalloc(reinterpret_cast<void**>(&p->arena), p->arena_bytes));
  DM_CUDA(cudaMemset(p->arena, 0, p->arena_bytes));
  ArenaPlanner ar;
  build(*p, false, ar);
  DM_CHECK(ar.high_water() <= need, "plan arena grew between dry and real build");
  Plan* raw = p.get();
  plans[key] = std::move(p);
  return raw;
}

void Engine::run_plan(Plan* p, cudaStream_t s) {
  const bool want_graph = use_graph && !debug_keep && !p->graph_failed;
  if (want_graph && p->graph) {
    DM_CUDA(cudaGraphLaunch(p->graph, s));
  } else if (want_graph && p->launches > 0 && p->eager_runs >= 1) {
    // second run of this plan: capture it (the first, eager run configured every kernel's attributes)
    cudaGraph_t g = nullptr;
    cudaError_t st = cudaStreamBeginCapture(cap_stream, cudaStreamCaptureModeThreadLocal);
    if (st == cudaSuccess) {
      try {
        for (auto& step : p->steps) step.run(cap_stream);
      } catch (...) {
        cudaStreamEndCapture(cap_stream, &g);
        if (g) cudaGraphDestroy(g);
        p->graph_failed = true;
        throw;
      }
      st = cudaStreamEndCapture(cap_stream, &g);
    }
    if (st == cudaSuccess && g) st = cudaGraphInstantiate(&p->graph, g, 0);
    if (g) cudaGraphDestroy(g);
    if (st != cudaSuccess || !p->graph) {
      cudaGetLastError();
      p->graph = nullptr;
      p->graph_failed = true;
      for (auto& step : p->steps) step.run(s);
    } else {
      DM_CUDA(cudaGraphLaunch(p->graph, s));
    }
  } else {
    for (auto& step : p->steps) step.run(s);
    p->eager_runs += 1;
  }
  launch_count += p->launches;
  flop_count += p->flops_igemm + p->flops_attn;
}

}  // namespace dm
