// Engine: packed SD-1.5 weights, context K/V caches, per-shape execution plans (a static list of prepared
// kernel launches over a private arena), replayed eagerly or as a CUDA graph.
#pragma once
#include <functional>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "ops.h"

namespace dm {

constexpr int kMaxCtxSlots = 64;
constexpr int kCtxTokens = 77;
constexpr int kCtxDim = 768;
constexpr int kTimeDim = 1280;

struct HostTensor {
  std::vector<__half> data;
  std::vector<int64_t> shape;
  int64_t numel() const {
    int64_t n = 1;
    for (auto s : shape) n *= s;
    return n;
  }
};

// first-fit offset allocator with coalescing free list (plan-build time only)
class ArenaPlanner {
 public:
  size_t alloc(size_t bytes);
  void free(size_t off);
  size_t high_water() const { return high_; }

 private:
  std::map<size_t, size_t> free_;   // offset -> size
  std::map<size_t, size_t> live_;   // offset -> size
  size_t top_ = 0, high_ = 0;
};

struct Act {  // NHWC fp16 activation inside a plan's arena
  size_t off = 0;
  int N = 0, H = 0, W = 0, C = 0;
  bool valid = false;
  // GroupNorm statistics record left by the producing conv's epilogue (igemm.cuh: IgGn), in the arena next to the data
  size_t st_off = 0;
  bool stats = false;
  long long pixels() const { return static_cast<long long>(N) * H * W; }
  size_t bytes() const { return static_cast<size_t>(pixels()) * C * sizeof(__half); }
};

enum StepClass { kStepIgemm = 0, kStepAttn = 1, kStepOther = 2 };
struct Step {
  std::function<void(cudaStream_t)> run;
  StepClass cls;
  double flops;
  int launches;
  std::string name;
};

enum PlanKind { kPlanUnet = 0, kPlanDift = 1, kPlanVae = 2 };
struct PlanKey {
  int kind, B, h, w, aux;
  bool operator<(const PlanKey& o) const {
    return std::tie(kind, B, h, w, aux) < std::tie(o.kind, o.B, o.h, o.w, o.aux);
  }
};

struct Plan {
  PlanKey key{};
  char* arena = nullptr;
  size_t arena_bytes = 0;
  std::vector<Step> steps;
  // boundary buffers (inside the arena)
  __half* a_in = nullptr;      // [B*h*w, 64] patch matrix of the input conv
  __half* temb_sin = nullptr;  // [B, 320]
  int* ctx_idx = nullptr;      // [B] context slot per row
  IgLossArgs* loss_args = nullptr;  // U-Net plans: conv_out's fused typicality epilogue (rewritten before every replay)
  __half* out = nullptr;       // U-Net: pred [B*h*w, 16]; DIFT: feature map NHWC; VAE: conv_out [B*h*w, 16]
  int out_H = 0, out_W = 0, out_C = 0;
  std::map<std::string, Act> taps;  // debug_keep only
  cudaGraphExec_t graph = nullptr;
  int eager_runs = 0;
  bool graph_failed = false;
  double flops_igemm = 0, flops_attn = 0;
  int launches = 0;
  ~Plan();
};

struct Engine {
  int device = 0, num_sms = 148;
  bool finalized = false;
  bool debug_keep = false;
  bool use_graph = true;
  std::map<std::string, HostTensor> staging;
  // packed device parameters
  std::map<std::string, __half*> wh;  // fp16 matrices [N, K]
  std::map<std::string, float*> wf;   // fp32 vectors
  std::map<std::string, size_t> wh_n, wf_n;  // element counts of the packed buffers (packed-weight cache)
  std::map<std::string, int> tproj_off;  // resnet name -> column offset in the stacked time_emb_proj output
  int tproj_total = 0;
  std::vector<void*> owned;           // every cudaMalloc to free
  float *sched_a = nullptr, *sched_b = nullptr;
  int sched_n = 0;
  // context K/V caches: per cross-attention layer, [kMaxCtxSlots, 77, 2C] fp16 (K | V)
  std::vector<std::string> xattn_layers;
  std::map<std::string, __half*> kv_cache;
  std::map<std::string, int> kv_C;
  std::map<PlanKey, std::unique_ptr<Plan>> plans;
  Plan* last_unet_plan = nullptr;
  float* gn_partial = nullptr;  // shared GroupNorm scratch
  size_t gn_partial_floats = 0;
  float* gn_ab = nullptr;          // GroupNorm per-(image, channel) scale/shift scratch
  size_t gn_ab_floats = 0;
  unsigned* gn_tickets = nullptr;  // GroupNorm last-block tickets (always left zero)
  size_t gn_ticket_count = 0;
  cudaStream_t cap_stream = nullptr;
  int64_t launch_count = 0;
  double flop_count = 0;
  bool has_unet = false, has_vae = false;
  // sticky device-side error flag (zero-copy pinned host word): set by kernels that had to clamp an out-of-range
  // timestep; reported by the next ABI call on this engine
  int* err_host = nullptr;
  int* err_dev = nullptr;
  void check_async_error();

  ~Engine();
  void* dmalloc(size_t bytes);
  const __half* H(const std::string& k) const;
  const float* F(const std::string& k) const;
  bool hasF(const std::string& k) const { return wf.count(k) != 0; }
  bool hasH(const std::string& k) const { return wh.count(k) != 0; }

  void load_tensor(const std::string& key, const void* host, int dtype, int ndim, const int64_t* shape);
  void finalize();
  void finish_setup();                       // scratch, default schedule, capture stream (after packing or load_packed)
  void save_packed(const std::string& path) const;  // packed device buffers -> one file
  void load_packed(const std::string& path);        // replaces load_tensor... + finalize
  void set_schedule(const float* a, const float* b, int n);
  void set_context(int slot, const float* ctx, cudaStream_t s);

  Plan* get_plan(PlanKey key);
  void run_plan(Plan* p, cudaStream_t s);
};

// plan builders (plan_unet.cu / plan_vae.cu)
void build_unet_plan(Engine& e, Plan& p, bool dry, ArenaPlanner& ar);
void build_vae_plan(Engine& e, Plan& p, bool dry, ArenaPlanner& ar);

}  // namespace dm
