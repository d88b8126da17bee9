// Shared helper for the plan builders: allocates activations in the plan arena and appends prepared steps.
#pragma once
#include "engine.h"

namespace dm {

struct Builder {
  Engine& e;
  Plan& p;
  bool dry;
  ArenaPlanner& ar;
  Builder(Engine& e_, Plan& p_, bool dry_, ArenaPlanner& ar_) : e(e_), p(p_), dry(dry_), ar(ar_) {}

  template <class T>
  T* at(size_t off) const { return reinterpret_cast<T*>(p.arena + off); }
  __half* hp(const Act& a) const { return at<__half>(a.off); }

  Act alloc(int N, int H, int W, int C) {
    Act a;
    a.N = N; a.H = H; a.W = W; a.C = C; a.valid = true;
    a.off = ar.alloc((a.bytes() + 1023) & ~size_t(1023));
    return a;
  }
  size_t alloc_bytes(size_t bytes) { return ar.alloc((bytes + 1023) & ~size_t(1023)); }
  void release(Act& a) {
    if (!a.valid) return;
    if (!e.debug_keep) {
      ar.free(a.off);
      if (a.stats) ar.free(a.st_off);
    }
    a.valid = false;
    a.stats = false;
  }
  // ---- GroupNorm statistics in the producer's epilogue.  kind: 1 = 3x3 conv, 2 = 1x1 conv / Linear (variant gn_epilogue)
  bool gn_fusable(int kind, int H, int W, int C) const {
    return (gn_epilogue_mode() & kind) != 0 && igemm_gn_fusable(H, W, C);
  }
  // reserve the statistics record of output `o` and point the conv at it
  void attach_stats(Act& o, IgemmDesc& d, bool flattened) {
    o.st_off = alloc_bytes(gn_record_floats(o.N, o.H * o.W, o.C) * sizeof(float));
    o.stats = true;
    d.gn.rec = at<float>(o.st_off);
    d.gn.tiles_img = flattened ? o.H * o.W / 128 : 0;
  }
  void tap(const std::string& name, const Act& a) {
    if (e.debug_keep && !dry) p.taps[name] = a;
  }
  ActView view(const Act& a) const { return ActView{hp(a), a.N, a.H, a.W, a.C, a.C}; }

  void push(Step s) {
    if (dry) return;
    if (s.cls == kStepIgemm) p.flops_igemm += s.flops;
    if (s.cls == kStepAttn) p.flops_attn += s.flops;
    p.launches += s.launches;
    p.steps.push_back(std::move(s));
  }
  void add_igemm(const std::string& name, const IgemmDesc& d) {
    if (dry) return;
    IgemmOp op = igemm_prepare(d, e.num_sms);
    push(Step{[op](cudaStream_t s) { igemm_launch(op, s); }, kStepIgemm, op.flops, 1, name});
  }
  void add_attn(const std::string& name, const AttnDesc& d) {
    if (dry) return;
    AttnOp op = attn_prepare(d);
    push(Step{[op](cudaStream_t s) { attn_launch(op, s); }, kStepAttn, op.flops, 1, name});
  }

  // ---- GroupNorm(+SiLU) over one or two sources -> dense normalised tensor.  When every source carries the statistics
  // record its producer's epilogue formed, one cluster kernel folds the records and applies y = x * a + b (1 read + 1
  // write); otherwise the stand-alone kernels read the sources twice.
  Act groupnorm(const std::string& name, const Act& x0, const Act* x1, const std::string& wkey, float eps, bool silu) {
    const int C = x0.C + (x1 ? x1->C : 0);
    Act o = alloc(x0.N, x0.H, x0.W, C);
    const bool from_stats = x0.stats && (!x1 || x1->stats) && gn_fold_apply_supported(x0.H * x0.W, x0.C, x1 ? x1->C : 0);
    if (!dry) {
      GnDesc d;
      d.src0 = hp(x0); d.C0 = x0.C; d.ps0 = x0.C;
      if (x1) { d.src1 = hp(*x1); d.C1 = x1->C; d.ps1 = x1->C; }
      d.Nimg = x0.N; d.HW = x0.H * x0.W;
      d.gamma = e.F(wkey + ".weight"); d.beta = e.F(wkey + ".bias");
      d.eps = eps; d.silu = silu ? 1 : 0;
      d.out = hp(o);
      if (from_stats) {
        const float* r0 = at<float>(x0.st_off);
        const float* r1 = x1 ? at<float>(x1->st_off) : nullptr;
        push(Step{[d, r0, r1](cudaStream_t s) { gn_fold_apply_launch(d, r0, r1, s); }, kStepOther, 0, 1, name + ".apply"});
      } else {
        d.partial = e.gn_partial; d.ab = e.gn_ab; d.tickets = e.gn_tickets;
        DM_CHECK(static_cast<size_t>(2) * C * x0.N * gn_splits(x0.N, d.HW) <= e.gn_partial_floats &&
                     static_cast<size_t>(2) * x0.N * C <= e.gn_ab_floats && static_cast<size_t>(x0.N) <= e.gn_ticket_count,
                 "GroupNorm scratch too small");
        push(Step{[d](cudaStream_t s) { gn_launch(d, s); }, kStepOther, 0, gn_launch_count(d.HW, C), name});
      }
    }
    return o;
  }

  // ---- 3x3 stride-1 conv over (x0 [+ x1]); epilogue options.  gn_stats: the output will be read by a GroupNorm, so the
  // epilogue also forms its statistics record where the shape allows (Act::stats)
  Act conv3x3(const std::string& name, const Act& x0, const Act* x1, const std::string& wkey, int Cout,
              const __half* rowbias, int ld_rowbias, const Act* residual, bool gn_stats = false) {
    Act o = alloc(x0.N, x0.H, x0.W, Cout);
    IgemmDesc d;
    d.Nimg = x0.N; d.H = x0.H; d.W = x0.W;
    d.nsrc = x1 ? 2 : 1;
    d.src[0] = view(x0);
    if (x1) d.src[1] = view(*x1);
    const int Cin = x0.C + (x1 ? x1->C : 0);
    seg_conv3x3(d, Cin, x0.C);
    d.Wt = dry ? nullptr : e.H(wkey + ".weight");
    d.N = Cout; d.K = 9 * Cin;
    d.bias = dry ? nullptr : e.F(wkey + ".bias");
    d.rowbias = rowbias; d.ld_rowbias = ld_rowbias;
    if (residual) { d.residual = hp(*residual); d.ld_res = residual->C; }
    d.out = hp(o); d.ld_out = Cout;
    if (gn_stats && gn_fusable(1, o.H, o.W, Cout)) attach_stats(o, d, false);
    add_igemm(name, d);
    return o;
  }

  // ---- 1x1 conv / Linear over (x0 [+ x1]) viewed as [M, C]
  Act linear(const std::string& name, const Act& x0, const Act* x1, const std::string& wkey, int Nout, bool has_bias,
             const Act* residual, bool geglu = false, bool silu = false, bool gn_stats = false) {
    const int Cout = geglu ? Nout / 2 : Nout;
    Act o = alloc(x0.N, x0.H, x0.W, Cout);
    IgemmDesc d;
    // flatten pixels into one long row so M-tiles never straddle anything
    const long long M = x0.pixels();
    d.Nimg = 1; d.H = 1; d.W = static_cast<int>(M);
    d.nsrc = x1 ? 2 : 1;
    d.src[0] = ActView{hp(x0), 1, 1, static_cast<int>(M), x0.C, x0.C};
    if (x1) d.src[1] = ActView{hp(*x1), 1, 1, static_cast<int>(M), x1->C, x1->C};
    seg_1x1(d, x0.C, x1 ? x1->C : 0);
    d.Wt = dry ? nullptr : e.H(wkey + ".weight");
    d.N = Nout; d.K = x0.C + (x1 ? x1->C : 0);
    d.bias = (dry || !has_bias) ? nullptr : e.F(wkey + ".bias");
    if (residual) { d.residual = hp(*residual); d.ld_res = residual->C; }
    d.out = hp(o); d.ld_out = Cout;
    d.geglu = geglu ? 1 : 0; d.act_silu = silu ? 1 : 0;
    // gn_stats: the output will be read by a GroupNorm -- form its statistics record in the epilogue where the shape allows
    if (gn_stats && !geglu && gn_fusable(2, o.H, o.W, Cout)) attach_stats(o, d, true);
    add_igemm(name, d);
    return o;
  }

  Act layernorm(const std::string& name, const Act& x, const std::string& wkey) {
    Act o = alloc(x.N, x.H, x.W, x.C);
    if (!dry) {
      const __half* in = hp(x);
      __half* out = hp(o);
      const float *g = e.F(wkey + ".weight"), *b = e.F(wkey + ".bias");
      const long long rows = x.pixels();
      const int C = x.C;
      push(Step{[=](cudaStream_t s) { layernorm_launch(in, C, g, b, 1e-5f, rows, C, out, C, s); }, kStepOther, 0, 1, name});
    }
    return o;
  }
};

}  // namespace dm
