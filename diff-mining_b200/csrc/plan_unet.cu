// Execution plan of the SD-1.5 U-Net noise-prediction forward (and its DIFT early exit).
// Topology follows the in-repo copy of UNet2DConditionModel.forward at
// /root/reference/diffmining/typicality/dift.py:84-165 and SURVEY.md appendix B; every op below is one of the
// hand-written kernels (igemm / attention / norm / misc).
#include "plan_builder.h"

namespace dm {

namespace {

struct UnetBuilder : Builder {
  int Bf = 0;
  __half* tproj = nullptr;  // [Bf, tproj_total] stacked time_emb_proj outputs
  int tproj_row_mult = 1;   // > 1 while building the shared prefix: row u of the activations <-> row u*G of tproj
  using Builder::Builder;

  // ResnetBlock2D (reference op order: applications/parallel-dataset/pnp.py:282-359).  Every conv whose output is read by
  // a GroupNorm (conv1 -> norm2; conv2 -> the next block's norm / norm1, directly or as a skip connection) also forms that
  // GroupNorm's statistics in its epilogue, so the normalisation is a single fold + apply pass.
  Act resnet(const std::string& key, const Act& x0, const Act* x1, int Cout) {
    const int Cin = x0.C + (x1 ? x1->C : 0);
    Act n1 = groupnorm(key + ".norm1", x0, x1, key + ".norm1", 1e-5f, true);
    const __half* rb = dry ? nullptr : tproj + e.tproj_off.at(key);
    Act h1 = conv3x3(key + ".conv1", n1, nullptr, key + ".conv1", Cout, rb, e.tproj_total * tproj_row_mult, nullptr, true);
    release(n1);
    Act n2 = groupnorm(key + ".norm2", h1, nullptr, key + ".norm2", 1e-5f, true);
    release(h1);
    Act sc;
    const Act* res = &x0;
    if (Cin != Cout || x1) {
      sc = linear(key + ".conv_shortcut", x0, x1, key + ".conv_shortcut", Cout, true, nullptr);
      res = &sc;
    }
    Act out = conv3x3(key + ".conv2", n2, nullptr, key + ".conv2", Cout, nullptr, 0, res, true);
    release(n2);
    release(sc);
    tap(key, out);
    return out;
  }

  // Transformer2DModel with one BasicTransformerBlock (self-attn, cross-attn, GEGLU FF), in two halves: everything up
  // to and including the self-attention residual depends on (x_t, t) only; the context enters in the tail.
  Act transformer_head(const std::string& key, const Act& x) {
    const int C = x.C, D = C / 8, T = x.H * x.W;
    const std::string t = key + ".transformer_blocks.0";
    Act n = groupnorm(key + ".norm", x, nullptr, key + ".norm", 1e-6f, false);
    Act h = linear(key + ".proj_in", n, nullptr, key + ".proj_in", C, true, nullptr);
    release(n);
    // --- self attention
    Act ln1 = layernorm(t + ".norm1", h, t + ".norm1");
    Act qkv = linear(t + ".attn1.qkv", ln1, nullptr, t + ".attn1.qkv", 3 * C, false, nullptr);
    release(ln1);
    Act ao = alloc(x.N, x.H, x.W, C);
    {
      AttnDesc d;
      d.B = x.N; d.heads = 8; d.D = D; d.Tq = T; d.Tk = T;
      d.q = hp(qkv); d.k = hp(qkv) + C; d.v = hp(qkv) + 2 * C;
      d.ld_q = d.ld_k = d.ld_v = 3 * C;
      d.bs_q = d.bs_k = d.bs_v = static_cast<long long>(T) * 3 * C;
      d.out = hp(ao); d.ld_out = C;
      add_attn(t + ".attn1", d);
    }
    release(qkv);
    Act h2 = linear(t + ".attn1.to_out.0", ao, nullptr, t + ".attn1.to_out.0", C, true, &h);
    release(ao);
    release(h);
    return h2;
  }

  // consumes h2 (released); x is the block input (residual of proj_out)
  Act transformer_tail(const std::string& key, const Act& x, Act& h2) {
    const int C = x.C, D = C / 8, T = x.H * x.W;
    const std::string t = key + ".transformer_blocks.0";
    // --- cross attention against the cached per-slot K/V
    Act ln2 = layernorm(t + ".norm2", h2, t + ".norm2");
    Act q2 = linear(t + ".attn2.to_q", ln2, nullptr, t + ".attn2.to_q", C, false, nullptr);
    release(ln2);
    Act ao2 = alloc(x.N, x.H, x.W, C);
    {
      AttnDesc d;
      d.B = x.N; d.heads = 8; d.D = D; d.Tq = T; d.Tk = kCtxTokens;
      const __half* kv = dry ? nullptr : e.kv_cache.at(t + ".attn2");
      d.q = hp(q2); d.k = kv; d.v = kv + C;
      d.ld_q = C; d.ld_k = d.ld_v = 2 * C;
      d.bs_q = static_cast<long long>(T) * C;
      d.bs_k = d.bs_v = static_cast<long long>(kCtxTokens) * 2 * C;
      d.kv_batches = kMaxCtxSlots;
      d.kv_index = p.ctx_idx;
      d.out = hp(ao2); d.ld_out = C;
      add_attn(t + ".attn2", d);
    }
    release(q2);
    Act h3 = linear(t + ".attn2.to_out.0", ao2, nullptr, t + ".attn2.to_out.0", C, true, &h2);
    release(ao2);
    release(h2);
    // --- GEGLU feed-forward
    Act ln3 = layernorm(t + ".norm3", h3, t + ".norm3");
    Act g = linear(t + ".ff.net.0.proj", ln3, nullptr, t + ".ff.net.0.proj", 8 * C, true, nullptr, /*geglu=*/true);
    release(ln3);
    Act h4 = linear(t + ".ff.net.2", g, nullptr, t + ".ff.net.2", C, true, &h3);
    release(g);
    release(h3);
    Act out = linear(key + ".proj_out", h4, nullptr, key + ".proj_out", C, true, &x, false, false, true);
    release(h4);
    tap(key, out);
    return out;
  }

  Act transformer(const std::string& key, const Act& x) {
    Act h2 = transformer_head(key, x);
    return transformer_tail(key, x, h2);
  }

  // replicate every row of x (N = groups) G times -> N*G rows (cond/uncond prefix sharing); with_stats: the copy will be
  // read by a GroupNorm (skip connection), so x's statistics record is replicated with it
  Act repeat_rows(const std::string& name, const Act& x, int G, bool with_stats = false) {
    Act o = alloc(x.N * G, x.H, x.W, x.C);
    if (with_stats && x.stats) {
      o.st_off = alloc_bytes(gn_record_floats(o.N, o.H * o.W, o.C) * sizeof(float));
      o.stats = true;
    }
    if (!dry) {
      const __half* in = hp(x);
      __half* out = hp(o);
      const long long rows = x.N, elems = static_cast<long long>(x.H) * x.W * x.C;
      push(Step{[=](cudaStream_t s) { repeat_rows_launch(in, rows, elems, G, out, s); }, kStepOther, 0, 1, name});
      if (o.stats) {
        const __half* sin = at<__half>(x.st_off);
        __half* sout = at<__half>(o.st_off);
        const long long selems = static_cast<long long>(gn_record_floats(1, x.H * x.W, x.C)) * 2;  // floats as half pairs
        push(Step{[=](cudaStream_t s) { repeat_rows_launch(sin, rows, selems, G, sout, s); }, kStepOther, 0, 1, name + ".stats"});
      }
    }
    return o;
  }

  Act downsample(const std::string& key, const Act& x) {
    const int H2 = (x.H + 1) / 2, W2 = (x.W + 1) / 2;
    Act planes = alloc(4 * x.N, H2, W2, x.C);
    if (!dry) {
      const __half* in = hp(x);
      __half* out = hp(planes);
      const int N = x.N, H = x.H, W = x.W, C = x.C;
      push(Step{[=](cudaStream_t s) { space_to_planes_launch(in, N, H, W, C, H2, W2, out, s); }, kStepOther, 0, 1,
                key + ".planes"});
    }
    Act o = alloc(x.N, H2, W2, x.C);
    IgemmDesc d;
    d.Nimg = x.N; d.H = H2; d.W = W2;
    d.nsrc = 1;
    d.src[0] = view(planes);
    seg_conv3x3_s2(d, x.C, x.N, false);
    d.Wt = dry ? nullptr : e.H(key + ".conv.weight");
    d.N = x.C; d.K = 9 * x.C;
    d.bias = dry ? nullptr : e.F(key + ".conv.bias");
    d.out = hp(o); d.ld_out = x.C;
    if (gn_fusable(1, H2, W2, x.C)) attach_stats(o, d, false);  // read by the next block's norm1 and, as a skip, by the up path
    add_igemm(key + ".conv", d);
    release(planes);
    tap(key, o);
    return o;
  }

  Act upsample(const std::string& key, const Act& x, int Ho, int Wo) {
    Act up = alloc(x.N, Ho, Wo, x.C);
    if (!dry) {
      const __half* in = hp(x);
      __half* out = hp(up);
      const int N = x.N, H = x.H, W = x.W, C = x.C;
      push(Step{[=](cudaStream_t s) { upsample_nearest_launch(in, N, H, W, C, Ho, Wo, out, s); }, kStepOther, 0, 1,
                key + ".nearest"});
    }
    Act o = conv3x3(key + ".conv", up, nullptr, key + ".conv", x.C, nullptr, 0, nullptr, true);
    release(up);
    tap(key, o);
    return o;
  }
};

}  // namespace

// key.kind: kPlanUnet (aux unused) -> p.out = conv_out prediction [Bf*h*w, 16] fp16
//           kPlanDift (aux = up_ft_index) -> p.out = activation after up_blocks[aux] (+upsampler), NHWC fp16
void build_unet_plan(Engine& e, Plan& p, bool dry, ArenaPlanner& ar) {
  UnetBuilder b(e, p, dry, ar);
  const int Bf = p.key.B, h = p.key.h, w = p.key.w;
  const int up_ft = p.key.kind == kPlanDift ? p.key.aux : -1;
  // kPlanUnet with aux = G > 1: rows come in groups of G consecutive forwards that share (x_t, t) and differ only in
  // the context (dm_typicality: cond fastest).  Everything before the first cross-attention -- conv_in,
  // down_blocks.0.resnets.0 and the first transformer up to its self-attention residual -- is computed once per
  // group on Bu = Bf / G rows and replicated; results are bit-identical to the unshared plan.
  const int G = (p.key.kind == kPlanUnet && p.key.aux > 1) ? p.key.aux : 1;
  DM_CHECK(Bf % G == 0, "shared-prefix plan: batch is not a multiple of the group size");
  const int Bu = Bf / G;
  b.Bf = Bf;
  const std::string U = "unet.";
  static const int ch[4] = {320, 640, 1280, 1280};

  // ---- boundary buffers
  const size_t off_ain = b.alloc_bytes(static_cast<size_t>(Bu) * h * w * 64 * sizeof(__half));
  const size_t off_sin = b.alloc_bytes(static_cast<size_t>(Bf) * 320 * sizeof(__half));
  const size_t off_ctx = b.alloc_bytes(static_cast<size_t>(Bf) * sizeof(int));
  p.a_in = b.at<__half>(off_ain);
  p.temb_sin = b.at<__half>(off_sin);
  p.ctx_idx = b.at<int>(off_ctx);
  const size_t off_loss = b.alloc_bytes(sizeof(IgLossArgs));  // zero-initialised with the arena: epilogue off
  p.loss_args = b.at<IgLossArgs>(off_loss);

  // ---- time embedding: sinusoid -> Linear+SiLU -> Linear(+SiLU for the ResNets) -> all 22 time_emb_proj at once
  Act sin_a; sin_a.off = off_sin; sin_a.N = 1; sin_a.H = 1; sin_a.W = Bf; sin_a.C = 320; sin_a.valid = true;
  Act e1 = b.linear("time_embedding.linear_1", sin_a, nullptr, U + "time_embedding.linear_1", kTimeDim, true, nullptr, false, true);
  Act semb = b.linear("time_embedding.linear_2", e1, nullptr, U + "time_embedding.linear_2", kTimeDim, true, nullptr, false, true);
  b.release(e1);
  Act tp = b.linear("time_emb_proj.all", semb, nullptr, U + "time_emb_proj_all", e.tproj_total, true, nullptr);
  b.release(semb);
  b.tproj = b.hp(tp);

  // ---- conv_in as a K=64 GEMM over the 36-wide 3x3x4 patch matrix
  Act ain; ain.off = off_ain; ain.N = Bu; ain.H = h; ain.W = w; ain.C = 64; ain.valid = true;
  Act x = b.linear("conv_in", ain, nullptr, U + "conv_in", 320, true, nullptr, false, false, true);
  b.tap("conv_in", x);

  std::vector<Act> skips;
  if (G == 1) skips.push_back(x);
  // ---- down path
  for (int i = 0; i < 4; ++i) {
    for (int j = 0; j < 2; ++j) {
      const std::string rk = U + "down_blocks." + std::to_string(i) + ".resnets." + std::to_string(j);
      const std::string ak = U + "down_blocks." + std::to_string(i) + ".attentions." + std::to_string(j);
      if (G > 1 && i == 0 && j == 0) {
        // shared prefix on Bu rows, then fan out to the Bf condition rows
        b.tproj_row_mult = G;
        Act r_u = b.resnet(rk, x, nullptr, ch[i]);
        b.tproj_row_mult = 1;
        Act h2_u = b.transformer_head(ak, r_u);
        Act s0 = b.repeat_rows("conv_in.fanout", x, G, true);
        b.release(x);
        skips.push_back(s0);
        Act r_f = b.repeat_rows(rk + ".fanout", r_u, G);
        b.release(r_u);
        Act h2_f = b.repeat_rows(ak + ".attn1.fanout", h2_u, G);
        b.release(h2_u);
        Act a = b.transformer_tail(ak, r_f, h2_f);
        b.release(r_f);
        x = a;
        skips.push_back(x);
        continue;
      }
      Act r = b.resnet(rk, x, nullptr, ch[i]);
      if (i < 3) {
        Act a = b.transformer(ak, r);
        b.release(r);
        r = a;
      }
      x = r;
      skips.push_back(x);
    }
    if (i < 3) {
      x = b.downsample(U + "down_blocks." + std::to_string(i) + ".downsamplers.0", x);
      skips.push_back(x);
    }
  }
  // ---- mid
  {
    Act r0 = b.resnet(U + "mid_block.resnets.0", x, nullptr, 1280);
    Act a = b.transformer(U + "mid_block.attentions.0", r0);
    b.release(r0);
    Act r1 = b.resnet(U + "mid_block.resnets.1", a, nullptr, 1280);
    b.release(a);
    x = r1;  // (x's previous value is skips.back(): still owned by the skip stack)
    b.tap("mid_block", x);
  }
  // ---- up path
  static const int up_out[4] = {1280, 1280, 640, 320};
  bool done = false;
  for (int i = 0; i < 4 && !done; ++i) {
    for (int j = 0; j < 3; ++j) {
      Act skip = skips.back();
      skips.pop_back();
      const std::string rk = U + "up_blocks." + std::to_string(i) + ".resnets." + std::to_string(j);
      const std::string ak = U + "up_blocks." + std::to_string(i) + ".attentions." + std::to_string(j);
      Act r = b.resnet(rk, x, &skip, up_out[i]);
      b.release(x);
      b.release(skip);
      if (i > 0) {
        Act a = b.transformer(ak, r);
        b.release(r);
        r = a;
      }
      x = r;
    }
    if (i < 3) {
      // nearest 2x, or nearest-to-the-skip's-size when the latent is not a multiple of 8 (dift.py:48-57,144-147)
      const Act& nxt = skips.back();
      Act u = b.upsample(U + "up_blocks." + std::to_string(i) + ".upsamplers.0", x, nxt.H, nxt.W);
      b.release(x);
      x = u;
    }
    if (i == up_ft) done = true;
  }
  if (up_ft >= 0) {
    p.out = b.hp(x);
    p.out_H = x.H; p.out_W = x.W; p.out_C = x.C;
    return;
  }
  // ---- out: GN + SiLU -> conv 320 -> 4 (weights zero-padded to 16 rows)
  Act n = b.groupnorm("conv_norm_out", x, nullptr, U + "conv_norm_out", 1e-5f, true);
  b.release(x);
  // conv 320 -> 4 (rows zero-padded to 16); its epilogue also forms (pred - eps)^2 into the raw fp16 grid when the
  // call is dm_typicality (the eps-MSE "fused into the final store" of the north star)
  Act pred = b.alloc(n.N, n.H, n.W, 16);
  {
    IgemmDesc d;
    d.Nimg = n.N; d.H = n.H; d.W = n.W;
    d.nsrc = 1;
    d.src[0] = b.view(n);
    seg_conv3x3(d, n.C, n.C);
    d.Wt = dry ? nullptr : e.H(U + "conv_out.weight");
    d.N = 16; d.K = 9 * n.C;
    d.bias = dry ? nullptr : e.F(U + "conv_out.bias");
    d.out = b.hp(pred); d.ld_out = 16;
    d.loss = p.loss_args;
    b.add_igemm("conv_out", d);
  }
  b.release(n);
  b.tap("conv_out", pred);
  p.out = b.hp(pred);
  p.out_H = h; p.out_W = w; p.out_C = 16;
}

}  // namespace dm
