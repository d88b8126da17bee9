// Execution plan of the SD-1.5 VAE encoder (AutoencoderKL.encode up to the posterior moments), reached from
// SD.encode_vae (/root/reference/diffmining/typicality/compute.py:91-93) and OneStepSDPipeline (dift.py:187).
// Same kernel family as the U-Net at C = 128/256/512; the single 512-d attention head of the mid block is run
// unfused (S = QK^T in fp32 -> row softmax -> PV) through the implicit-GEMM kernel, one image at a time.
#include "plan_builder.h"

namespace dm {

namespace {

struct VaeBuilder : Builder {
  using Builder::Builder;

  Act resnet(const std::string& key, const Act& x, int Cout) {
    Act n1 = groupnorm(key + ".norm1", x, nullptr, key + ".norm1", 1e-6f, true);
    Act h1 = conv3x3(key + ".conv1", n1, nullptr, key + ".conv1", Cout, nullptr, 0, nullptr);
    release(n1);
    Act n2 = groupnorm(key + ".norm2", h1, nullptr, key + ".norm2", 1e-6f, true);
    release(h1);
    Act sc;
    const Act* res = &x;
    if (x.C != Cout) {
      sc = linear(key + ".conv_shortcut", x, nullptr, key + ".conv_shortcut", Cout, true, nullptr);
      res = &sc;
    }
    Act out = conv3x3(key + ".conv2", n2, nullptr, key + ".conv2", Cout, nullptr, 0, res);
    release(n2);
    release(sc);
    return out;
  }

  // zero-pad (right, bottom) by one then 3x3 stride-2 pad-0 conv -> floor(H/2) x floor(W/2)
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
    Act o = alloc(x.N, x.H / 2, x.W / 2, x.C);
    IgemmDesc d;
    d.Nimg = x.N; d.H = x.H / 2; d.W = x.W / 2;
    d.nsrc = 1;
    d.src[0] = view(planes);
    seg_conv3x3_s2(d, x.C, x.N, true);
    d.Wt = dry ? nullptr : e.H(key + ".conv.weight");
    d.N = x.C; d.K = 9 * x.C;
    d.bias = dry ? nullptr : e.F(key + ".conv.bias");
    d.out = hp(o); d.ld_out = x.C;
    add_igemm(key + ".conv", d);
    release(planes);
    return o;
  }

  // Attention(heads=1, dim 512, GroupNorm, q/k/v/out with bias, residual) over T = H*W tokens per image
  Act mid_attention(const std::string& key, const Act& x) {
    const int C = x.C, T = x.H * x.W, B = x.N;
    DM_CHECK(T % 8 == 0, "VAE attention needs (H/8)*(W/8) to be a multiple of 8; got " + std::to_string(T));
    Act n = groupnorm(key + ".group_norm", x, nullptr, key + ".group_norm", 1e-6f, false);
    Act qk = linear(key + ".qk", n, nullptr, key + ".qk", 2 * C, true, nullptr);  // [M, q | k]
    Act vT = alloc(B, 1, C, T);     // per image [C, T] = V^T without bias (bias re-added after PV: softmax rows sum to 1)
    Act o = alloc(B, x.H, x.W, C);  // attention output tokens
    const size_t s_off = alloc_bytes(static_cast<size_t>(T) * T * sizeof(float));
    const size_t p_off = alloc_bytes(static_cast<size_t>(T) * T * sizeof(__half));
    const int kchunks = (T + 63) / 64;
    for (int b = 0; b < B; ++b) {
      const __half* nb = hp(n) + static_cast<size_t>(b) * T * C;
      const __half* qb = hp(qk) + static_cast<size_t>(b) * T * 2 * C;
      {  // V^T[c, t] = sum_i Wv[c, i] * n[t, i]   (A = Wv as a 512-row "activation", B = tokens)
        IgemmDesc d;
        d.Nimg = 1; d.H = 1; d.W = C;
        d.nsrc = 1;
        d.src[0] = ActView{dry ? nullptr : e.H(key + ".to_v.weight"), 1, 1, C, C, C};
        seg_1x1(d, C, 0);
        d.Wt = nb; d.N = T; d.K = C;
        d.out = hp(vT) + static_cast<size_t>(b) * C * T; d.ld_out = T;
        add_igemm(key + ".vT", d);
      }
      {  // S = q k^T  (fp32)
        IgemmDesc d;
        d.Nimg = 1; d.H = 1; d.W = T;
        d.nsrc = 1;
        d.src[0] = ActView{qb, 1, 1, T, C, 2 * C};
        seg_1x1(d, C, 0);
        d.Wt = qb + C; d.N = T; d.K = C;
        d.w_ld = 2 * C;
        d.out = at<float>(s_off); d.ld_out = T; d.out_f32 = 1;
        add_igemm(key + ".qk^T", d);
      }
      if (!dry) {
        const float* S = at<float>(s_off);
        __half* P = at<__half>(p_off);
        const float scale = 1.0f / sqrtf(static_cast<float>(C));
        push(Step{[=](cudaStream_t s) { softmax_rows_launch(S, T, T, T, scale, P, T, s); }, kStepOther, 0, 1,
                  key + ".softmax"});
      }
      {  // O = P V + b_v
        IgemmDesc d;
        d.Nimg = 1; d.H = 1; d.W = T;
        d.nsrc = 1;
        d.src[0] = ActView{at<__half>(p_off), 1, 1, T, T, T};
        d.nseg = 1;
        d.seg[0] = IgSeg{0, 0, 0, 0, 0, 0, kchunks};
        d.Wt = hp(vT) + static_cast<size_t>(b) * C * T; d.N = C; d.K = T;
        d.k_ragged = 1;
        d.bias = dry ? nullptr : e.F(key + ".to_v.bias");
        d.out = hp(o) + static_cast<size_t>(b) * T * C; d.ld_out = C;
        add_igemm(key + ".pv", d);
      }
    }
    ar.free(s_off);
    ar.free(p_off);
    release(n);
    release(qk);
    release(vT);
    Act out = linear(key + ".to_out.0", o, nullptr, key + ".to_out.0", C, true, &x);
    release(o);
    return out;
  }
};

}  // namespace

// key: kind = kPlanVae, B images, h/w = IMAGE height/width.  p.out = conv_out [B*(h/8)*(w/8), 16] fp16 (8 valid).
void build_vae_plan(Engine& e, Plan& p, bool dry, ArenaPlanner& ar) {
  VaeBuilder b(e, p, dry, ar);
  const int B = p.key.B, H = p.key.h, W = p.key.w;
  const std::string V = "vae.encoder.";
  static const int ch[4] = {128, 256, 512, 512};
  const size_t off_ain = b.alloc_bytes(static_cast<size_t>(B) * H * W * 64 * sizeof(__half));
  p.a_in = b.at<__half>(off_ain);
  Act ain; ain.off = off_ain; ain.N = B; ain.H = H; ain.W = W; ain.C = 64; ain.valid = true;
  Act x = b.linear("conv_in", ain, nullptr, V + "conv_in", 128, true, nullptr);
  ar.free(off_ain);  // patch matrix is dead after conv_in (it is rewritten before every run)
  for (int i = 0; i < 4; ++i) {
    for (int j = 0; j < 2; ++j) {
      Act r = b.resnet(V + "down_blocks." + std::to_string(i) + ".resnets." + std::to_string(j), x, ch[i]);
      b.release(x);
      x = r;
    }
    b.tap("encoder.down_blocks." + std::to_string(i), x);
    if (i < 3) {
      Act d = b.downsample(V + "down_blocks." + std::to_string(i) + ".downsamplers.0", x);
      b.release(x);
      x = d;
    }
  }
  Act r0 = b.resnet(V + "mid_block.resnets.0", x, 512);
  b.release(x);
  Act a = b.mid_attention(V + "mid_block.attentions.0", r0);
  b.release(r0);
  b.tap("encoder.mid_block.attentions.0", a);
  Act r1 = b.resnet(V + "mid_block.resnets.1", a, 512);
  b.release(a);
  Act n = b.groupnorm("conv_norm_out", r1, nullptr, V + "conv_norm_out", 1e-6f, true);
  b.release(r1);
  Act mo = b.conv3x3("conv_out", n, nullptr, V + "conv_out", 16, nullptr, 0, nullptr);
  b.release(n);
  p.out = b.hp(mo);
  p.out_H = mo.H; p.out_W = mo.W; p.out_C = 16;
}

}  // namespace dm
