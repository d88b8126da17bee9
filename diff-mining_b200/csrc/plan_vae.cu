// Execution plan of the SD-1.5 VAE encoder (AutoencoderKL.encode up to the posterior moments), reached from
// SD.encode_vae (/root/reference/diffmining/typicality/compute.py:91-93) and OneStepSDPipeline (dift.py:187).
// Same kernel family as the U-Net at C = 128/256/512; the single 512-wide attention head of the mid block runs on its own
// flash kernel (attention.cuh: vattention_kernel), batched over images, for any token count.
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

  // Attention(heads=1, dim 512, GroupNorm, q/k/v/out with bias, residual) over T = H*W tokens per image: one fused
  // q|k|v projection, the single-head flash kernel (vattention_kernel: no T x T buffer, any T, all images in one launch),
  // output projection + residual
  Act mid_attention(const std::string& key, const Act& x) {
    const int C = x.C, T = x.H * x.W, B = x.N;
    Act n = groupnorm(key + ".group_norm", x, nullptr, key + ".group_norm", 1e-6f, false);
    Act qkv = linear(key + ".qkv", n, nullptr, key + ".qkv", 3 * C, true, nullptr);  // [M, q | k | v]
    release(n);
    Act o = alloc(B, x.H, x.W, C);
    {
      AttnDesc d;
      d.B = B; d.heads = 1; d.D = C; d.Tq = T; d.Tk = T;
      d.q = hp(qkv); d.k = hp(qkv) + C; d.v = hp(qkv) + 2 * C;
      d.ld_q = d.ld_k = d.ld_v = 3 * C;
      d.bs_q = d.bs_k = d.bs_v = static_cast<long long>(T) * 3 * C;
      d.out = hp(o); d.ld_out = C;
      add_attn(key, d);
    }
    release(qkv);
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
