// Host-side preparation + launch of every kernel (see ops.h).
#include "ops.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <type_traits>

#include "misc.cuh"
#include "norm.cuh"

namespace dm {

// ------------------------------------------------------------------ tensor maps
using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  DM_CHECK(fn != nullptr, "cuTensorMapEncodeTiled not available (no CUDA driver?)");
  return fn;
}

// fp16 tensor, dims innermost-first, strides (bytes) for dims 1..rank-1, 128B swizzle, zero OOB fill
static void make_tmap(CUtensorMap* m, const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides_b,
                      const uint32_t* box, CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B) {
  cuuint64_t gd[5], gs[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) {
    gd[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
    DM_CHECK(dims[i] > 0 && box[i] > 0 && box[i] <= 256, "tensor map: bad dim/box");
  }
  for (int i = 0; i < rank - 1; ++i) {
    gs[i] = strides_b[i];
    DM_CHECK(gs[i] % 16 == 0, "tensor map: stride not a multiple of 16 bytes");
  }
  DM_CHECK((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, "tensor map: base not 16-byte aligned");
  CUresult r = encode_fn()(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, rank, const_cast<void*>(ptr), gd, gs, bx, es,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DM_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with code " + std::to_string(static_cast<int>(r)));
}


// cudaFuncSetAttribute is per device: one-time kernel configuration is keyed by the current device ordinal, so a second
// engine on another GPU of the same process configures its own copy of every kernel
static bool first_use_on_this_device(bool (&done)[64]) {
  int dev = 0;
  DM_CUDA(cudaGetDevice(&dev));
  DM_CHECK(dev >= 0 && dev < 64, "device ordinal out of range");
  if (done[dev]) return false;
  done[dev] = true;
  return true;
}

// ------------------------------------------------------------------ kernel-variant switches
static int g_igemm_pair = -1, g_gn_fused = -1;
static int env_or(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}
static int variant_igemm_pair() { return g_igemm_pair >= 0 ? g_igemm_pair : env_or("DM_IGEMM_PAIR", 1); }
static int variant_gn_fused() { return g_gn_fused >= 0 ? g_gn_fused : env_or("DM_GN_FUSED", 1); }
static int g_xattn = -1, g_prefix = -1, g_ng4 = -1, g_attn3 = -1;
static int variant_attn3() { return g_attn3 >= 0 ? g_attn3 : env_or("DM_ATTN3", 1); }
static int variant_igemm_ng4() { return g_ng4 >= 0 ? g_ng4 : env_or("DM_IGEMM_NG4", 1); }
int variant_prefix_share() { return g_prefix >= 0 ? g_prefix : env_or("DM_PREFIX_SHARE", 1); }
static int g_gn_epi = -1, g_igemm_ws = -1;
static int variant_igemm_ws() { return g_igemm_ws >= 0 ? g_igemm_ws : env_or("DM_IGEMM_WS", 0); }
int gn_epilogue_mode() { return g_gn_epi >= 0 ? g_gn_epi : env_or("DM_GN_EPILOGUE", 3); }
static int variant_xattn() { return g_xattn >= 0 ? g_xattn : env_or("DM_XATTN", 2); }
void set_variant(const std::string& name, int value) {
  if (name == "igemm_pair") g_igemm_pair = value;
  else if (name == "gn_fused") g_gn_fused = value;
  else if (name == "xattn") g_xattn = value;
  else if (name == "prefix_share") g_prefix = value;
  else if (name == "igemm_ng4") g_ng4 = value;
  else if (name == "attn3") g_attn3 = value;
  else if (name == "gn_epilogue") g_gn_epi = value;
  else if (name == "igemm_ws") g_igemm_ws = value;
  else DM_CHECK(false, "unknown kernel variant '" + name + "'");
}

// ------------------------------------------------------------------ igemm
void seg_conv3x3(IgemmDesc& d, int Cin_total, int C0) {
  d.nseg = 0;
  for (int tap = 0; tap < 9; ++tap) {
    const int dy = tap / 3 - 1, dx = tap % 3 - 1;
    IgSeg s{};
    s.src = 0; s.dy = dy; s.dx = dx; s.dn = 0; s.chan0 = 0; s.nchunks = C0 / 64;
    d.seg[d.nseg++] = s;
    if (Cin_total > C0) {
      s.src = 1; s.nchunks = (Cin_total - C0) / 64;
      d.seg[d.nseg++] = s;
    }
  }
}
void seg_conv3x3_s2(IgemmDesc& d, int C, int Nimg, bool vae_pad) {
  // src0 = parity planes [(py*2+px)*Nimg + n, H2, W2, C]
  static const int unet_par[3] = {1, 0, 1}, unet_off[3] = {-1, 0, 0};  // pad 1:   in = 2*out + r - 1
  static const int vae_par[3] = {0, 1, 0}, vae_off[3] = {0, 0, 1};     // pad 0/1: in = 2*out + r
  const int* par = vae_pad ? vae_par : unet_par;
  const int* off = vae_pad ? vae_off : unet_off;
  d.nseg = 0;
  for (int tap = 0; tap < 9; ++tap) {
    const int r = tap / 3, c = tap % 3;
    IgSeg s{};
    s.src = 0; s.dy = off[r]; s.dx = off[c]; s.dn = (par[r] * 2 + par[c]) * Nimg; s.chan0 = 0; s.nchunks = C / 64;
    d.seg[d.nseg++] = s;
  }
}
void seg_1x1(IgemmDesc& d, int C0, int C1) {
  d.nseg = 0;
  IgSeg s{};
  s.src = 0; s.nchunks = C0 / 64;
  d.seg[d.nseg++] = s;
  if (C1 > 0) {
    s.src = 1; s.nchunks = C1 / 64;
    d.seg[d.nseg++] = s;
  }
}

static int pick_bn(int N) {
  static const int cand[] = {256, 160, 128, 64, 32, 16};
  int best = 16;
  long long best_pad = -1;
  for (int bn : cand) {
    const long long pad = static_cast<long long>((N + bn - 1) / bn) * bn;
    if (best_pad < 0 || pad < best_pad) { best_pad = pad; best = bn; }
  }
  return best;
}

// M-tile shape: 2^wt x 2^ht x 2^nt = 128 pixels, minimal padding, widest rows preferred
static void pick_mtile(int Nimg, int H, int W, int& wt_log, int& ht_log, int& nt_log) {
  long long best = -1;
  for (int wl = 7; wl >= 0; --wl)
    for (int hl = 7 - wl; hl >= 0; --hl) {
      const int nl = 7 - wl - hl;
      const long long wt = 1 << wl, ht = 1 << hl, nt = 1 << nl;
      const long long pad = ((W + wt - 1) / wt * wt) * ((H + ht - 1) / ht * ht) * ((Nimg + nt - 1) / nt * nt);
      if (best < 0 || pad < best) { best = pad; wt_log = wl; ht_log = hl; nt_log = nl; }
    }
}


bool igemm_gn_fusable(int H, int W, int N) {
  if (N % 64 != 0 || (static_cast<long long>(H) * W) % IG_BM != 0) return false;
  int wl = 0, hl = 0, nl = 0;
  pick_mtile(2, H, W, wl, hl, nl);  // conv layout: the m-tile must be 2^hl x 2^wl pixels of ONE image, tiling it exactly
  const int bn = pick_bn(N);
  return nl == 0 && (W % (1 << wl)) == 0 && (H % (1 << hl)) == 0 && N % bn == 0 && bn >= 32;
}
size_t gn_record_floats(int Nimg, int HW, int C) { return static_cast<size_t>(Nimg) * (2 * (HW / IG_BM * 4) + 1) * C; }

IgemmOp igemm_prepare(const IgemmDesc& d, int num_sms) {
  IgemmOp op{};
  IgParams& p = op.p;
  DM_CHECK(d.nseg > 0 && d.nseg <= IG_MAX_SEG, "igemm: bad segment count");
  DM_CHECK(d.N % 8 == 0 && d.N >= 16, "igemm: N must be a multiple of 8 and >= 16");
  int kit = 0;
  for (int i = 0; i < d.nseg; ++i) {
    DM_CHECK(d.seg[i].src < d.nsrc, "igemm: segment source out of range");
    kit += d.seg[i].nchunks;
    p.seg[i] = d.seg[i];
  }
  if (d.k_ragged)
    DM_CHECK(kit == (d.K + IG_BK - 1) / IG_BK && d.nseg == 1 && d.K % 8 == 0, "igemm: bad ragged-K description");
  else
    DM_CHECK(kit * IG_BK == d.K, "igemm: K (" + std::to_string(d.K) + ") != 64 * chunks (" + std::to_string(kit) + ")");
  p.nseg = d.nseg;
  p.k_iters = kit;
  p.Nimg = d.Nimg; p.H = d.H; p.W = d.W;
  pick_mtile(d.Nimg, d.H, d.W, p.wt_log, p.ht_log, p.nt_log);
  const int wt = 1 << p.wt_log, ht = 1 << p.ht_log, nt = 1 << p.nt_log;
  p.tiles_x = (d.W + wt - 1) / wt;
  p.tiles_y = (d.H + ht - 1) / ht;
  p.tiles_n = (d.Nimg + nt - 1) / nt;
  p.m_tiles = p.tiles_x * p.tiles_y * p.tiles_n;
  op.bn = d.bn ? d.bn : pick_bn(d.N);
  p.n_tiles = (d.N + op.bn - 1) / op.bn;
  p.N = d.N;
  p.bias = d.bias; p.rowbias = d.rowbias; p.ld_rowbias = d.ld_rowbias;
  p.residual = d.residual; p.ld_res = d.ld_res;
  p.out = d.out; p.ld_out = d.ld_out;
  p.out_f32 = d.out_f32; p.geglu = d.geglu; p.act_silu = d.act_silu;
  p.loss = d.loss;
  DM_CHECK(d.out != nullptr && d.ld_out % 8 == 0, "igemm: bad output");
  // staged epilogue (TMA store / residual prefetch) unless the output is fp32 or the N-tile is narrower than a chunk
  op.direct = (d.out_f32 || op.bn < 32) ? 1 : 0;
  // CTA pairs (cta_group::2, 256-row tiles) when there is enough work to keep every SM pair busy
  const int pair_mode = variant_igemm_pair();
  // (pair_mode 2 = use pairs whenever the tile shape allows it: unit tests)
  op.cg = (pair_mode && !op.direct && op.bn >= 128 &&
           (pair_mode == 2 || d.cg == 2 ||
            (d.cg == 0 && p.m_tiles >= 2 && kit >= 16 &&  // short-K layers are epilogue-bound: pairs only couple them
             static_cast<long long>((p.m_tiles + 1) / 2) * p.n_tiles >= num_sms / 2)))
              ? 2 : 1;
  // weight-stationary CTA pairs for the short-K (K <= 320) Linears with many M-tiles: the B tile of ONE N-tile stays in shared
  // memory while the pair walks down M (igemm.cuh: IgCfg).  ws_mode 2 = wherever the shape allows (unit tests).
  {
    const int ws_mode = variant_igemm_ws();
    const int m_units = (p.m_tiles + 1) / 2;
    const int per_n = p.n_tiles > 0 ? (num_sms / 2) / p.n_tiles : 0;  // pairs per N-tile
    op.ws = (ws_mode && !op.direct && !d.k_ragged && kit <= IG_WS_KCHUNKS && (op.bn == 256 || op.bn == 160) && per_n >= 1 &&
             p.m_tiles >= 2 && (ws_mode == 2 || m_units >= 8 * per_n)) ? 1 : 0;
    if (op.ws) op.cg = 2;
  }
  DM_CHECK(d.loss == nullptr || (op.direct && !d.out_f32 && d.N >= 4), "igemm: the fused loss epilogue needs the direct fp16 epilogue");
  if (d.gn.rec != nullptr) {
    IgGn& g = p.gn;
    g.rec = d.gn.rec;
    g.tiles_img = d.gn.tiles_img ? d.gn.tiles_img : p.tiles_x * p.tiles_y;
    // whole 128-pixel tiles, each inside one image; whole N-tiles; the staged fp16 epilogue
    DM_CHECK(!op.direct && !d.geglu && d.bn == 0 && d.N % op.bn == 0 && d.N % 64 == 0 && p.nt_log == 0 &&
                 d.W % (1 << p.wt_log) == 0 && d.H % (1 << p.ht_log) == 0 && p.m_tiles % g.tiles_img == 0 &&
                 (d.gn.tiles_img == 0 || (d.Nimg == 1 && d.H == 1)),
             "igemm: GroupNorm statistics in the epilogue need whole 128-pixel tiles per image and whole N-tiles");
  }
  DM_CHECK(!d.geglu || (op.bn % 64 == 0 && !op.direct), "igemm: GEGLU needs an N-tile that is a multiple of 64");
  DM_CHECK(!d.geglu || (!d.residual && !d.rowbias && !d.act_silu), "igemm: GEGLU excludes the other epilogue options");
  if (!op.direct) {
    const int Cout = d.geglu ? d.N / 2 : d.N;
    auto cmap = [&](CUtensorMap* m, const void* ptr, long long ld) {
      DM_CHECK(ld % 8 == 0 && ld >= Cout, "igemm: output / residual row stride must be a multiple of 8 elements");
      const uint64_t dims[4] = {static_cast<uint64_t>(Cout), static_cast<uint64_t>(d.W), static_cast<uint64_t>(d.H),
                                static_cast<uint64_t>(d.Nimg)};
      const uint64_t st[3] = {static_cast<uint64_t>(ld) * 2, static_cast<uint64_t>(ld) * 2 * d.W,
                              static_cast<uint64_t>(ld) * 2 * d.W * d.H};
      const uint32_t box[4] = {IG_CW, static_cast<uint32_t>(1 << p.wt_log), static_cast<uint32_t>(1 << p.ht_log),
                               static_cast<uint32_t>(1 << p.nt_log)};
      make_tmap(m, ptr, 4, dims, st, box, CU_TENSOR_MAP_SWIZZLE_64B);
    };
    cmap(&op.maps.c, d.out, d.ld_out);
    if (d.residual) cmap(&op.maps.r, d.residual, d.ld_res);
    else op.maps.r = op.maps.c;
  }
  for (int s = 0; s < d.nsrc; ++s) {
    const ActView& a = d.src[s];
    DM_CHECK(a.ptr && a.C > 0 && a.pix_stride >= a.C, "igemm: bad source view");
    const uint64_t dims[4] = {static_cast<uint64_t>(a.C), static_cast<uint64_t>(a.W), static_cast<uint64_t>(a.H),
                              static_cast<uint64_t>(a.N)};
    const uint64_t st[3] = {static_cast<uint64_t>(a.pix_stride) * 2, static_cast<uint64_t>(a.pix_stride) * 2 * a.W,
                            static_cast<uint64_t>(a.pix_stride) * 2 * a.W * a.H};
    const uint32_t box[4] = {64, static_cast<uint32_t>(wt), static_cast<uint32_t>(ht), static_cast<uint32_t>(nt)};
    make_tmap(&op.maps.a[s], a.ptr, 4, dims, st, box);
  }
  if (d.nsrc == 1) op.maps.a[1] = op.maps.a[0];
  {
    const uint64_t dims[2] = {static_cast<uint64_t>(d.K), static_cast<uint64_t>(d.N)};
    const uint64_t st[1] = {static_cast<uint64_t>(d.w_ld ? d.w_ld : d.K) * 2};
    const uint32_t box[2] = {64, static_cast<uint32_t>(op.bn / op.cg)};  // a CTA of a pair loads half of the B tile
    make_tmap(&op.maps.b, d.Wt, 2, dims, st, box);
  }
  // four epilogue warpgroups where the epilogue math bounds the tile: GEGLU with K = 320 (-14 % measured; the plain
  // K = 320 / 640 Linears are HBM-bound and measured 2-5 % slower with four groups)
  const int ng_mode = variant_igemm_ng4();
  op.ng = (ng_mode && !op.direct && op.cg == 1 && (op.bn == 256 || op.bn == 160) &&
           (ng_mode == 2 || (d.geglu && kit <= 5 && static_cast<long long>(p.m_tiles) * p.n_tiles >= 2 * num_sms))) ? 4 : 2;
  const int tiles = ((p.m_tiles + op.cg - 1) / op.cg) * p.n_tiles;
  op.grid = std::min(tiles, num_sms / op.cg) * op.cg;
  if (op.ws) {
    // units = a multiple of n_tiles, so that unit + k * units never changes its N-tile
    const int per_n = std::min((num_sms / 2) / p.n_tiles, (p.m_tiles + 1) / 2);
    op.grid = 2 * per_n * p.n_tiles;
    op.ng = (d.geglu && variant_igemm_ng4()) ? 4 : 2;
  }
  op.flops = 2.0 * d.Nimg * d.H * d.W * static_cast<double>(d.N) * d.K;
  return op;
}

template <int BN, bool DIRECT, int CG, int NG = 2, bool WS = false>
static void igemm_launch_bn(const IgemmOp& op, cudaStream_t s) {
  static bool configured[64] = {};
  using Cfg = IgCfg<BN, DIRECT, CG, NG, WS>;
  if (first_use_on_this_device(configured)) {
    DM_CUDA(cudaFuncSetAttribute(igemm_kernel<BN, DIRECT, CG, NG, WS>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
  }
  if (CG == 1) {
    igemm_kernel<BN, DIRECT, CG, NG, WS><<<op.grid, Cfg::THREADS, Cfg::SMEM_BYTES, s>>>(op.maps, op.p);
    DM_CUDA(cudaGetLastError());
  } else {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(op.grid, 1, 1);
    cfg.blockDim = dim3(Cfg::THREADS, 1, 1);
    cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CG;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    DM_CUDA(cudaLaunchKernelEx(&cfg, igemm_kernel<BN, DIRECT, CG, NG, WS>, op.maps, op.p));
  }
}

void igemm_launch(const IgemmOp& op, cudaStream_t s) {
  if (op.direct) {
    switch (op.bn) {
      case 256: igemm_launch_bn<256, true, 1>(op, s); break;
      case 160: igemm_launch_bn<160, true, 1>(op, s); break;
      case 128: igemm_launch_bn<128, true, 1>(op, s); break;
      case 64: igemm_launch_bn<64, true, 1>(op, s); break;
      case 32: igemm_launch_bn<32, true, 1>(op, s); break;
      case 16: igemm_launch_bn<16, true, 1>(op, s); break;
      default: DM_CHECK(false, "igemm: unsupported BN " + std::to_string(op.bn) + " for the direct epilogue");
    }
    return;
  }
  if (op.ws) {
    if (op.ng == 4) {
      switch (op.bn) {
        case 256: igemm_launch_bn<256, false, 2, 4, true>(op, s); break;
        default: DM_CHECK(false, "igemm: unsupported BN " + std::to_string(op.bn) + " for weight-stationary pairs with four groups");
      }
      return;
    }
    switch (op.bn) {
      case 256: igemm_launch_bn<256, false, 2, 2, true>(op, s); break;
      case 160: igemm_launch_bn<160, false, 2, 2, true>(op, s); break;
      default: DM_CHECK(false, "igemm: unsupported BN " + std::to_string(op.bn) + " for weight-stationary pairs");
    }
    return;
  }
  if (op.cg == 2) {
    switch (op.bn) {
      case 256: igemm_launch_bn<256, false, 2>(op, s); break;
      case 160: igemm_launch_bn<160, false, 2>(op, s); break;
      case 128: igemm_launch_bn<128, false, 2>(op, s); break;
      default: DM_CHECK(false, "igemm: unsupported BN " + std::to_string(op.bn) + " for CTA pairs");
    }
    return;
  }
  if (op.ng == 4) {
    switch (op.bn) {
      case 256: igemm_launch_bn<256, false, 1, 4>(op, s); break;
      case 160: igemm_launch_bn<160, false, 1, 4>(op, s); break;
      default: DM_CHECK(false, "igemm: unsupported BN " + std::to_string(op.bn) + " for four epilogue groups");
    }
    return;
  }
  switch (op.bn) {
    case 256: igemm_launch_bn<256, false, 1>(op, s); break;
    case 160: igemm_launch_bn<160, false, 1>(op, s); break;
    case 128: igemm_launch_bn<128, false, 1>(op, s); break;
    case 64: igemm_launch_bn<64, false, 1>(op, s); break;
    case 32: igemm_launch_bn<32, false, 1>(op, s); break;
    default: DM_CHECK(false, "igemm: unsupported BN " + std::to_string(op.bn));
  }
}

// ------------------------------------------------------------------ attention
AttnOp attn_prepare(const AttnDesc& d) {
  AttnOp op{};
  DM_CHECK(d.D == 40 || d.D == 80 || d.D == 160 || (d.D == 512 && d.heads == 1),
           "attention: head_dim must be 40, 80 or 160 (or one 512-wide head: the VAE mid block)");
  op.vattn = d.D == 512 ? 1 : 0;
  DM_CHECK(d.Tq > 0 && d.Tk > 0 && d.B > 0, "attention: empty problem");
  op.D = d.D;
  // long self-attention (head_dim 40 / 80) -> warp-specialised kernel; short key sets (cross-attention, 77 keys)
  // and head_dim 160 stay on the one-tile kernel.  DM_ATTN2=0 disables, DM_ATTN2=2 also routes cross-attention.
  static const int attn2_mode = [] { const char* e = getenv("DM_ATTN2"); return e ? atoi(e) : 1; }();
  op.v2 = !op.vattn && (d.D == 40 || d.D == 80) && attn2_mode > 0 && (d.Tk > 128 || attn2_mode > 1) ? 1 : 0;
  // short key sets (text cross-attention, <= 80 keys): single-tile kernel with P kept in tensor memory
  op.xattn = (!op.vattn && !op.v2 && d.Tk <= 80 && variant_xattn()) ? 1 : 0;
  const int bkv = op.vattn ? VAttnCfg::BKV : op.xattn ? 80 : d.D == 40 ? 128 : 64;
  auto mk = [&](CUtensorMap* m, const __half* ptr, long long ld, long long bs, int T, int nb, int rows) {
    const uint64_t dims[4] = {static_cast<uint64_t>(d.D), static_cast<uint64_t>(d.heads), static_cast<uint64_t>(T),
                              static_cast<uint64_t>(nb)};
    const uint64_t st[3] = {static_cast<uint64_t>(d.D) * 2, static_cast<uint64_t>(ld) * 2, static_cast<uint64_t>(bs) * 2};
    const uint32_t box[4] = {64, 1, static_cast<uint32_t>(rows), 1};
    make_tmap(m, ptr, 4, dims, st, box);
  };
  const int kvb = d.kv_batches ? d.kv_batches : d.B;
  mk(&op.maps.q, d.q, d.ld_q, d.bs_q, d.Tq, d.B, 128);
  mk(&op.maps.k, d.k, d.ld_k, d.bs_k, d.Tk, kvb, bkv);
  mk(&op.maps.v, d.v, d.ld_v, d.bs_v, d.Tk, kvb, bkv);
  op.p.Tq = d.Tq; op.p.Tk = d.Tk; op.p.B = d.B; op.p.heads = d.heads;
  op.p.kv_index = d.kv_index;
  op.p.out = d.out; op.p.ld_out = d.ld_out;
  op.p.scale_log2 = static_cast<float>(1.0 / std::sqrt(static_cast<double>(d.D)) * 1.4426950408889634);
  op.v3 = (op.v2 && variant_attn3()) ? 1 : 0;
  if (op.vattn) {
    DM_CHECK(d.kv_index == nullptr, "attention: the 512-wide kernel takes no K/V indirection");
    op.grid = dim3((d.Tq + 127) / 128, 2, d.B);
  } else if (op.v3) {
    int dev = 0, sms = 0;
    DM_CUDA(cudaGetDevice(&dev));
    DM_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const long long work = static_cast<long long>((d.Tq + 255) / 256) * d.heads * d.B;
    op.grid = dim3(static_cast<unsigned>(std::min<long long>(work, sms)), 1, 1);  // persistent CTAs
  } else {
    op.grid = op.v2 ? dim3((d.Tq + 255) / 256, d.heads, d.B) : dim3((d.Tq + 127) / 128, d.heads, d.B);
  }
  op.flops = 4.0 * d.B * d.heads * static_cast<double>(d.Tq) * d.Tk * d.D;
  return op;
}

template <int D, int BKV>
static void attn_launch_d(const AttnOp& op, cudaStream_t s) {
  static bool configured[64] = {};
  using Cfg = AttnCfg<D, BKV>;
  if (first_use_on_this_device(configured)) {
    DM_CUDA(cudaFuncSetAttribute(attention_kernel<D, BKV>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
  }
  attention_kernel<D, BKV><<<op.grid, 128, Cfg::SMEM_BYTES, s>>>(op.maps, op.p);
  DM_CUDA(cudaGetLastError());
}
template <int D, int BKV, int ST>
static void attn2_launch_d(const AttnOp& op, cudaStream_t s) {
  static bool configured[64] = {};
  using Cfg = Attn2Cfg<D, BKV, ST>;
  if (first_use_on_this_device(configured)) {
    DM_CUDA(cudaFuncSetAttribute(attention2_kernel<D, BKV, ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
  }
  attention2_kernel<D, BKV, ST><<<op.grid, Cfg::THREADS, Cfg::SMEM_BYTES, s>>>(op.maps, op.p);
  DM_CUDA(cudaGetLastError());
}
template <int D, int BKV, int ST, int TPR>
static void attn3_launch_d(const AttnOp& op, cudaStream_t s) {
  static bool configured[64] = {};
  using Cfg = Attn3Cfg<D, BKV, ST, TPR>;
  if (first_use_on_this_device(configured)) {
    DM_CUDA(cudaFuncSetAttribute(attention3_kernel<D, BKV, ST, TPR>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
  }
  attention3_kernel<D, BKV, ST, TPR><<<op.grid, Cfg::THREADS, Cfg::SMEM_BYTES, s>>>(op.maps, op.p);
  DM_CUDA(cudaGetLastError());
}
template <int D>
static void xattn_launch_d(const AttnOp& op, cudaStream_t s) {
  static bool configured[64] = {};
  using Cfg = XAttnCfg<D>;
  if (first_use_on_this_device(configured)) {
    DM_CUDA(cudaFuncSetAttribute(xattention_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
  }
  xattention_kernel<D><<<op.grid, 128, Cfg::SMEM_BYTES, s>>>(op.maps, op.p);
  DM_CUDA(cudaGetLastError());
}
template <int D>
static void xattn2_launch_d(const AttnOp& op, cudaStream_t s) {
  static bool configured[64] = {};
  using Cfg = XAttn2Cfg<D>;
  if (first_use_on_this_device(configured)) {
    DM_CUDA(cudaFuncSetAttribute(xattention2_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
  }
  int dev = 0, sms = 0;
  DM_CUDA(cudaGetDevice(&dev));
  DM_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int nq = (op.p.Tq + 127) / 128;
  const long long total_q = static_cast<long long>(nq) * op.p.B;  // (batch, query block) pairs; every CTA serves one head
  DM_CHECK(total_q < (1ll << 31), "cross-attention: too many query blocks");
  const long long ranges = std::max<long long>(1, std::min<long long>(total_q, static_cast<long long>(sms) * Cfg::CTAS_PER_SM / op.p.heads));
  const unsigned grid = static_cast<unsigned>(ranges * op.p.heads);
  xattention2_kernel<D><<<grid, 128, Cfg::SMEM_BYTES, s>>>(op.maps, op.p, nq, static_cast<int>(total_q));
  DM_CUDA(cudaGetLastError());
}
static void vattn_launch(const AttnOp& op, cudaStream_t s) {
  static bool configured[64] = {};
  if (first_use_on_this_device(configured)) {
    DM_CUDA(cudaFuncSetAttribute(vattention_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, VAttnCfg::SMEM_BYTES));
  }
  vattention_kernel<2><<<op.grid, 128, VAttnCfg::SMEM_BYTES, s>>>(op.maps, op.p);
  DM_CUDA(cudaGetLastError());
}
void attn_launch(const AttnOp& op, cudaStream_t s) {
  if (op.vattn) {
    vattn_launch(op, s);
    return;
  }
  if (op.xattn) {
    // xattn = 2 (default): persistent kernel at head_dim 40 (0.150 vs 0.185 ms per 64x64 layer at Bf 54); at head_dim 80 it
    // holds two CTAs per SM and measured 10 % slower than one CTA per query block, so it is only used there with xattn = 3
    // (tests).  xattn = 1: one CTA per query block everywhere.
    const int xv = variant_xattn();
    switch (op.D) {
      case 40: xv >= 2 ? xattn2_launch_d<40>(op, s) : xattn_launch_d<40>(op, s); break;
      case 80: xv >= 3 ? xattn2_launch_d<80>(op, s) : xattn_launch_d<80>(op, s); break;
      case 160: xattn_launch_d<160>(op, s); break;
      default: DM_CHECK(false, "attention: unsupported head_dim");
    }
    return;
  }
  if (op.v3) {
    // attn3 = 1: two softmax threads per query row (16 softmax warps, default); attn3 = 2: one thread per row (8 softmax
    // warps, 168 registers): equal at head_dim 80, 14 % slower at head_dim 40 (gpurun_out/r02_ab_attn5.log)
    const bool one = variant_attn3() == 2;
    if (op.D == 40) one ? attn3_launch_d<40, 128, 2, 1>(op, s) : attn3_launch_d<40, 128, 2, 2>(op, s);
    else one ? attn3_launch_d<80, 64, 3, 1>(op, s) : attn3_launch_d<80, 64, 3, 2>(op, s);
    return;
  }
  if (op.v2) {
    if (op.D == 40) attn2_launch_d<40, 128, 2>(op, s);
    else attn2_launch_d<80, 64, 3>(op, s);
    return;
  }
  switch (op.D) {
    case 40: attn_launch_d<40, 128>(op, s); break;
    case 80: attn_launch_d<80, 64>(op, s); break;
    case 160: attn_launch_d<160, 64>(op, s); break;
    default: DM_CHECK(false, "attention: unsupported head_dim");
  }
}

// ------------------------------------------------------------------ norms / misc
static int grid_for(long long total, int block = 256, int cap = 148 * 16) {
  long long g = (total + block - 1) / block;
  return static_cast<int>(std::max<long long>(1, std::min<long long>(g, cap)));
}

struct GnGeom {
  int VT, R, threads, px_stats, splits, px_apply, blocks_apply;
};
// The geometry depends on (HW, C) only: the partial-sum grouping (hence every output bit) must not change with the
// batch size, so an image scores identically alone, inside any micro-batch, and on any rank.
static GnGeom gn_geom(int HW, int C) {
  GnGeom g;
  g.VT = C / 8;
  g.R = std::max(1, 384 / g.VT);
  g.threads = std::max(256, (g.VT * g.R + 31) / 32 * 32);
  g.px_stats = std::min(64, std::max(16, HW / 16));
  if ((HW + g.px_stats - 1) / g.px_stats > 64) g.px_stats = (HW + 63) / 64;
  g.splits = (HW + g.px_stats - 1) / g.px_stats;
  g.px_apply = std::min(128, std::max(16, HW / 16));
  g.blocks_apply = (HW + g.px_apply - 1) / g.px_apply;
  return g;
}
int gn_splits(int Nimg, int HW) {
  (void)Nimg;
  return gn_geom(HW, 320).splits;  // C does not enter the split count
}

static bool gn_use_fused(int HW, int C) {
  return variant_gn_fused() && static_cast<long long>(HW) * C * 2 <= (12ll << 20);
}
int gn_launch_count(int HW, int C) { return gn_use_fused(HW, C) ? 1 : 2; }

void gn_launch(const GnDesc& d, cudaStream_t s) {
  const int C = d.C0 + d.C1;
  DM_CHECK(C % 32 == 0 && d.C0 % 8 == 0 && d.C1 % 8 == 0, "groupnorm: channel counts must be multiples of 8 / 32");
  DM_CHECK(C <= 3072, "groupnorm: more than 3072 channels");
  DM_CHECK(d.Nimg <= 65535, "groupnorm: more than 65535 images in one call");
  DM_CHECK(d.partial && d.ab && d.tickets, "groupnorm: missing scratch");
  const int cpg = C / 32;
  const GnGeom g = gn_geom(d.HW, C);
  NormSrc s0{d.src0, d.C0, d.ps0}, s1{d.src1, d.C1, d.ps1};
  const size_t smem = gn_smem_floats(g.VT, g.R) * sizeof(float);
  // images that fit in L2 comfortably: one fused kernel, one cluster per image (1 HBM read + 1 write)
  if (gn_use_fused(d.HW, C)) {
    // cluster size depends on HW only (batch invariance); 16 = non-portable size, allowed on sm_100
    static const int cl_max = env_or("DM_GN_CLUSTER", 8);  // 16 measured 8 % slower (r01)
    int CL = d.HW >= 2048 ? 16 : d.HW >= 512 ? 8 : d.HW >= 256 ? 4 : d.HW >= 128 ? 2 : 1;
    CL = std::min(CL, cl_max);
    static const int gn_var = env_or("DM_GN_VAR", 2);  // (CTAs per SM the registers are capped for, loads in flight): 0 = (1, 8), 1 = (3, 4), 2 = (2, 8) [default: -4.6 % measured], 3 = (3, 6)
    const int px_per = (d.HW + CL - 1) / CL;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(CL, d.Nimg, 1);
    cfg.blockDim = dim3(g.threads, 1, 1);
    cfg.dynamicSmemBytes = std::max<size_t>(smem, 256);
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    auto go = [&](auto kernel) {
      static bool configured[64] = {};
      if (first_use_on_this_device(configured)) DM_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
      DM_CUDA(cudaLaunchKernelEx(&cfg, kernel, s0, s1, d.HW, cpg, px_per, g.VT, g.R, d.gamma, d.beta, d.eps, d.silu, d.out));
    };
    switch (gn_var) {
      case 1: go(gn_fused_kernel<3, 4>); break;
      case 2: go(gn_fused_kernel<2, 8>); break;
      case 3: go(gn_fused_kernel<3, 6>); break;
      default: go(gn_fused_kernel<1, 8>); break;
    }
    return;
  }
  gn_stats_kernel<<<dim3(g.splits, d.Nimg), g.threads, std::max<size_t>(smem, 256), s>>>(
      s0, s1, d.HW, cpg, g.splits, g.px_stats, g.VT, g.R, d.partial, d.gamma, d.beta, d.eps,
      reinterpret_cast<float2*>(d.ab), d.tickets);
  DM_CUDA(cudaGetLastError());
  gn_apply_kernel<<<dim3(g.blocks_apply, d.Nimg), g.threads, 0, s>>>(s0, s1, d.HW, g.px_apply, g.VT, g.R,
                                                                    reinterpret_cast<const float2*>(d.ab), d.silu, d.out);
  DM_CUDA(cudaGetLastError());
}

bool gn_fold_apply_supported(int HW, int C0, int C1) {
  const int C = C0 + C1;
  return HW % 128 == 0 && C % 32 == 0 && C0 % 8 == 0 && C1 % 8 == 0 && C <= GN_FOLD_MAXC && C / 8 <= 384;
}

void gn_fold_apply_launch(const GnDesc& d, const float* rec0, const float* rec1, cudaStream_t s) {
  const int C = d.C0 + d.C1;
  DM_CHECK(gn_fold_apply_supported(d.HW, d.C0, d.C1) && d.Nimg <= 65535 && rec0 && (d.C1 == 0 || (rec1 && d.src1)) &&
               d.ps0 == d.C0 && (d.C1 == 0 || d.ps1 == d.C1), "groupnorm fold+apply: bad arguments");
  // (threads, CTAs per SM the registers are capped for, loads in flight, cluster size): the geometry is a function of the
  // variant and of C alone, never of the batch
  static const int var = env_or("DM_GNFA_VAR", 2);
  static const int cl_env = env_or("DM_GNFA_CL", 8);
  const int cl_base = (cl_env == 4 || cl_env == 16) ? cl_env : 8;
  const int E = d.HW / 128 * 4;
  const int VT = C / 8;
  GnStatSrc s0{d.src0, rec0, d.C0}, s1{d.src1, rec1, d.C1};
  auto go = [&](auto kernel, int threads) {
    int CL = cl_base;
    // a CTA folds C / CL channels with one thread per (slice, channel) and reduces its 32 / CL groups with one warp each
    while (C / CL > threads || 32 / CL > threads / 32) CL *= 2;
    DM_CHECK(CL <= 16 && VT <= threads, "groupnorm fold+apply: too many channels for this variant");
    static bool configured[64] = {};
    if (first_use_on_this_device(configured)) DM_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaLaunchConfig_t cfg{};
    // clusters per image: a function of HW alone (batch invariance); 128x128 latents come few per micro-batch
    const int splits = d.HW >= 16384 ? 4 : d.HW >= 8192 ? 2 : 1;
    cfg.gridDim = dim3(CL, d.Nimg, splits);
    cfg.blockDim = dim3(threads, 1, 1);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    const int R = std::max(1, threads / VT);
    const int px_per = (d.HW + CL * splits - 1) / (CL * splits);
    DM_CUDA(cudaLaunchKernelEx(&cfg, kernel, s0, s1, d.HW, C / 32, px_per, VT, R, E, d.gamma, d.beta, d.eps, d.silu, d.out));
  };
  if (VT > 256) {  // 2560 channels (up_blocks.1 at 16x16): one thread per channel vector needs the 384-thread build
    go(gn_fold_apply_kernel<384, 2, 8>, 384);
    return;
  }
  switch (var) {
    case 1: go(gn_fold_apply_kernel<256, 3, 8>, 256); break;
    case 4: go(gn_fold_apply_kernel<256, 5, 4>, 256); break;
    case 0: go(gn_fold_apply_kernel<384, 2, 8>, 384); break;
    default: go(gn_fold_apply_kernel<256, 4, 6>, 256); break;
  }
}

void layernorm_launch(const __half* x, long long ld_x, const float* gamma, const float* beta, float eps, long long rows,
                      int C, __half* out, long long ld_out, cudaStream_t s) {
  DM_CHECK(C % 8 == 0 && C <= 1280, "layernorm: C must be a multiple of 8 and <= 1280");
  // persistent warps: 8 per block, exactly as many blocks as are co-resident (a second wave would start late and
  // leave the tail of the row range to half the machine)
  const long long want = (rows + 7) / 8;
  const int maxv = (C / 8 + 31) / 32;
  const size_t smem = 2 * static_cast<size_t>(C) * sizeof(float);
  auto launch = [&](auto kernel, int& occ_cache) {
    if (occ_cache == 0) {
      DM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_cache, kernel, 256, smem));
      occ_cache = std::max(1, occ_cache);
    }
    const unsigned blocks = static_cast<unsigned>(std::max<long long>(1, std::min<long long>(want, 148ll * occ_cache)));
    kernel<<<blocks, 256, smem, s>>>(x, ld_x, gamma, beta, eps, rows, C, out, ld_out);
  };
  static int occ2 = 0, occ3 = 0, occ5 = 0;
  // the three Transformer2DModel widths: every lane busy (LPR lanes x 5 vectors per row)
  static const int ln_var = env_or("DM_LN_VAR", 2);
  if (ln_var && (C == 320 || C == 640 || C == 1280)) {
    auto launch5 = [&](auto kernel, int rpw, int& occ_cache, size_t dyn_smem) {
      if (occ_cache == 0) {
        DM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_cache, kernel, 256, dyn_smem));
        occ_cache = std::max(1, occ_cache);
      }
      const long long want5 = (rows + 8 * rpw - 1) / (8 * rpw);
      const unsigned blocks = static_cast<unsigned>(std::max<long long>(1, std::min<long long>(want5, 148ll * occ_cache)));
      kernel<<<blocks, 256, dyn_smem, s>>>(x, ld_x, gamma, beta, eps, rows, out, ld_out);
    };
    static int o8 = 0, o16 = 0, o32 = 0, p8 = 0, p16 = 0, p32 = 0;
    if (ln_var == 2) {  // default: occupancy instead of a register prefetch, fp16 gamma / beta in (static) shared memory
      if (C == 320) launch5(layernorm5b_kernel<8, 4>, 4, p8, 0);
      else if (C == 640) launch5(layernorm5b_kernel<16, 4>, 2, p16, 0);
      else launch5(layernorm5b_kernel<32, 4>, 1, p32, 0);
    } else {
      if (C == 320) launch5(layernorm5_kernel<8>, 4, o8, smem);
      else if (C == 640) launch5(layernorm5_kernel<16>, 2, o16, smem);
      else launch5(layernorm5_kernel<32>, 1, o32, smem);
    }
    DM_CUDA(cudaGetLastError());
    return;
  }
  if (maxv <= 2) launch(layernorm_kernel<2>, occ2);
  else if (maxv == 3) launch(layernorm_kernel<3>, occ3);
  else launch(layernorm_kernel<5>, occ5);
  DM_CUDA(cudaGetLastError());
}

void patch3x3_launch(const float* x0, const int* x_index, const float* noise, const int* noise_index, const long long* t,
                     const float* ca, const float* cb, int Bf, int Cin, int H, int W, __half* out, cudaStream_t s,
                     int index_stride, int sched_n, int* err_flag) {
  DM_CHECK(9 * Cin <= 64, "patch3x3: Cin too large");
  patch3x3_kernel<<<grid_for(static_cast<long long>(Bf) * H * W * 8), 256, 0, s>>>(x0, x_index, noise, noise_index, t, ca,
                                                                                   cb, sched_n, err_flag, Bf, Cin, H, W,
                                                                                   index_stride, out);
  DM_CUDA(cudaGetLastError());
}
void repeat_rows_launch(const __half* in, long long rows, long long row_elems, int G, __half* out, cudaStream_t s) {
  DM_CHECK(row_elems % 8 == 0 && G >= 1, "repeat_rows: row length must be a multiple of 8");
  repeat_rows_kernel<<<grid_for(rows * (row_elems / 8)), 256, 0, s>>>(reinterpret_cast<const uint4*>(in), rows, row_elems / 8, G,
                                                                     reinterpret_cast<uint4*>(out));
  DM_CUDA(cudaGetLastError());
}
void timestep_embed_launch(const long long* t, const int* t_index, int Bf, __half* out, cudaStream_t s) {
  timestep_embed_kernel<<<Bf, 160, 0, s>>>(t, t_index, Bf, out);
  DM_CUDA(cudaGetLastError());
}
void upsample_nearest_launch(const __half* in, int N, int H, int W, int C, int Ho, int Wo, __half* out, cudaStream_t s) {
  upsample_nearest_kernel<<<grid_for(static_cast<long long>(N) * Ho * Wo * (C / 8)), 256, 0, s>>>(in, N, H, W, C, Ho, Wo,
                                                                                                  out);
  DM_CUDA(cudaGetLastError());
}
void space_to_planes_launch(const __half* in, int N, int H, int W, int C, int H2, int W2, __half* out, cudaStream_t s) {
  space_to_planes_kernel<<<grid_for(4ll * N * H2 * W2 * (C / 8)), 256, 0, s>>>(in, N, H, W, C, H2, W2, out);
  DM_CUDA(cudaGetLastError());
}
void loss_launch(const __half* pred, int ld_pred, const float* noise, const int* noise_index, const int* grid_row,
                 float* loss_f32, __half* grid_f16, float* eps_f32, int Bf, int HW, cudaStream_t s) {
  LossMap mp{noise_index, grid_row, loss_f32, grid_f16, eps_f32};
  loss_kernel<<<grid_for(static_cast<long long>(Bf) * HW), 256, 0, s>>>(pred, ld_pred, noise, mp, Bf, HW);
  DM_CUDA(cudaGetLastError());
}
void tmap_launch(const __half* grid, int Bi, int N, int n_cond, int HW, float* T, cudaStream_t s) {
  tmap_kernel<<<grid_for(static_cast<long long>(Bi) * (n_cond - 1) * HW), 256, 0, s>>>(grid, Bi, N, n_cond, HW, T);
  DM_CUDA(cudaGetLastError());
}
void vae_sample_launch(const __half* h16, int ld_h, const __half* wq, const float* bq, const float* eps, float scaling,
                       int B, int HW, float* z, float* mean_out, float* logvar_out, cudaStream_t s) {
  vae_sample_kernel<<<grid_for(static_cast<long long>(B) * HW), 256, 0, s>>>(h16, ld_h, wq, bq, eps, scaling, B, HW, z,
                                                                            mean_out, logvar_out);
  DM_CUDA(cudaGetLastError());
}
void softmax_rows_launch(const float* S, long long ld_s, int rows, int cols, float scale, __half* P, long long ld_p,
                         cudaStream_t s) {
  softmax_rows_kernel<<<rows, 256, 0, s>>>(S, ld_s, cols, scale, P, ld_p);
  DM_CUDA(cudaGetLastError());
}
void nhwc_to_nchw_mean_launch(const __half* in, int B, int E, int HW, int C, float* out, cudaStream_t s) {
  nhwc_to_nchw_mean_kernel<<<dim3((HW + 31) / 32, (C + 31) / 32, B), 256, 0, s>>>(in, B, E, HW, C, out);
  DM_CUDA(cudaGetLastError());
}

}  // namespace dm
