// Implicit-GEMM on tcgen05/TMEM: the one dense-contraction kernel of the engine.
//
//   out[m, n] = epilogue( sum_k A[m, k] * Wt[n, k] )          fp16 x fp16 -> fp32 (TMEM) -> fp16/fp32
//
// A is never materialised: rows m are output pixels (img, y, x) of an NHWC activation tensor, and the K
// axis is a list of "segments" (source tensor, pixel offset (dn,dy,dx), first channel, #64-channel chunks).
// A 3x3 conv is 9 segments (one per tap), a 1x1 conv / Linear is 1 segment, a conv over cat(h, skip) uses
// two source tensors, a stride-2 conv addresses the four parity planes produced by space_to_planes().
// Each (segment, chunk) is ONE 4-D TMA box [nt][ht][wt][64ch] whose out-of-bounds part is zero-filled by
// the TMA unit -- that is the conv zero padding -- landing in shared memory directly in the 128B-swizzled
// K-major layout tcgen05.mma consumes.  Weights Wt[N, K] (K ordered like the segment list) arrive by 2-D TMA.
//
// Persistent CTAs, warp-specialised: warp0 = TMA producer, warp1 = MMA issuer (+TMEM alloc),
// warps 2..9 = two epilogue warpgroups.  Two TMEM accumulator buffers let the epilogue of tile i overlap the
// main loop of tile i+1.
//
// Epilogue (staged): the accumulator tile is drained in chunks of 32 output columns, alternating between the two
// warpgroups.  A chunk lives in a 128 x 32 fp16 shared-memory buffer (64-byte swizzle) of the warpgroup's ring:
// the residual tile is PREFETCHED into it by TMA (several chunks ahead, so its HBM latency is never exposed), each
// thread (= one accumulator row) combines TMEM + bias + row-bias + activation + residual in packed fp16 arithmetic
// (HADD2 / HMUL2 round exactly like the reference's fp16 tensor adds), writes the result back in place, and one
// elected thread hands the buffer to a TMA store, which clips the M / N tails.  No per-thread strided global
// access remains.  The "direct" variant (fp32 output, N-tiles narrower than 32) keeps plain global stores.
#pragma once
#include "ptx.cuh"

namespace dm {

constexpr int IG_MAX_SEG = 24;
constexpr int IG_BM = 128;
constexpr int IG_BK = 64;
constexpr int IG_CW = 32;                           // output columns per epilogue chunk
constexpr int IG_CHUNK_BYTES = IG_BM * IG_CW * 2;   // 8 KB
constexpr int IG_THREADS = 64 + 256;   // default: two epilogue warpgroups (NG = 2)

struct IgSeg {
  int16_t src, dy, dx, pad_;
  int32_t dn, chan0, nchunks;
};

struct alignas(64) IgMaps {
  CUtensorMap a[2];
  CUtensorMap b;
  CUtensorMap c;  // output   [Nimg, H, W, ld_out], box [nt][ht][wt][32], 64B swizzle (staged epilogue only)
  CUtensorMap r;  // residual [Nimg, H, W, ld_res], same box
};

// Typicality epilogue of conv_out (reference: F.mse_loss(noise_pred.float(), noise, 'none') at
// /root/reference/diffmining/typicality/compute.py:101, stored as fp16 [.., 4, h, w] rows, compute.py:155-160): lives in
// DEVICE memory because the pointers change per micro-batch while the prepared launch (and its CUDA graph) does not.
struct IgLossArgs {
  const float* noise;      // [*, 4, H, W] fp32 draws
  const int* noise_index;  // [Nimg] forward row -> draw
  const int* grid_row;     // [Nimg] forward row -> row of the raw grid, or null (identity offset 0)
  __half* grid_f16;        // [rows, 4, H, W] fp16; null = epilogue disabled for this replay
};

// GroupNorm statistics of the OUTPUT tensor, formed in the producing conv's epilogue (the following GroupNorm then only has
// to fold them and apply y = x * a + b: reference op order conv1 -> +temb -> norm2 -> SiLU,
// applications/parallel-dataset/pnp.py:318-345).  After a chunk (128 rows x 32 columns of final fp16 values) is staged in
// shared memory for its TMA store, the 128 threads of the epilogue group re-read it COLUMN-wise (thread = column pair x
// 16 alternate rows of its warp's 32; conflict-free on the 64-byte-swizzled chunk) and store the per-column shifted sums
// S = sum(x - k), Q = sum((x - k)^2) per (image, m-tile, warp quarter, channel); k = bias + time-embedding bias + the
// residual's value at the image's first pixel (known to every tile of the image, so the sums are additive) is written once
// per (image, channel) by the image's first tile.  No atomics, no completion protocol: gn_fold_apply_kernel folds the
// partials in a fixed order.  Requires whole 128-pixel tiles per image (conv layout: nt_log == 0; flattened Linear layout:
// H * W % 128 == 0) and N a multiple of the N-tile.
struct IgGn {
  float* rec;           // per image: [E = tiles_img * 4][N][2] (S, Q) partials, then [N] shifts k -- (2E + 1) * N floats
  int tiles_img;        // m-tiles per image (m-tile mt belongs to image mt / tiles_img)
};

struct IgParams {
  int Nimg, H, W;                 // OUTPUT pixel grid; M = Nimg*H*W
  int wt_log, ht_log, nt_log;     // M-tile = 2^nt images x 2^ht rows x 2^wt cols (=128 pixels)
  int tiles_x, tiles_y, tiles_n;  // ceil-div tile counts
  int m_tiles, n_tiles;
  int N;                          // GEMM N (weight rows)
  int nseg, k_iters;
  IgSeg seg[IG_MAX_SEG];
  const float* bias;              // [N] fp32 or null
  const __half* rowbias;          // [Nimg, ld_rowbias] per-image bias (time-embedding projection) or null
  int ld_rowbias;
  const __half* residual;         // [M, ld_res] or null
  long long ld_res;
  void* out;                      // [M, ld_out] fp16 (or fp32 when out_f32)
  long long ld_out;
  int out_f32, geglu, act_silu;
  const IgLossArgs* loss;         // direct epilogue only: fused (pred - eps)^2 -> fp16 grid (null = off)
  IgGn gn;                        // staged epilogue only: GroupNorm statistics of the output (gn.rec == null = off)
};

// CG = CTAs per tile: 1, or 2 = a CTA pair (cta_group::2) computing a 256 x BN tile with each CTA holding its 128
// rows of A and HALF of the B tile, which cuts the per-SM shared-memory operand traffic and the L2 -> SM weight
// traffic by the B half.
// NG = epilogue warpgroups (2, or 4 for the short-K layers whose epilogue, not the MMA, bounds the tile time).
// WS = weight-stationary (CTA pairs, K <= 64 * IG_WS_KCHUNKS): the unit's B tile (all k-chunks of ONE N-tile, half per CTA)
// is loaded once and stays in shared memory while the unit walks down the M dimension; the pipeline stages carry A only.
// The K = 320 Linears at 64x64 are bound by L2 -> SM traffic (the LTS throughput cap), 55 % of which is the weight tile
// being re-fetched for every output tile.
constexpr int IG_WS_KCHUNKS = 5;
template <int BN, bool DIRECT, int CG = 1, int NG = 2, bool WS = false>
struct IgCfg {
  static constexpr int THREADS = 64 + 128 * NG;
  static constexpr int A_BYTES = IG_BM * IG_BK * 2;
  static constexpr int B_BYTES = (BN / CG) * IG_BK * 2;
  static_assert(CG == 1 || (CG == 2 && !DIRECT && BN % 32 == 0), "CTA pairs: staged epilogue, BN multiple of 32");
  static_assert(!WS || (CG == 2 && !DIRECT), "weight-stationary: CTA pairs, staged epilogue");
  static constexpr int W_BYTES = WS ? IG_WS_KCHUNKS * B_BYTES : 0;  // resident weight chunks
  static constexpr int STAGE_BYTES = WS ? A_BYTES : A_BYTES + B_BYTES;
#ifdef IG_RING_SMALL  // experiment: two chunk buffers per group everywhere (one more pipeline stage at BN = 160)
  static constexpr int NBG = DIRECT ? 0 : 2;
#else
  static constexpr int NBG = DIRECT ? 0 : (BN >= 256 || NG > 2) ? 2 : 4;  // chunk buffers per epilogue warpgroup
#endif
  static constexpr int LOOKAHEAD = NBG >= 4 ? 2 : 1;             // residual prefetch distance (chunks)
  static constexpr int RING_BYTES = NG * NBG * IG_CHUNK_BYTES;
  static constexpr int BIAS_BYTES = NG * BN * 4;
  static constexpr int BAR_BYTES = 256;
  static constexpr int RAW_STAGES = (232448 - RING_BYTES - BIAS_BYTES - BAR_BYTES - W_BYTES) / STAGE_BYTES;
  static constexpr int STAGES = RAW_STAGES > 8 ? 8 : RAW_STAGES;
  static constexpr int TMEM_COLS = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + W_BYTES + RING_BYTES + BIAS_BYTES + BAR_BYTES;
  static_assert(STAGES >= 3, "pipeline too shallow");
  static_assert(2 * STAGES + 4 + NG * (NBG > 0 ? NBG : 1) + 2 <= BAR_BYTES / 8, "barrier area");
  static_assert(NG == 2 || (NG == 4 && !DIRECT && (CG == 1 || WS)), "four epilogue groups: staged epilogue, single CTA or weight-stationary pair");
};

// Phi(x) * x with erfc from Abramowitz-Stegun 7.1.26 (|abs err| < 4.3e-7 on the result: within one fp16 ulp of
// the erf formulation everywhere, closer to the exact value than 0.5*x*(1+erff(x/sqrt2)) on the negative tail).
// Branch-free: 11 scalar FMA-pipe ops + 2 MUFU + 1 ALU op.  (Scalar on purpose: on B200 a packed f32x2 FMA issues
// every 3.1 cycles per sub-partition against 1.08 for a scalar FFMA -- tools/microbench/fma.cu -- and the GEGLU
// epilogue is FMA-pipe-bound.)
__device__ __forceinline__ float gelu_fast(float x) {
  const float z = fabsf(x) * 0.70710678118654752f;
  const float t = fast_rcp(fmaf(0.3275911f, z, 1.f));
  float pl = fmaf(t, 0.5f * 1.061405429f, 0.5f * -1.453152027f);
  pl = fmaf(t, pl, 0.5f * 1.421413741f);
  pl = fmaf(t, pl, 0.5f * -0.284496736f);
  pl = fmaf(t, pl, 0.5f * 0.254829592f);
  const float e = fast_exp2((x * -0.72134752044448170f) * x);  // exp(-x^2 / 2)
  const float q = (pl * t) * e;                                 // 0.5 * erfc(|x| / sqrt 2)
  // x * Phi(x) = x * q for x < 0 and x - x * q for x >= 0, i.e. relu(x) - |x| * q for both signs: one FMNMX + one FFMA
  // (a single rounding) instead of a subtract, a compare, a select and a multiply
  return fmaf(-fabsf(x), q, fmaxf(x, 0.f));
}

__device__ __forceinline__ __half2 u2h(uint32_t u) { return *reinterpret_cast<__half2*>(&u); }
__device__ __forceinline__ uint32_t h2u(__half2 h) { return *reinterpret_cast<uint32_t*>(&h); }

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

template <int BN, bool DIRECT, int CG, int NG, bool WS = false>
__global__ void __launch_bounds__(64 + 128 * NG, 1) igemm_kernel(const __grid_constant__ IgMaps maps, const IgParams p) {
  using Cfg = IgCfg<BN, DIRECT, CG, NG, WS>;
  constexpr int STAGES = Cfg::STAGES;
  constexpr int NBG = Cfg::NBG;
  extern __shared__ __align__(1024) uint8_t smem[];  // 128B-swizzled tiles need 1024-byte alignment
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* smA = smem;
  uint8_t* smB = smem + STAGES * Cfg::A_BYTES;
  uint8_t* smC = smem + STAGES * Cfg::STAGE_BYTES + Cfg::W_BYTES;
  float* smBias = reinterpret_cast<float*>(smC + Cfg::RING_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smC + Cfg::RING_BYTES + Cfg::BIAS_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + STAGES;
  uint64_t* tfull = bars + 2 * STAGES;
  uint64_t* tempty = bars + 2 * STAGES + 2;
  uint64_t* rfull = bars + 2 * STAGES + 4;  // [NG groups][NBG]
  uint64_t* wfull = bars + 2 * STAGES + 4 + NG * (NBG > 0 ? NBG : 1);  // weight-stationary: the resident B tile has landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wfull + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int cta_rank = CG == 2 ? static_cast<int>(cluster_ctarank()) : 0;
  const int unit = static_cast<int>(blockIdx.x) / CG;     // tile-processing unit: a CTA, or a CTA pair
  const int nunits = static_cast<int>(gridDim.x) / CG;

  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full[i], 1);   // CG == 2: only the leader's is used (both CTAs' loads complete_tx on it)
      mbar_init(&empty[i], 1);
    }
    mbar_init(&tfull[0], 1);
    mbar_init(&tfull[1], 1);
    mbar_init(&tempty[0], 4 * NG * CG);  // CG == 2: the leader's collects the epilogue warps of both CTAs
    mbar_init(&tempty[1], 4 * NG * CG);
    for (int i = 0; i < NG * (NBG > 0 ? NBG : 1); ++i) mbar_init(&rfull[i], 1);
    mbar_init(wfull, 1);
    fence_barrier_init();
    tma_prefetch_desc(&maps.a[0]);
    tma_prefetch_desc(&maps.b);
    if constexpr (!DIRECT) tma_prefetch_desc(&maps.c);
  }
  if (warp == 1) {
    if constexpr (CG == 2) {
      tmem_alloc_pair(tmem_slot, Cfg::TMEM_COLS);
      tmem_relinquish_pair();
    } else {
      tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all();  // the peer's barriers must be initialised before anything arrives on them
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // tiles are (m-unit, n-tile): an m-unit is CG consecutive 128-row M-tiles (an odd tail unit has a dummy second
  // tile whose loads are zero-filled and whose stores are clipped by the tensor maps)
  const int num_tiles = ((p.m_tiles + CG - 1) / CG) * p.n_tiles;

  if (warp == 0) {
    // ============================== TMA producer ==============================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      if constexpr (WS) {
        // the launcher makes the number of units a multiple of n_tiles, so unit + k * nunits keeps ONE N-tile: its weights
        // (this CTA's half, every k-chunk) are fetched once
        if (unit < num_tiles) {
          const int ntile = unit % p.n_tiles;
          if (cta_rank == 0) mbar_arrive_expect_tx(wfull, 2 * p.k_iters * Cfg::B_BYTES);
          for (int kit = 0; kit < p.k_iters; ++kit)
            tma_load_2d_pair(smB + kit * Cfg::B_BYTES, &maps.b, wfull, kit * IG_BK, ntile * BN + cta_rank * (BN / 2));
        }
      }
      for (int tile = unit; tile < num_tiles; tile += nunits) {
        const int mt = (tile / p.n_tiles) * CG + cta_rank, ntile = tile % p.n_tiles;
        const int tx = mt % p.tiles_x, ty = (mt / p.tiles_x) % p.tiles_y, tn = mt / (p.tiles_x * p.tiles_y);
        const int x0 = tx << p.wt_log, y0 = ty << p.ht_log, n0 = tn << p.nt_log;
        int kit = 0;
        for (int s = 0; s < p.nseg; ++s) {
          const IgSeg sg = p.seg[s];
          const CUtensorMap* am = &maps.a[sg.src];
          for (int c = 0; c < sg.nchunks; ++c, ++kit) {
            mbar_wait(&empty[stage], phase ^ 1);
            if constexpr (WS) {
              if (cta_rank == 0) mbar_arrive_expect_tx(&full[stage], 2 * Cfg::A_BYTES);
              tma_load_4d_pair(smA + stage * Cfg::A_BYTES, am, &full[stage], sg.chan0 + c * IG_BK, x0 + sg.dx,
                               y0 + sg.dy, n0 + sg.dn);
            } else if constexpr (CG == 2) {
              // the leader arms its barrier for the bytes of BOTH CTAs; each CTA loads its A rows and its B half
              if (cta_rank == 0) mbar_arrive_expect_tx(&full[stage], 2 * Cfg::STAGE_BYTES);
              tma_load_4d_pair(smA + stage * Cfg::A_BYTES, am, &full[stage], sg.chan0 + c * IG_BK, x0 + sg.dx,
                               y0 + sg.dy, n0 + sg.dn);
              tma_load_2d_pair(smB + stage * Cfg::B_BYTES, &maps.b, &full[stage], kit * IG_BK,
                               ntile * BN + cta_rank * (BN / 2));
            } else {
              mbar_arrive_expect_tx(&full[stage], Cfg::STAGE_BYTES);
              tma_load_4d(smA + stage * Cfg::A_BYTES, am, &full[stage], sg.chan0 + c * IG_BK, x0 + sg.dx, y0 + sg.dy,
                          n0 + sg.dn);
              tma_load_2d(smB + stage * Cfg::B_BYTES, &maps.b, &full[stage], kit * IG_BK, ntile * BN);
            }
            if (++stage == STAGES) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ============================== MMA issuer ==============================
    if (lane == 0 && cta_rank == 0) {
      constexpr uint32_t idesc = umma_idesc_f16(BN, false, 128 * CG);
      int stage = 0;
      uint32_t phase = 0;
      int lt = 0;
      if constexpr (WS) {
        if (unit < num_tiles) mbar_wait(wfull, 0);
      }
      for (int tile = unit; tile < num_tiles; tile += nunits, ++lt) {
        const int acc = lt & 1;
        const uint32_t acc_phase = (lt >> 1) & 1;
        mbar_wait(&tempty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kit = 0; kit < p.k_iters; ++kit) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint64_t ad = umma_desc_kmajor_sw128(smem_u32(smA + stage * Cfg::A_BYTES));
          const uint64_t bd = umma_desc_kmajor_sw128(smem_u32(smB + (WS ? kit : stage) * Cfg::B_BYTES));
#pragma unroll
          for (int k = 0; k < IG_BK / 16; ++k) {
            // advance 16 fp16 = 32 B along K inside the 128-B swizzle row: +2 in (addr>>4) units
            if constexpr (CG == 2) umma_f16_pair(d_tmem, ad + 2 * k, bd + 2 * k, idesc, (kit | k) != 0 ? 1u : 0u);
            else umma_f16(d_tmem, ad + 2 * k, bd + 2 * k, idesc, (kit | k) != 0 ? 1u : 0u);
          }
          if constexpr (CG == 2) umma_commit_pair(&empty[stage]);
          else umma_commit(&empty[stage]);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        if constexpr (CG == 2) umma_commit_pair(&tfull[acc]);
        else umma_commit(&tfull[acc]);
      }
    }
  } else {
    // ============================== epilogue (2 warpgroups) ==============================
    const int ew = warp - 2;        // 0..7
    const int eg = ew >> 2;         // warpgroup
    const int quarter = warp & 3;   // TMEM lane quarter this warp may access
    const int r = quarter * 32 + lane;
    if constexpr (DIRECT) {
      constexpr int CH = (BN % 32 == 0) ? 32 : 16;
      int lt = 0;
      for (int tile = unit; tile < num_tiles; tile += nunits, ++lt) {
        const int acc = lt & 1;
        const uint32_t acc_phase = (lt >> 1) & 1;
        const int mt = tile / p.n_tiles, ntile = tile % p.n_tiles;
        const int tx = mt % p.tiles_x, ty = (mt / p.tiles_x) % p.tiles_y, tn = mt / (p.tiles_x * p.tiles_y);
        const int x = (tx << p.wt_log) + (r & ((1 << p.wt_log) - 1));
        const int y = (ty << p.ht_log) + ((r >> p.wt_log) & ((1 << p.ht_log) - 1));
        const int n = (tn << p.nt_log) + (r >> (p.wt_log + p.ht_log));
        const bool row_ok = (x < p.W) && (y < p.H) && (n < p.Nimg);
        const long long m = (static_cast<long long>(n) * p.H + y) * p.W + x;

        mbar_wait(&tfull[acc], acc_phase);
        tc_fence_after();
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BN;
#pragma unroll 1
        for (int c0 = eg * CH; c0 < BN; c0 += NG * CH) {
          uint32_t raw[CH];
          if constexpr (CH == 32) tmem_ld_x32(taddr + c0, raw);
          else tmem_ld_x16(taddr + c0, raw);
          tmem_wait_ld();
          const int col0 = ntile * BN + c0;
          int nvalid = p.N - col0;
          nvalid = nvalid > CH ? CH : nvalid;
          if (!row_ok || nvalid <= 0) continue;
          float v[CH];
#pragma unroll
          for (int i = 0; i < CH; ++i) v[i] = __uint_as_float(raw[i]);
          if (p.bias) {
#pragma unroll
            for (int i = 0; i < CH; i += 4) {
              if (i < nvalid) {
                const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + i));
                v[i] += b.x; v[i + 1] += b.y; v[i + 2] += b.z; v[i + 3] += b.w;
              }
            }
          }
          if (p.out_f32) {
            float* o = reinterpret_cast<float*>(p.out) + m * p.ld_out + col0;
#pragma unroll
            for (int i = 0; i < CH; i += 4)
              if (i < nvalid) *reinterpret_cast<float4*>(o + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
            continue;
          }
#pragma unroll
          for (int i = 0; i < CH; ++i) v[i] = round_h(v[i]);
          if (p.loss != nullptr && col0 == 0) {
            // conv_out of the typicality path: loss = (float(fp16 pred) - eps)^2 straight into the raw fp16 grid; threads of a
            // warp hold consecutive pixels, so every channel plane is written with coalesced 2-byte stores
            const IgLossArgs la = *p.loss;
            if (la.grid_f16 != nullptr) {
              const long long HW = static_cast<long long>(p.H) * p.W, px = static_cast<long long>(y) * p.W + x;
              const long long nrow = la.noise_index ? la.noise_index[n] : n, grow = la.grid_row ? la.grid_row[n] : n;
#pragma unroll
              for (int c = 0; c < 4; ++c) {
                const float dlt = v[c] - __ldg(la.noise + (nrow * 4 + c) * HW + px);
                la.grid_f16[(grow * 4 + c) * HW + px] = __float2half_rn(dlt * dlt);
              }
            }
          }
          if (p.rowbias) {
            const __half* rb = p.rowbias + static_cast<long long>(n) * p.ld_rowbias + col0;
#pragma unroll
            for (int i = 0; i < CH; i += 8) {
              if (i < nvalid) {
                const uint4 q = __ldg(reinterpret_cast<const uint4*>(rb + i));
                const __half2* h = reinterpret_cast<const __half2*>(&q);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const float2 f = __half22float2(h[j]);
                  v[i + 2 * j] = round_h(v[i + 2 * j] + f.x);
                  v[i + 2 * j + 1] = round_h(v[i + 2 * j + 1] + f.y);
                }
              }
            }
          }
          if (p.act_silu) {
#pragma unroll
            for (int i = 0; i < CH; ++i) v[i] = round_h(silu_f(v[i]));
          }
          if (p.residual) {
            const __half* rs = p.residual + m * p.ld_res + col0;
#pragma unroll
            for (int i = 0; i < CH; i += 8) {
              if (i < nvalid) {
                const uint4 q = *reinterpret_cast<const uint4*>(rs + i);
                const __half2* h = reinterpret_cast<const __half2*>(&q);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const float2 f = __half22float2(h[j]);
                  v[i + 2 * j] += f.x;
                  v[i + 2 * j + 1] += f.y;
                }
              }
            }
          }
          __half* o = reinterpret_cast<__half*>(p.out) + m * p.ld_out + col0;
#pragma unroll
          for (int i = 0; i < CH; i += 8) {
            if (i < nvalid) {
              *reinterpret_cast<uint4*>(o + i) = make_uint4(pack_h2(v[i], v[i + 1]), pack_h2(v[i + 2], v[i + 3]),
                                                            pack_h2(v[i + 4], v[i + 5]), pack_h2(v[i + 6], v[i + 7]));
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[acc]);
      }
    } else {
      // ---------------- staged epilogue ----------------
      const bool leader = (ew & 3) == 0 && lane == 0;
      const int gtid = (ew & 3) * 32 + lane;
      const int bar_id = 1 + eg;
      uint8_t* ring = smC + eg * NBG * IG_CHUNK_BYTES;
      float* sbias = smBias + eg * BN;
      uint64_t* rf = rfull + eg * NBG;
      const bool has_res = p.residual != nullptr;
      const int acc_per_chunk = p.geglu ? 2 * IG_CW : IG_CW;
      const int nchunk = BN / acc_per_chunk;
      const int out_tile_cols = p.geglu ? BN / 2 : BN;
      // byte offset of 16-byte piece j of this thread's row inside a chunk buffer (64B swizzle: Swizzle<2,4,3>)
      const uint32_t row_off = static_cast<uint32_t>(r) * 64u;
      const uint32_t sw = (static_cast<uint32_t>(r) >> 1) & 3u;

      // iterator over this warpgroup's chunks: chunk c of local tile lt belongs to group (lt*nchunk + c) % NG
      struct It {
        int tile, lt, c;
      };
      auto first_c = [&](int lt_) { return ((eg - lt_ * nchunk) % NG + NG) % NG; };
      auto settle = [&](It& it) {  // skip tiles in which this group owns no chunk
        while (it.tile < num_tiles && it.c >= nchunk) {
          it.tile += nunits;
          it.lt += 1;
          it.c = first_c(it.lt);
        }
      };
      auto advance = [&](It& it) {
        it.c += NG;
        settle(it);
      };
      auto coords = [&](const It& it, int& col, int& x0, int& y0, int& n0) {
        const int mt = (it.tile / p.n_tiles) * CG + cta_rank, ntile = it.tile % p.n_tiles;
        const int tx = mt % p.tiles_x, ty = (mt / p.tiles_x) % p.tiles_y, tn = mt / (p.tiles_x * p.tiles_y);
        x0 = tx << p.wt_log; y0 = ty << p.ht_log; n0 = tn << p.nt_log;
        col = ntile * out_tile_cols + it.c * IG_CW;
      };
      It pf{unit, 0, first_c(0)};  // residual prefetch cursor (leader thread only)
      settle(pf);
      int pf_k = 0;
      auto prefetch_one = [&]() {
        if (pf.tile < num_tiles) {
          int col, x0, y0, n0;
          coords(pf, col, x0, y0, n0);
          const int b = pf_k % NBG;
          mbar_arrive_expect_tx(&rf[b], IG_CHUNK_BYTES);
          tma_load_4d(ring + b * IG_CHUNK_BYTES, &maps.r, &rf[b], col, x0, y0, n0);
        }
        ++pf_k;
        advance(pf);
      };
      if (leader && has_res) {
        tma_prefetch_desc(&maps.r);
        for (int i = 0; i < Cfg::LOOKAHEAD; ++i) prefetch_one();
      }

      int k = 0;  // chunks processed by this warpgroup
      int lt = 0;
      for (int tile = unit; tile < num_tiles; tile += nunits, ++lt) {
        const int acc = lt & 1;
        const uint32_t acc_phase = (lt >> 1) & 1;
        const int ntile = tile % p.n_tiles;
        const int mt = (tile / p.n_tiles) * CG + cta_rank;  // this CTA's M-tile of the unit
        if (p.bias) {
          for (int i = gtid; i < BN; i += 128) {
            const int col = ntile * BN + i;
            sbias[i] = col < p.N ? __ldg(p.bias + col) : 0.f;
          }
          named_bar_sync(bar_id, 128);
        }
        const __half* rb = nullptr;
        if (p.rowbias) {
          const int tn = mt / (p.tiles_x * p.tiles_y);
          int n = (tn << p.nt_log) + (r >> (p.wt_log + p.ht_log));
          n = n < p.Nimg ? n : p.Nimg - 1;
          rb = p.rowbias + static_cast<long long>(n) * p.ld_rowbias + ntile * BN;
        }
        mbar_wait(&tfull[acc], acc_phase);
        tc_fence_after();
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BN;
        const int c_first = first_c(lt);
#pragma unroll 1
        for (int c = c_first; c < nchunk; c += NG, ++k) {
          const int b = k % NBG;
          uint8_t* buf = ring + b * IG_CHUNK_BYTES + row_off;
          // fused GroupNorm statistics (see IgGn): this thread's two columns' shift, fetched ahead of its use
          float k0 = 0.f, k1 = 0.f;
          int n_img = 0, t_img = 0;
          if (p.gn.rec != nullptr && mt < p.m_tiles) {
            n_img = mt / p.gn.tiles_img;
            t_img = mt % p.gn.tiles_img;
            const int col = ntile * BN + c * IG_CW + 2 * (lane & 15);  // first of this thread's two output columns
            if (p.bias) { k0 = sbias[c * 32 + 2 * (lane & 15)]; k1 = sbias[c * 32 + 2 * (lane & 15) + 1]; }
            if (p.rowbias) {
              const float2 a = __half22float2(__ldg(reinterpret_cast<const __half2*>(
                  p.rowbias + static_cast<long long>(n_img) * p.ld_rowbias + col)));
              k0 += a.x; k1 += a.y;
            }
            if (has_res) {  // the residual's value at the image's first pixel
              const float2 a = __half22float2(__ldg(reinterpret_cast<const __half2*>(
                  p.residual + static_cast<long long>(n_img) * p.gn.tiles_img * IG_BM * p.ld_res + col)));
              k0 += a.x; k1 += a.y;
            }
          }
          uint32_t pk[16];  // 32 output halves
          if (p.geglu) {
            // 64 accumulator columns = (value, gate) x 32 outputs, drained in two 32-column TMEM reads
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
              uint32_t raw[32];
              tmem_ld_x32(taddr + c * 64 + hh * 32, raw);
              tmem_wait_ld();
              const float4* sb4 = reinterpret_cast<const float4*>(sbias + c * 64 + hh * 32);
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                // accumulator columns 4i..4i+3 = (value, gate, value, gate) -> output columns 2i, 2i+1
                float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
                if (p.bias) bb = sb4[i];
                const __half2 val = __floats2half2_rn(__uint_as_float(raw[4 * i]) + bb.x, __uint_as_float(raw[4 * i + 2]) + bb.z);
                const __half2 gate = __floats2half2_rn(__uint_as_float(raw[4 * i + 1]) + bb.y, __uint_as_float(raw[4 * i + 3]) + bb.w);
                const float2 gf = __half22float2(gate);
                const __half2 act = __floats2half2_rn(gelu_fast(gf.x), gelu_fast(gf.y));
                pk[hh * 8 + i] = h2u(__hmul2(val, act));
              }
            }
          } else {
            uint32_t raw[32];
            tmem_ld_x32(taddr + c * 32, raw);
            tmem_wait_ld();
            const float4* sb4 = reinterpret_cast<const float4*>(sbias + c * 32);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
              if (p.bias) bb = sb4[i];
              pk[2 * i] = h2u(__floats2half2_rn(__uint_as_float(raw[4 * i]) + bb.x, __uint_as_float(raw[4 * i + 1]) + bb.y));
              pk[2 * i + 1] = h2u(__floats2half2_rn(__uint_as_float(raw[4 * i + 2]) + bb.z, __uint_as_float(raw[4 * i + 3]) + bb.w));
            }
            if (rb) {
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                if (ntile * BN + c * 32 + j * 8 >= p.N) break;
                const uint4 q = __ldg(reinterpret_cast<const uint4*>(rb + c * 32 + j * 8));
                pk[4 * j] = h2u(__hadd2(u2h(pk[4 * j]), u2h(q.x)));
                pk[4 * j + 1] = h2u(__hadd2(u2h(pk[4 * j + 1]), u2h(q.y)));
                pk[4 * j + 2] = h2u(__hadd2(u2h(pk[4 * j + 2]), u2h(q.z)));
                pk[4 * j + 3] = h2u(__hadd2(u2h(pk[4 * j + 3]), u2h(q.w)));
              }
            }
            if (p.act_silu) {
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const float2 f = __half22float2(u2h(pk[i]));
                pk[i] = h2u(__floats2half2_rn(silu_f(f.x), silu_f(f.y)));
              }
            }
          }
          if (c + NG >= nchunk) {  // last TMEM read of this tile by this warp: hand the accumulator back
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              if constexpr (CG == 2) mbar_arrive_cluster(&tempty[acc], 0);  // the issuer lives in the leader CTA
              else mbar_arrive(&tempty[acc]);
            }
          }
          if (has_res) {
            mbar_wait(&rf[b], (k / NBG) & 1);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint4 q = *reinterpret_cast<const uint4*>(buf + ((static_cast<uint32_t>(j) ^ sw) << 4));
              pk[4 * j] = h2u(__hadd2(u2h(pk[4 * j]), u2h(q.x)));
              pk[4 * j + 1] = h2u(__hadd2(u2h(pk[4 * j + 1]), u2h(q.y)));
              pk[4 * j + 2] = h2u(__hadd2(u2h(pk[4 * j + 2]), u2h(q.z)));
              pk[4 * j + 3] = h2u(__hadd2(u2h(pk[4 * j + 3]), u2h(q.w)));
            }
          }
#pragma unroll
          for (int j = 0; j < 4; ++j)
            *reinterpret_cast<uint4*>(buf + ((static_cast<uint32_t>(j) ^ sw) << 4)) =
                make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
          fence_proxy_async_smem();
          if (leader) {
            // buffer (k + LOOKAHEAD) % NBG must have been drained by the store of chunk k + LOOKAHEAD - NBG
            tma_store_wait_read<NBG - Cfg::LOOKAHEAD - 1>();
          }
          named_bar_sync(bar_id, 128);
          if (leader) {
            const int tx = mt % p.tiles_x, ty = (mt / p.tiles_x) % p.tiles_y, tn = mt / (p.tiles_x * p.tiles_y);
            tma_store_4d(&maps.c, ring + b * IG_CHUNK_BYTES, ntile * out_tile_cols + c * IG_CW, tx << p.wt_log,
                         ty << p.ht_log, tn << p.nt_log);
            tma_store_commit();
            if (has_res) prefetch_one();
          }
          if (p.gn.rec != nullptr && mt < p.m_tiles) {
            // GroupNorm statistics of the staged chunk (see IgGn).  Safe against the ring: this buffer is next written by
            // a residual prefetch / a later chunk only after a later named barrier, which every thread reaches after
            // these reads in program order.
            const int cp = lane & 15, hrow = lane >> 4;
            const int col = ntile * BN + c * IG_CW + 2 * cp;
            const uint8_t* cb = ring + b * IG_CHUNK_BYTES + static_cast<uint32_t>(quarter * 32 + hrow) * 64u +
                                static_cast<uint32_t>(cp & 3) * 4u;
            const uint32_t unit16 = static_cast<uint32_t>(cp >> 2);
            float S0 = 0.f, S1 = 0.f, Q0 = 0.f, Q1 = 0.f;
#pragma unroll
            for (int i = 0; i < 16; ++i) {  // row quarter*32 + hrow + 2i: its swizzle term (row >> 1) & 3 is i & 3
              const uint32_t u = *reinterpret_cast<const uint32_t*>(cb + i * 128 + ((unit16 ^ static_cast<uint32_t>(i & 3)) << 4));
              const float2 f = __half22float2(u2h(u));
              const float d0 = f.x - k0, d1 = f.y - k1;
              S0 += d0; S1 += d1;
              Q0 = fmaf(d0, d0, Q0); Q1 = fmaf(d1, d1, Q1);
            }
            S0 += __shfl_xor_sync(0xffffffffu, S0, 16); S1 += __shfl_xor_sync(0xffffffffu, S1, 16);
            Q0 += __shfl_xor_sync(0xffffffffu, Q0, 16); Q1 += __shfl_xor_sync(0xffffffffu, Q1, 16);
            if (hrow == 0) {
              const int E = p.gn.tiles_img * 4;
              float* rec = p.gn.rec + static_cast<long long>(n_img) * (2 * E + 1) * p.N;
              *reinterpret_cast<float4*>(rec + (static_cast<long long>(t_img * 4 + quarter) * p.N + col) * 2) =
                  make_float4(S0, Q0, S1, Q1);
              if (t_img == 0 && quarter == 0)
                *reinterpret_cast<float2*>(rec + static_cast<long long>(2 * E) * p.N + col) = make_float2(k0, k1);
            }
          }
        }
        if (c_first >= nchunk) {  // this group owns no chunk of the tile: still release the accumulator
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if constexpr (CG == 2) mbar_arrive_cluster(&tempty[acc], 0);
            else mbar_arrive(&tempty[acc]);
          }
        }
      }
      if (leader) tma_store_wait_all();
    }
  }

  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all();  // nobody leaves while its peer may still signal it or read its operands
  else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if constexpr (CG == 2) tmem_dealloc_pair(tmem_base, Cfg::TMEM_COLS);
    else tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

}  // namespace dm
