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
// warps 2..5 = epilogue (TMEM -> regs -> bias/temb/SiLU/GEGLU/residual -> global).  Two TMEM accumulator
// buffers let the epilogue of tile i overlap the main loop of tile i+1.
#pragma once
#include "ptx.cuh"

namespace dm {

constexpr int IG_MAX_SEG = 24;
constexpr int IG_BM = 128;
constexpr int IG_BK = 64;

struct IgSeg {
  int16_t src, dy, dx, pad_;
  int32_t dn, chan0, nchunks;
};

struct alignas(64) IgMaps {
  CUtensorMap a[2];
  CUtensorMap b;
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
};

template <int BN>
struct IgCfg {
  static constexpr int A_BYTES = IG_BM * IG_BK * 2;
  static constexpr int B_BYTES = BN * IG_BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int RAW_STAGES = (220 * 1024) / STAGE_BYTES;
  static constexpr int STAGES = RAW_STAGES > 8 ? 8 : RAW_STAGES;
  static constexpr int TMEM_COLS = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
  static constexpr int THREADS = 192;
};

template <int BN>
__global__ void __launch_bounds__(192, 1) igemm_kernel(const __grid_constant__ IgMaps maps, const IgParams p) {
  using Cfg = IgCfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smA = smem;
  uint8_t* smB = smem + STAGES * Cfg::A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + STAGES;
  uint64_t* tfull = bars + 2 * STAGES;
  uint64_t* tempty = bars + 2 * STAGES + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(&tfull[0], 1);
    mbar_init(&tfull[1], 1);
    mbar_init(&tempty[0], 4);
    mbar_init(&tempty[1], 4);
    fence_barrier_init();
    tma_prefetch_desc(&maps.a[0]);
    tma_prefetch_desc(&maps.b);
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int num_tiles = p.m_tiles * p.n_tiles;

  if (warp == 0) {
    // ============================== TMA producer ==============================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int mt = tile / p.n_tiles, ntile = tile % p.n_tiles;
        const int tx = mt % p.tiles_x, ty = (mt / p.tiles_x) % p.tiles_y, tn = mt / (p.tiles_x * p.tiles_y);
        const int x0 = tx << p.wt_log, y0 = ty << p.ht_log, n0 = tn << p.nt_log;
        int kit = 0;
        for (int s = 0; s < p.nseg; ++s) {
          const IgSeg sg = p.seg[s];
          const CUtensorMap* am = &maps.a[sg.src];
          for (int c = 0; c < sg.nchunks; ++c, ++kit) {
            mbar_wait(&empty[stage], phase ^ 1);
            mbar_arrive_expect_tx(&full[stage], Cfg::STAGE_BYTES);
            tma_load_4d(smA + stage * Cfg::A_BYTES, am, &full[stage], sg.chan0 + c * IG_BK, x0 + sg.dx, y0 + sg.dy,
                        n0 + sg.dn);
            tma_load_2d(smB + stage * Cfg::B_BYTES, &maps.b, &full[stage], kit * IG_BK, ntile * BN);
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
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_f16(BN);
      int stage = 0;
      uint32_t phase = 0;
      int lt = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++lt) {
        const int acc = lt & 1;
        const uint32_t acc_phase = (lt >> 1) & 1;
        mbar_wait(&tempty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kit = 0; kit < p.k_iters; ++kit) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint64_t ad = umma_desc_kmajor_sw128(smem_u32(smA + stage * Cfg::A_BYTES));
          const uint64_t bd = umma_desc_kmajor_sw128(smem_u32(smB + stage * Cfg::B_BYTES));
#pragma unroll
          for (int k = 0; k < IG_BK / 16; ++k) {
            // advance 16 fp16 = 32 B along K inside the 128-B swizzle row: +2 in (addr>>4) units
            umma_f16(d_tmem, ad + 2 * k, bd + 2 * k, idesc, (kit | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty[stage]);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&tfull[acc]);
      }
    }
  } else {
    // ============================== epilogue (4 warps) ==============================
    const int quarter = warp & 3;  // TMEM lane quarter this warp may access
    const int r = quarter * 32 + lane;
    constexpr int CH = (BN % 32 == 0) ? 32 : 16;
    int lt = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++lt) {
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
      for (int c0 = 0; c0 < BN; c0 += CH) {
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
        if (p.geglu) {
          // weight rows interleaved: even column = value, odd column = gate -> out[m, col/2]
          __half* o = reinterpret_cast<__half*>(p.out) + m * p.ld_out + (col0 >> 1);
#pragma unroll
          for (int i = 0; i < CH; i += 16) {
            if (i < nvalid) {
              uint32_t pk[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float a0 = round_h(v[i + 4 * j] * round_h(gelu_erf_f(v[i + 4 * j + 1])));
                const float a1 = round_h(v[i + 4 * j + 2] * round_h(gelu_erf_f(v[i + 4 * j + 3])));
                pk[j] = pack_h2(a0, a1);
              }
              *reinterpret_cast<uint4*>(o + (i >> 1)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            }
          }
          continue;
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
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

}  // namespace dm
