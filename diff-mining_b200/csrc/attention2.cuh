// Warp-specialised flash attention for the long self-attention layers (head_dim 40 / 80), sm_100a.
// Same op as attention.cuh (xformers.memory_efficient_attention, no mask, scale d^-0.5; call shape witnessed at
// /root/reference/diffmining/applications/parallel-dataset/pnp.py:440-442), restructured so that the softmax --
// which is MUFU.EX2-bound at these head dims (128x128 exps per tile = 1024 SM cycles vs <= 640 tensor cycles) --
// never waits for the tensor pipe:
//
//   one CTA = one (batch, head, 256-query block) = two 128-row Q tiles, one per softmax group
//   warp 0       TMA producer: Q once, then K / V tiles through ST-deep rings
//   warps 1, 2   MMA issuers, one per Q tile: S_q = Q_q K_j^T  and  O_q += P_q V_j   (tcgen05, fp32 in TMEM)
//   warp 3       idle
//   warps 4-11   softmax group 0 (tile 0): TWO threads per query row, each owning half of the key columns
//   warps 12-19  softmax group 1 (tile 1)
// Sixteen softmax warps = four per SM sub-partition: while one warp sits in a MUFU burst, waits for S or pulls TMEM,
// three others can issue, which is what keeps the XU pipe (the bound at these head dims) busy.  The two threads
// of a row exchange their half-row maxima / sums through shared memory (one named barrier per KV tile).
//
// A softmax thread pulls its half S row into registers in one TMEM pass and releases S at once (s_free), so
// QK^T of tile j+1 is issued while the exponentials of tile j are still being computed; P is double-buffered in
// shared memory so P_q V_j runs under the softmax of tile j+1 without a wait; the two groups hand an "XU token"
// (named barriers 4 / 5) back and forth so that their exponential phases alternate instead of colliding.  Row max via 3-input FMNMX, scale/offset and row sums via packed f32x2 FMA/ADD.
// O is rescaled in TMEM only when the row max grew by more than 2^8 (lazy rescale: P stays <= 256 in fp16, the
// final O / l is unchanged up to fp32 rounding).
#pragma once
#include "attention.cuh"

#ifndef DM_ATTN_PINGPONG
#define DM_ATTN_PINGPONG 1
#endif

namespace dm {

template <int D, int BKV, int ST>
struct Attn2Cfg {
  static constexpr int DK = (D + 15) / 16 * 16;
  static constexpr int NCH = (D + 63) / 64;
  static constexpr int Q_TILE_BYTES = NCH * 128 * 128;
  static constexpr int KV_BYTES = NCH * BKV * 128;
  static constexpr int P_TILE_BYTES = 128 * BKV * 2;  // one P buffer; two per Q tile
  static constexpr int NBAR = 1 + 4 * ST + 12;
  static constexpr int SMEM_BYTES = 2 * Q_TILE_BYTES + 2 * ST * KV_BYTES + 4 * P_TILE_BYTES + NBAR * 8 + 64;
  // Half-row exchange slots live in the never-read tail of the Q tile: logical 16-byte chunk 6 of every 128-byte
  // row of the last 64-channel chunk holds channels >= 48 (+64*(NCH-1)), which no MMA touches.
  static_assert(D - 64 * (NCH - 1) <= 48, "no free chunk in the Q tile for the exchange slots");
  static constexpr int HC = BKV / 2;  // key columns per softmax thread
  static constexpr int O_COL = 2 * BKV;  // S_q at columns [q*BKV, (q+1)*BKV), O_q at O_COL + q*DK
  static constexpr int TMEM_COLS = 512;
  static constexpr int THREADS = 128 + 512;
  static_assert(2 * BKV + 2 * DK <= 512, "TMEM budget");
  static_assert(BKV == 64 || BKV == 128, "BKV");
  static_assert(SMEM_BYTES <= 232448, "shared memory budget");
};

__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
__device__ __forceinline__ uint64_t pack_f2(float a, float b) {
  uint64_t d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(a), "f"(b));
  return d;
}
__device__ __forceinline__ void unpack_f2(uint64_t v, float& a, float& b) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ uint64_t fma_f2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t add_f2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

__device__ __forceinline__ void named_bar_sync256(int id) {
  asm volatile("bar.sync %0, 256;" ::"r"(id) : "memory");
}

template <int D, int BKV, int ST>
__global__ void __launch_bounds__(640, 1) attention2_kernel(const __grid_constant__ AttnMaps maps, const AttnParams p) {
  using Cfg = Attn2Cfg<D, BKV, ST>;
  constexpr int DK = Cfg::DK, NCH = Cfg::NCH;
  extern __shared__ __align__(1024) uint8_t smem[];  // 128B-swizzled tiles need 1024-byte alignment
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + 2 * Cfg::Q_TILE_BYTES;
  uint8_t* sV = sK + ST * Cfg::KV_BYTES;
  uint8_t* sP = sV + ST * Cfg::KV_BYTES;  // [q][buf]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 4 * Cfg::P_TILE_BYTES);
  uint64_t* q_full = bars;
  uint64_t* k_full = bars + 1;
  uint64_t* k_empty = k_full + ST;
  uint64_t* v_full = k_empty + ST;
  uint64_t* v_empty = v_full + ST;
  uint64_t* s_full = v_empty + ST;  // [q]
  uint64_t* s_free = s_full + 2;    // [q]
  uint64_t* p_full = s_free + 2;    // [q][buf]: per P buffer (a single barrier could be lapped: the softmax may finish tile j+1
                                    // before the issuer has waited for tile j)
  uint64_t* o_full = p_full + 4;    // [q][buf]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // batches in reverse launch order: the qkv GEMM that ran just before wrote the highest batch indices last, so
  // those rows are still L2-resident when the first CTAs start
  const int q0 = blockIdx.x * 256, head = blockIdx.y, b = static_cast<int>(gridDim.z) - 1 - static_cast<int>(blockIdx.z);
  const int kvb = p.kv_index ? p.kv_index[b] : b;
  const int nkv = (p.Tk + BKV - 1) / BKV;

  if (threadIdx.x == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < ST; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 2);  // one commit from each MMA issuer
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 2);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&s_free[i], 8);
      mbar_init(&p_full[2 * i], 8);
      mbar_init(&p_full[2 * i + 1], 8);
      mbar_init(&o_full[2 * i], 1);
      mbar_init(&o_full[2 * i + 1], 1);
    }
    fence_barrier_init();
    // this thread is also the TMA producer: start the Q / first K / first V loads before the TMEM allocation and the
    // CTA-wide sync, their latency (>= 2000 cycles on a cold tile) is the longest part of the prologue
    mbar_arrive_expect_tx(q_full, 2 * Cfg::Q_TILE_BYTES);
    for (int qq = 0; qq < 2; ++qq)
      for (int c = 0; c < NCH; ++c)
        tma_load_4d(sQ + qq * Cfg::Q_TILE_BYTES + c * 16384, &maps.q, q_full, c * 64, head, q0 + qq * 128, b);
    mbar_arrive_expect_tx(&k_full[0], Cfg::KV_BYTES);
    for (int c = 0; c < NCH; ++c) tma_load_4d(sK + c * BKV * 128, &maps.k, &k_full[0], c * 64, head, 0, kvb);
    mbar_arrive_expect_tx(&v_full[0], Cfg::KV_BYTES);
    for (int c = 0; c < NCH; ++c) tma_load_4d(sV + c * BKV * 128, &maps.v, &v_full[0], c * 64, head, 0, kvb);
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ============================== TMA producer ==============================
    if (lane == 0) {
      for (int j = 1; j < nkv; ++j) {  // Q and tile 0 were issued in the prologue
        const int st = j % ST;
        const uint32_t ph = (j / ST) & 1;
        mbar_wait(&k_empty[st], ph ^ 1);
        mbar_arrive_expect_tx(&k_full[st], Cfg::KV_BYTES);
        for (int c = 0; c < NCH; ++c)
          tma_load_4d(sK + st * Cfg::KV_BYTES + c * BKV * 128, &maps.k, &k_full[st], c * 64, head, j * BKV, kvb);
        mbar_wait(&v_empty[st], ph ^ 1);
        mbar_arrive_expect_tx(&v_full[st], Cfg::KV_BYTES);
        for (int c = 0; c < NCH; ++c)
          tma_load_4d(sV + st * Cfg::KV_BYTES + c * BKV * 128, &maps.v, &v_full[st], c * 64, head, j * BKV, kvb);
      }
    }
  } else if (warp == 1 || warp == 2) {
    // ============================== MMA issuer of Q tile q (one thread) ==============================
    // Each Q tile has its own in-order issuer so the two softmax warpgroups never wait for each other.
    if (lane == 0) {
      const int q = warp - 1;
      constexpr uint32_t idesc_s = umma_idesc_f16(BKV, false);
      constexpr uint32_t idesc_o = umma_idesc_f16(DK, true);
      auto issue_qk = [&](int st) {
#pragma unroll
        for (int ks = 0; ks < DK / 16; ++ks) {
          const uint64_t ad =
              umma_desc_kmajor_sw128(smem_u32(sQ + q * Cfg::Q_TILE_BYTES + (ks >> 2) * 16384)) + 2 * (ks & 3);
          const uint64_t bd =
              umma_desc_kmajor_sw128(smem_u32(sK + st * Cfg::KV_BYTES + (ks >> 2) * BKV * 128)) + 2 * (ks & 3);
          umma_f16(tmem_base + q * BKV, ad, bd, idesc_s, ks != 0 ? 1u : 0u);
        }
      };
      auto issue_pv = [&](int st, int buf, bool first) {
        const uint8_t* pb = sP + (2 * q + buf) * Cfg::P_TILE_BYTES;
#pragma unroll
        for (int ks = 0; ks < BKV / 16; ++ks) {
          const uint64_t ad = umma_desc_kmajor_sw128(smem_u32(pb + (ks >> 2) * 16384)) + 2 * (ks & 3);
          const uint64_t bd = umma_desc_mnmajor_sw128(smem_u32(sV + st * Cfg::KV_BYTES + ks * 2048), BKV * 128);
          umma_f16(tmem_base + Cfg::O_COL + q * DK, ad, bd, idesc_o, (first && ks == 0) ? 0u : 1u);
        }
      };
      mbar_wait(q_full, 0);
      mbar_wait(&k_full[0], 0);
      tc_fence_after();
      issue_qk(0);
      umma_commit(&s_full[q]);
      umma_commit(&k_empty[0]);
      for (int j = 0; j < nkv; ++j) {
        const uint32_t jp = j & 1;
        if (j + 1 < nkv) {
          const int st1 = (j + 1) % ST;
          mbar_wait(&k_full[st1], ((j + 1) / ST) & 1);
          mbar_wait(&s_free[q], jp);  // the softmax warpgroup holds S_q(j) in registers
          tc_fence_after();
          issue_qk(st1);
          umma_commit(&s_full[q]);
          umma_commit(&k_empty[st1]);
        }
        const int st = j % ST;
        mbar_wait(&v_full[st], (j / ST) & 1);
        mbar_wait(&p_full[2 * q + jp], (j >> 1) & 1);  // P_q(j) in smem buffer j&1, O_q rescaled
        tc_fence_after();
        issue_pv(st, jp, j == 0);
        umma_commit(&o_full[2 * q + jp]);
        umma_commit(&v_empty[st]);
      }
    }
  } else if (warp >= 4) {
    // ============================== softmax groups ==============================
    constexpr int HC = Cfg::HC;
    const int wg = (warp - 4) >> 3;
    const int half = ((warp - 4) >> 2) & 1;  // which half of the key columns of the tile
    const int quarter = warp & 3;            // TMEM lane quarter this warp may access
    const int row = quarter * 32 + lane;
    const int bar_id = 2 + wg;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    const uint32_t t_s = t_lane + wg * BKV + half * HC;
    const uint32_t t_o = t_lane + Cfg::O_COL + wg * DK;
    uint8_t* sPq = sP + 2 * wg * Cfg::P_TILE_BYTES + row * 128;
    uint64_t* o_full_q = o_full + 2 * wg;
    // exchange slots [parity][half] of this row (see Attn2Cfg)
    float* xrow = reinterpret_cast<float*>(sQ + wg * Cfg::Q_TILE_BYTES + (NCH - 1) * 16384 + row * 128 +
                                           ((6 ^ (row & 7)) << 4));
    float* xmine = xrow + half;
    float* xpeer = xrow + (half ^ 1);
    const float sc = p.scale_log2;
    const float thr = 8.f / sc;  // lazy rescale threshold in raw-score units (2^8 headroom)
    const uint64_t sc2 = pack_f2(sc, sc);
    float m_ref = -INFINITY;
    uint64_t l2 = pack_f2(0.f, 0.f);
#if DM_ATTN_PINGPONG
    if (wg == 1) asm volatile("bar.arrive 4, 512;" ::: "memory");  // group 0 exponentiates first
#endif
    // O columns (in 16-column TMEM chunks) this thread rescales on the rare path
    constexpr int NCH16 = DK / 16;
    const int ch_lo = half == 0 ? 0 : (NCH16 + 1) / 2, ch_hi = half == 0 ? (NCH16 + 1) / 2 : NCH16;

    for (int j = 0; j < nkv; ++j) {
      const uint32_t jp = j & 1;
      mbar_wait(&s_full[wg], jp);
      tc_fence_after();
      uint32_t raw[HC];
#pragma unroll
      for (int c0 = 0; c0 < HC; c0 += 32) tmem_ld_x32(t_s + c0, raw + c0);
      tmem_wait_ld();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_free[wg]);

      const int kbase = j * BKV + half * HC;
      if (kbase + HC > p.Tk) {  // ragged last tile (cross-attention: 77 keys)
#pragma unroll
        for (int i = 0; i < HC; ++i)
          if (kbase + i >= p.Tk) raw[i] = 0xff800000u;  // -inf
      }
      float mx0 = __uint_as_float(raw[0]), mx1 = __uint_as_float(raw[1]);
#pragma unroll
      for (int i = 2; i < HC; i += 4) {
        mx0 = fmax3(mx0, __uint_as_float(raw[i]), __uint_as_float(raw[i + 1]));
        if (i + 2 < HC) mx1 = fmax3(mx1, __uint_as_float(raw[i + 2]), __uint_as_float(raw[i + 3]));
      }
      // the row's other half lives in the partner thread (warp + 4): exchange the half-row maxima
      xmine[jp * 2] = fmaxf(mx0, mx1);
      named_bar_sync256(bar_id);
      const float mx = fmaxf(fmaxf(mx0, mx1), xpeer[jp * 2]);
      const bool grow = mx > m_ref + thr;  // true on the first tile (m_ref = -inf); identical in both halves
      const float m_new = grow ? mx : m_ref;
      const float alpha = grow ? exp2f((m_ref - m_new) * sc) : 1.f;
      if (j > 0 && __any_sync(0xffffffffu, grow)) {
        // rare: O_q must be rescaled, so PV(j-1) has to be complete first
        mbar_wait(&o_full_q[jp ^ 1], ((j - 1) >> 1) & 1);
        tc_fence_after();
        for (int ch = ch_lo; ch < ch_hi; ++ch) {
          uint32_t o[16];
          tmem_ld_x16(t_o + ch * 16, o);
          tmem_wait_ld();
#pragma unroll
          for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
          tmem_st_x16(t_o + ch * 16, o);
        }
        tmem_wait_st();
      }
      // P buffer j&1 was last read by PV(j-2).  No wait is needed: this thread observed s_full(j), i.e. the completion
      // of QK(j), and tcgen05.commit tracks ALL earlier MMAs of the issuing thread -- PV(j-2) was issued before QK(j).
      {
        float la, lb;
        unpack_f2(l2, la, lb);
        l2 = pack_f2(la * alpha, lb * alpha);
      }
      m_ref = m_new;
#if DM_ATTN_PINGPONG
      // ping-pong: the MUFU-bound exponential phases of the two groups alternate, so one group's TMEM reads, row
      // maxima and barrier waits run under the other group's exponentials (token = named barrier 4 + group)
      if (wg == 0) asm volatile("bar.sync 4, 512;" ::: "memory");
      else asm volatile("bar.sync 5, 512;" ::: "memory");
#endif
      const float nmoff = -m_ref * sc;
      const uint64_t off2 = pack_f2(nmoff, nmoff);
      uint8_t* sPb = sPq + jp * Cfg::P_TILE_BYTES;
#pragma unroll
      for (int c0 = 0; c0 < HC; c0 += 8) {
        uint32_t pk[4];
#pragma unroll
        for (int i = 0; i < 8; i += 2) {
          const uint64_t x =
              fma_f2(pack_f2(__uint_as_float(raw[c0 + i]), __uint_as_float(raw[c0 + i + 1])), sc2, off2);
          float e0, e1;
          unpack_f2(x, e0, e1);
          e0 = fast_exp2(e0);
          e1 = fast_exp2(e1);
          l2 = add_f2(l2, pack_f2(e0, e1));
          pk[i >> 1] = pack_h2(e0, e1);
        }
        const int c = half * HC + c0;  // column inside the BKV-wide P tile
        const uint32_t off = (c >> 6) * 16384 + ((((c & 63) >> 3) ^ (row & 7)) << 4);
        *reinterpret_cast<uint4*>(sPb + off) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      }
#if DM_ATTN_PINGPONG
      // hand the XU token to the other group (group 1 keeps its last one: group 0 has no tile left to wait for)
      if (wg == 0) asm volatile("bar.arrive 5, 512;" ::: "memory");
      else if (j + 1 < nkv) asm volatile("bar.arrive 4, 512;" ::: "memory");
#endif
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[2 * wg + jp]);
    }

    // ---- epilogue: O / l -> fp16; the two threads of a row add their half-row sums and split the head dim
    float la, lb;
    unpack_f2(l2, la, lb);
    const uint32_t xp = nkv & 1;
    xmine[xp * 2] = la + lb;
    named_bar_sync256(bar_id);
    const float inv = 1.f / (la + lb + xpeer[xp * 2]);
    mbar_wait(&o_full_q[(nkv - 1) & 1], ((nkv - 1) >> 1) & 1);
    tc_fence_after();
    const int q = q0 + wg * 128 + row;
    __half* orow = p.out + (static_cast<long long>(b) * p.Tq + q) * p.ld_out + head * D;
    constexpr int N8 = D / 8;
    const int c8_lo = half == 0 ? 0 : (N8 + 1) / 2, c8_hi = half == 0 ? (N8 + 1) / 2 : N8;
    for (int c8 = c8_lo; c8 < c8_hi; ++c8) {
      uint32_t o[8];
      tmem_ld_x8(t_o + c8 * 8, o);
      tmem_wait_ld();
      if (q < p.Tq) {
        *reinterpret_cast<uint4*>(orow + c8 * 8) =
            make_uint4(pack_h2(__uint_as_float(o[0]) * inv, __uint_as_float(o[1]) * inv),
                       pack_h2(__uint_as_float(o[2]) * inv, __uint_as_float(o[3]) * inv),
                       pack_h2(__uint_as_float(o[4]) * inv, __uint_as_float(o[5]) * inv),
                       pack_h2(__uint_as_float(o[6]) * inv, __uint_as_float(o[7]) * inv));
      }
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
