// Persistent warp-specialised flash attention for the long self-attention layers (head_dim 40 / 80), sm_100a.
// Same op and same tile structure as attention2.cuh (xformers.memory_efficient_attention, no mask, scale d^-0.5; call
// shape witnessed at /root/reference/diffmining/applications/parallel-dataset/pnp.py:440-442), rebuilt around what the
// round-1 profile showed: the softmax is MUFU.EX2-bound (128x128 exps per tile = 1024 XU cycles vs <= 640 tensor cycles),
// but with only four softmax warps per SM sub-partition the ISSUE port has to stay > 70 % busy to keep the XU fed
// (~370 instructions per warp and 64-column half-row, counted in SASS), and every CTA paid its own prologue/epilogue.
//
//   grid = min(#work, #SMs) PERSISTENT CTAs; work item = (batch, head, 256-query block) = two 128-row Q tiles
//   warp 0       TMA producer: Q of the next work item as soon as the last QK^T of the current one has completed,
//                K / V tiles through ST-deep rings that run across work items
//   warps 1, 2   MMA issuers, one per Q tile: S_q = Q_q K_j^T  and  [O_q | l_q] += P_q [V_j | 1]   (tcgen05, fp32 in TMEM)
//   warp 3       writes a column of ONES at channel D of every landed V tile (the padding column TMA zero-filled), so the
//                PV MMA itself accumulates the row sums l = sum_k P in TMEM column D of O: the softmax warps carry no
//                row-sum arithmetic and no second exchange (and l is the sum of the fp16-rounded P the numerator uses)
//   TPR = 2:  warps 4-11 softmax group 0 (tile 0), warps 12-19 group 1 (tile 1): TWO threads per query row, each owning
//             half of the key columns; the two threads of a row agree on the running maximum through one 16-bit
//             shared-memory slot each (the half-row max rounded UP to bf16 precision: any common value >= the true max
//             keeps the softmax exact after normalisation)
//   TPR = 1:  warps 4-7 / 8-11: ONE thread per query row (no exchange, fewer instructions per element, up to 168
//             registers per thread)
// S is pulled into registers in one TMEM pass and released at once, P is double-buffered in shared memory, O is rescaled
// in TMEM only when the row max grew by more than 2^8 (lazy rescale), and the two groups alternate their exponential
// phases through an "XU token" (named barriers 4 / 5) exactly as in attention2.
#pragma once
#include "attention2.cuh"

namespace dm {

template <int D, int BKV, int ST, int TPR = 2>
struct Attn3Cfg {
  static constexpr int DK = (D + 15) / 16 * 16;           // K extent of QK^T
  static constexpr int DKL = DK > D ? DK : DK + 16;       // N extent of the PV MMA: head_dim + the ones column at channel D
  static constexpr int NCH = (D + 63) / 64;               // 64-wide d chunks of Q / K
  static constexpr int NCHV = (DKL + 63) / 64;            // ... of V (must cover channel D)
  static constexpr int Q_TILE_BYTES = NCH * 128 * 128;
  static constexpr int K_BYTES = NCH * BKV * 128;
  static constexpr int V_BYTES = NCHV * BKV * 128;
  static constexpr int P_TILE_BYTES = 128 * BKV * 2;      // one P buffer; two per Q tile
  static constexpr int XCH_BYTES = 2 * 2 * 128 * 2 * 2;   // [group][parity][row][half] bf16 half-row maxima (TPR = 2)
  static constexpr int NBAR = 2 + 5 * ST + 14;
  static constexpr int SMEM_BYTES = 2 * Q_TILE_BYTES + ST * (K_BYTES + V_BYTES) + 4 * P_TILE_BYTES + XCH_BYTES + NBAR * 8 + 64;
  static constexpr int HC = BKV / TPR;                    // key columns per softmax thread
  static constexpr int O_COL = 2 * BKV;                   // S_q at columns [q*BKV, (q+1)*BKV), O_q at O_COL + q*DKL
  static constexpr int TMEM_COLS = 512;
  static constexpr int THREADS = 128 + 256 * TPR;
  // position of the ones column inside a V tile row: chunk, 16-byte piece, byte inside the piece
  static constexpr int ONE_CHUNK = D / 64, ONE_PIECE = (D % 64) / 8, ONE_BYTE = (D % 8) * 2;
  static_assert(TPR == 1 || TPR == 2, "threads per query row");
  static_assert(NCHV == NCH, "V tile must have the same chunk count as K (shared ring stride)");
  static_assert(2 * BKV + 2 * DKL <= 512, "TMEM budget");
  static_assert(BKV == 64 || BKV == 128, "BKV");
  static_assert(SMEM_BYTES <= 232448, "shared memory budget");
};

// round a float UP to bf16 precision (keeps the sign, -inf stays -inf); returned as the 16-bit pattern
__device__ __forceinline__ uint32_t bf16_round_up_bits(float x) {
  const uint32_t u = __float_as_uint(x);
  return (static_cast<int32_t>(u) >= 0) ? ((u + 0xFFFFu) >> 16) : (u >> 16);  // negatives: truncation moves toward +inf
}
template <int ID, int NT>
__device__ __forceinline__ void nbar_sync() {
  asm volatile("bar.sync %0, %1;" ::"n"(ID), "n"(NT) : "memory");
}
template <int ID, int NT>
__device__ __forceinline__ void nbar_arrive() {
  asm volatile("bar.arrive %0, %1;" ::"n"(ID), "n"(NT) : "memory");
}

template <int D, int BKV, int ST, int TPR>
__global__ void __launch_bounds__(128 + 256 * TPR, 1) attention3_kernel(const __grid_constant__ AttnMaps maps, const AttnParams p) {
  using Cfg = Attn3Cfg<D, BKV, ST, TPR>;
  constexpr int DK = Cfg::DK, DKL = Cfg::DKL, NCH = Cfg::NCH;
  constexpr int NSW = 4 * TPR;        // softmax warps per group
  constexpr int NTOK = 2 * 128 * TPR;  // threads on an XU-token barrier (both groups)
  extern __shared__ __align__(1024) uint8_t smem[];  // 128B-swizzled tiles need 1024-byte alignment
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + 2 * Cfg::Q_TILE_BYTES;
  uint8_t* sV = sK + ST * Cfg::K_BYTES;
  uint8_t* sP = sV + ST * Cfg::V_BYTES;  // [q][buf]
  uint16_t* xch = reinterpret_cast<uint16_t*>(sP + 4 * Cfg::P_TILE_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 4 * Cfg::P_TILE_BYTES + Cfg::XCH_BYTES);
  uint64_t* q_full = bars;
  uint64_t* q_free = bars + 1;
  uint64_t* k_full = bars + 2;
  uint64_t* k_empty = k_full + ST;
  uint64_t* v_full = k_empty + ST;
  uint64_t* v_ready = v_full + ST;
  uint64_t* v_empty = v_ready + ST;
  uint64_t* s_full = v_empty + ST;  // [q]
  uint64_t* s_free = s_full + 2;    // [q]
  uint64_t* p_full = s_free + 2;    // [q][buf]: one barrier per P buffer -- the softmax may run a whole tile ahead of the
                                    // issuer, and a single barrier two completions ahead would alias its phase parity
  uint64_t* o_full = p_full + 4;    // [q][buf]
  uint64_t* o_free = o_full + 4;    // [q]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_free + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nkv = (p.Tk + BKV - 1) / BKV;
  const int nqb = (p.Tq + 255) / 256;
  const int total = nqb * p.heads * p.B;
  // work w -> (query block, head, batch); consecutive w share K/V (L2 reuse); batches in reverse order: the qkv GEMM that
  // ran just before wrote the highest batch indices last, so those rows are still L2-resident when the first CTAs start
  auto decode = [&](int w, int& q0, int& head, int& b) {
    q0 = (w % nqb) * 256;
    head = (w / nqb) % p.heads;
    b = p.B - 1 - w / (nqb * p.heads);
  };

  if (threadIdx.x == 0) {
    mbar_init(q_full, 1);
    mbar_init(q_free, 2);  // one commit from each MMA issuer
    for (int i = 0; i < ST; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 2);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_ready[i], 1);
      mbar_init(&v_empty[i], 2);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&s_free[i], NSW);
      mbar_init(&p_full[2 * i], NSW);
      mbar_init(&p_full[2 * i + 1], NSW);
      mbar_init(&o_full[2 * i], 1);
      mbar_init(&o_full[2 * i + 1], 1);
      mbar_init(&o_free[i], NSW);
    }
    fence_barrier_init();
    tma_prefetch_desc(&maps.q);
    tma_prefetch_desc(&maps.k);
    tma_prefetch_desc(&maps.v);
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
      int g = 0;  // K/V tiles loaded so far (ring position runs across work items)
      int lw = 0;
      for (int w = blockIdx.x; w < total; w += gridDim.x, ++lw) {
        int q0, head, b;
        decode(w, q0, head, b);
        const int kvb = p.kv_index ? p.kv_index[b] : b;
        if (lw > 0) mbar_wait(q_free, (lw - 1) & 1);  // every QK^T of the previous work item has completed
        mbar_arrive_expect_tx(q_full, 2 * Cfg::Q_TILE_BYTES);
        for (int qq = 0; qq < 2; ++qq)
          for (int c = 0; c < NCH; ++c)
            tma_load_4d(sQ + qq * Cfg::Q_TILE_BYTES + c * 16384, &maps.q, q_full, c * 64, head, q0 + qq * 128, b);
        for (int j = 0; j < nkv; ++j, ++g) {
          const int st = g % ST;
          const uint32_t ph = (g / ST) & 1;
          mbar_wait(&k_empty[st], ph ^ 1);
          mbar_arrive_expect_tx(&k_full[st], Cfg::K_BYTES);
          for (int c = 0; c < NCH; ++c)
            tma_load_4d(sK + st * Cfg::K_BYTES + c * BKV * 128, &maps.k, &k_full[st], c * 64, head, j * BKV, kvb);
          mbar_wait(&v_empty[st], ph ^ 1);
          mbar_arrive_expect_tx(&v_full[st], Cfg::V_BYTES);
          for (int c = 0; c < Cfg::NCHV; ++c)
            tma_load_4d(sV + st * Cfg::V_BYTES + c * BKV * 128, &maps.v, &v_full[st], c * 64, head, j * BKV, kvb);
        }
      }
    }
  } else if (warp == 1 || warp == 2) {
    // ============================== MMA issuer of Q tile q (one thread) ==============================
    if (lane == 0) {
      const int q = warp - 1;
      constexpr uint32_t idesc_s = umma_idesc_f16(BKV, false);
      constexpr uint32_t idesc_o = umma_idesc_f16(DKL, true);
      auto issue_qk = [&](int st) {
#pragma unroll
        for (int ks = 0; ks < DK / 16; ++ks) {
          const uint64_t ad =
              umma_desc_kmajor_sw128(smem_u32(sQ + q * Cfg::Q_TILE_BYTES + (ks >> 2) * 16384)) + 2 * (ks & 3);
          const uint64_t bd =
              umma_desc_kmajor_sw128(smem_u32(sK + st * Cfg::K_BYTES + (ks >> 2) * BKV * 128)) + 2 * (ks & 3);
          umma_f16(tmem_base + q * BKV, ad, bd, idesc_s, ks != 0 ? 1u : 0u);
        }
      };
      auto issue_pv = [&](int st, int buf, bool first) {
        const uint8_t* pb = sP + (2 * q + buf) * Cfg::P_TILE_BYTES;
#pragma unroll
        for (int ks = 0; ks < BKV / 16; ++ks) {
          const uint64_t ad = umma_desc_kmajor_sw128(smem_u32(pb + (ks >> 2) * 16384)) + 2 * (ks & 3);
          const uint64_t bd = umma_desc_mnmajor_sw128(smem_u32(sV + st * Cfg::V_BYTES + ks * 2048), BKV * 128);
          umma_f16(tmem_base + Cfg::O_COL + q * DKL, ad, bd, idesc_o, (first && ks == 0) ? 0u : 1u);
        }
      };
      int G = 0;  // tiles processed so far by this issuer (runs across work items: barrier parities follow it)
      int lw = 0;
      for (int w = blockIdx.x; w < total; w += gridDim.x, ++lw) {
        mbar_wait(q_full, lw & 1);
        mbar_wait(&k_full[G % ST], (G / ST) & 1);
        if (G > 0) mbar_wait(&s_free[q], (G - 1) & 1);  // the softmax group holds S_q of the previous tile in registers
        tc_fence_after();
        issue_qk(G % ST);
        umma_commit(&s_full[q]);
        umma_commit(&k_empty[G % ST]);
        if (nkv == 1) umma_commit(q_free);
        for (int j = 0; j < nkv; ++j, ++G) {
          const uint32_t gp = G & 1;
          if (j + 1 < nkv) {
            const int st1 = (G + 1) % ST;
            mbar_wait(&k_full[st1], ((G + 1) / ST) & 1);
            mbar_wait(&s_free[q], gp);
            tc_fence_after();
            issue_qk(st1);
            umma_commit(&s_full[q]);
            umma_commit(&k_empty[st1]);
            if (j + 2 == nkv) umma_commit(q_free);  // last QK^T of this work item: Q may be overwritten once it completes
          }
          const int st = G % ST;
          mbar_wait(&v_ready[st], (G / ST) & 1);            // V tile landed and its ones column written
          mbar_wait(&p_full[2 * q + gp], (G >> 1) & 1);      // P_q in smem buffer G&1, O_q rescaled
          if (j == 0 && lw > 0) mbar_wait(&o_free[q], (lw - 1) & 1);  // the epilogue of the previous work item has read O_q
          tc_fence_after();
          issue_pv(st, gp, j == 0);
          umma_commit(&o_full[2 * q + gp]);
          umma_commit(&v_empty[st]);
        }
      }
    }
  } else if (warp == 3) {
    // ============================== ones column of V ==============================
    int g = 0;
    for (int w = blockIdx.x; w < total; w += gridDim.x) {
      for (int j = 0; j < nkv; ++j, ++g) {
        const int st = g % ST;
        mbar_wait(&v_full[st], (g / ST) & 1);
        uint8_t* vt = sV + st * Cfg::V_BYTES + Cfg::ONE_CHUNK * BKV * 128;
        for (int k = lane; k < BKV; k += 32)
          *reinterpret_cast<uint16_t*>(vt + k * 128 + ((Cfg::ONE_PIECE ^ (k & 7)) << 4) + Cfg::ONE_BYTE) = 0x3C00u;  // fp16 1.0
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&v_ready[st]);
      }
    }
  } else {
    // ============================== softmax groups ==============================
    constexpr int HC = Cfg::HC;
    const int wg = (warp - 4) / NSW;
    const int half = TPR == 2 ? ((warp - 4) >> 2) & 1 : 0;  // which half of the key columns of the tile (TPR = 2)
    const int quarter = warp & 3;                            // TMEM lane quarter this warp may access
    const int row = quarter * 32 + lane;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    const uint32_t t_s = t_lane + wg * BKV + half * HC;
    const uint32_t t_o = t_lane + Cfg::O_COL + wg * DKL;
    // P row of this thread: 128-byte swizzled rows inside 64-column chunks of 16 KB; 16-byte piece i of a chunk sits at
    // (row base | (row & 7) << 4) ^ (i << 4).  The thread's first column is half * HC.
    const uint32_t p_row = (smem_u32(sP) + 2 * wg * Cfg::P_TILE_BYTES + row * 128 + ((row & 7) << 4) + ((half * HC) >> 6) * 16384) ^
                           (static_cast<uint32_t>(((half * HC) & 63) >> 3) << 4);
    uint64_t* o_full_q = o_full + 2 * wg;
    uint16_t* xg = xch + wg * 2 * 128 * 2 + row * 2;  // [parity][row][half]
    const float sc = p.scale_log2;
    const float thr = 8.f / sc;  // lazy rescale threshold in raw-score units (2^8 headroom)
    const uint64_t sc2 = pack_f2(sc, sc);
    // O columns (in 16-column TMEM chunks, the l column included) this thread rescales on the rare path, and the 8-column
    // output pieces it stores in the epilogue (split between the two threads of a row when TPR = 2)
    constexpr int NCH16 = DKL / 16, N8 = D / 8;
    const int ch_lo = (TPR == 2 && half) ? (NCH16 + 1) / 2 : 0, ch_hi = (TPR == 2 && !half) ? (NCH16 + 1) / 2 : NCH16;
    const int c8_lo = (TPR == 2 && half) ? (N8 + 1) / 2 : 0, c8_hi = (TPR == 2 && !half) ? (N8 + 1) / 2 : N8;

    int G = 0;
    for (int w = blockIdx.x; w < total; w += gridDim.x) {
      int q0, head, b;
      decode(w, q0, head, b);
      float m_ref = -INFINITY;
      if (wg == 1) nbar_arrive<4, NTOK>();  // group 0 exponentiates first
      for (int j = 0; j < nkv; ++j, ++G) {
        const uint32_t gp = G & 1;
        mbar_wait(&s_full[wg], gp);
        tc_fence_after();
        uint32_t raw[HC];
#pragma unroll
        for (int c0 = 0; c0 < HC; c0 += 32) tmem_ld_x32(t_s + c0, raw + c0);
        tmem_wait_ld();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_free[wg]);

        const int kbase = j * BKV + half * HC;
        if (kbase + HC > p.Tk) {  // ragged last tile
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
        float mx = fmaxf(mx0, mx1);
        if constexpr (TPR == 2) {
          // the row's other half lives in the partner thread (warp + 4): both publish their half-row maximum rounded up to
          // bf16 and take the larger one, so the two threads always agree on the reference maximum
          const uint32_t mine = bf16_round_up_bits(mx);
          xg[gp * 256 + half] = static_cast<uint16_t>(mine);
          if (wg == 0) nbar_sync<2, 256>();
          else nbar_sync<3, 256>();
          const uint32_t peer = xg[gp * 256 + (half ^ 1)];
          mx = fmaxf(__uint_as_float(mine << 16), __uint_as_float(peer << 16));
        }
        const bool grow = mx > m_ref + thr;  // true on the first tile (m_ref = -inf); identical in both halves
        const float alpha = grow ? exp2f((m_ref - mx) * sc) : 1.f;
        if (j > 0 && __any_sync(0xffffffffu, grow)) {
          // rare: O_q (and its l column) must be rescaled, so PV of the previous tile has to be complete first
          mbar_wait(&o_full_q[gp ^ 1], ((G - 1) >> 1) & 1);
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
        m_ref = grow ? mx : m_ref;
        // P buffer G&1 was last read by PV(G-2).  No wait is needed: this thread observed s_full(G), i.e. the completion
        // of QK(G), and tcgen05.commit tracks ALL earlier MMAs of the issuing thread -- PV(G-2) was issued before QK(G).
        // ping-pong: the MUFU-bound exponential phases of the two groups alternate (token = named barrier 4 + group);
        // measured: without the token -19 %, handing it over early -3 %; evaluating 2 or 3 of every 8 exponentials with a
        // degree-3 polynomial on the FMA pipe: +-0 % / -6 %
        if (wg == 0) nbar_sync<4, NTOK>();
        else nbar_sync<5, NTOK>();
        const float nmoff = -m_ref * sc;
        const uint64_t off2 = pack_f2(nmoff, nmoff);
        const uint32_t pbase = p_row + gp * Cfg::P_TILE_BYTES;
#pragma unroll
        for (int c0 = 0; c0 < HC; c0 += 8) {
          uint32_t pk[4];
#pragma unroll
          for (int i = 0; i < 8; i += 2) {
            const uint64_t x =
                fma_f2(pack_f2(__uint_as_float(raw[c0 + i]), __uint_as_float(raw[c0 + i + 1])), sc2, off2);
            float e0, e1;
            unpack_f2(x, e0, e1);
            pk[i >> 1] = pack_h2(fast_exp2(e0), fast_exp2(e1));
          }
          // column c0 of this thread -> chunk (c0 >> 6), piece ((c0 & 63) >> 3) relative to the thread's first column
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"((pbase + (c0 >> 6) * 16384) ^ static_cast<uint32_t>(((c0 & 63) >> 3) << 4)),
                       "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3])
                       : "memory");
        }
        // hand the XU token to the other group (group 1 keeps its last one: group 0 has no tile left to wait for)
        if (wg == 0) nbar_arrive<5, NTOK>();
        else if (j + 1 < nkv) nbar_arrive<4, NTOK>();
        fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[2 * wg + gp]);
      }

      // ---- epilogue of the work item: O / l -> fp16; l = TMEM column D (accumulated by the PV MMA from the ones column)
      const int Gl = G - 1;  // last tile of this work item
      mbar_wait(&o_full_q[Gl & 1], (Gl >> 1) & 1);
      tc_fence_after();
      uint32_t lcol[8];
      tmem_ld_x8(t_o + D, lcol);
      tmem_wait_ld();
      const float inv = 1.f / __uint_as_float(lcol[0]);
      const int q = q0 + wg * 128 + row;
      __half* orow = p.out + (static_cast<long long>(b) * p.Tq + q) * p.ld_out + head * D;
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
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&o_free[wg]);  // O_q may be overwritten by the next work item's first PV
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
