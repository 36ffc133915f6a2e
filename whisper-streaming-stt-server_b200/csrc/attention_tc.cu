// K5: encoder self-attention (non-causal, T = 1500, head dim 64) on tcgen05 tensor cores.
//
// One persistent CTA per SM walks over (window, head, 256-query pair) items.  Per 128-key tile j and 128-query tile:
//     S_j = Q K_j^T        tcgen05.mma M128 N128 K64  -> TMEM
//     softmax warps: one thread per query row reads its S row with tcgen05.ld (no shuffles), raises the row's
//                    reference maximum lazily, writes P_j (bf16, truncated) into a SWIZZLE_128B K-major smem tile
//     O_j += P_j V_j       tcgen05.mma M128 N80 K128, accumulating in TMEM; V is consumed straight from the TMA tile as
//                          an MN-major B operand extended by a block of ones, so the row sum comes out of the tensor core
// The two query tiles of an item run in ping-pong (one tile's exp phase overlaps the other's MMAs).  The comments in
// front of the constants below record what each earlier version measured and why it was replaced (v3 .. v5 kernels
// themselves were removed in round 2; profiles/r1_attn_v*_ncu.txt keep their counters).
// Replaces F.scaled_dot_product_attention in upstream MultiHeadAttention.qkv_attention
// (reached from reference torch_whisper.py:55); SURVEY.md section 2.2 row K5.
#include <cuda.h>
#include <stdlib.h>

#include <algorithm>

#include "kernels.cuh"

namespace bw {

CUtensorMap make_operand_map(const void* ptr, int rows, int K, int ld, int Z, long long zstride, int box_rows);

namespace {

constexpr int TK = 128, HD = 64;  // key tile, head dim


// ------------------------------------------------------------------------------------------------
// v4: same ping-pong structure as v3, with the softmax inner loop cut from ~13 to ~4.5 issue slots per score
// (ncu of v3, profiles/r1_attn_v3_ncu.txt: 17 instructions per MUFU.EX2, XU pipe only 37 % busy):
//   * the row sum comes out of the tensor core: V is extended by a block of ones (N = 80: columns 64..79 of the
//     PV accumulator all hold sum_k P[q][k]), so no FADD per score and both column halves get the full-row sum;
//   * P is TRUNCATED to bf16 (one PRMT per pair).  The sum is taken over the same truncated values, so the
//     truncation bias cancels in O / l and what remains is the same +-2^-8 noise as round-to-nearest;
//   * the exponent argument x * scale - max is an FFMA2 (two scores per issue slot), O rescale-accumulate too;
//   * key masking (only the last key tile of a window has padding) is a separate code path;
//   * mbarrier waits carry a suspend-time hint (idle TMA / MMA warps no longer poll every ~100 cycles).
constexpr int V4_Q = 0;
constexpr int V4_KV = 32768;                       // 3 stages x (K 16 KB + V 16 KB)
constexpr int V4_STAGES = 3;
constexpr int V4_P = V4_KV + V4_STAGES * 32768;    // 2 x 32 KB
constexpr int V4_ONES = V4_P + 2 * 32768;          // 16 keys x 128 B of bf16 1.0: the second MN atom of every V k-slice
constexpr int V4_BAR = V4_ONES + 2048;
constexpr int V4_MX = V4_BAR + 256;                // row-max exchange: [2 tiles][2 parity][2 halves][128] floats
constexpr int V4_TOTAL = V4_MX + 2 * 512 * 4 + 1024;
constexpr int V4_ON = 80;                          // PV accumulator columns: 64 outputs + 16 copies of the row sum

template <bool MASK>
__device__ __forceinline__ float v4_chunk_max(const uint32_t* r, int c, int kvalid, float mx) {
  if (!MASK) {
#pragma unroll
    for (int i = 0; i < 32; i += 2) mx = fmaxf(mx, fmaxf(__uint_as_float(r[i]), __uint_as_float(r[i + 1])));
  } else {
#pragma unroll
    for (int i = 0; i < 32; ++i) mx = fmaxf(mx, (c * 32 + i < kvalid) ? __uint_as_float(r[i]) : -INFINITY);
  }
  return mx;
}

#ifdef V6_MAX4
// Row maximum over 64 scores with four independent accumulators (a single FMNMX3 chain is 32 dependent ops deep).
template <bool MASK>
__device__ __forceinline__ float v6_row_max(const uint32_t* sv, int kvalid) {
  float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
  if (!MASK) {
#pragma unroll
    for (int i = 0; i < 64; i += 8)
#pragma unroll
      for (int q = 0; q < 4; ++q) mx[q] = fmaxf(mx[q], fmaxf(__uint_as_float(sv[i + 2 * q]), __uint_as_float(sv[i + 2 * q + 1])));
  } else {
#pragma unroll
    for (int i = 0; i < 64; ++i) mx[i & 3] = fmaxf(mx[i & 3], (i < kvalid) ? __uint_as_float(sv[i]) : -INFINITY);
  }
  return fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
}
#endif
#ifdef V6_NARROW_PP
// The exponential of a row segment in three separately callable phases (see the ping-pong comment in v6): only the
// MUFU block needs the XU pipe; the FFMA2 block before it and the pack / store block after it overlap the OTHER
// query tile's MUFU block.
__device__ __forceinline__ void v6_exp_prep(uint32_t* sv, float sl2, float neg_m) {
#pragma unroll
  for (int i = 0; i < 64; i += 2) {
    float a0, a1;
    ffma2(a0, a1, __uint_as_float(sv[i]), __uint_as_float(sv[i + 1]), sl2, sl2, neg_m, neg_m);
    sv[i] = __float_as_uint(a0); sv[i + 1] = __float_as_uint(a1);
  }
}
__device__ __forceinline__ void v6_exp_mufu(uint32_t* sv) {
#pragma unroll
  for (int i = 0; i < 64; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+r"(sv[i]));
}
template <bool MASK>
__device__ __forceinline__ void v6_exp_store(const uint32_t* sv, int kvalid, uint32_t prow, int row) {
#pragma unroll
  for (int g = 0; g < 8; ++g) {
    uint32_t pk[4];
#pragma unroll
    for (int i = 0; i < 8; i += 2) {
      uint32_t e0 = sv[g * 8 + i], e1 = sv[g * 8 + i + 1];
      if (MASK) {
        if (g * 8 + i >= kvalid) e0 = 0u;
        if (g * 8 + i + 1 >= kvalid) e1 = 0u;
      }
      asm volatile("prmt.b32 %0, %1, %2, 0x7632;" : "=r"(pk[i >> 1]) : "r"(e0), "r"(e1));  // truncation (see v4 header)
    }
    st_shared_v4(prow + (uint32_t)((g ^ (row & 7)) << 4), pk[0], pk[1], pk[2], pk[3]);
  }
}
#endif

// One tile row segment (64 scores) -> 64 truncated-bf16 probabilities in the swizzled P tile, in three explicit
// phases so that the 64 MUFU.EX2 of a warp issue back to back (in-order issue: a PRMT scheduled right behind its
// MUFU stalls the warp for the MUFU latency and the XU pipe runs at half rate -- 16 instead of 8 cycles per
// warp-wide MUFU in the interleaved version).
template <bool MASK>
__device__ __forceinline__ void v5_row_exp(uint32_t* sv, int kvalid, float sl2, float neg_m, uint32_t prow, int row) {
#pragma unroll
  for (int i = 0; i < 64; i += 2) {
    float a0, a1;
    ffma2(a0, a1, __uint_as_float(sv[i]), __uint_as_float(sv[i + 1]), sl2, sl2, neg_m, neg_m);
    sv[i] = __float_as_uint(a0); sv[i + 1] = __float_as_uint(a1);
  }
#ifdef V6_POLY_PERIOD
  // Experiment kept for the record (FlashAttention-4's trick): one pair in every V6_POLY_PERIOD / 2 pairs computed by
  // exp2_poly2 on the FMA / ALU pipes instead of MUFU.EX2.  Measured at batch 16: 50 % -> 354 us, 25 % -> 318 us,
  // 12.5 % -> 316 us, 0 % -> 315 us per layer: the exp phase is not XU-throughput-bound here, the extra issue slots cost
  // more than the XU cycles they free (profiles/r1_notes.md).
#pragma unroll
  for (int i = 0; i < 64; i += 2) {
    if ((i % (2 * V6_POLY_PERIOD / 2)) == V6_POLY_PERIOD - 2) {
      exp2_poly2(sv[i], sv[i + 1], __uint_as_float(sv[i]), __uint_as_float(sv[i + 1]));
    } else {
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+r"(sv[i]));
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+r"(sv[i + 1]));
    }
  }
#else
#pragma unroll
  for (int i = 0; i < 64; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+r"(sv[i]));
#endif
#pragma unroll
  for (int g = 0; g < 8; ++g) {
    uint32_t pk[4];
#pragma unroll
    for (int i = 0; i < 8; i += 2) {
      uint32_t e0 = sv[g * 8 + i], e1 = sv[g * 8 + i + 1];
      if (MASK) {
        if (g * 8 + i >= kvalid) e0 = 0u;
        if (g * 8 + i + 1 >= kvalid) e1 = 0u;
      }
      asm volatile("prmt.b32 %0, %1, %2, 0x7632;" : "=r"(pk[i >> 1]) : "r"(e0), "r"(e1));  // truncation (see v4 header)
    }
    st_shared_v4(prow + (uint32_t)((g ^ (row & 7)) << 4), pk[0], pk[1], pk[2], pk[3]);
  }
}




// ------------------------------------------------------------------------------------------------
// v5: v4 + the accumulators never leave TMEM.  ncu of v4 (profiles/r1_attn_v4_ncu.txt): XU 44 %, the softmax
// warps stall on tcgen05.ld (S is read twice per tile, O once) -- latency-, not issue-bound.  Here
//   * O and the row sum (ones block) ACCUMULATE in TMEM across key tiles (tcgen05.mma accumulate), against a
//     per-row reference maximum that is only raised when a tile's maximum exceeds it by more than 2^8: then, and
//     only then, the softmax warps rescale their accumulator columns in place (tcgen05.ld / st).  P stays <= 256,
//     exact in the fp32 accumulators; O / l is formed once at the end;
//   * S is read from TMEM once per tile into 64 registers (the 32 registers of the old O copy are gone);
//   * the two warps that share a TMEM lane quarter exchange their half-row maxima through a 64-thread named
//     barrier instead of a 256-thread one.
constexpr float kV5Slack = 8.f;  // log2 units
constexpr int V5_MMA_B = 18;     // warp index of the second MMA issuer
constexpr int V5_THREADS = 608;  // TMA, MMA A, 16 softmax warps, MMA B



// ------------------------------------------------------------------------------------------------
// v6: v5 made persistent.  Timing v5 with 12 / 6 / 3 / 2 key tiles per CTA (tools/attn_bench.py) showed a fixed
// cost of ~11 K cycles per CTA (launch, TMEM allocation, first Q / K / V round trip, final PV + store, teardown)
// next to ~3 K cycles per key tile: a quarter of the kernel, and with one 204 KB CTA per SM nothing overlaps it.
// Here one CTA per SM walks over (window, head, query-pair) items; every role keeps a flattened tile counter, so
// the TMA warp is already loading the next item's Q / K / V, and the MMA warps issue the next item's first S,
// while the softmax warps finish the current one.  Q is single-buffered: the last S of an item is issued two
// tiles early (see v5), so the buffer is free long before the item ends.
// Timeline of CTA 0 (tools/attn_trace.py, profiles/r1_attn_v6_timeline.txt) then showed, per key tile and query
// tile: exp phase 1485 cycles (1024 of MUFU + FFMA2 prologue + PRMT / st.shared tail, all inside the ping-pong
// critical section), row max + exchange 583, wait for the PV product 390, and S_{t+1} issued ~2 K cycles late because
// K_{t+1} shared its ring stage with V_{t-2}.  Hence: K and V have their own 3-deep rings (a K stage is released as
// soon as both S products have read it), only the 64 MUFU per thread sit between the ping-pong barriers, and the
// row maximum uses four independent chains.
template <bool PINGPONG, bool TRACE>
__global__ void __launch_bounds__(V5_THREADS, 1)
attn_encoder_v6_kernel(const __grid_constant__ CUtensorMap tm, bf16* __restrict__ out, int T_len, int d, int n_head, int n_qp,
                       int n_items, unsigned long long* trace_buf, int pv_n) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + V4_BAR);
  uint64_t* q_full = bars;           // 1
  uint64_t* k_full = bars + 1;       // 3
  uint64_t* k_empty = bars + 4;      // 3 (2 arrivals: one per MMA warp, after its S product)
  uint64_t* s_full = bars + 7;       // 2 (per tile)
  uint64_t* p_full = bars + 9;       // 2 (256 arrivals each)
  uint64_t* o_full = bars + 11;      // 2
  uint64_t* s_free = bars + 13;      // 2 (256 arrivals each)
  uint64_t* q_empty = bars + 15;     // 1 (2 arrivals: both MMA warps have issued the item's last S)
  uint64_t* v_full = bars + 16;      // 3
  uint64_t* v_empty = bars + 19;     // 3 (2 arrivals: one per MMA warp, after its PV product)
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 22);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_kt = (T_len + TK - 1) / TK;
  // debug timeline (CTA 0, lane 0 of the TMA warp, both MMA warps and one softmax warp per tile): tag = warp << 16 | event << 8 | tile index
  // (fire-and-forget clock64 stores into fixed slots [warp slot][tile < 40][event < 20]: the atomics-based trace_mark
  //  costs ~0.5 us per mark, more than the phases measured here)
  unsigned long long* const trace = (blockIdx.x == 0 && lane == 0 && (warp <= 2 || warp == 10 || warp == V5_MMA_B)) ? trace_buf : nullptr;
  const int tslot = warp <= 2 ? warp : (warp == 10 ? 3 : 4);
  const int my_items = (n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  auto item_coords = [&](int it, int& row_base, int& q0, int& h) {
    const int item = blockIdx.x + it * gridDim.x;
    const int qp = item % n_qp;
    h = (item / n_qp) % n_head;
    row_base = (item / (n_qp * n_head)) * T_len;
    q0 = qp * 256;
  };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm);
    mbar_init(q_full, 1);
    mbar_init(q_empty, 2);
    for (int i = 0; i < V4_STAGES; ++i) { mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 2); mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 2); }
    for (int i = 0; i < 2; ++i) { mbar_init(&s_full[i], 1); mbar_init(&p_full[i], 256); mbar_init(&o_full[i], 1); mbar_init(&s_free[i], 256); }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr_smem, 512);
    tmem_relinquish();
  }
  if (threadIdx.x >= 64 && threadIdx.x < 64 + 128) {  // 2 KB of bf16 1.0
    reinterpret_cast<uint4*>(smem + V4_ONES)[threadIdx.x - 64] = make_uint4(0x3f803f80u, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u);
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const uint32_t tm_S = tmem_base;         // + 128 * tile
  const uint32_t tm_O = tmem_base + 256;   // + 128 * tile (80 columns used)

  if (warp == 0) {
    if (lane == 0) {
      // K runs ahead of V: a K stage is free once both S products have read it (two tiles before the matching V stage
      // is), so the next scores never wait for a load.  One thread drives both rings: V of tile g is requested when the
      // K ring has moved on by kKLead tiles (or at the end), which keeps the waits deadlock-free and in order.
      constexpr int kKLead = 2;
      const int total = my_items * n_kt;
      auto coords = [&](int g, int& row_base, int& h, int& j) {
        int q0;
        item_coords(g / n_kt, row_base, q0, h);
        j = g % n_kt;
      };
      auto load_v = [&](int g) {
        int row_base, h, j;
        coords(g, row_base, h, j);
        const int s = g % V4_STAGES;
        mbar_wait(&v_empty[s], ((g / V4_STAGES) & 1) ^ 1);
        mbar_arrive_expect_tx(&v_full[s], 16384);
        tma_load_3d(smem + V4_KV + s * 32768 + 16384, &tm, &v_full[s], 2 * d + h * HD, row_base + j * TK, 0);
      };
      for (int g = 0; g < total; ++g) {
        int row_base, h, j;
        coords(g, row_base, h, j);
        if (g >= kKLead) load_v(g - kKLead);  // before the Q wait below: V of the previous item's last tiles must not queue behind it
        if (j == 0) {  // first tile of an item: its Q
          const int it = g / n_kt;
          int rb, q0, hh;
          item_coords(it, rb, q0, hh);
          if (it > 0) mbar_wait(q_empty, (it - 1) & 1);
          mbar_arrive_expect_tx(q_full, 32768);
          tma_load_3d(smem + V4_Q, &tm, q_full, hh * HD, rb + q0, 0);
          tma_load_3d(smem + V4_Q + 16384, &tm, q_full, hh * HD, rb + q0 + 128, 0);
        }
        const int s = g % V4_STAGES;
        mbar_wait(&k_empty[s], ((g / V4_STAGES) & 1) ^ 1);
        mbar_arrive_expect_tx(&k_full[s], 16384);
        tma_load_3d(smem + V4_KV + s * 32768, &tm, &k_full[s], d + h * HD, row_base + j * TK, 0);
      }
      for (int g = (total > kKLead ? total - kKLead : 0); g < total; ++g) load_v(g);
    }
    __syncwarp();
  } else if (warp == 1 || warp == V5_MMA_B) {
    if (lane == 0) {
      const int tile = warp == 1 ? 0 : 1;
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128, 0, 0);
      const uint32_t idesc_o = umma_idesc_bf16(128, pv_n, 0, 1);  // B (= V | ones) is MN-major
      const uint32_t ones = smem_u32(smem + V4_ONES);
      const int total = my_items * n_kt;
      // S of flattened tile t (item t / n_kt, key tile t % n_kt); the caller made sure its Q and K have landed
      auto issue_S = [&](int t) {
        const uint64_t adesc = umma_smem_desc_sw128(smem_u32(smem + V4_Q + tile * 16384), 16, 1024);
        const uint64_t bdesc = umma_smem_desc_sw128(smem_u32(smem + V4_KV + (t % V4_STAGES) * 32768), 16, 1024);
#pragma unroll
        for (int k = 0; k < HD / 16; ++k) umma_f16(tm_S + tile * 128, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc_s, k != 0);
        umma_commit(&s_full[tile]);
        umma_commit(&k_empty[t % V4_STAGES]);            // this tile's S no longer needs K_t (one of two arrivals)
        if (t % n_kt == n_kt - 1) umma_commit(q_empty);  // the item's last read of Q (one of two arrivals)
      };
      mbar_wait(q_full, 0);
      mbar_wait(&k_full[0], 0);
      tc_fence_after();
      issue_S(0);
      for (int t = 0; t < total; ++t) {
        const int j = t % n_kt;
        if (t + 1 < total) {  // next scores, possibly the next item's first tile: issued before P_t exists
          if (j == n_kt - 1) mbar_wait(q_full, ((t + 1) / n_kt) & 1);
          mbar_wait(&k_full[(t + 1) % V4_STAGES], ((t + 1) / V4_STAGES) & 1);
          mbar_wait(&s_free[tile], t & 1);
          tc_fence_after();
          issue_S(t + 1);
          if (TRACE && trace && (t + 1) < 40) trace[1 + (tslot * 40 + (t + 1)) * 20 + 3] = (unsigned long long)clock64();
        }
        mbar_wait(&v_full[t % V4_STAGES], (t / V4_STAGES) & 1);
        mbar_wait(&p_full[tile], t & 1);
        if (TRACE && trace && (t) < 40) trace[1 + (tslot * 40 + (t)) * 20 + 4] = (unsigned long long)clock64();
        tc_fence_after();
        const uint32_t sp = smem_u32(smem + V4_P + tile * 32768);
        const uint32_t sv = smem_u32(smem + V4_KV + (t % V4_STAGES) * 32768 + 16384);
#pragma unroll
        for (int k = 0; k < TK / 16; ++k) {
          const uint64_t adesc = umma_smem_desc_sw128(sp + (k >> 2) * 16384 + (k & 3) * 32, 16, 1024);
          const uint32_t vk = sv + k * 2048;
          const uint64_t bdesc = umma_smem_desc_sw128(vk, ones - vk, 1024);  // MN atom 1 (at +LBO) = the block of ones
          umma_f16(tm_O + tile * 128, adesc, bdesc, idesc_o, (j | k) != 0);   // accumulates across the item's key tiles
        }
        umma_commit(&o_full[tile]);
        umma_commit(&v_empty[t % V4_STAGES]);
      }
    }
    __syncwarp();
  } else {
    const int tile = (warp - 2) >> 3;
    const int half = ((warp - 2) >> 2) & 1;
    const int wq = warp & 3;
    const int row = wq * 32 + lane;
    const uint32_t lane_off = (uint32_t)(wq * 32) << 16;
    const float sl2 = 0.125f * 1.4426950408889634f;
    const uint32_t my_S = tm_S + tile * 128 + lane_off + half * 64, my_O = tm_O + tile * 128 + lane_off + half * 32;
    const uint32_t my_L = tm_O + tile * 128 + lane_off + 64;
    const int pair_bar = 1 + tile * 4 + wq;
    const uint32_t prow = smem_u32(smem + V4_P + tile * 32768 + half * 16384 + row * 128);
    const uint32_t mxbuf = smem_u32(smem + V4_MX) + tile * 2048;
    constexpr int kPingPongBar = 9;
    if (PINGPONG && tile == 1) asm volatile("bar.arrive %0, 512;" ::"r"(kPingPongBar + 0) : "memory");
    const int total = my_items * n_kt;
    int t = 0;
    for (int it = 0; it < my_items; ++it) {
      int row_base, q0, h;
      item_coords(it, row_base, q0, h);
      float m = -INFINITY;
      for (int j = 0; j < n_kt; ++j, ++t) {
        if (TRACE && trace && (t) < 40) trace[1 + (tslot * 40 + (t)) * 20 + 10] = (unsigned long long)clock64();
        mbar_wait(&s_full[tile], t & 1);
        if (TRACE && trace && (t) < 40) trace[1 + (tslot * 40 + (t)) * 20 + 11] = (unsigned long long)clock64();
        tc_fence_after();
        uint32_t sv[64];
        tmem_ld_32x32b_x32(my_S, sv);
        tmem_ld_32x32b_x32(my_S + 32, sv + 32);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(&s_free[tile]);
        if (TRACE && trace && (t) < 40) trace[1 + (tslot * 40 + (t)) * 20 + 12] = (unsigned long long)clock64();
        const int kvalid = T_len - j * TK - half * 64;
        const bool masked = kvalid < 64;
#ifndef V6_MAX4
        float mx = -INFINITY;
        if (!masked) { mx = v4_chunk_max<false>(sv, 0, kvalid, mx); mx = v4_chunk_max<false>(sv + 32, 1, kvalid, mx); }
        else { mx = v4_chunk_max<true>(sv, 0, kvalid, mx); mx = v4_chunk_max<true>(sv + 32, 1, kvalid, mx); }
#else
        const float mx = masked ? v6_row_max<true>(sv, kvalid) : v6_row_max<false>(sv, kvalid);
#endif
        st_shared_f32(mxbuf + (uint32_t)((((t & 1) * 2 + half) * 128 + row) * 4), mx);
        asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
        const float tmax = fmaxf(mx, ld_shared_f32(mxbuf + (uint32_t)((((t & 1) * 2 + (half ^ 1)) * 128 + row) * 4))) * sl2;
        if (TRACE && trace && (t) < 40) trace[1 + (tslot * 40 + (t)) * 20 + 13] = (unsigned long long)clock64();
        if (j == 0) {
          m = tmax;  // the previous item's epilogue waited for its last PV: P and the accumulators are free
        } else {
          mbar_wait(&o_full[tile], (t - 1) & 1);
          tc_fence_after();
          const bool grow = tmax > m + kV5Slack;
          if (__any_sync(0xffffffffu, grow)) {
            const float m_new = grow ? tmax : m;
            const float f = fast_exp2(m - m_new);
#pragma unroll 1
            for (int c = 0; c < (half == 0 ? 3 : 2); ++c) {
              const uint32_t addr = (c < 2) ? my_O + c * 16 : my_L;
              uint32_t r[16];
              tmem_ld_32x32b_x16(addr, r);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 16; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * f);
              tmem_st_32x32b_x16(addr, r);
            }
            tmem_st_wait();
            m = m_new;
          }
        }
#ifndef V6_NARROW_PP
        if (PINGPONG) asm volatile("bar.sync %0, 512;" ::"r"(kPingPongBar + tile) : "memory");
        if (!masked) v5_row_exp<false>(sv, kvalid, sl2, -m, prow, row);
        else v5_row_exp<true>(sv, kvalid, sl2, -m, prow, row);
        if (PINGPONG && !(tile == 1 && t == total - 1)) asm volatile("bar.arrive %0, 512;" ::"r"(kPingPongBar + (tile ^ 1)) : "memory");
#else
        v6_exp_prep(sv, sl2, -m);
        if (TRACE && trace && (t) < 40) trace[1 + (tslot * 40 + (t)) * 20 + 14] = (unsigned long long)clock64();
        if (PINGPONG) asm volatile("bar.sync %0, 512;" ::"r"(kPingPongBar + tile) : "memory");
        if (TRACE && trace && (t) < 40) trace[1 + (tslot * 40 + (t)) * 20 + 15] = (unsigned long long)clock64();
        v6_exp_mufu(sv);
        if (PINGPONG && !(tile == 1 && t == total - 1)) asm volatile("bar.arrive %0, 512;" ::"r"(kPingPongBar + (tile ^ 1)) : "memory");
        if (TRACE && trace && (t) < 40) trace[1 + (tslot * 40 + (t)) * 20 + 16] = (unsigned long long)clock64();
        if (!masked) v6_exp_store<false>(sv, kvalid, prow, row);
        else v6_exp_store<true>(sv, kvalid, prow, row);
#endif
        fence_proxy_async_smem();
        tc_fence_before();
        mbar_arrive(&p_full[tile]);
        if (TRACE && trace && (t) < 40) trace[1 + (tslot * 40 + (t)) * 20 + 17] = (unsigned long long)clock64();
      }
      // item epilogue: O / l -> global
      mbar_wait(&o_full[tile], (t - 1) & 1);
      if (TRACE && trace && (t) < 40) trace[1 + (tslot * 40 + (t)) * 20 + 18] = (unsigned long long)clock64();
      tc_fence_after();
      uint32_t r[32];
      tmem_ld_32x32b_x32(my_O, r);
      const float l = __uint_as_float(tmem_ld_32x32b_x1(my_L));
      tmem_ld_wait();
      const int qrow = q0 + tile * 128 + row;
      if (qrow < T_len) {
        const float inv = 1.f / l;
        bf16* orow = out + (long long)(row_base + qrow) * d + h * HD + half * 32;
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          uint4 o4;
          o4.x = pack_bf16x2(__uint_as_float(r[i]) * inv, __uint_as_float(r[i + 1]) * inv);
          o4.y = pack_bf16x2(__uint_as_float(r[i + 2]) * inv, __uint_as_float(r[i + 3]) * inv);
          o4.z = pack_bf16x2(__uint_as_float(r[i + 4]) * inv, __uint_as_float(r[i + 5]) * inv);
          o4.w = pack_bf16x2(__uint_as_float(r[i + 6]) * inv, __uint_as_float(r[i + 7]) * inv);
          *reinterpret_cast<uint4*>(orow + i) = o4;
        }
      }
      tc_fence_before();
      if (TRACE && trace && (t) < 40) trace[1 + (tslot * 40 + (t)) * 20 + 19] = (unsigned long long)clock64();
    }
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace

void attn_encoder_tc(const bf16* qkv, bf16* out, int batch, int T_len, int n_head, cudaStream_t stream) {
  const int d = n_head * HD;
  int dev = 0;
  BW_CUDA(cudaGetDevice(&dev));
  CUtensorMap tm = make_operand_map(qkv, batch * T_len, 3 * d, 3 * d, 1, 0, 128);
  static std::atomic<unsigned long long> v6_set{0};
  static int sm_count = 0;
  if (!(v6_set.load() >> dev & 1ull)) {
    BW_CUDA(cudaFuncSetAttribute(attn_encoder_v6_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, V4_TOTAL));
    BW_CUDA(cudaFuncSetAttribute(attn_encoder_v6_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, V4_TOTAL));
    BW_CUDA(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev));
    v6_set.fetch_or(1ull << dev);
  }
  const int n_qp = (T_len + 255) / 256, n_items = n_qp * n_head * batch;
  const int grid = std::min(n_items, sm_count > 0 ? sm_count : 148);
  const int pv_n = V4_ON;
  if (g_trace_dev) attn_encoder_v6_kernel<true, true><<<grid, V5_THREADS, V4_TOTAL, stream>>>(tm, out, T_len, d, n_head, n_qp, n_items, g_trace_dev, pv_n);
  else attn_encoder_v6_kernel<true, false><<<grid, V5_THREADS, V4_TOTAL, stream>>>(tm, out, T_len, d, n_head, n_qp, n_items, nullptr, pv_n);
  BW_CUDA(cudaGetLastError());
  ++g_kernel_launches;
}

}  // namespace bw
