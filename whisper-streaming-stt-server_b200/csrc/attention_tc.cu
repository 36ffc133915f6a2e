// K5: encoder self-attention (non-causal, T = 1500, head dim 64) on tcgen05 tensor cores.
//
// One CTA = one (batch, head, 128-query tile).  Per 128-key tile j:
//     S_j = Q K_j^T        tcgen05.mma M128 N128 K64  -> TMEM (double buffered)
//     softmax warps: one thread per query row reads its S row with tcgen05.ld (no shuffles),
//                    online max / sum, writes P_j (bf16) into a SWIZZLE_128B K-major smem tile
//     O_j = P_j V_j        tcgen05.mma M128 N64 K128, V consumed straight from the TMA tile as an
//                          MN-major B operand (no transposed copy of V) -> TMEM (double buffered)
//     softmax warps fold O_j into fp32 registers: o = o * alpha + O_j
// Warp roles: 0 = TMA producer, 1 = MMA issuer + TMEM allocator, 2..9 = softmax / epilogue: two warps per
// TMEM lane quarter (= per SM sub-partition), each owning 64 of the 128 key columns of its 32 query rows and
// 32 of the 64 output columns, so every scheduler has two warps to hide tcgen05.ld / MUFU latency.
// Replaces F.scaled_dot_product_attention in upstream MultiHeadAttention.qkv_attention
// (reached from reference torch_whisper.py:55); SURVEY.md section 2.2 row K5.
#include <cuda.h>
#include <stdlib.h>

#include "kernels.cuh"

namespace bw {

CUtensorMap make_operand_map(const void* ptr, int rows, int K, int ld, int Z, long long zstride, int box_rows);

namespace {

constexpr int TQ = 128, TK = 128, HD = 64;
constexpr int SM_Q = 0;
constexpr int SM_K = 16384;                 // 2 stages x 16 KB
constexpr int SM_V = SM_K + 2 * 16384;      // 2 stages x 16 KB
constexpr int SM_P = SM_V + 2 * 16384;      // 2 buffers x 32 KB (two 64-column panels each)
constexpr int SM_BAR = SM_P + 2 * 32768;
constexpr int SM_MX = SM_BAR + 256;             // row-max exchange: [2 parity][2 halves][128] floats
constexpr int SM_TOTAL = SM_MX + 2 * 2 * 128 * 4 + 1024;
constexpr int kThreads = 320;

__global__ void __launch_bounds__(kThreads, 1)
attn_encoder_tc_kernel(const __grid_constant__ CUtensorMap tm, bf16* __restrict__ out, int T_len, int d) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SM_BAR);
  uint64_t* q_full = bars;          // 1
  uint64_t* kv_full = bars + 1;     // 2
  uint64_t* kv_empty = bars + 3;    // 2
  uint64_t* s_full = bars + 5;      // 2
  uint64_t* p_full = bars + 7;      // 2 (256 arrivals)
  uint64_t* o_full = bars + 9;      // 2
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 11);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * TQ, h = blockIdx.y, b = blockIdx.z;
  const int row_base = b * T_len;
  const int n_kt = (T_len + TK - 1) / TK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm);
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 256); mbar_init(&o_full[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr_smem, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const uint32_t tm_S = tmem_base;          // + 128 * buf
  const uint32_t tm_O = tmem_base + 256;    // + 64 * buf

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, 16384);
      tma_load_3d(smem + SM_Q, &tm, q_full, h * HD, row_base + q0, 0);
      for (int j = 0; j < n_kt; ++j) {
        const int s = j & 1;
        mbar_wait(&kv_empty[s], ((j >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&kv_full[s], 32768);
        tma_load_3d(smem + SM_K + s * 16384, &tm, &kv_full[s], d + h * HD, row_base + j * TK, 0);
        tma_load_3d(smem + SM_V + s * 16384, &tm, &kv_full[s], 2 * d + h * HD, row_base + j * TK, 0);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128, 0, 0);
      constexpr uint32_t idesc_o = umma_idesc_bf16(128, 64, 0, 1);  // B (= V) is MN-major
      const uint32_t sq = smem_u32(smem + SM_Q);
      auto issue_S = [&](int j) {
        const int s = j & 1;
        mbar_wait(&kv_full[s], (j >> 1) & 1);
        tc_fence_after();
        const uint64_t adesc = umma_smem_desc_sw128(sq, 16, 1024);
        const uint64_t bdesc = umma_smem_desc_sw128(smem_u32(smem + SM_K + s * 16384), 16, 1024);
#pragma unroll
        for (int k = 0; k < HD / 16; ++k) umma_f16(tm_S + s * 128, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc_s, k != 0);
        umma_commit(&s_full[s]);
      };
      mbar_wait(q_full, 0);
      issue_S(0);
      for (int j = 0; j < n_kt; ++j) {
        const int s = j & 1;
        if (j + 1 < n_kt) issue_S(j + 1);
        mbar_wait(&p_full[s], (j >> 1) & 1);
        tc_fence_after();
        const uint32_t sp = smem_u32(smem + SM_P + s * 32768);
        const uint32_t sv = smem_u32(smem + SM_V + s * 16384);
#pragma unroll
        for (int k = 0; k < TK / 16; ++k) {
          // A: P panel (k / 4), +32 B per 16 columns; B: V rows [16k, 16k+16) = 2 swizzle atoms of 1024 B
          const uint64_t adesc = umma_smem_desc_sw128(sp + (k >> 2) * 16384 + (k & 3) * 32, 16, 1024);
          const uint64_t bdesc = umma_smem_desc_sw128(sv + k * 2048, 1024, 1024);
          umma_f16(tm_O + s * 64, adesc, bdesc, idesc_o, k != 0);
        }
        umma_commit(&o_full[s]);
        umma_commit(&kv_empty[s]);
      }
    }
    __syncwarp();
  } else {
    const int wq = warp & 3;             // TMEM lane quarter this warp may touch
    const int half = (warp - 2) >> 2;    // which 64 key columns / 32 output columns
    const int row = wq * 32 + lane;
    const uint32_t lane_off = (uint32_t)(wq * 32) << 16;
    const float sl2 = 0.125f * 1.4426950408889634f;
    float* mxbuf = reinterpret_cast<float*>(smem + SM_MX);
    float m = -INFINITY, l = 0.f, alpha_prev = 0.f;
    float o[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) o[i] = 0.f;
    for (int j = 0; j < n_kt; ++j) {
      const int s = j & 1;
      mbar_wait(&s_full[s], (j >> 1) & 1);
      tc_fence_after();
      uint32_t sv[64];
      tmem_ld_32x32b_x32(tm_S + s * 128 + lane_off + half * 64, sv);
      tmem_ld_32x32b_x32(tm_S + s * 128 + lane_off + half * 64 + 32, sv + 32);
      tmem_ld_wait();
      const int kvalid = T_len - j * TK - half * 64;  // my columns >= kvalid are padding / the next window
      float mx = -INFINITY;
      if (kvalid >= 64) {
#pragma unroll
        for (int i = 0; i < 64; ++i) mx = fmaxf(mx, __uint_as_float(sv[i]));
      } else {
#pragma unroll
        for (int i = 0; i < 64; ++i) {
          if (i >= kvalid) sv[i] = 0xff800000u;  // -inf
          mx = fmaxf(mx, __uint_as_float(sv[i]));
        }
      }
      mxbuf[(s * 2 + half) * 128 + row] = mx;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      mx = fmaxf(mx, mxbuf[(s * 2 + (half ^ 1)) * 128 + row]);
      const float m_new = fmaxf(m, mx * sl2);
      const float alpha = fast_exp2(m - m_new);  // m = -inf on the first tile -> 0
      float lsum = 0.f;
      uint8_t* prow = smem + SM_P + s * 32768 + half * 16384 + row * 128;
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        float p[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          p[i] = fast_exp2(fmaf(__uint_as_float(sv[g * 8 + i]), sl2, -m_new));
          lsum += p[i];
        }
        uint4 t;
        t.x = pack_bf16x2(p[0], p[1]);
        t.y = pack_bf16x2(p[2], p[3]);
        t.z = pack_bf16x2(p[4], p[5]);
        t.w = pack_bf16x2(p[6], p[7]);
        *reinterpret_cast<uint4*>(prow + ((g ^ (row & 7)) << 4)) = t;
      }
      l = l * alpha + lsum;
      m = m_new;
      fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
      tc_fence_before();
      mbar_arrive(&p_full[s]);
      if (j > 0) {
        const int so = (j - 1) & 1;
        mbar_wait(&o_full[so], ((j - 1) >> 1) & 1);
        tc_fence_after();
        uint32_t r[32];
        tmem_ld_32x32b_x32(tm_O + so * 64 + lane_off + half * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) o[i] = fmaf(o[i], alpha_prev, __uint_as_float(r[i]));
      }
      alpha_prev = alpha;
    }
    {
      const int so = (n_kt - 1) & 1;
      mbar_wait(&o_full[so], ((n_kt - 1) >> 1) & 1);
      tc_fence_after();
      uint32_t r[32];
      tmem_ld_32x32b_x32(tm_O + so * 64 + lane_off + half * 32, r);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) o[i] = fmaf(o[i], alpha_prev, __uint_as_float(r[i]));
    }
    // total row sum = this half + the other half (both track the same running max)
    asm volatile("bar.sync 1, 256;" ::: "memory");
    mxbuf[half * 128 + row] = l;
    asm volatile("bar.sync 1, 256;" ::: "memory");
    l += mxbuf[(half ^ 1) * 128 + row];
    if (q0 + row < T_len) {
      const float inv = 1.f / l;
      bf16* orow = out + (long long)(row_base + q0 + row) * d + h * HD + half * 32;
#pragma unroll
      for (int i = 0; i < 32; i += 8) {
        uint4 t;
        t.x = pack_bf16x2(o[i] * inv, o[i + 1] * inv);
        t.y = pack_bf16x2(o[i + 2] * inv, o[i + 3] * inv);
        t.z = pack_bf16x2(o[i + 4] * inv, o[i + 5] * inv);
        t.w = pack_bf16x2(o[i + 6] * inv, o[i + 7] * inv);
        *reinterpret_cast<uint4*>(orow + i) = t;
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}


// ------------------------------------------------------------------------------------------------
// v3: two 128-query tiles (A, B) per CTA in ping-pong.  K/V tiles are loaded once for both; while the softmax
// warps of one tile work on CUDA cores / MUFU, the tensor core runs the other tile's S = Q K^T and O = P V, so
// neither pipe idles on the other's latency.  8 softmax warps per tile (4 per SM sub-partition overall), TMEM:
// S_A | S_B (128 columns each) + O_A | O_B (64 each); shared: Q_A,Q_B 32 KB + 3 K|V stages 96 KB + P_A,P_B 64 KB.
constexpr int PP_Q = 0;
constexpr int PP_KV = 32768;                       // 3 stages x (K 16 KB + V 16 KB)
constexpr int PP_STAGES = 3;
constexpr int PP_P = PP_KV + PP_STAGES * 32768;    // 2 x 32 KB
constexpr int PP_BAR = PP_P + 2 * 32768;
constexpr int PP_MX = PP_BAR + 256;               // row-max exchange: [2 tiles][2 parity][2 halves][128] floats
constexpr int PP_TOTAL = PP_MX + 2 * 512 * 4 + 1024;
constexpr int PP_THREADS = 576;

__global__ void __launch_bounds__(PP_THREADS, 1)
attn_encoder_pp_kernel(const __grid_constant__ CUtensorMap tm, bf16* __restrict__ out, int T_len, int d) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + PP_BAR);
  uint64_t* q_full = bars;           // 1
  uint64_t* kv_full = bars + 1;      // 3
  uint64_t* kv_empty = bars + 4;     // 3
  uint64_t* s_full = bars + 7;       // 2 (per tile)
  uint64_t* p_full = bars + 9;       // 2 (128 arrivals each)
  uint64_t* o_full = bars + 11;      // 2
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 13);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 256, h = blockIdx.y, b = blockIdx.z;
  const int row_base = b * T_len;
  const int n_kt = (T_len + TK - 1) / TK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm);
    mbar_init(q_full, 1);
    for (int i = 0; i < PP_STAGES; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&s_full[i], 1); mbar_init(&p_full[i], 256); mbar_init(&o_full[i], 1); }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr_smem, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const uint32_t tm_S = tmem_base;         // + 128 * tile
  const uint32_t tm_O = tmem_base + 256;   // + 64 * tile

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, 32768);
      tma_load_3d(smem + PP_Q, &tm, q_full, h * HD, row_base + q0, 0);
      tma_load_3d(smem + PP_Q + 16384, &tm, q_full, h * HD, row_base + q0 + 128, 0);
      for (int j = 0; j < n_kt; ++j) {
        const int s = j % PP_STAGES;
        mbar_wait(&kv_empty[s], ((j / PP_STAGES) & 1) ^ 1);
        mbar_arrive_expect_tx(&kv_full[s], 32768);
        tma_load_3d(smem + PP_KV + s * 32768, &tm, &kv_full[s], d + h * HD, row_base + j * TK, 0);
        tma_load_3d(smem + PP_KV + s * 32768 + 16384, &tm, &kv_full[s], 2 * d + h * HD, row_base + j * TK, 0);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128, 0, 0);
      constexpr uint32_t idesc_o = umma_idesc_bf16(128, 64, 0, 1);  // B (= V) is MN-major
      auto issue_S = [&](int tile, int j) {  // caller made sure K_j has landed
        const uint64_t adesc = umma_smem_desc_sw128(smem_u32(smem + PP_Q + tile * 16384), 16, 1024);
        const uint64_t bdesc = umma_smem_desc_sw128(smem_u32(smem + PP_KV + (j % PP_STAGES) * 32768), 16, 1024);
#pragma unroll
        for (int k = 0; k < HD / 16; ++k) umma_f16(tm_S + tile * 128, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc_s, k != 0);
        umma_commit(&s_full[tile]);
      };
      auto issue_PV = [&](int tile, int j) {
        mbar_wait(&p_full[tile], j & 1);
        tc_fence_after();
        const uint32_t sp = smem_u32(smem + PP_P + tile * 32768);
        const uint32_t sv = smem_u32(smem + PP_KV + (j % PP_STAGES) * 32768 + 16384);
#pragma unroll
        for (int k = 0; k < TK / 16; ++k) {
          const uint64_t adesc = umma_smem_desc_sw128(sp + (k >> 2) * 16384 + (k & 3) * 32, 16, 1024);
          const uint64_t bdesc = umma_smem_desc_sw128(sv + k * 2048, 1024, 1024);
          umma_f16(tm_O + tile * 64, adesc, bdesc, idesc_o, k != 0);
        }
        umma_commit(&o_full[tile]);
      };
      mbar_wait(q_full, 0);
      mbar_wait(&kv_full[0], 0);
      tc_fence_after();
      issue_S(0, 0);
      issue_S(1, 0);
      for (int j = 0; j < n_kt; ++j) {
        const bool more = j + 1 < n_kt;
        issue_PV(0, j);
        if (more) {
          mbar_wait(&kv_full[(j + 1) % PP_STAGES], ((j + 1) / PP_STAGES) & 1);
          tc_fence_after();
          issue_S(0, j + 1);
        }
        issue_PV(1, j);
        umma_commit(&kv_empty[j % PP_STAGES]);  // K_j / V_j fully consumed by both tiles
        if (more) issue_S(1, j + 1);
      }
    }
    __syncwarp();
  } else {
    // 16 softmax warps: warps 2..9 -> tile A, 10..17 -> tile B; inside a tile two warps share each TMEM lane
    // quarter (= SM sub-partition) and split the 128 key columns / 64 output columns between them.
    const int tile = (warp - 2) >> 3;
    const int half = ((warp - 2) >> 2) & 1;
    const int wq = warp & 3;
    const int row = wq * 32 + lane;
    const uint32_t lane_off = (uint32_t)(wq * 32) << 16;
    const float sl2 = 0.125f * 1.4426950408889634f;
    const uint32_t my_S = tm_S + tile * 128 + lane_off + half * 64, my_O = tm_O + tile * 64 + lane_off + half * 32;
    uint8_t* prow = smem + PP_P + tile * 32768 + half * 16384 + row * 128;  // panel `half` of P
    float* mxbuf = reinterpret_cast<float*>(smem + PP_MX) + tile * 512;     // [2 parity][2 halves][128]
    float m = -INFINITY, l = 0.f, alpha_prev = 0.f;
    float o[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) o[i] = 0.f;
    for (int j = 0; j < n_kt; ++j) {
      mbar_wait(&s_full[tile], j & 1);
      tc_fence_after();
      const int kvalid = T_len - j * TK - half * 64;  // my 64 columns: those >= kvalid are padding / next window
      float mx = -INFINITY;
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(my_S + c * 32, r);
        tmem_ld_wait();
        if (kvalid >= 64) {
#pragma unroll
          for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(r[i]));
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) mx = fmaxf(mx, (c * 32 + i < kvalid) ? __uint_as_float(r[i]) : -INFINITY);
        }
      }
      mxbuf[((j & 1) * 2 + half) * 128 + row] = mx;
      asm volatile("bar.sync %0, 256;" ::"r"(1 + tile) : "memory");
      mx = fmaxf(mx, mxbuf[((j & 1) * 2 + (half ^ 1)) * 128 + row]);
      const float m_new = fmaxf(m, mx * sl2);
      const float alpha = fast_exp2(m - m_new);
      if (j > 0) {  // PV_{j-1} has retired: fold it in (and P may be overwritten below)
        mbar_wait(&o_full[tile], (j - 1) & 1);
        tc_fence_after();
        uint32_t r[32];
        tmem_ld_32x32b_x32(my_O, r);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) o[i] = fmaf(o[i], alpha_prev, __uint_as_float(r[i]));
      }
      float lsum = 0.f;
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(my_S + c * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint32_t pb[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float e = fast_exp2(fmaf(__uint_as_float(r[g * 8 + i]), sl2, -m_new));
            if (kvalid < 64 && c * 32 + g * 8 + i >= kvalid) e = 0.f;
            pb[i] = bf16_round_bits(e);  // integer-ALU rounding; the row sum uses the ROUNDED values
            lsum += __uint_as_float(pb[i]);
          }
          *reinterpret_cast<uint4*>(prow + (((c * 4 + g) ^ (row & 7)) << 4)) =
              make_uint4(pack_bf16x2_bits(pb[0], pb[1]), pack_bf16x2_bits(pb[2], pb[3]), pack_bf16x2_bits(pb[4], pb[5]),
                         pack_bf16x2_bits(pb[6], pb[7]));
        }
      }
      l = l * alpha + lsum;
      m = m_new;
      alpha_prev = alpha;
      fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
      tc_fence_before();
      mbar_arrive(&p_full[tile]);
    }
    mbar_wait(&o_full[tile], (n_kt - 1) & 1);
    tc_fence_after();
    {
      uint32_t r[32];
      tmem_ld_32x32b_x32(my_O, r);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) o[i] = fmaf(o[i], alpha_prev, __uint_as_float(r[i]));
    }
    // total row sum = both column halves (they track the same running max)
    asm volatile("bar.sync %0, 256;" ::"r"(1 + tile) : "memory");
    mxbuf[half * 128 + row] = l;
    asm volatile("bar.sync %0, 256;" ::"r"(1 + tile) : "memory");
    l += mxbuf[(half ^ 1) * 128 + row];
    const int qrow = q0 + tile * 128 + row;
    if (qrow < T_len) {
      const float inv = 1.f / l;
      bf16* orow = out + (long long)(row_base + qrow) * d + h * HD + half * 32;
#pragma unroll
      for (int i = 0; i < 32; i += 8) {
        uint4 t;
        t.x = pack_bf16x2(o[i] * inv, o[i + 1] * inv); t.y = pack_bf16x2(o[i + 2] * inv, o[i + 3] * inv);
        t.z = pack_bf16x2(o[i + 4] * inv, o[i + 5] * inv); t.w = pack_bf16x2(o[i + 6] * inv, o[i + 7] * inv);
        *reinterpret_cast<uint4*>(orow + i) = t;
      }
    }
    tc_fence_before();
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
  static std::atomic<unsigned long long> attr_set{0};
  int dev = 0;
  BW_CUDA(cudaGetDevice(&dev));
  if (!(attr_set.load() >> dev & 1ull)) {
    BW_CUDA(cudaFuncSetAttribute(attn_encoder_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TOTAL));
    attr_set.fetch_or(1ull << dev);
  }
  CUtensorMap tm = make_operand_map(qkv, batch * T_len, 3 * d, 3 * d, 1, 0, 128);
  static const bool v2 = getenv("B200W_ATTN_V2") != nullptr;
  if (!v2) {
    static std::atomic<unsigned long long> pp_set{0};
    if (!(pp_set.load() >> dev & 1ull)) {
      BW_CUDA(cudaFuncSetAttribute(attn_encoder_pp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PP_TOTAL));
      pp_set.fetch_or(1ull << dev);
    }
    dim3 grid_pp((T_len + 255) / 256, n_head, batch);
    attn_encoder_pp_kernel<<<grid_pp, PP_THREADS, PP_TOTAL, stream>>>(tm, out, T_len, d);
    BW_CUDA(cudaGetLastError());
    ++g_kernel_launches;
    return;
  }
  dim3 grid((T_len + TQ - 1) / TQ, n_head, batch);
  attn_encoder_tc_kernel<<<grid, kThreads, SM_TOTAL, stream>>>(tm, out, T_len, d);
  BW_CUDA(cudaGetLastError());
  ++g_kernel_launches;
}

}  // namespace bw
