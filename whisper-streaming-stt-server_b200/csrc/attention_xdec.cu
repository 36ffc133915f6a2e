// Decoder cross-attention, bf16 product path: the HBM-bound kernel of the batched decoder step.
//
// The cached encoder K/V of a segment ([T_enc, 2d] per layer) is streamed ONCE per step for all hypotheses
// (beams) of that segment.  The SM's only job is to keep loads in flight, so the data path is
//   TMA (cp.async.bulk.tensor, SWIZZLE_128B, 6-stage mbarrier ring of 64-key K|V tiles = 96 KB per CTA)
//   -> ldmatrix -> mma.sync.m16n8k16 (q.K^T and P.V; rows 0..7 of the M=16 fragment are the <= 8 hypotheses)
// which needs ~7x fewer instructions per byte than the SIMT version (profiles/r1_xattn_simt_ncu_full.txt).
// One CTA = (head, segment group, T split); 4 consumer warps each own 16 of the 64 keys of a tile and keep
// their own running (max, sum, O); a 5th warp is the TMA producer.  Partials over warps are merged through
// shared memory, partials over T splits through the workspace + dec_cross_combine kernel (attention.cu).
// Upstream: whisper/model.py MultiHeadAttention.forward with cached cross K/V (kv_cache hooks).
#include <cuda.h>

#include "kernels.cuh"

namespace bw {

CUtensorMap make_box_map(const void* ptr, int inner, int rows, int outer, long long row_stride, long long outer_stride,
                         int box_inner, int box_rows);

namespace {

constexpr int XT = 64;          // keys per tile
constexpr int XSTAGES = 6;
constexpr int XTILE_BYTES = XT * 128;  // one 64x64 bf16 tile
constexpr int XSM_BAR = XSTAGES * 2 * XTILE_BYTES;
constexpr int XSM_RED = XSM_BAR + 128;
constexpr int XSM_TOTAL = XSM_RED + 4 * 8 * 66 * 4 + 1024;
constexpr int kMaxSplitX = 8;

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float* c, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

__global__ void __launch_bounds__(160, 2)
dec_cross_attention_mma_kernel(const __grid_constant__ CUtensorMap tm, const int* __restrict__ group_first_row,
                               const int* __restrict__ group_n_rows, const int* __restrict__ group_xslot,
                               const float* __restrict__ q, int T_enc, int n_layer, int layer, int d, int n_split,
                               bf16* __restrict__ out, float* __restrict__ ws, unsigned long long* trace_buf) {
  extern __shared__ uint8_t smem_raw[];
  unsigned long long* trace = nullptr;
  unsigned ttag = 4u << 24;
  if (threadIdx.x == 0) {
    if ((blockIdx.x | blockIdx.y | blockIdx.z) == 0) trace = trace_buf;
    else if (blockIdx.x == gridDim.x - 1 && blockIdx.y == gridDim.y - 1 && blockIdx.z == gridDim.z - 1) { trace = trace_buf; ttag = 5u << 24; }
  }
  trace_mark(trace, ttag | 1);
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + XSM_BAR);
  uint64_t* empty_bar = full_bar + XSTAGES;
  float* red = reinterpret_cast<float*>(smem + XSM_RED);  // [4 warps][8 rows][66]: m, l, o[64]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.x, g = blockIdx.y, sp = blockIdx.z, n_head = gridDim.x;
  const int row0 = group_first_row[g], nq = group_n_rows[g];
  const int chunk = ((T_enc + n_split - 1) / n_split + XT - 1) / XT * XT;
  const int t0 = sp * chunk, t1 = min(T_enc, t0 + chunk);
  const int n_tiles = (t1 - t0 + XT - 1) / XT;  // may be 0 for a trailing split

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm);
    for (int s = 0; s < XSTAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 4); }
    fence_barrier_init();
  }
  __syncthreads();
  // PDL: the cached encoder K/V is never written during a decoder step, so the producer starts streaming it
  // right away; only the consumers (which read q and write out / ws) wait for the preceding kernels.
  pdl_trigger();

  if (warp == 4) {
    if (lane == 0) {
      const int zc = group_xslot[g] * n_layer + layer;
      for (int it = 0; it < n_tiles; ++it) {
        const int s = it % XSTAGES;
        mbar_wait(&empty_bar[s], ((it / XSTAGES) & 1) ^ 1);
        mbar_arrive_expect_tx(&full_bar[s], 2 * XTILE_BYTES);
        uint8_t* ks = smem + s * 2 * XTILE_BYTES;
        tma_load_3d(ks, &tm, &full_bar[s], h * 64, t0 + it * XT, zc);
        tma_load_3d(ks + XTILE_BYTES, &tm, &full_bar[s], d + h * 64, t0 + it * XT, zc);
      }
    }
    __syncwarp();
  } else {
    pdl_wait();
    trace_mark(trace, ttag | 2);
    const int gid = lane >> 2, tig = lane & 3;  // query row (hypothesis) and column pair inside an n-tile
    // A fragments of q (rows 8..15 are zero): 4 k-steps of 16 dims; 1/sqrt(64) folded in (exact in bf16)
    uint32_t qa[4][2];
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      float v0 = 0.f, v1 = 0.f, v2 = 0.f, v3 = 0.f;
      if (gid < nq) {
        const float* qr = q + (long long)(row0 + gid) * d + h * 64 + ks * 16 + tig * 2;
        v0 = qr[0] * 0.125f; v1 = qr[1] * 0.125f; v2 = qr[8] * 0.125f; v3 = qr[9] * 0.125f;
      }
      qa[ks][0] = pack_bf16x2(v0, v1);
      qa[ks][1] = pack_bf16x2(v2, v3);
    }
    const float LOG2E = 1.4426950408889634f;
    float m = -INFINITY, l = 0.f;  // running max (log2 domain) and this thread's partial row sum
    float o[8][4];
#pragma unroll
    for (int nd = 0; nd < 8; ++nd) { o[nd][0] = o[nd][1] = o[nd][2] = o[nd][3] = 0.f; }
    // ldmatrix lane roles
    const int lrow = lane & 7, lmat = lane >> 3;
    for (int it = 0; it < n_tiles; ++it) {
      const int s = it % XSTAGES;
      mbar_wait(&full_bar[s], (it / XSTAGES) & 1);
      const uint32_t kbase = smem_u32(smem + s * 2 * XTILE_BYTES);
      const uint32_t vbase = kbase + XTILE_BYTES;
      // ---- S[16 x 16 keys] = Q . K^T for this warp's keys [warp*16, +16) ----
      float sc[2][4];
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
        sc[nt][0] = sc[nt][1] = sc[nt][2] = sc[nt][3] = 0.f;
        const int krow = warp * 16 + nt * 8 + lrow;
#pragma unroll
        for (int half = 0; half < 2; ++half) {  // dims [32*half, +32): matrices = 16-byte chunks 4*half + lmat
          uint32_t b0, b1, b2, b3;
          ldsm_x4(kbase + krow * 128 + (((half * 4 + lmat) ^ (krow & 7)) << 4), b0, b1, b2, b3);
          mma_bf16(sc[nt], qa[half * 2][0], 0u, qa[half * 2][1], 0u, b0, b1);
          mma_bf16(sc[nt], qa[half * 2 + 1][0], 0u, qa[half * 2 + 1][1], 0u, b2, b3);
        }
      }
      // ---- online softmax for query row gid over these 16 keys (4 of them in this thread) ----
      const int key0 = t0 + it * XT + warp * 16 + tig * 2;
      float sv[4];
      sv[0] = (key0 < t1) ? sc[0][0] * LOG2E : -INFINITY;
      sv[1] = (key0 + 1 < t1) ? sc[0][1] * LOG2E : -INFINITY;
      sv[2] = (key0 + 8 < t1) ? sc[1][0] * LOG2E : -INFINITY;
      sv[3] = (key0 + 9 < t1) ? sc[1][1] * LOG2E : -INFINITY;
      float mx = fmaxf(fmaxf(sv[0], sv[1]), fmaxf(sv[2], sv[3]));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
      const float m_new = fmaxf(m, mx);
      float p[4] = {0.f, 0.f, 0.f, 0.f};
      if (m_new != -INFINITY) {
        const float alpha = fast_exp2(m - m_new);
#pragma unroll
        for (int i = 0; i < 4; ++i) p[i] = fast_exp2(sv[i] - m_new);
        l = l * alpha + (p[0] + p[1]) + (p[2] + p[3]);
        if (alpha != 1.f) {
#pragma unroll
          for (int nd = 0; nd < 8; ++nd) { o[nd][0] *= alpha; o[nd][1] *= alpha; }
        }
        m = m_new;
      }
      const uint32_t pa0 = pack_bf16x2(p[0], p[1]), pa2 = pack_bf16x2(p[2], p[3]);
      // ---- O[16 x 64] += P[16 x 16 keys] . V[16 keys x 64 dims] ----
#pragma unroll
      for (int np = 0; np < 4; ++np) {  // pairs of 8-dim n-tiles
        const int vrow = warp * 16 + (lmat & 1) * 8 + lrow;
        uint32_t b0, b1, b2, b3;
        ldsm_x4_t(vbase + vrow * 128 + (((np * 2 + (lmat >> 1)) ^ (vrow & 7)) << 4), b0, b1, b2, b3);
        mma_bf16(o[np * 2], pa0, 0u, pa2, 0u, b0, b1);
        mma_bf16(o[np * 2 + 1], pa0, 0u, pa2, 0u, b2, b3);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_bar[s]);
    }
    // row sum over the 4 threads of a quad, then park this warp's partial
    l += __shfl_xor_sync(0xffffffffu, l, 1);
    l += __shfl_xor_sync(0xffffffffu, l, 2);
    float* mine = red + (warp * 8 + gid) * 66;
    if (tig == 0) { mine[0] = m; mine[1] = l; }
#pragma unroll
    for (int nd = 0; nd < 8; ++nd) { mine[2 + nd * 8 + tig * 2] = o[nd][0]; mine[2 + nd * 8 + tig * 2 + 1] = o[nd][1]; }
  }
  __syncthreads();
  // merge the 4 warps: thread -> (row, 4 dims)
  if (threadIdx.x < 128) {
    const int r = threadIdx.x >> 4, c4 = (threadIdx.x & 15) * 4;
    if (r < nq) {
      float M = -INFINITY;
#pragma unroll
      for (int w = 0; w < 4; ++w) M = fmaxf(M, red[(w * 8 + r) * 66]);
      float num[4] = {0.f, 0.f, 0.f, 0.f}, den = 0.f;
      if (M != -INFINITY) {
#pragma unroll
        for (int w = 0; w < 4; ++w) {
          const float* pr = red + (w * 8 + r) * 66;
          const float e = fast_exp2(pr[0] - M);
          den = fmaf(e, pr[1], den);
#pragma unroll
          for (int i = 0; i < 4; ++i) num[i] = fmaf(e, pr[2 + c4 + i], num[i]);
        }
      }
      const int row = row0 + r;
      if (n_split == 1) {
        const float inv = 1.f / den;
        uint2 t;
        t.x = pack_bf16x2(num[0] * inv, num[1] * inv);
        t.y = pack_bf16x2(num[2] * inv, num[3] * inv);
        *reinterpret_cast<uint2*>(out + (long long)row * d + h * 64 + c4) = t;
      } else {
        float* w = ws + (((long long)row * n_head + h) * kMaxSplitX + sp) * 66;
        if (c4 == 0) { w[0] = (M == -INFINITY) ? -INFINITY : M * 0.6931471805599453f; w[1] = den; }  // natural-log max
#pragma unroll
        for (int i = 0; i < 4; ++i) w[2 + c4 + i] = num[i];
      }
    }
  }
  trace_mark(trace, ttag | 8);
}

}  // namespace

void dec_cross_attention_mma(const int* group_first_row, const int* group_n_rows, const int* group_xslot, int n_groups,
                             const float* q, const CrossKV& kv, int n_layer, int layer, int d, int n_head, int n_split,
                             bf16* out, float* ws, cudaStream_t stream) {
  static std::atomic<unsigned long long> attr_set{0};
  int dev = 0;
  BW_CUDA(cudaGetDevice(&dev));
  if (!(attr_set.load() >> dev & 1ull)) {
    BW_CUDA(cudaFuncSetAttribute(dec_cross_attention_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, XSM_TOTAL));
    attr_set.fetch_or(1ull << dev);
  }
  // one map per (cache pointer, geometry); cheap enough to rebuild, but the decoder calls this 32x per step
  static thread_local const void* cached_ptr = nullptr;
  static thread_local int cached_geo[4] = {0, 0, 0, 0};
  static thread_local CUtensorMap cached_tm;
  if (cached_ptr != kv.cache || cached_geo[0] != d || cached_geo[1] != kv.T_enc || cached_geo[2] != kv.n_slots || cached_geo[3] != n_layer) {
    cached_tm = make_box_map(kv.cache, 2 * d, kv.T_enc, kv.n_slots * n_layer, 2LL * d, (long long)kv.T_enc * 2 * d, 64, XT);
    cached_ptr = kv.cache; cached_geo[0] = d; cached_geo[1] = kv.T_enc; cached_geo[2] = kv.n_slots; cached_geo[3] = n_layer;
  }
  const CUtensorMap tm = cached_tm;
  dim3 grid(n_head, n_groups, n_split);
  launch_kernel(dec_cross_attention_mma_kernel, grid, dim3(160), XSM_TOTAL, stream, tm, group_first_row, group_n_rows, group_xslot, q,
                kv.T_enc, n_layer, layer, d, n_split, out, ws, g_trace_dev);
  ++g_kernel_launches;
}

}  // namespace bw
