// LayerNorm, conv2 im2col, embedding lookup, KV append and dtype conversion kernels.
// Upstream counterparts: whisper/model.py LayerNorm (fp32 statistics), Conv1d stem, TextDecoder
// token/positional embedding (reached from reference torch_whisper.py:55).
#include "kernels.cuh"

namespace bw {
namespace {

// One 128-thread block per row, float4 loads (d % 4 == 0, d <= 1536), two-pass mean / centred variance
// in registers like torch's kernel.  Many small blocks keep enough loads in flight both for the encoder
// (tens of thousands of rows) and for the decoder step (a few hundred rows).
template <typename TOut, bool GATHER>
__global__ void __launch_bounds__(128)
layernorm_kernel(const float* __restrict__ x, const int* __restrict__ rows_idx, const float* __restrict__ gamma,
                 const float* __restrict__ beta, TOut* __restrict__ out, int rows, int d, unsigned long long* trace_buf) {
  __shared__ float red[4];
  unsigned long long* const trace = (blockIdx.x == 0 && threadIdx.x == 0) ? trace_buf : nullptr;
  trace_mark(trace, (2u << 24) | 1);
  pdl_trigger();
  pdl_wait();
  trace_mark(trace, (2u << 24) | 2);
  const int row = blockIdx.x;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const float4* xr = reinterpret_cast<const float4*>(x + (long long)(GATHER ? rows_idx[row] : row) * d);
  const int nv = d >> 2;
  float4 v[3];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const int c = tid + i * 128;
    v[i] = (c < nv) ? xr[c] : make_float4(0.f, 0.f, 0.f, 0.f);
    s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  }
  s = warp_sum(s);
  if (lane == 0) red[warp] = s;
  __syncthreads();
  const float mean = (red[0] + red[1] + red[2] + red[3]) / d;
  __syncthreads();
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const int c = tid + i * 128;
    if (c < nv) {
      const float a = v[i].x - mean, b = v[i].y - mean, cc = v[i].z - mean, dd = v[i].w - mean;
      q += (a * a + b * b) + (cc * cc + dd * dd);
    }
  }
  q = warp_sum(q);
  if (lane == 0) red[warp] = q;
  __syncthreads();
  const float rstd = rsqrtf((red[0] + red[1] + red[2] + red[3]) / d + 1e-5f);
  TOut* orow = out + (long long)row * d;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const int c = tid + i * 128;
    if (c < nv) {
      const float4 g = reinterpret_cast<const float4*>(gamma)[c], b = reinterpret_cast<const float4*>(beta)[c];
      const float o0 = (v[i].x - mean) * rstd * g.x + b.x, o1 = (v[i].y - mean) * rstd * g.y + b.y;
      const float o2 = (v[i].z - mean) * rstd * g.z + b.z, o3 = (v[i].w - mean) * rstd * g.w + b.w;
      if constexpr (sizeof(TOut) == 4) {
        reinterpret_cast<float4*>(orow)[c] = make_float4(o0, o1, o2, o3);
      } else {
        uint2 t;
        t.x = pack_bf16x2(o0, o1);
        t.y = pack_bf16x2(o2, o3);
        reinterpret_cast<uint2*>(orow)[c] = t;
      }
    }
  }
  trace_mark(trace, (2u << 24) | 8);
}

template <typename T>
__global__ void im2col_conv2_kernel(const T* __restrict__ y1, T* __restrict__ A2, int batch, int d) {
  // one thread per 8-element (or 4 for float) vector; consecutive threads -> consecutive channels
  constexpr int VEC = 16 / sizeof(T);
  const long long dv = d / VEC;
  const long long total = (long long)batch * 1500 * 3 * dv;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int cv = (int)(i % dv);
  const int k = (int)((i / dv) % 3);
  const int t = (int)((i / (3 * dv)) % 1500);
  const int b = (int)(i / (3 * dv * 1500));
  const int src_t = 2 * t + k - 1;
  uint4 v = make_uint4(0, 0, 0, 0);
  if (src_t >= 0 && src_t < 3000) v = *reinterpret_cast<const uint4*>(y1 + ((long long)b * 3000 + src_t) * d + cv * VEC);
  *reinterpret_cast<uint4*>(A2 + ((long long)b * 1500 + t) * 3 * d + (long long)k * d + cv * VEC) = v;
}

template <typename T>
__global__ void convert_kernel(const float* __restrict__ src, T* __restrict__ dst, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = from_f<T>(src[i]);
}
__global__ void f32_from_bf16_kernel(const bf16* __restrict__ src, float* __restrict__ dst, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = __bfloat162float(src[i]);
}
template <typename T>
__global__ void permute_conv_kernel(const float* __restrict__ src, T* __restrict__ dst, int co, int ci) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)co * ci * 3) return;
  const int c = (int)(i % ci);
  const int k = (int)((i / ci) % 3);
  const int o = (int)(i / (3LL * ci));
  dst[i] = from_f<T>(src[((long long)o * ci + c) * 3 + k]);
}

template <typename T>
__global__ void dec_embed_kernel(const int* __restrict__ row_seq, const int* __restrict__ row_pos,
                                 const int* __restrict__ row_tok, const int* __restrict__ next_tok,
                                 const T* __restrict__ tok_emb, const T* __restrict__ pos_emb, float* __restrict__ x, int d,
                                 const int* __restrict__ row_page, int* __restrict__ page_table, int n_blocks) {
  pdl_trigger();
  pdl_wait();
  const int r = blockIdx.x;
  int tok = row_tok[r];
  if (tok < 0) tok = next_tok[row_seq[r]];
  const int pos = row_pos[r];
  if (threadIdx.x == 0 && page_table) page_table[row_seq[r] * n_blocks + pos / kPageTokens] = row_page[r];
  for (int c = threadIdx.x; c < d; c += blockDim.x)
    x[(long long)r * d + c] = to_f(tok_emb[(long long)tok * d + c]) + to_f(pos_emb[(long long)pos * d + c]);
}

// Decoder LayerNorm fusion, row side: bf16 copy of a freshly written residual row + its per-64-column-tile LayerNorm
// partials (mean, M2 of the bf16-ROUNDED values), the format gemm_tc_rows' consumer epilogue combines.  `vals` = the
// row in shared memory; warp w reduces tiles w, w + n_warps, ...
__device__ __forceinline__ void row_xb_and_ln_partials(const float* vals, int d, bf16* xb_row, float2* st_row) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, n_warps = blockDim.x >> 5;
  for (int tile = warp; tile < d / 64; tile += n_warps) {
    const bf16 b0 = __float2bfloat16(vals[tile * 64 + lane]), b1 = __float2bfloat16(vals[tile * 64 + 32 + lane]);
    xb_row[tile * 64 + lane] = b0;
    xb_row[tile * 64 + 32 + lane] = b1;
    const float r0 = __bfloat162float(b0), r1 = __bfloat162float(b1);
    const float tm = warp_sum(r0 + r1) * (1.f / 64.f);
    const float dv = warp_sum((r0 - tm) * (r0 - tm) + (r1 - tm) * (r1 - tm));
    if (lane == 0) st_row[tile] = make_float2(tm, dv);
  }
}

template <typename T>
__global__ void dec_embed_ln_kernel(const int* __restrict__ row_seq, const int* __restrict__ row_pos,
                                    const int* __restrict__ row_tok, const int* __restrict__ next_tok,
                                    const T* __restrict__ tok_emb, const T* __restrict__ pos_emb, float* __restrict__ x, int d,
                                    bf16* __restrict__ xb, float2* __restrict__ stats, const int* __restrict__ row_page,
                                    int* __restrict__ page_table, int n_blocks) {
  extern __shared__ float embed_row[];
  pdl_trigger();
  pdl_wait();
  const int r = blockIdx.x;
  int tok = row_tok[r];
  if (tok < 0) tok = next_tok[row_seq[r]];
  const int pos = row_pos[r];
  if (threadIdx.x == 0 && page_table) page_table[row_seq[r] * n_blocks + pos / kPageTokens] = row_page[r];
  for (int c = threadIdx.x; c < d; c += blockDim.x) {
    const float v = to_f(tok_emb[(long long)tok * d + c]) + to_f(pos_emb[(long long)pos * d + c]);
    x[(long long)r * d + c] = v;
    embed_row[c] = v;
  }
  __syncthreads();
  row_xb_and_ln_partials(embed_row, d, xb + (long long)r * d, stats + (long long)r * (d / 64));
}

// fp32 rows -> bf16 copy + LayerNorm partials (test hook / any producer that is not a row GEMM)
__global__ void rows_ln_partials_kernel(const float* __restrict__ x, int d, bf16* __restrict__ xb, float2* __restrict__ stats) {
  extern __shared__ float embed_row[];
  const int r = blockIdx.x;
  for (int c = threadIdx.x; c < d; c += blockDim.x) embed_row[c] = x[(long long)r * d + c];
  __syncthreads();
  row_xb_and_ln_partials(embed_row, d, xb + (long long)r * d, stats + (long long)r * (d / 64));
}

// Fold a LayerNorm into the Linear that consumes it (one block per output row n):
//   Wf[n][k] = bf16(W[n][k] * gamma[k]),  c1[n] = sum_k float(Wf[n][k]),  c2[n] = sum_k beta[k] * W[n][k] + bias[n]
// so that LN(x).W^T + b = rstd * (x.Wf^T - mean * c1) + c2 with mean / rstd of the row of x.
__global__ void __launch_bounds__(256)
fold_ln_kernel(const float* __restrict__ W, const float* __restrict__ gamma, const float* __restrict__ beta,
               const float* __restrict__ bias, int K, bf16* __restrict__ Wf, float* __restrict__ c1, float* __restrict__ c2) {
  __shared__ float red[2][8];
  const int n = blockIdx.x;
  const float* w = W + (long long)n * K;
  float s1 = 0.f, s2 = 0.f;
  for (int k = threadIdx.x; k < K; k += 256) {
    const float wk = w[k];
    const bf16 f = __float2bfloat16(wk * gamma[k]);
    Wf[(long long)n * K + k] = f;
    s1 += __bfloat162float(f);
    s2 = fmaf(beta[k], wk, s2);
  }
  s1 = warp_sum(s1);
  s2 = warp_sum(s2);
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = s1; red[1][threadIdx.x >> 5] = s2; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, b = 0.f;
    for (int i = 0; i < 8; ++i) { a += red[0][i]; b += red[1][i]; }
    c1[n] = a;
    c2[n] = b + (bias ? bias[n] : 0.f);
  }
}


inline unsigned blocks_for(long long n, int t) { return (unsigned)((n + t - 1) / t); }

}  // namespace

template <typename T>
void layernorm(const float* x, const float* gamma, const float* beta, T* out, int rows, int d, cudaStream_t stream) {
  BW_CHECK(d <= 1536 && d % 4 == 0, "layernorm supports d <= 1536, d % 4 == 0");
  if (rows <= 0) return;
  launch_kernel(layernorm_kernel<T, false>, dim3(rows), dim3(128), 0, stream, x, (const int*)nullptr, gamma, beta, out, rows, d, g_trace_dev);
  ++g_kernel_launches;
}
template void layernorm<float>(const float*, const float*, const float*, float*, int, int, cudaStream_t);
template void layernorm<bf16>(const float*, const float*, const float*, bf16*, int, int, cudaStream_t);

void layernorm_f32out(const float* x, const float* gamma, const float* beta, float* out, int rows, int d, cudaStream_t stream) {
  layernorm<float>(x, gamma, beta, out, rows, d, stream);
}

template <typename T>
void layernorm_gather(const float* x, const int* rows_idx, const float* gamma, const float* beta, T* out, int n, int d,
                      cudaStream_t stream) {
  BW_CHECK(d <= 1536 && d % 4 == 0, "layernorm supports d <= 1536, d % 4 == 0");
  if (n <= 0) return;
  launch_kernel(layernorm_kernel<T, true>, dim3(n), dim3(128), 0, stream, x, rows_idx, gamma, beta, out, n, d, g_trace_dev);
  ++g_kernel_launches;
}
template void layernorm_gather<float>(const float*, const int*, const float*, const float*, float*, int, int, cudaStream_t);
template void layernorm_gather<bf16>(const float*, const int*, const float*, const float*, bf16*, int, int, cudaStream_t);

template <typename T>
void im2col_conv2(const T* y1, T* A2, int batch, int d, cudaStream_t stream) {
  constexpr int VEC = 16 / sizeof(T);
  BW_CHECK(d % VEC == 0, "d must be a multiple of the vector width");
  const long long total = (long long)batch * 1500 * 3 * (d / VEC);
  im2col_conv2_kernel<T><<<blocks_for(total, 256), 256, 0, stream>>>(y1, A2, batch, d);
  BW_CUDA(cudaGetLastError());
  ++g_kernel_launches;
}
template void im2col_conv2<float>(const float*, float*, int, int, cudaStream_t);
template void im2col_conv2<bf16>(const bf16*, bf16*, int, int, cudaStream_t);

template <typename T>
void convert_f32(const float* src, T* dst, long long n, cudaStream_t stream) {
  if (n <= 0) return;
  convert_kernel<T><<<blocks_for(n, 256), 256, 0, stream>>>(src, dst, n);
  BW_CUDA(cudaGetLastError());
}
template void convert_f32<float>(const float*, float*, long long, cudaStream_t);
template void convert_f32<bf16>(const float*, bf16*, long long, cudaStream_t);

void f32_from_bf16(const bf16* src, float* dst, long long n, cudaStream_t stream) {
  if (n <= 0) return;
  f32_from_bf16_kernel<<<blocks_for(n, 256), 256, 0, stream>>>(src, dst, n);
  BW_CUDA(cudaGetLastError());
}

template <typename T>
void permute_conv_weight(const float* src, T* dst, int co, int ci, cudaStream_t stream) {
  permute_conv_kernel<T><<<blocks_for((long long)co * ci * 3, 256), 256, 0, stream>>>(src, dst, co, ci);
  BW_CUDA(cudaGetLastError());
}
template void permute_conv_weight<float>(const float*, float*, int, int, cudaStream_t);
template void permute_conv_weight<bf16>(const float*, bf16*, int, int, cudaStream_t);

template <typename T>
void dec_embed(const DecRows& rows, const int* next_tok, const T* tok_emb, const T* pos_emb, float* x, int d, int* page_table,
               int n_blocks, cudaStream_t stream) {
  if (rows.n_rows <= 0) return;
  launch_kernel(dec_embed_kernel<T>, dim3(rows.n_rows), dim3(128), 0, stream, rows.row_seq, rows.row_pos, rows.row_tok, next_tok, tok_emb, pos_emb, x, d,
                rows.row_page, rows.row_page ? page_table : nullptr, n_blocks);
  ++g_kernel_launches;
}
template void dec_embed<float>(const DecRows&, const int*, const float*, const float*, float*, int, int*, int, cudaStream_t);
template void dec_embed<bf16>(const DecRows&, const int*, const bf16*, const bf16*, float*, int, int*, int, cudaStream_t);

template <typename T>
void dec_embed_ln(const DecRows& rows, const int* next_tok, const T* tok_emb, const T* pos_emb, float* x, int d, bf16* xb,
                  float2* stats, int* page_table, int n_blocks, cudaStream_t stream) {
  if (rows.n_rows <= 0) return;
  BW_CHECK(d % 64 == 0, "LayerNorm-fused decoder needs d % 64 == 0");
  launch_kernel(dec_embed_ln_kernel<T>, dim3(rows.n_rows), dim3(128), (size_t)d * 4, stream, rows.row_seq, rows.row_pos, rows.row_tok,
                next_tok, tok_emb, pos_emb, x, d, xb, stats, rows.row_page, rows.row_page ? page_table : nullptr, n_blocks);
  ++g_kernel_launches;
}
template void dec_embed_ln<bf16>(const DecRows&, const int*, const bf16*, const bf16*, float*, int, bf16*, float2*, int*, int, cudaStream_t);

void rows_ln_partials(const float* x, int rows, int d, bf16* xb, float2* stats, cudaStream_t stream) {
  if (rows <= 0) return;
  BW_CHECK(d % 64 == 0, "d % 64 == 0 required");
  rows_ln_partials_kernel<<<rows, 128, (size_t)d * 4, stream>>>(x, d, xb, stats);
  BW_CUDA(cudaGetLastError());
  ++g_kernel_launches;
}

void fold_layernorm(const float* W, const float* gamma, const float* beta, const float* bias, int N, int K, bf16* Wf, float* c1,
                    float* c2, cudaStream_t stream) {
  fold_ln_kernel<<<N, 256, 0, stream>>>(W, gamma, beta, bias, K, Wf, c1, c2);
  BW_CUDA(cudaGetLastError());
}

}  // namespace bw
