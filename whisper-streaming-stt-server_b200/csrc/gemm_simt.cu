// SIMT tiled GEMM (64x64x16 tiles, 4x4 micro-tile per thread, fp32 accumulate in k order).
// T = float is the fp32 validation mode's GEMM (token-id parity vs the oracle); T = bf16 is the
// cross-check / debugging fallback for the tcgen05 kernel.  Same GemmArgs contract (gemm.cuh).
#include "gemm.cuh"

namespace bw {
namespace {

template <typename T>
__global__ void __launch_bounds__(256) gemm_simt_kernel(const GemmArgs g) {
  __shared__ float As[16][64 + 4];
  __shared__ float Bs[16][64 + 4];
  const int z = blockIdx.z;
  const T* A = reinterpret_cast<const T*>(g.A) + (long long)z * g.a_zstride;
  const T* B = reinterpret_cast<const T*>(g.B) + (long long)z * g.b_zstride;
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  float acc[4][4] = {};
  const int lr = tid >> 2;        // 0..63 : tile row loaded by this thread
  const int lk = (tid & 3) * 4;   // 0,4,8,12
  for (int k0 = 0; k0 < g.K; k0 += 16) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int k = k0 + lk + q;
      const int am = m0 + lr, bn = n0 + lr;
      As[lk + q][lr] = (am < g.M && k < g.K) ? to_f(A[(long long)am * g.lda + k]) : 0.f;
      Bs[lk + q][lr] = (bn < g.N && k < g.K) ? to_f(B[(long long)bn * g.ldb + k]) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      float a[4], b[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) { a[q] = As[k][ty * 4 + q]; b[q] = Bs[k][tx * 4 + q]; }
#pragma unroll
      for (int ii = 0; ii < 4; ++ii)
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) acc[ii][jj] = fmaf(a[ii], b[jj], acc[ii][jj]);
    }
    __syncthreads();
  }
  const float* bias = g.bias ? g.bias + (long long)z * g.bias_zstride : nullptr;
  const float* res = g.residual ? g.residual + (long long)z * g.res_zstride : nullptr;
#pragma unroll
  for (int ii = 0; ii < 4; ++ii) {
    const int i = m0 + ty * 4 + ii;
    if (i >= g.M) continue;
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      const int j = n0 + tx * 4 + jj;
      if (j >= g.N) continue;
      float x = acc[ii][jj];
      if (bias) x += bias[g.transposed ? i : j];
      if (g.gelu) x = gelu_erf(x);
      const long long oi = g.transposed ? (long long)j * g.ldc + i : (long long)i * g.ldc + j;
      if (res) x += res[g.transposed ? (long long)j * g.ldres + i : (long long)i * g.ldres + j];
      if (g.out_fp32) (reinterpret_cast<float*>(g.C) + (long long)z * g.c_zstride)[oi] = x;
      else (reinterpret_cast<T*>(g.C) + (long long)z * g.c_zstride)[oi] = from_f<T>(x);
    }
  }
}

}  // namespace

template <typename T>
void gemm_simt(const GemmArgs& g, cudaStream_t stream) {
  BW_CHECK(g.M > 0 && g.N > 0 && g.K > 0 && g.Z > 0, "empty GEMM");
  dim3 grid((g.N + 63) / 64, (g.M + 63) / 64, g.Z);
  gemm_simt_kernel<T><<<grid, 256, 0, stream>>>(g);
  BW_CUDA(cudaGetLastError());
  ++g_kernel_launches;
}
template void gemm_simt<float>(const GemmArgs&, cudaStream_t);
template void gemm_simt<bf16>(const GemmArgs&, cudaStream_t);

}  // namespace bw
