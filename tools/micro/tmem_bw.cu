// Microbenchmark: tcgen05.ld throughput per SM (B200), MUFU.EX2 throughput, and both together.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_bw tmem_bw.cu && ./tmem_bw
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../whisper-streaming-stt-server_b200/csrc/common.cuh"
namespace bw { unsigned long long* g_trace_dev = nullptr; thread_local bool tl_pdl = false; }
using namespace bw;

__global__ void __launch_bounds__(512, 1) k(int mode, int iters, long long* out, float* sink) {
  __shared__ uint32_t tptr;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) { tmem_alloc(&tptr, 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t base = tptr + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 64 % 512;
  float acc = 0.f;
  uint32_t r[32];
  float x = threadIdx.x * 1e-3f;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    if (mode & 1) {  // TMEM read: 2 x (32 lanes x 32 columns x 4 B) = 8 KB per warp per iteration
      tmem_ld_32x32b_x32(base, r);
      tmem_ld_32x32b_x32(base + 32, r);
      tmem_ld_wait();
      acc += __uint_as_float(r[i & 31]);
    }
    if (mode & 2) {  // 64 MUFU.EX2 per thread per iteration
#pragma unroll
      for (int u = 0; u < 64; ++u) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x + u)); acc += y * 1e-30f; }
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  __syncthreads();
  if (warp == 0) tmem_dealloc(tptr, 512);
}

int main() {
  long long* out; float* sink;
  cudaMalloc(&out, 8 * 148); cudaMalloc(&sink, 4 * 148 * 512);
  const int iters = 2000;
  for (int warps : {4, 8, 16}) for (int mode : {1, 2, 3}) {
    k<<<148, warps * 32>>>(mode, iters, out, sink);
    cudaError_t e = cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
    const double cyc = (double)h / iters;
    printf("warps %2d mode %d (%s): %.0f cycles/iter  -> TMEM %.1f B/clk/SM, MUFU %.2f lanes/clk/SM  [%s]\n", warps, mode,
           mode == 1 ? "ldtm" : mode == 2 ? "mufu" : "both", cyc, (mode & 1) ? warps * 8192.0 / cyc : 0.0,
           (mode & 2) ? warps * 32 * 64.0 / cyc : 0.0, cudaGetErrorString(e));
  }
  return 0;
}
