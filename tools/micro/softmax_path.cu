// Microbenchmark: the softmax warps' per-tile instruction stream of the encoder attention kernel WITHOUT the MMA
// coupling, adding one ingredient at a time (16 warps / SM, 64 scores per thread per iteration).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../whisper-streaming-stt-server_b200/csrc/common.cuh"
namespace bw { unsigned long long* g_trace_dev = nullptr; thread_local bool tl_pdl = false; }
using namespace bw;

__global__ void __launch_bounds__(512, 1) k(int mode, int iters, long long* out, float* sink) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint32_t tptr;
  __shared__ uint64_t bar;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) { tmem_alloc(&tptr, 512); tmem_relinquish(); }
  if (threadIdx.x == 0) { mbar_init(&bar, 512); fence_barrier_init(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const int tile = warp >> 3, half = (warp >> 2) & 1, wq = warp & 3, row = wq * 32 + lane;
  const uint32_t my_S = tptr + ((uint32_t)(wq * 32) << 16) + tile * 128 + half * 64;
  const uint32_t prow = smem_u32(smem + tile * 32768 + half * 16384 + row * 128);
  const uint32_t mxbuf = smem_u32(smem + 65536) + tile * 2048;
  const int pair_bar = 1 + tile * 4 + wq;
  float m = 0.f, acc = 0.f;
  __syncthreads();
  const long long t0 = clock64();
  for (int j = 0; j < iters; ++j) {
    uint32_t sv[64];
    tmem_ld_32x32b_x32(my_S, sv);
    tmem_ld_32x32b_x32(my_S + 32, sv + 32);
    tmem_ld_wait();
    if (mode >= 1) {  // row max
      float mx = -INFINITY;
#pragma unroll
      for (int i = 0; i < 64; i += 2) mx = fmaxf(mx, fmaxf(__uint_as_float(sv[i]), __uint_as_float(sv[i + 1])));
      if (mode >= 3) {  // exchange with the partner warp
        st_shared_f32(mxbuf + (uint32_t)((((j & 1) * 2 + half) * 128 + row) * 4), mx);
        asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
        mx = fmaxf(mx, ld_shared_f32(mxbuf + (uint32_t)((((j & 1) * 2 + (half ^ 1)) * 128 + row) * 4)));
      }
      m = fmaxf(m, mx * 1e-30f);
    }
    // exp
#pragma unroll
    for (int i = 0; i < 64; i += 2) {
      float a0, a1;
      ffma2(a0, a1, __uint_as_float(sv[i]), __uint_as_float(sv[i + 1]), 0.18f, 0.18f, -m, -m);
      sv[i] = __float_as_uint(a0); sv[i + 1] = __float_as_uint(a1);
    }
#pragma unroll
    for (int i = 0; i < 64; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+r"(sv[i]));
    if (mode >= 2) {  // pack + store P
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        uint32_t pk[4];
#pragma unroll
        for (int i = 0; i < 8; i += 2) asm volatile("prmt.b32 %0, %1, %2, 0x7632;" : "=r"(pk[i >> 1]) : "r"(sv[g * 8 + i]), "r"(sv[g * 8 + i + 1]));
        st_shared_v4(prow + (uint32_t)((g ^ (row & 7)) << 4), pk[0], pk[1], pk[2], pk[3]);
      }
    } else {
#pragma unroll
      for (int i = 0; i < 64; ++i) acc += __uint_as_float(sv[i]);
    }
    if (mode >= 4) { fence_proxy_async_smem(); tc_fence_before(); mbar_arrive(&bar); }
    if (mode >= 5) { mbar_wait(&bar, j & 1); tc_fence_after(); }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc + m;
  __syncthreads();
  if (warp == 0) tmem_dealloc(tptr, 512);
}

int main() {
  long long* out; float* sink;
  cudaMalloc(&out, 8 * 148); cudaMalloc(&sink, 4 * 148 * 512);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024);
  const int iters = 2000;
  const char* names[] = {"ldtm + ffma2 + 64 mufu (+ 64 fadd)", "+ row max (fmnmx3)", "+ prmt + P st.shared (no fadd)", "+ pair-barrier max exchange",
                         "+ fence.proxy.async + mbarrier arrive", "+ mbarrier wait (all 512 threads in lockstep)"};
  for (int mode = 0; mode <= 5; ++mode) {
    k<<<148, 512, 80 * 1024>>>(mode, iters, out, sink);
    cudaError_t e = cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
    printf("mode %d %-48s: %.0f cycles/iter (MUFU-bound floor 2048)  [%s]\n", mode, names[mode], (double)h / iters, cudaGetErrorString(e));
  }
  return 0;
}
