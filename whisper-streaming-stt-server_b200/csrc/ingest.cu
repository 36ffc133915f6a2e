// Raw audio ingest on the device: PCM16 -> float32 and resampling to 16 kHz, in front of the log-mel kernel.
// Replaces, for callers that hand over the received bytes directly, the host-side numpy / torchaudio work of
// reference stt_server/utils/audio.py:6-30 (called from ModelWorker._decode, stt_server/model/worker.py:118-121).
// The filter bank is torchaudio's sinc_interp_hann kernel (lowpass_filter_width 6, rolloff 0.99), built on the host
// (b200_whisper/ingest.py) exactly as `_get_sinc_resample_kernel` does and registered per source rate.
// HBM-bound by contract and tiny in practice: 2 B read + 4 B * 16000 / rate written per input sample.
#include "kernels.cuh"

namespace bw {
namespace {

__global__ void __launch_bounds__(256)
pcm16_to_f32_kernel(const int16_t* __restrict__ pcm, long long n, float* __restrict__ out) {
  // 8 samples (one 16-byte load) per thread; the tail is handled sample by sample
  const long long i8 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 8;
  if (i8 >= n) return;
  if (i8 + 8 <= n && (reinterpret_cast<uintptr_t>(pcm + i8) & 15) == 0) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(pcm + i8));
    const short* s = reinterpret_cast<const short*>(&u);
    float v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = (float)s[k] * (1.0f / 32768.0f);
    if ((reinterpret_cast<uintptr_t>(out + i8) & 15) == 0) {
      *reinterpret_cast<float4*>(out + i8) = make_float4(v[0], v[1], v[2], v[3]);
      *reinterpret_cast<float4*>(out + i8 + 4) = make_float4(v[4], v[5], v[6], v[7]);
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k) out[i8 + k] = v[k];
    }
  } else {
    for (long long i = i8; i < n && i < i8 + 8; ++i) out[i] = (float)pcm[i] * (1.0f / 32768.0f);
  }
}

// One thread per output sample; consecutive threads read overlapping input windows (L1 hits) and walk the taps of
// their own phase.  The 2^-15 scale is applied to the sum: exact, it commutes with fp32 rounding.
__global__ void __launch_bounds__(256)
pcm16_resample_kernel(const int16_t* __restrict__ pcm, long long n, const float* __restrict__ taps, const int2* __restrict__ ranges,
                      int orig, int nw, int K, int width, float* __restrict__ out, long long n_out) {
  const long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= n_out) return;
  const int i = (int)(o % nw);
  const long long base = (o / nw) * orig - width;
  const int2 r = ranges[i];
  const float* tp = taps + (long long)i * K;
  float acc = 0.f;
  for (int k = r.x; k < r.y; ++k) {
    const long long idx = base + k;
    if (idx >= 0 && idx < n) acc = fmaf(__ldg(tp + k), (float)__ldg(pcm + idx), acc);
  }
  out[o] = acc * (1.0f / 32768.0f);
}

}  // namespace

void pcm16_to_f32(const int16_t* pcm, long long n, float* out, cudaStream_t stream) {
  if (n <= 0) return;
  const long long threads = (n + 7) / 8;
  launch_kernel(pcm16_to_f32_kernel, dim3((unsigned)((threads + 255) / 256)), dim3(256), 0, stream, pcm, n, out);
  ++g_kernel_launches;
}

void pcm16_resample(const int16_t* pcm, long long n, const float* taps, const int2* ranges, int orig, int nw, int K, int width,
                    float* out, long long n_out, cudaStream_t stream) {
  if (n_out <= 0) return;
  launch_kernel(pcm16_resample_kernel, dim3((unsigned)((n_out + 255) / 256)), dim3(256), 0, stream, pcm, n, taps, ranges, orig, nw, K,
                width, out, n_out);
  ++g_kernel_launches;
}

}  // namespace bw
