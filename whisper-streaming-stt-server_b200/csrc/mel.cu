// K1: log-mel frontend (upstream whisper/audio.py log_mel_spectrogram, reached from reference
// stt_server/model/backends/torch_whisper.py:55).
//
//   phase 1  mel_power_kernel : reflect pad + periodic Hann + 400-point real DFT (even/odd folded,
//            fp32 direct sum) + |.|^2 + Slaney mel filterbank (sparse rows) + log10(max(x,1e-10)),
//            plus a grid-wide max (atomicMax on the bit pattern of value+16 > 0).
//            Frames that only see the 30 s zero padding are never computed (they equal -10).
//   phase 2  mel_normalize_kernel / mel_window_kernel : max(x, gmax-8), (x+4)/4, written either as
//            the reference layout f32 [n_mels, frames] or as the encoder's conv1 operand
//            (time-major im2col rows [3000, 3*n_mels], zero conv padding, zeros past the segment).
#include "kernels.cuh"

namespace bw {
namespace {

// frames per CTA (FT): 16 for long calls; 8 / 4 when a call has too few frames to give every SM a CTA (a 6 s streaming partial
// is 600 frames = 38 CTAs of 16)
constexpr int NBIN = 201;
constexpr int MEL_THREADS = 224;
constexpr int PSTRIDE = 209;  // odd stride: conflict-free column reads of the power tile

__device__ __forceinline__ float sample_at(const float* __restrict__ pcm, long long idx, long long n, long long L) {
  // reflect padding of the zero-extended signal of length L (torch.stft center=True, pad_mode="reflect")
  if (idx < 0) idx = -idx;
  if (idx >= L) idx = 2 * (L - 1) - idx;
  return (idx < n) ? __ldg(pcm + idx) : 0.f;
}

template <int FT>
__global__ void __launch_bounds__(MEL_THREADS)
mel_power_kernel(const float* __restrict__ pcm, long long n, long long L, int n_real, int total_frames,
                 const float* __restrict__ tables,   // cos[425] sin[425] win[201]
                 const float* __restrict__ filters,  // [n_mels][201]
                 const int2* __restrict__ ranges,    // [n_mels] (lo, hi)
                 int n_mels, float* __restrict__ logmel, int ld, int* __restrict__ gmax_bits) {
  extern __shared__ float sm[];
  float* raw = sm;                              // FT*160 + 240
  float* xe = raw + (FT * 160 + 240);           // [201][FT]
  float* xo = xe + NBIN * FT;                   // [201][FT]
  float* cs = xo + NBIN * FT;                   // 425
  float* sn = cs + 425;                         // 425
  float* pw = sn + 425;                         // [FT][PSTRIDE]
  __shared__ float red[8];

  const int tid = threadIdx.x;
  const int f0 = blockIdx.x * FT;
  const long long base = (long long)f0 * 160 - 200;
  const int nraw = FT * 160 + 240;
  if (base >= 0 && base + nraw <= n && ((base & 3) == 0)) {
    const float4* src = reinterpret_cast<const float4*>(pcm + base);
    for (int i = tid; i < nraw / 4; i += MEL_THREADS) reinterpret_cast<float4*>(raw)[i] = __ldg(src + i);
  } else {
    for (int i = tid; i < nraw; i += MEL_THREADS) raw[i] = sample_at(pcm, base + i, n, L);
  }
  for (int i = tid; i < 850; i += MEL_THREADS) cs[i] = __ldg(tables + i);  // cos and sin are adjacent
  __syncthreads();
  const float* win = tables + 850;
  for (int i = tid; i < NBIN * FT; i += MEL_THREADS) {
    const int nn = i / FT, f = i % FT;
    const float w = __ldg(win + nn);
    const float a = raw[f * 160 + nn];
    float e, o;
    if (nn == 0 || nn == 200) { e = w * a; o = 0.f; }
    else { const float b = raw[f * 160 + 400 - nn]; e = w * (a + b); o = w * (a - b); }
    xe[i] = e; xo[i] = o;
  }
  __syncthreads();

  const int k = tid;
  if (k < NBIN) {
    float re[FT], im[FT];
#pragma unroll
    for (int f = 0; f < FT; ++f) { re[f] = 0.f; im[f] = 0.f; }
    int idx = 0;
#pragma unroll 1
    for (int nn = 0; nn <= 200; ++nn) {
      const int ti = idx + (idx >> 4);
      const float c = cs[ti], s = sn[ti];
      const float4* pe = reinterpret_cast<const float4*>(xe + nn * FT);
      const float4* po = reinterpret_cast<const float4*>(xo + nn * FT);
#pragma unroll
      for (int q = 0; q < FT / 4; ++q) {
        const float4 e = pe[q], o = po[q];
        re[4 * q + 0] = fmaf(e.x, c, re[4 * q + 0]); im[4 * q + 0] = fmaf(o.x, s, im[4 * q + 0]);
        re[4 * q + 1] = fmaf(e.y, c, re[4 * q + 1]); im[4 * q + 1] = fmaf(o.y, s, im[4 * q + 1]);
        re[4 * q + 2] = fmaf(e.z, c, re[4 * q + 2]); im[4 * q + 2] = fmaf(o.z, s, im[4 * q + 2]);
        re[4 * q + 3] = fmaf(e.w, c, re[4 * q + 3]); im[4 * q + 3] = fmaf(o.w, s, im[4 * q + 3]);
      }
      idx += k;
      if (idx >= 400) idx -= 400;
    }
#pragma unroll
    for (int f = 0; f < FT; ++f) pw[f * PSTRIDE + k] = re[f] * re[f] + im[f] * im[f];
  }
  __syncthreads();

  float lmax = 0.f;  // values are stored +16 (> 0)
  for (int i = tid; i < FT * n_mels; i += MEL_THREADS) {
    const int m = i / FT, f = i % FT;
    if (f0 + f >= n_real) continue;
    const int2 r = __ldg(ranges + m);
    const float* fr = filters + m * NBIN;
    const float* pr = pw + f * PSTRIDE;
    float acc = 0.f;
    for (int kk = r.x; kk < r.y; ++kk) acc = fmaf(__ldg(fr + kk), pr[kk], acc);
    const float v = log10f(fmaxf(acc, 1e-10f));
    logmel[(long long)m * ld + f0 + f] = v;
    lmax = fmaxf(lmax, v + 16.f);
  }
  lmax = warp_max(lmax);
  if ((tid & 31) == 0) red[tid >> 5] = lmax;
  __syncthreads();
  if (tid == 0) {
    float mx = 0.f;
    for (int w = 0; w < MEL_THREADS / 32; ++w) mx = fmaxf(mx, red[w]);
    if (blockIdx.x == 0 && n_real < total_frames) mx = fmaxf(mx, 6.f);  // all-zero frames: log10(1e-10)+16
    atomicMax(gmax_bits, __float_as_int(mx));
  }
}

__device__ __forceinline__ float norm_mel(float raw_log, float gmax) {
  return (fmaxf(raw_log, gmax - 8.f) + 4.f) * 0.25f;
}

// reference layout: out[m][f] f32, f < total_frames
__global__ void mel_normalize_kernel(const float* __restrict__ logmel, int ld, int n_real, const int* __restrict__ gmax_bits,
                                     int n_mels, int total_frames, float* __restrict__ out) {
  const float gmax = __int_as_float(*gmax_bits) - 16.f;
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  const int m = blockIdx.y;
  if (f >= total_frames) return;
  const float v = (f < n_real) ? logmel[(long long)m * ld + f] : -10.f;
  out[(long long)m * total_frames + f] = norm_mel(v, gmax);
}

// conv1 operand for the window [seek, seek+3000): A1[t][k*n_mels + m] = x[t+k-1][m],
// x[f][m] = normalised mel(seek+f) for 0 <= f < segment_size, else 0 (pad_or_trim zeros and conv padding).
template <typename T>
__global__ void mel_window_kernel(const float* __restrict__ logmel, int ld, int n_real, const int* __restrict__ gmax_bits,
                                  int n_mels, int seek, int segment_size, T* __restrict__ A1) {
  __shared__ float tile[32][33];
  // gmax_bits == nullptr: `logmel` already holds normalised values (stage-level bw_encode input)
  const float gmax = gmax_bits ? __int_as_float(*gmax_bits) - 16.f : 0.f;
  const int fb = blockIdx.x * 32 - 1;  // window frame of tile column 0 (covers f = -1 .. 3000)
  const int mb = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int m = mb + r, f = fb + threadIdx.x;
    float v = 0.f;
    if (m < n_mels && f >= 0 && f < segment_size) {
      const int gf = seek + f;
      const float raw = gf < n_real ? logmel[(long long)m * ld + gf] : -10.f;
      v = gmax_bits ? norm_mel(raw, gmax) : raw;
    }
    tile[r][threadIdx.x] = v;
  }
  __syncthreads();
  const int K3 = 3 * n_mels;
  for (int c = threadIdx.y; c < 32; c += blockDim.y) {
    const int f = fb + c;  // source frame
    const int m = mb + threadIdx.x;
    if (m >= n_mels) continue;
    const T v = from_f<T>(tile[threadIdx.x][c]);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int t = f - k + 1;  // x[t+k-1] = x[f]
      if (t >= 0 && t < 3000) A1[(long long)t * K3 + k * n_mels + m] = v;
    }
  }
}

}  // namespace

size_t mel_tables_floats() { return 850 + 201; }

void mel_fill_tables(float* host /* 1051 floats */) {
  const double two_pi = 6.283185307179586476925286766559;
  for (int i = 0; i < 850; ++i) host[i] = 0.f;
  for (int idx = 0; idx < 400; ++idx) {
    const int ti = idx + (idx >> 4);
    host[ti] = (float)cos(two_pi * idx / 400.0);
    host[425 + ti] = (float)sin(two_pi * idx / 400.0);
  }
  for (int n = 0; n <= 200; ++n) host[850 + n] = (float)(0.5 - 0.5 * cos(two_pi * n / 400.0));
}

void mel_power(const float* pcm_dev, long long n, long long padding, const float* tables, const float* filters,
               const int2* ranges, int n_mels, float* logmel, int ld, int n_real, int total_frames, int* gmax_bits,
               cudaStream_t stream) {
  BW_CUDA(cudaMemsetAsync(gmax_bits, 0, sizeof(int), stream));
  if (n_real <= 0) return;
  static int n_sm[64];
  static std::atomic<unsigned long long> attr_set{0};
  int dev = 0;
  BW_CUDA(cudaGetDevice(&dev));
  auto smem_of = [](int ft) { return sizeof(float) * ((size_t)(ft * 160 + 240) + 2 * NBIN * ft + 850 + ft * PSTRIDE); };
  if (!(attr_set.load() >> dev & 1ull)) {
    BW_CUDA(cudaFuncSetAttribute(mel_power_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_of(16)));
    BW_CUDA(cudaDeviceGetAttribute(&n_sm[dev], cudaDevAttrMultiProcessorCount, dev));
    attr_set.fetch_or(1ull << dev);
  }
  static const int forced = getenv("B200W_MEL_FT") ? atoi(getenv("B200W_MEL_FT")) : 0;
  const int ft = forced ? forced : (n_real >= 16 * n_sm[dev] ? 16 : n_real >= 8 * n_sm[dev] ? 8 : 4);
  const int grid = (n_real + ft - 1) / ft;
  auto launch = [&](auto kern) {
    kern<<<grid, MEL_THREADS, smem_of(ft), stream>>>(pcm_dev, n, n + padding, n_real, total_frames, tables, filters, ranges, n_mels,
                                                     logmel, ld, gmax_bits);
  };
  if (ft == 16) launch(mel_power_kernel<16>);
  else if (ft == 8) launch(mel_power_kernel<8>);
  else launch(mel_power_kernel<4>);
  BW_CUDA(cudaGetLastError());
  ++g_kernel_launches;
}

void mel_normalize_f32(const float* logmel, int ld, int n_real, const int* gmax_bits, int n_mels, int total_frames,
                       float* out, cudaStream_t stream) {
  if (total_frames <= 0) return;
  dim3 grid((total_frames + 255) / 256, n_mels);
  mel_normalize_kernel<<<grid, 256, 0, stream>>>(logmel, ld, n_real, gmax_bits, n_mels, total_frames, out);
  BW_CUDA(cudaGetLastError());
  ++g_kernel_launches;
}

template <typename T>
void mel_window(const float* logmel, int ld, int n_real, const int* gmax_bits, int n_mels, int seek, int segment_size,
                T* A1, cudaStream_t stream) {
  dim3 grid((3002 + 31) / 32, (n_mels + 31) / 32), block(32, 8);
  mel_window_kernel<T><<<grid, block, 0, stream>>>(logmel, ld, n_real, gmax_bits, n_mels, seek, segment_size, A1);
  BW_CUDA(cudaGetLastError());
  ++g_kernel_launches;
}
template void mel_window<float>(const float*, int, int, const int*, int, int, int, float*, cudaStream_t);
template void mel_window<bf16>(const float*, int, int, const int*, int, int, int, bf16*, cudaStream_t);

}  // namespace bw
