// a1: Whisper log-mel front end on the GPU (whisper.log_mel_spectrogram semantics, as called from
// datasets/speech_dataset.py:101-103): reflect-pad(200) -> hann(400) STFT hop 160 -> |X|^2 (drop last
// frame) -> slaney mel filterbank -> log10(clamp 1e-10) -> max(., utterance_max - 8) -> (. + 4) / 4,
// written time-major [B, n_frames, n_mels] (the dataset's .permute(1, 0)).
//
// One CTA = FR consecutive frames of one utterance: windowed frames staged in shared memory, a direct
// 400-point real DFT per (bin, frame) with a shared twiddle table (exact index arithmetic k*n mod 400),
// then the mel contraction from shared memory.  HBM traffic = read wav once + write mel once.
#include <math_constants.h>

#include "../../include/slam_b200.h"
#include "common.cuh"
#include "host.cuh"

namespace slam {

constexpr int NFFT = 400;
constexpr int HOP = 160;
constexpr int NBIN = 201;
constexpr int FR = 8;  // frames per CTA
static_assert(FR == 8, "the DFT loop reads the 8 frames of a sample as two float4");

__device__ __forceinline__ int float_key(float f) {
  const int i = __float_as_int(f);
  return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float key_float(int k) {
  return __int_as_float(k >= 0 ? k : k ^ 0x7fffffff);
}

__global__ void logmel_init_kernel(float* scratch, int batch) {
  pdl_trigger();
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < batch) reinterpret_cast<int*>(scratch)[i] = float_key(-CUDART_INF_F);
}

__global__ void __launch_bounds__(256) logmel_kernel(const float* __restrict__ wav, int n_samples_max, int n_frames, const int* __restrict__ lengths,
                                                     const float* __restrict__ filt_t, int n_mels, float* __restrict__ out, float* __restrict__ scratch) {
  pdl_trigger();
  pdl_wait();
  __shared__ float s_cos[NFFT];
  __shared__ float s_sin[NFFT];
  __shared__ __align__(16) float s_x[NFFT][FR];      // windowed frames, sample-major: the FR frames of one sample are two 16-byte broadcast loads
  __shared__ float s_pow[FR][NBIN + 3];
  __shared__ float s_red[8];
  const int b = blockIdx.y;
  const int f0 = blockIdx.x * FR;
  const float* w = wav + static_cast<int64_t>(b) * n_samples_max;
  // variable-length batches (dynamic-frame recipes): utterance b has lengths[b] real samples; frames beyond its own
  // n_valid/160 are zero padding of the MEL (the reference pads mel, not audio: speech_dataset_large.py:196-199)
  const int n_samples = lengths != nullptr ? min(lengths[b], n_samples_max) : n_samples_max;
  const int n_valid_frames = n_samples / HOP;

  for (int n = threadIdx.x; n < NFFT; n += blockDim.x) {
    float sn, cs;
    sincospif(2.0f * static_cast<float>(n) / static_cast<float>(NFFT), &sn, &cs);
    s_cos[n] = cs;
    s_sin[n] = sn;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < FR * NFFT; i += blockDim.x) {
    const int f = i / NFFT, n = i - f * NFFT;
    const int t = f0 + f;
    float v = 0.0f;
    if (t < n_valid_frames) {
      int j = t * HOP + n - NFFT / 2;  // index into the un-padded waveform
      if (j < 0) j = -j;               // reflect padding (torch.stft center=True, pad_mode="reflect")
      if (j >= n_samples) j = 2 * (n_samples - 1) - j;
      const float hann = 0.5f - 0.5f * s_cos[n];  // periodic hann window (torch.hann_window(400))
      v = w[j] * hann;
    }
    s_x[n][f] = v;
  }
  __syncthreads();

  // direct DFT: thread k computes bin k of all FR frames
  if (threadIdx.x < NBIN) {
    const int k = threadIdx.x;
    float re[FR], im[FR];
#pragma unroll
    for (int f = 0; f < FR; ++f) re[f] = im[f] = 0.0f;
    int idx = 0;
    for (int n = 0; n < NFFT; ++n) {
      const float c = s_cos[idx], s = s_sin[idx];
      const float4 x0 = *reinterpret_cast<const float4*>(&s_x[n][0]), x1 = *reinterpret_cast<const float4*>(&s_x[n][4]);
      const float xv[FR] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
#pragma unroll
      for (int f = 0; f < FR; ++f) {
        re[f] = fmaf(xv[f], c, re[f]);
        im[f] = fmaf(xv[f], s, im[f]);
      }
      idx += k;
      if (idx >= NFFT) idx -= NFFT;
    }
#pragma unroll
    for (int f = 0; f < FR; ++f) s_pow[f][k] = re[f] * re[f] + im[f] * im[f];
  }
  __syncthreads();

  // mel contraction + log10; filt_t is [201, n_mels] so consecutive threads read consecutive mels
  float lmax = -CUDART_INF_F;
  for (int i = threadIdx.x; i < FR * n_mels; i += blockDim.x) {
    const int f = i / n_mels, m = i - f * n_mels;
    const int t = f0 + f;
    if (t >= n_frames) continue;
    if (t >= n_valid_frames) {
      out[(static_cast<int64_t>(b) * n_frames + t) * n_mels + m] = 0.0f;
      continue;
    }
    float acc = 0.0f;
    for (int k = 0; k < NBIN; ++k) acc = fmaf(filt_t[k * n_mels + m], s_pow[f][k], acc);
    const float lg = log10f(fmaxf(acc, 1e-10f));
    out[(static_cast<int64_t>(b) * n_frames + t) * n_mels + m] = lg;
    lmax = fmaxf(lmax, lg);
  }
  lmax = warp_max(lmax);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = lmax;
  __syncthreads();
  if (threadIdx.x == 0) {
    float m = s_red[0];
    for (int i = 1; i < (blockDim.x >> 5); ++i) m = fmaxf(m, s_red[i]);
    atomicMax(reinterpret_cast<int*>(scratch) + b, float_key(m));
  }
}

__global__ void logmel_norm_kernel(float* __restrict__ out, const float* __restrict__ scratch, int64_t per_utt, const int* __restrict__ lengths,
                                   int n_samples_max, int n_mels) {
  pdl_trigger();
  pdl_wait();
  const int b = blockIdx.y;
  const float gmax = key_float(reinterpret_cast<const int*>(scratch)[b]);
  float* o = out + static_cast<int64_t>(b) * per_utt;
  const int64_t valid = lengths != nullptr ? static_cast<int64_t>(min(lengths[b], n_samples_max) / HOP) * n_mels : per_utt;
  int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (; i < valid; i += stride) o[i] = (fmaxf(o[i], gmax - 8.0f) + 4.0f) / 4.0f;
}

}  // namespace slam

extern "C" int slam_logmel(const float* wav, int32_t batch, int32_t n_samples, const int32_t* lengths, const float* filters_t, int32_t n_mels,
                           float* out, float* scratch_max, void* stream) {
  using namespace slam;
  SLAM_CHECK_ARG(batch > 0 && n_samples > NFFT && n_mels > 0 && n_mels <= 256, "logmel: bad shape batch=%d n_samples=%d n_mels=%d", batch,
                 n_samples, n_mels);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int n_frames = n_samples / HOP;
  launch_pdl(logmel_init_kernel, static_cast<unsigned>(ceil_div(batch, 128)), 128, 0, st, scratch_max, batch);
  SLAM_LAUNCH_CHECK("slam_logmel.init");
  dim3 grid(static_cast<unsigned>(ceil_div(n_frames, FR)), batch);
  launch_pdl(logmel_kernel, grid, 256, 0, st, wav, n_samples, n_frames, lengths, filters_t, n_mels, out, scratch_max);
  SLAM_LAUNCH_CHECK("slam_logmel");
  const int64_t per_utt = static_cast<int64_t>(n_frames) * n_mels;
  dim3 g2(static_cast<unsigned>(ceil_div(per_utt, 256 * 4) > 296 ? 296 : ceil_div(per_utt, 256 * 4)), batch);
  launch_pdl(logmel_norm_kernel, g2, 256, 0, st, out, scratch_max, per_utt, lengths, n_samples, n_mels);
  SLAM_LAUNCH_CHECK("slam_logmel.norm");
  return 0;
}
