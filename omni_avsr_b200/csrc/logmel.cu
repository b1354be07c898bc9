// Whisper log-mel front end on the GPU (replaces the host-side WhisperFeatureExtractor call and its D2H/H2D round trip
// at Omni_AVSR/modeling_OmniAVSR.py:531-534; algorithm = transformers 4.43.1 feature_extraction_whisper
// `_torch_extract_fbank_features`):
//   zero-pad / truncate to 30 s -> reflect-padded STFT (periodic hann 400, n_fft 400, hop 160), |.|^2, drop last frame
//   -> 80 slaney mel bins -> log10(clamp 1e-10) -> max(x, max_utt - 8) -> (x + 4) / 4  -> bf16 [B, 80, 3000]
// The 400-point DFT is evaluated directly in fp32 from shared-memory twiddles (400 x 201 MACs per frame is ~1 GFLOP per
// utterance: negligible, and it needs no FFT library); pass 1 writes fp32 log-mel + the per-utterance maximum, pass 2
// applies the dynamic-range clamp and the affine map.
#include "common.cuh"
#include "../../include/omni_avsr.h"

namespace omni {

constexpr int LM_NFFT = 400;
constexpr int LM_HOP = 160;
constexpr int LM_BINS = 201;
constexpr int LM_MELS = 80;
constexpr int LM_NSAMP = 480000;
constexpr int LM_FRAMES = 3000;
constexpr int LM_FPB = 16;       // frames per block
constexpr int LM_THREADS = 256;

__device__ __forceinline__ float lm_sample(const void* audio, int is_bf16, long long base, int T, int j) {
  // j indexes the 30 s zero-padded signal after reflect padding has been resolved
  if (j < 0) j = -j;
  if (j >= LM_NSAMP) j = 2 * (LM_NSAMP - 1) - j;
  if (j >= T) return 0.f;
  return is_bf16 ? __bfloat162float(reinterpret_cast<const bf16*>(audio)[base + j])
                 : reinterpret_cast<const float*>(audio)[base + j];
}

// atomic max on floats via ordered-int encoding
__device__ __forceinline__ void atomic_max_float(float* addr, float v) {
  int* ia = reinterpret_cast<int*>(addr);
  int old = *ia;
  while (__int_as_float(old) < v) {
    const int assumed = old;
    old = atomicCAS(ia, assumed, __float_as_int(v));
    if (old == assumed) break;
  }
}

__global__ void __launch_bounds__(LM_THREADS)
logmel_pass1_kernel(const void* __restrict__ audio, int is_bf16, long long audio_bs, int T,
                    const float* __restrict__ mel_filters,   // [80, 201]
                    float* __restrict__ logmel,               // [B, 80, 3000]
                    float* __restrict__ utt_max) {            // [B], pre-filled with -inf
  __shared__ float s_x[(LM_FPB - 1) * LM_HOP + LM_NFFT];      // windowless samples of the frame span
  __shared__ float s_win[LM_NFFT];
  __shared__ float s_cos[LM_NFFT], s_sin[LM_NFFT];
  __shared__ float s_pow[LM_FPB][LM_BINS + 1];
  __shared__ float s_max[LM_THREADS / 32];

  const int b = blockIdx.y;
  const int f0 = blockIdx.x * LM_FPB;
  const int tid = threadIdx.x;
  const long long base = static_cast<long long>(b) * audio_bs;

  for (int i = tid; i < LM_NFFT; i += LM_THREADS) {
    const float ang = 6.283185307179586f * static_cast<float>(i) / static_cast<float>(LM_NFFT);
    float sn, cs;
    sincosf(ang, &sn, &cs);
    s_cos[i] = cs;
    s_sin[i] = sn;
    s_win[i] = 0.5f - 0.5f * cs;   // periodic hann
  }
  const int span = (LM_FPB - 1) * LM_HOP + LM_NFFT;
  for (int i = tid; i < span; i += LM_THREADS)
    s_x[i] = lm_sample(audio, is_bf16, base, T, f0 * LM_HOP + i - LM_NFFT / 2);
  __syncthreads();

  // power spectrum: (frame, bin) pairs
  for (int p = tid; p < LM_FPB * LM_BINS; p += LM_THREADS) {
    const int f = p / LM_BINS;
    const int k = p - f * LM_BINS;
    const float* x = s_x + f * LM_HOP;
    float re = 0.f, im = 0.f;
    int idx = 0;
    for (int n = 0; n < LM_NFFT; ++n) {
      const float v = x[n] * s_win[n];
      re = fmaf(v, s_cos[idx], re);
      im = fmaf(v, s_sin[idx], im);
      idx += k;
      if (idx >= LM_NFFT) idx -= LM_NFFT;
    }
    s_pow[f][k] = re * re + im * im;
  }
  __syncthreads();

  // mel projection + log10, track the block maximum
  float local_max = -INFINITY;
  for (int p = tid; p < LM_FPB * LM_MELS; p += LM_THREADS) {
    const int m = p / LM_FPB;
    const int f = p - m * LM_FPB;
    const int frame = f0 + f;
    if (frame >= LM_FRAMES) continue;
    const float* w = mel_filters + m * LM_BINS;
    float acc = 0.f;
    for (int k = 0; k < LM_BINS; ++k) acc = fmaf(__ldg(w + k), s_pow[f][k], acc);
    const float lv = log10f(fmaxf(acc, 1e-10f));
    logmel[(static_cast<long long>(b) * LM_MELS + m) * LM_FRAMES + frame] = lv;
    local_max = fmaxf(local_max, lv);
  }
  local_max = warp_max(local_max);
  if ((tid & 31) == 0) s_max[tid >> 5] = local_max;
  __syncthreads();
  if (tid == 0) {
    float mx = s_max[0];
    for (int i = 1; i < LM_THREADS / 32; ++i) mx = fmaxf(mx, s_max[i]);
    if (mx > -INFINITY) atomic_max_float(utt_max + b, mx);
  }
}

__global__ void __launch_bounds__(256)
logmel_pass2_kernel(const float* __restrict__ logmel, const float* __restrict__ utt_max, bf16* __restrict__ out,
                    long long per_utt, long long total) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long b = i / per_utt;
    const float v = fmaxf(logmel[i], utt_max[b] - 8.0f);
    out[i] = __float2bfloat16_rn((v + 4.0f) / 4.0f);
  }
}

__global__ void fill_neg_inf_kernel(float* p, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = -INFINITY;
}

}  // namespace omni

extern "C" int64_t omni_logmel_workspace_bytes(int32_t B) {
  if (B <= 0) return 0;
  return static_cast<int64_t>(B) * omni::LM_MELS * omni::LM_FRAMES * 4 + static_cast<int64_t>(B) * 4 + 256;
}

extern "C" int omni_logmel(const void* audio, int32_t audio_is_bf16, int64_t audio_bs, int32_t B, int32_t T,
                           const float* mel_filters, void* out, void* workspace, int64_t workspace_bytes, void* stream) {
  using namespace omni;
  OMNI_CHECK_ARG(audio && mel_filters && out && workspace && B > 0 && T > 0 && audio_bs >= T);
  if (workspace_bytes < omni_logmel_workspace_bytes(B)) return OMNI_ERR_WORKSPACE;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  float* logmel = reinterpret_cast<float*>(workspace);
  float* utt_max = logmel + static_cast<long long>(B) * LM_MELS * LM_FRAMES;
  fill_neg_inf_kernel<<<ceil_div(B, 128), 128, 0, st>>>(utt_max, B);
  OMNI_LAUNCH_CHECK();
  dim3 grid(ceil_div(LM_FRAMES, LM_FPB), B);
  logmel_pass1_kernel<<<grid, LM_THREADS, 0, st>>>(audio, audio_is_bf16, audio_bs, T < LM_NSAMP ? T : LM_NSAMP,
                                                   mel_filters, logmel, utt_max);
  OMNI_LAUNCH_CHECK();
  const long long per_utt = static_cast<long long>(LM_MELS) * LM_FRAMES;
  const long long total = per_utt * B;
  long long blocks = ceil_div_ll(total, 256);
  if (blocks > kNumSMs * 16LL) blocks = kNumSMs * 16LL;
  logmel_pass2_kernel<<<static_cast<int>(blocks), 256, 0, st>>>(logmel, utt_max, reinterpret_cast<bf16*>(out), per_utt,
                                                                total);
  OMNI_LAUNCH_CHECK();
  return OMNI_OK;
}
