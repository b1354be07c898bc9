// Whisper log-mel front end on the GPU (replaces the host-side WhisperFeatureExtractor call and its D2H/H2D round trip
// at Omni_AVSR/modeling_OmniAVSR.py:531-534; algorithm = transformers 4.43.1 feature_extraction_whisper
// `_torch_extract_fbank_features`):
//   zero-pad / truncate to 30 s -> reflect-padded STFT (periodic hann 400, n_fft 400, hop 160), |.|^2, drop last frame
//   -> 80 slaney mel bins -> log10(clamp 1e-10) -> max(x, max_utt - 8) -> (x + 4) / 4  -> bf16 [B, 80, 3000]
// The 400-point DFT is evaluated directly in fp32 from shared-memory twiddles, folded in half by the real-input symmetry
// and register-tiled over 8 frames per thread (no FFT library); frames that lie entirely in the zero padding behind the
// utterance are written as the constant they evaluate to.  Pass 1 writes fp32 log-mel + the per-utterance maximum,
// pass 2 applies the dynamic-range clamp and the affine map.
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
  // Real-input DFT folded in half: with e[n] = xw[n] + xw[400-n], o[n] = xw[n] - xw[400-n] (n = 1..199; e[0] = xw[0],
  // e[200] = xw[200], o[0] = o[200] = 0):  Re X[k] = sum_{n<=200} e[n] cos(2 pi n k / 400),  Im X[k] = sum o[n] sin(..).
  // Thread = one bin k, 8 frames at a time in registers; e / o are stored frame-fastest so one 16-byte shared load
  // feeds four frames and the twiddle loads are amortised over eight.
  constexpr int HALF = LM_NFFT / 2;                            // 200
  constexpr int SPAN = (LM_FPB - 1) * LM_HOP + LM_NFFT;
  constexpr int POW_LD = LM_BINS + 1;
  __shared__ float s_xp[SPAN > LM_FPB * POW_LD ? SPAN : LM_FPB * POW_LD];   // samples, later the power spectrum
  __shared__ float s_cos[LM_NFFT], s_sin[LM_NFFT];
  __shared__ __align__(16) float s_e[(HALF + 1) * LM_FPB];     // [n][frame]
  __shared__ __align__(16) float s_o[(HALF + 1) * LM_FPB];
  float* s_x = s_xp;
  float (*s_pow)[POW_LD] = reinterpret_cast<float (*)[POW_LD]>(s_xp);
  __shared__ float s_max[LM_THREADS / 32];

  const int b = blockIdx.y;
  const int f0 = blockIdx.x * LM_FPB;
  const int tid = threadIdx.x;
  const long long base = static_cast<long long>(b) * audio_bs;

  // frames that only see the zero padding behind the utterance (a 16 s clip fills 1600 of the 3000 frames):
  // power 0 -> log10(clamp 1e-10) = -10, no transform needed
  if (T <= LM_NSAMP - LM_NFFT && f0 * LM_HOP - HALF >= T) {
    for (int p = tid; p < LM_FPB * LM_MELS; p += LM_THREADS) {
      const int m = p / LM_FPB, frame = f0 + (p - m * LM_FPB);
      if (frame < LM_FRAMES) logmel[(static_cast<long long>(b) * LM_MELS + m) * LM_FRAMES + frame] = -10.0f;
    }
    if (tid == 0) atomic_max_float(utt_max + b, -10.0f);
    return;
  }

  for (int i = tid; i < LM_NFFT; i += LM_THREADS) {
    const float ang = 6.283185307179586f * static_cast<float>(i) / static_cast<float>(LM_NFFT);
    float sn, cs;
    sincosf(ang, &sn, &cs);
    s_cos[i] = cs;
    s_sin[i] = sn;
  }
  for (int i = tid; i < SPAN; i += LM_THREADS)
    s_x[i] = lm_sample(audio, is_bf16, base, T, f0 * LM_HOP + i - HALF);
  __syncthreads();
  for (int p = tid; p < (HALF + 1) * LM_FPB; p += LM_THREADS) {
    const int n = p / LM_FPB, f = p - n * LM_FPB;
    const float* x = s_x + f * LM_HOP;
    const float wn = 0.5f - 0.5f * s_cos[n];                   // periodic hann (symmetric: w[400-n] = w[n])
    const float a = x[n] * wn;
    const float c = (n == 0 || n == HALF) ? 0.f : x[LM_NFFT - n] * wn;
    s_e[p] = a + c;
    s_o[p] = (n == 0 || n == HALF) ? 0.f : a - c;
  }
  __syncthreads();

  if (tid < LM_BINS) {
    const int k = tid;
#pragma unroll 1
    for (int g = 0; g < LM_FPB / 8; ++g) {
      float re[8], im[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) re[i] = im[i] = 0.f;
      int idx = 0;
#pragma unroll 2
      for (int n = 0; n <= HALF; ++n) {
        const float cs = s_cos[idx], sn = s_sin[idx];
        const float4 e0 = *reinterpret_cast<const float4*>(s_e + n * LM_FPB + g * 8);
        const float4 e1 = *reinterpret_cast<const float4*>(s_e + n * LM_FPB + g * 8 + 4);
        const float4 o0 = *reinterpret_cast<const float4*>(s_o + n * LM_FPB + g * 8);
        const float4 o1 = *reinterpret_cast<const float4*>(s_o + n * LM_FPB + g * 8 + 4);
        re[0] = fmaf(e0.x, cs, re[0]); re[1] = fmaf(e0.y, cs, re[1]); re[2] = fmaf(e0.z, cs, re[2]); re[3] = fmaf(e0.w, cs, re[3]);
        re[4] = fmaf(e1.x, cs, re[4]); re[5] = fmaf(e1.y, cs, re[5]); re[6] = fmaf(e1.z, cs, re[6]); re[7] = fmaf(e1.w, cs, re[7]);
        im[0] = fmaf(o0.x, sn, im[0]); im[1] = fmaf(o0.y, sn, im[1]); im[2] = fmaf(o0.z, sn, im[2]); im[3] = fmaf(o0.w, sn, im[3]);
        im[4] = fmaf(o1.x, sn, im[4]); im[5] = fmaf(o1.y, sn, im[5]); im[6] = fmaf(o1.z, sn, im[6]); im[7] = fmaf(o1.w, sn, im[7]);
        idx += k;
        if (idx >= LM_NFFT) idx -= LM_NFFT;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) s_pow[g * 8 + i][k] = re[i] * re[i] + im[i] * im[i];
    }
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
