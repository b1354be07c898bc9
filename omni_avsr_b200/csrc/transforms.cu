// On-device input pipeline: the per-utterance transforms of the reference's
// datamodule/transforms.py, as HBM-bound byte / float kernels.
//
//   video (VideoTransform, :83-104): uint8 frames [T, C, H, W]  ->  x / 255  ->  crop 88 x 88 at (i, j)  ->  grayscale
//     (torchvision rgb_to_grayscale: 0.2989 r + 0.587 g + 0.114 b, fp32, left to right)  ->  AdaptiveTimeMask (:36-56, whole
//     frames zeroed; spans sampled on the host with the reference's RNG calls)  ->  Normalize (x - 0.421) / 0.165.
//     Every operation is a correctly rounded fp32 op in the reference's order (no FMA contraction): bit-exact.
//   audio (AudioTransform, :107-131): waveform [T]  ->  AdaptiveTimeMask  ->  AddNoise (:59-80, torchaudio.functional.add_noise:
//     noise scaled to the requested SNR from the two signal energies)  ->  layer_norm over the whole utterance (eps 1e-8).
//     Two passes over the clip: five fp64 sums (s, n, s^2, n^2, s n) give the noise scale AND the mean / variance of the
//     mixed signal, the second pass writes the result.  Reductions cannot be bit-identical to torch's CPU summation order:
//     compared with a 1e-5 tolerance.
#include "common.cuh"
#include "../../include/omni_avsr.h"

namespace omni {

constexpr int TR_THREADS = 256;
constexpr int TR_MAX_SPANS = OMNI_MAX_MASK_SPANS;

struct Spans {
  int n;
  int start[TR_MAX_SPANS];
  int end[TR_MAX_SPANS];
};

__device__ __forceinline__ bool in_spans(const Spans& s, long long t) {
  bool m = false;
  for (int i = 0; i < s.n; ++i) m |= (t >= s.start[i] && t < s.end[i]);
  return m;
}

// one thread = 4 consecutive output pixels of one frame row (88 = 22 * 4)
__global__ void __launch_bounds__(TR_THREADS)
video_transform_kernel(const uint8_t* __restrict__ in, int T, int C, int H, int W, int ci, int cj, Spans spans,
                       float* __restrict__ out_f32, bf16* __restrict__ out_bf16, long long total) {
  const long long plane = static_cast<long long>(H) * W;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int q = static_cast<int>(idx % 22);
    const long long r = idx / 22;
    const int y = static_cast<int>(r % 88);
    const int t = static_cast<int>(r / 88);
    const bool masked = in_spans(spans, t);
    const uint8_t* base = in + (static_cast<long long>(t) * C) * plane + static_cast<long long>(ci + y) * W + cj + 4 * q;
    float v[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float g;
      if (C == 3) {
        const float rr = __fdiv_rn(static_cast<float>(base[e]), 255.0f);
        const float gg = __fdiv_rn(static_cast<float>(base[plane + e]), 255.0f);
        const float bb = __fdiv_rn(static_cast<float>(base[2 * plane + e]), 255.0f);
        g = __fadd_rn(__fadd_rn(__fmul_rn(0.2989f, rr), __fmul_rn(0.587f, gg)), __fmul_rn(0.114f, bb));
      } else {
        g = __fdiv_rn(static_cast<float>(base[e]), 255.0f);
      }
      if (masked) g = 0.0f;
      v[e] = __fdiv_rn(__fsub_rn(g, 0.421f), 0.165f);
    }
    const long long o = (static_cast<long long>(t) * 88 + y) * 88 + 4 * q;
    if (out_f32) {
      *reinterpret_cast<float4*>(out_f32 + o) = make_float4(v[0], v[1], v[2], v[3]);
    } else {
      uint2 p;
      p.x = f2_to_bf2(v[0], v[1]);
      p.y = f2_to_bf2(v[2], v[3]);
      *reinterpret_cast<uint2*>(out_bf16 + o) = p;
    }
  }
}

// pass 1: per-block partial sums (fp64) of s, n, s^2, n^2, s*n with s = masked waveform
__global__ void __launch_bounds__(TR_THREADS)
audio_sums_kernel(const float* __restrict__ wave, const float* __restrict__ noise, long long T, Spans spans,
                  double* __restrict__ partial) {
  double a[5] = {0, 0, 0, 0, 0};
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < T;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float s = in_spans(spans, i) ? 0.0f : wave[i];
    const float n = noise ? noise[i] : 0.0f;
    a[0] += s; a[1] += n;
    a[2] += static_cast<double>(s) * s; a[3] += static_cast<double>(n) * n; a[4] += static_cast<double>(s) * n;
  }
  __shared__ double sh[5][TR_THREADS / 32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    double v = a[k];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) sh[k][w] = v;
  }
  __syncthreads();
  if (threadIdx.x < 5) {
    double v = 0;
    for (int i = 0; i < TR_THREADS / 32; ++i) v += sh[threadIdx.x][i];
    partial[blockIdx.x * 5 + threadIdx.x] = v;
  }
}

// pass 2: every block re-reduces the (few) partials, derives scale / mean / rstd, writes its slice
__global__ void __launch_bounds__(TR_THREADS)
audio_apply_kernel(const float* __restrict__ wave, const float* __restrict__ noise, long long T, Spans spans, float snr_db,
                   const double* __restrict__ partial, int n_partial, float* __restrict__ out) {
  __shared__ double tot[5];
  __shared__ float par[3];
  if (threadIdx.x < 5) {
    double v = 0;
    for (int i = 0; i < n_partial; ++i) v += partial[i * 5 + threadIdx.x];
    tot[threadIdx.x] = v;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float scale = 0.0f;
    if (noise) {
      // torchaudio.functional.add_noise: fp32 arithmetic on the two energies
      const float es = static_cast<float>(tot[2]), en = static_cast<float>(tot[3]);
      const float snr0 = 10.0f * (log10f(es) - log10f(en));
      scale = powf(10.0f, (snr0 - snr_db) / 20.0f);
    }
    const double sc = scale;
    const double sum_y = tot[0] + sc * tot[1];
    const double sum_y2 = tot[2] + 2.0 * sc * tot[4] + sc * sc * tot[3];
    const double mean = sum_y / static_cast<double>(T);
    double var = sum_y2 / static_cast<double>(T) - mean * mean;      // biased variance, as F.layer_norm
    if (var < 0) var = 0;
    par[0] = scale;
    par[1] = static_cast<float>(mean);
    par[2] = static_cast<float>(1.0 / sqrt(var + 1e-8));
  }
  __syncthreads();
  const float scale = par[0], mean = par[1], rstd = par[2];
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < T;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float s = in_spans(spans, i) ? 0.0f : wave[i];
    const float y = noise ? __fadd_rn(s, __fmul_rn(scale, noise[i])) : s;
    out[i] = (y - mean) * rstd;
  }
}

static int fill_spans(Spans& s, const int32_t* spans, int32_t n) {
  if (n < 0 || n > TR_MAX_SPANS || (n > 0 && !spans)) return OMNI_ERR_BAD_ARG;
  s.n = n;
  for (int i = 0; i < n; ++i) {
    s.start[i] = spans[2 * i];
    s.end[i] = spans[2 * i + 1];
  }
  return OMNI_OK;
}

constexpr int AUDIO_BLOCKS = 148;

}  // namespace omni

using namespace omni;

extern "C" int omni_video_transform(const void* frames, int32_t T, int32_t C, int32_t H, int32_t W, int32_t crop_i,
                                    int32_t crop_j, const int32_t* spans_host, int32_t n_spans, void* out, int32_t out_bf16,
                                    void* stream) {
  OMNI_CHECK_ARG(frames && out && T > 0 && (C == 1 || C == 3) && H >= 88 && W >= 88);
  OMNI_CHECK_ARG(crop_i >= 0 && crop_j >= 0 && crop_i + 88 <= H && crop_j + 88 <= W);
  OMNI_CHECK_ARG((reinterpret_cast<uintptr_t>(out) & 15) == 0);
  Spans s;
  const int rc = fill_spans(s, spans_host, n_spans);
  if (rc) return rc;
  const long long total = static_cast<long long>(T) * 88 * 22;
  long long blocks = ceil_div_ll(total, TR_THREADS);
  if (blocks > kNumSMs * 16LL) blocks = kNumSMs * 16LL;
  video_transform_kernel<<<(int)blocks, TR_THREADS, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const uint8_t*>(frames), T, C, H, W, crop_i, crop_j, s, out_bf16 ? nullptr : reinterpret_cast<float*>(out),
      out_bf16 ? reinterpret_cast<bf16*>(out) : nullptr, total);
  OMNI_LAUNCH_CHECK();
  return OMNI_OK;
}

extern "C" int64_t omni_audio_transform_workspace_bytes(void) { return AUDIO_BLOCKS * 5 * sizeof(double); }

extern "C" int omni_audio_transform(const float* wave, const float* noise, int64_t T, float snr_db, const int32_t* spans_host,
                                    int32_t n_spans, float* out, void* workspace, int64_t workspace_bytes, void* stream) {
  OMNI_CHECK_ARG(wave && out && T > 0 && workspace && workspace_bytes >= omni_audio_transform_workspace_bytes());
  Spans s;
  const int rc = fill_spans(s, spans_host, n_spans);
  if (rc) return rc;
  long long blocks = ceil_div_ll(T, TR_THREADS * 4);
  if (blocks > AUDIO_BLOCKS) blocks = AUDIO_BLOCKS;
  if (blocks < 1) blocks = 1;
  double* partial = reinterpret_cast<double*>(workspace);
  audio_sums_kernel<<<(int)blocks, TR_THREADS, 0, (cudaStream_t)stream>>>(wave, noise, T, s, partial);
  OMNI_LAUNCH_CHECK();
  audio_apply_kernel<<<(int)blocks, TR_THREADS, 0, (cudaStream_t)stream>>>(wave, noise, T, s, snr_db, partial, (int)blocks,
                                                                          out);
  OMNI_LAUNCH_CHECK();
  return OMNI_OK;
}
