// ResNet-18 trunk of the AV-HuBERT video front-end (av_hubert/avhubert/resnet.py:35-74 BasicBlock, :77-129 ResNet,
// :156-164 ResEncoder.forward) without library convolutions: every 3x3 / 1x1 convolution is a tcgen05 GEMM.
//
// Activation layout: channels-last frames with a ONE-PIXEL ZERO RING, [N, H + 2, W + 2, C], flattened to rows of C
// channels.  For a 3x3 stride-1 convolution the nine taps of output row m are the input rows m + dy (W + 2) + dx, so an
// OVERLAPPING-ROW view of the activations (row stride C, row length (2 (W + 2) + 3) C, base one padded line + one pixel
// before row 0) holds all of them: the GEMM's main K loop reads the dy = -1 segment [0, 3C), two groups of K-extension
// blocks (the mechanism that carries the LoRA up-projection in the LLM) read the dy = 0 / +1 segments, against the
// tap-major filter matrix [C_out, 9 C_in] -- one omni_gemm_bf16 launch per convolution, no im2col buffer.  The rows of
// the ring come out as garbage and are re-zeroed by the PReLU kernel below, which every convolution is followed by.
// The three stride-2 3x3 convolutions and the three 1x1 stride-2 downsample convolutions have small outputs: a gather
// kernel writes their GEMM operand rows for every position of the ring-padded OUTPUT grid (ring rows: zeros).
//
// Kernels here (all HBM-bound, 16-byte vectors, grid-stride): PReLU (+ folded-BatchNorm bias, + residual) with ring
// re-zeroing, the two stride-2 gathers, and the final AdaptiveAvgPool2d(1) over the interior.
#include "common.cuh"
#include "../../include/omni_avsr.h"

namespace omni {

constexpr int RT_THREADS = 256;

__device__ __forceinline__ void rt_unpack8(const uint4& u, float (&f)[8]) {
  float2 t;
  t = bf2_to_f2(u.x); f[0] = t.x; f[1] = t.y;
  t = bf2_to_f2(u.y); f[2] = t.x; f[3] = t.y;
  t = bf2_to_f2(u.z); f[4] = t.x; f[5] = t.y;
  t = bf2_to_f2(u.w); f[6] = t.x; f[7] = t.y;
}
__device__ __forceinline__ uint4 rt_pack8(const float (&f)[8]) {
  uint4 o;
  o.x = f2_to_bf2(f[0], f[1]); o.y = f2_to_bf2(f[2], f[3]);
  o.z = f2_to_bf2(f[4], f[5]); o.w = f2_to_bf2(f[6], f[7]);
  return o;
}
__device__ __forceinline__ float rt_rbf(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

// x <- PReLU((x + bias) (+ residual + res_bias)) on the interior pixels, 0 on the ring (same arithmetic and rounding
// points as prelu_res_kernel in elementwise.cu: bf16(conv + shift), bf16(out + residual), PReLU)
__global__ void __launch_bounds__(RT_THREADS)
prelu_res_ring_kernel(bf16* __restrict__ x, const bf16* __restrict__ res, const bf16* __restrict__ slope,
                      const bf16* __restrict__ bias, const bf16* __restrict__ res_bias, int Hp, int Wp, int C8,
                      long long total8) {
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total8;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(idx % C8);
    const long long row = idx / C8;
    const int px = static_cast<int>(row % Wp);
    const int py = static_cast<int>((row / Wp) % Hp);
    if (px == 0 || py == 0 || px == Wp - 1 || py == Hp - 1) {
      reinterpret_cast<uint4*>(x)[idx] = make_uint4(0u, 0u, 0u, 0u);
      continue;
    }
    float f[8], s[8];
    rt_unpack8(reinterpret_cast<const uint4*>(x)[idx], f);
    rt_unpack8(__ldg(reinterpret_cast<const uint4*>(slope) + c), s);
    if (bias) {
      float b[8];
      rt_unpack8(__ldg(reinterpret_cast<const uint4*>(bias) + c), b);
#pragma unroll
      for (int i = 0; i < 8; ++i) f[i] = rt_rbf(f[i] + b[i]);
    }
    if (res) {
      float r[8];
      rt_unpack8(ld_nc_u4(reinterpret_cast<const uint4*>(res) + idx), r);
      if (res_bias) {
        float b[8];
        rt_unpack8(__ldg(reinterpret_cast<const uint4*>(res_bias) + c), b);
#pragma unroll
        for (int i = 0; i < 8; ++i) r[i] = rt_rbf(r[i] + b[i]);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) f[i] = rt_rbf(f[i] + r[i]);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] = f[i] > 0.f ? f[i] : f[i] * s[i];
    reinterpret_cast<uint4*>(x)[idx] = rt_pack8(f);
  }
}

// GEMM operand rows of a strided convolution for EVERY position of the ring-padded output grid [N, Ho + 2, Wo + 2]:
//   taps = 9: row = the 3x3 window (pad 1) around input pixel (s oy, s ox), tap-major [9][C]     (conv3x3, stride s)
//   taps = 1: row = input pixel (s oy, s ox)                                                   (1x1 downsample, stride s)
// (stride 2 in the trunk; stride 1 serves channel counts whose 3 C is not a multiple of the GEMM's 64-wide K block)
// ring positions get zero rows.  x is the ring-padded input [N, H + 2, W + 2, C]; one thread = 16 bytes.
__global__ void __launch_bounds__(RT_THREADS)
gather_s2_ring_kernel(const bf16* __restrict__ x, bf16* __restrict__ out, int H, int W, int Ho, int Wo, int C8, int taps,
                      int stride, long long total8) {
  const int Hp = H + 2, Wp = W + 2, Hop = Ho + 2, Wop = Wo + 2;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total8;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(idx % C8);
    long long r = idx / C8;
    const int tap = static_cast<int>(r % taps);
    r /= taps;
    const int ox = static_cast<int>(r % Wop);
    const int oy = static_cast<int>((r / Wop) % Hop);
    const long long n = r / (static_cast<long long>(Wop) * Hop);
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (ox > 0 && oy > 0 && ox < Wop - 1 && oy < Hop - 1) {
      // output pixel (oy - 1, ox - 1); input pixel (2 (oy - 1) + ky - 1, ..) in unpadded, + 1 in padded coordinates
      const int ky = taps == 9 ? tap / 3 : 1, kx = taps == 9 ? tap % 3 : 1;
      const int iy = stride * (oy - 1) + ky, ix = stride * (ox - 1) + kx;
      v = ld_nc_u4(reinterpret_cast<const uint4*>(x) + ((n * Hp + iy) * Wp + ix) * C8 + c);
    }
    st_na_u4(reinterpret_cast<uint4*>(out) + idx, v);
  }
}

// AdaptiveAvgPool2d(1) over the interior of [N, H + 2, W + 2, C]: fp32 sum / (H W), rounded to bf16 (ATen's mean on bf16)
__global__ void __launch_bounds__(RT_THREADS)
avgpool_ring_kernel(const bf16* __restrict__ x, bf16* __restrict__ out, int H, int W, int C8, long long total8) {
  const int Hp = H + 2, Wp = W + 2;
  const float inv = 1.0f / static_cast<float>(H * W);
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total8;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(idx % C8);
    const long long n = idx / C8;
    float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int y = 1; y <= H; ++y)
      for (int xx = 1; xx <= W; ++xx) {
        float f[8];
        rt_unpack8(ld_nc_u4(reinterpret_cast<const uint4*>(x) + ((n * Hp + y) * Wp + xx) * C8 + c), f);
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] += f[i];
      }
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] *= inv;
    reinterpret_cast<uint4*>(out)[idx] = rt_pack8(a);
  }
}

// mean over the P pixels of plain channels-last frames [N, P_alloc, C] (ops.FrameRows): same arithmetic as above
__global__ void __launch_bounds__(RT_THREADS)
avgpool_frames_kernel(const bf16* __restrict__ x, bf16* __restrict__ out, int P, int P_alloc, int C8, long long total8) {
  const float inv = 1.0f / static_cast<float>(P);
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total8;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(idx % C8);
    const long long n = idx / C8;
    float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int q = 0; q < P; ++q) {
      float f[8];
      rt_unpack8(ld_nc_u4(reinterpret_cast<const uint4*>(x) + (n * P_alloc + q) * C8 + c), f);
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] += f[i];
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] *= inv;
    reinterpret_cast<uint4*>(out)[idx] = rt_pack8(a);
  }
}

static int rt_grid(long long total) {
  long long blocks = ceil_div_ll(total, RT_THREADS);
  if (blocks > kNumSMs * 16LL) blocks = kNumSMs * 16LL;
  if (blocks < 1) blocks = 1;
  return static_cast<int>(blocks);
}

}  // namespace omni

extern "C" int omni_prelu_res_ring(void* x, const void* residual, const void* slope, const void* bias, const void* res_bias,
                                   int64_t N, int32_t H, int32_t W, int32_t C, void* stream) {
  using namespace omni;
  OMNI_CHECK_ARG(x && slope && N >= 0 && H > 0 && W > 0 && C > 0 && (C % 8) == 0);
  if (N == 0) return OMNI_OK;
  const long long total8 = N * (H + 2) * (W + 2) * (C / 8);
  prelu_res_ring_kernel<<<rt_grid(total8), RT_THREADS, 0, (cudaStream_t)stream>>>(
      (bf16*)x, (const bf16*)residual, (const bf16*)slope, (const bf16*)bias, (const bf16*)res_bias, H + 2, W + 2, C / 8, total8);
  OMNI_LAUNCH_CHECK();
  return OMNI_OK;
}

extern "C" int omni_gather_s2_ring(const void* x, void* out, int64_t N, int32_t H, int32_t W, int32_t C, int32_t taps,
                                   int32_t stride, void* stream) {
  using namespace omni;
  OMNI_CHECK_ARG(x && out && N >= 0 && H > 0 && W > 0 && C > 0 && (C % 8) == 0 && (taps == 1 || taps == 9));
  OMNI_CHECK_ARG(stride == 1 || stride == 2);
  if (N == 0) return OMNI_OK;
  const int Ho = (H - 1) / stride + 1, Wo = (W - 1) / stride + 1;   // 3x3 pad 1 and 1x1 pad 0 give the same output grid
  const long long total8 = N * (Ho + 2) * (Wo + 2) * taps * (C / 8);
  gather_s2_ring_kernel<<<rt_grid(total8), RT_THREADS, 0, (cudaStream_t)stream>>>((const bf16*)x, (bf16*)out, H, W, Ho, Wo, C / 8,
                                                                                 taps, stride, total8);
  OMNI_LAUNCH_CHECK();
  return OMNI_OK;
}

extern "C" int omni_avgpool_ring(const void* x, void* out, int64_t N, int32_t H, int32_t W, int32_t C, void* stream) {
  using namespace omni;
  OMNI_CHECK_ARG(x && out && N >= 0 && H > 0 && W > 0 && C > 0 && (C % 8) == 0);
  if (N == 0) return OMNI_OK;
  const long long total8 = N * (C / 8);
  avgpool_ring_kernel<<<rt_grid(total8), RT_THREADS, 0, (cudaStream_t)stream>>>((const bf16*)x, (bf16*)out, H, W, C / 8, total8);
  OMNI_LAUNCH_CHECK();
  return OMNI_OK;
}

extern "C" int omni_avgpool_frames(const void* x, void* out, int64_t N, int32_t P, int32_t P_alloc, int32_t C, void* stream) {
  using namespace omni;
  OMNI_CHECK_ARG(x && out && N >= 0 && P > 0 && P_alloc >= P && C > 0 && (C % 8) == 0);
  if (N == 0) return OMNI_OK;
  const long long total8 = N * (C / 8);
  avgpool_frames_kernel<<<rt_grid(total8), RT_THREADS, 0, (cudaStream_t)stream>>>((const bf16*)x, (bf16*)out, P, P_alloc, C / 8,
                                                                                total8);
  OMNI_LAUNCH_CHECK();
  return OMNI_OK;
}
