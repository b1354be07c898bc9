// Cross-entropy over bf16 logits (fp32 math, as `logits.float()` + CrossEntropyLoss at
// Omni_AVSR/Llama_LoRA.py:373-386 / Qwen_LoRA.py:179-192) and the fused grad-norm-clip + AdamW update of the flat
// trainable-parameter buffer (lightning_OmniAVSR.py:152-157, train_OmniAVSR.py:53 gradient_clip_val=10).
#include "common.cuh"
#include "../../include/omni_avsr.h"

namespace omni {

constexpr int CE_THREADS = 1024;

__device__ __forceinline__ void online_merge(float& m, float& s, float m2, float s2) {
  const float mn = fmaxf(m, m2);
  s = s * __expf(m - mn) + s2 * __expf(m2 - mn);
  m = mn;
}

// one CTA per row: lse[r] = logsumexp(logits[r, :V]); loss[r] = lse - logits[r, target] (0 if target == ignore)
__global__ void __launch_bounds__(CE_THREADS)
ce_fwd_kernel(const bf16* __restrict__ logits, const int64_t* __restrict__ targets, float* __restrict__ loss,
              float* __restrict__ lse_out, int V, long long ld, long long ignore_index) {
  __shared__ float sm[32], ss[32];
  const long long r = blockIdx.x;
  const bf16* row = logits + r * ld;
  float m = -INFINITY, s = 0.f;
  const int V8 = V / 8;
  for (int c = threadIdx.x; c < V8; c += CE_THREADS) {
    const uint4 u = ld_nc_u4(reinterpret_cast<const uint4*>(row) + c);
    float f[8];
    float2 t;
    t = bf2_to_f2(u.x); f[0] = t.x; f[1] = t.y;
    t = bf2_to_f2(u.y); f[2] = t.x; f[3] = t.y;
    t = bf2_to_f2(u.z); f[4] = t.x; f[5] = t.y;
    t = bf2_to_f2(u.w); f[6] = t.x; f[7] = t.y;
    float cm = f[0];
#pragma unroll
    for (int i = 1; i < 8; ++i) cm = fmaxf(cm, f[i]);
    float cs = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) cs += __expf(f[i] - cm);
    online_merge(m, s, cm, cs);
  }
  for (int v = V8 * 8 + threadIdx.x; v < V; v += CE_THREADS) online_merge(m, s, __bfloat162float(row[v]), 1.0f);
  // warp then block reduction of (m, s)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float m2 = __shfl_xor_sync(0xffffffffu, m, o);
    const float s2 = __shfl_xor_sync(0xffffffffu, s, o);
    if (m2 > -INFINITY) online_merge(m, s, m2, s2);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { sm[warp] = m; ss[warp] = s; }
  __syncthreads();
  if (warp == 0) {
    m = sm[lane];
    s = ss[lane];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float m2 = __shfl_xor_sync(0xffffffffu, m, o);
      const float s2 = __shfl_xor_sync(0xffffffffu, s, o);
      if (m2 > -INFINITY) online_merge(m, s, m2, s2);
    }
    if (lane == 0) {
      const float lse = m + logf(s);
      lse_out[r] = lse;
      const long long t = targets[r];
      loss[r] = (t == ignore_index) ? 0.f : (lse - __bfloat162float(row[t]));
    }
  }
}

// in place: logits[r, v] <- bf16( (exp(logits - lse) - [v == target]) * scale[r] ); ignored rows -> 0
__global__ void __launch_bounds__(CE_THREADS)
ce_bwd_kernel(bf16* __restrict__ logits, const int64_t* __restrict__ targets, const float* __restrict__ lse,
              const float* __restrict__ scale, int V, long long ld, long long ignore_index) {
  const long long r = blockIdx.x;
  bf16* row = logits + r * ld;
  const long long t = targets[r];
  const float sc = (t == ignore_index) ? 0.f : scale[r];
  const float l = lse[r];
  const int V8 = V / 8;
  for (int c = threadIdx.x; c < V8; c += CE_THREADS) {
    uint4* p = reinterpret_cast<uint4*>(row) + c;
    const uint4 u = *p;
    float f[8];
    float2 q;
    q = bf2_to_f2(u.x); f[0] = q.x; f[1] = q.y;
    q = bf2_to_f2(u.y); f[2] = q.x; f[3] = q.y;
    q = bf2_to_f2(u.z); f[4] = q.x; f[5] = q.y;
    q = bf2_to_f2(u.w); f[6] = q.x; f[7] = q.y;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float g = __expf(f[i] - l);
      if (static_cast<long long>(c) * 8 + i == t) g -= 1.0f;
      f[i] = g * sc;
    }
    uint4 o;
    o.x = f2_to_bf2(f[0], f[1]); o.y = f2_to_bf2(f[2], f[3]);
    o.z = f2_to_bf2(f[4], f[5]); o.w = f2_to_bf2(f[6], f[7]);
    *p = o;
  }
  for (int v = V8 * 8 + threadIdx.x; v < V; v += CE_THREADS) {
    float g = __expf(__bfloat162float(row[v]) - l);
    if (v == t) g -= 1.0f;
    row[v] = __float2bfloat16_rn(g * sc);
  }
}

// greedy argmax over fp32(logits[r, :V]) -> int64 (first maximal index, as torch.argmax)
__global__ void __launch_bounds__(CE_THREADS)
argmax_kernel(const bf16* __restrict__ logits, int64_t* __restrict__ out, int V, long long ld) {
  __shared__ float sv[32];
  __shared__ int si[32];
  const long long r = blockIdx.x;
  const bf16* row = logits + r * ld;
  float best = -INFINITY;
  int bi = 0x7fffffff;
  for (int v = threadIdx.x; v < V; v += CE_THREADS) {
    const float f = __bfloat162float(row[v]);
    if (f > best || (f == best && v < bi)) { best = f; bi = v; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float f2 = __shfl_xor_sync(0xffffffffu, best, o);
    const int i2 = __shfl_xor_sync(0xffffffffu, bi, o);
    if (f2 > best || (f2 == best && i2 < bi)) { best = f2; bi = i2; }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { sv[warp] = best; si[warp] = bi; }
  __syncthreads();
  if (warp == 0) {
    best = sv[lane];
    bi = si[lane];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float f2 = __shfl_xor_sync(0xffffffffu, best, o);
      const int i2 = __shfl_xor_sync(0xffffffffu, bi, o);
      if (f2 > best || (f2 == best && i2 < bi)) { best = f2; bi = i2; }
    }
    if (lane == 0) out[r] = bi;
  }
}

// One launch for the token bookkeeping of a greedy decode step (HF `_sample` semantics, transformers 4.43.1, as driven by
// modeling_OmniAVSR.py:313-322): argmax of the row (16-byte loads; first maximal index, as torch.argmax), pad for finished
// sequences, the output slot of this step, the unfinished flag (EOS), "any sequence still running", and the embedding row of
// the chosen token written straight into the next forward's input.  Replaces argmax + ~12 elementwise / index ATen kernels.
__global__ void __launch_bounds__(CE_THREADS)
decode_pick_kernel(const bf16* __restrict__ logits, int V, long long ld, long long* __restrict__ unfinished,
                   const long long* __restrict__ eos, const long long* __restrict__ pad, const long long* __restrict__ step_idx,
                   long long* __restrict__ out, int B, long long* __restrict__ alive, const bf16* __restrict__ embed,
                   long long ld_embed, bf16* __restrict__ x_next, long long ld_x, int H8) {
  __shared__ float sv[32];
  __shared__ int si[32];
  __shared__ long long s_tok;
  const int b = blockIdx.x;
  const bf16* row = logits + static_cast<long long>(b) * ld;
  float best = -INFINITY;
  int bi = 0x7fffffff;
  auto take = [&](float f, int v) {
    if (f > best || (f == best && v < bi)) { best = f; bi = v; }
  };
  const int V8 = ((reinterpret_cast<uintptr_t>(row) & 15) == 0) ? (V >> 3) : 0;      // 16-byte chunks when the row is aligned
  for (int c = threadIdx.x; c < V8; c += CE_THREADS) {
    const uint4 u = ld_nc_u4(reinterpret_cast<const uint4*>(row) + c);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 f = bf2_to_f2(w[i]);
      take(f.x, c * 8 + 2 * i);
      take(f.y, c * 8 + 2 * i + 1);
    }
  }
  for (int v = V8 * 8 + threadIdx.x; v < V; v += CE_THREADS) take(__bfloat162float(row[v]), v);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float f2 = __shfl_xor_sync(0xffffffffu, best, o);
    const int i2 = __shfl_xor_sync(0xffffffffu, bi, o);
    if (f2 > best || (f2 == best && i2 < bi)) { best = f2; bi = i2; }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { sv[warp] = best; si[warp] = bi; }
  __syncthreads();
  if (warp == 0) {
    best = lane < CE_THREADS / 32 ? sv[lane] : -INFINITY;
    bi = lane < CE_THREADS / 32 ? si[lane] : 0x7fffffff;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float f2 = __shfl_xor_sync(0xffffffffu, best, o);
      const int i2 = __shfl_xor_sync(0xffffffffu, bi, o);
      if (f2 > best || (f2 == best && i2 < bi)) { best = f2; bi = i2; }
    }
    if (lane == 0) {
      const long long step = *step_idx;
      const long long u = unfinished[b];
      const long long tok = u ? static_cast<long long>(bi) : *pad;       // next = argmax * unfinished + pad * (1 - unfinished)
      out[step * B + b] = tok;
      const long long u2 = (u && tok != *eos) ? 1 : 0;
      unfinished[b] = u2;
      if (u2) atomicMax(reinterpret_cast<unsigned long long*>(alive + step), 1ull);
      s_tok = tok;
    }
  }
  __syncthreads();
  const uint4* src = reinterpret_cast<const uint4*>(embed + s_tok * ld_embed);
  uint4* dst = reinterpret_cast<uint4*>(x_next + static_cast<long long>(b) * ld_x);
  for (int c = threadIdx.x; c < H8; c += CE_THREADS) dst[c] = __ldg(src + c);
}

// step_idx += 1, len_idx += 1, pos[i] += 1: the device-side counters of the captured decode step, after its forward
__global__ void decode_advance_kernel(long long* step_idx, long long* len_idx, int* pos, int n_pos) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_pos) pos[i] += 1;
  if (i == 0) {
    *step_idx += 1;
    *len_idx += 1;
  }
}

// sum of squares of a bf16 buffer into *acc (fp32, atomically) -- global grad norm
__global__ void __launch_bounds__(256)
sumsq_kernel(const bf16* __restrict__ g, long long n8, long long n, float* __restrict__ acc) {
  float s = 0.f;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n8;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const uint4 u = ld_nc_u4(reinterpret_cast<const uint4*>(g) + i);
    float2 t;
    t = bf2_to_f2(u.x); s += t.x * t.x + t.y * t.y;
    t = bf2_to_f2(u.y); s += t.x * t.x + t.y * t.y;
    t = bf2_to_f2(u.z); s += t.x * t.x + t.y * t.y;
    t = bf2_to_f2(u.w); s += t.x * t.x + t.y * t.y;
  }
  if (blockIdx.x == 0) {
    for (long long i = n8 * 8 + threadIdx.x; i < n; i += blockDim.x) {
      const float f = __bfloat162float(g[i]);
      s += f * f;
    }
  }
  s = warp_sum(s);
  __shared__ float sh[8];
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 8) {
    s = sh[threadIdx.x];
    s += __shfl_xor_sync(0xffu, s, 4);
    s += __shfl_xor_sync(0xffu, s, 2);
    s += __shfl_xor_sync(0xffu, s, 1);
    if (threadIdx.x == 0) atomicAdd(acc, s);
  }
}

// AdamW (decoupled weight decay, torch semantics) on bf16 params with fp32 moments; the gradient is scaled by
// grad_scale * min(1, max_norm / (sqrt(*sumsq * grad_scale^2) + 1e-6)) when max_norm > 0.
__global__ void __launch_bounds__(256)
adamw_kernel(bf16* __restrict__ p, const bf16* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
             long long n, float lr, float beta1, float beta2, float eps, float wd, float bc1, float bc2,
             float grad_scale, float max_norm, const float* __restrict__ sumsq) {
  float coef = grad_scale;
  if (max_norm > 0.f && sumsq) {
    const float total = sqrtf(*sumsq) * grad_scale;
    coef *= fminf(1.0f, max_norm / (total + 1e-6f));
  }
  const float step = lr / bc1;
  const float inv_sqrt_bc2 = rsqrtf(bc2);
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float gi = __bfloat162float(g[i]) * coef;
    float pi = __bfloat162float(p[i]);
    pi *= (1.0f - lr * wd);
    const float mi = beta1 * m[i] + (1.0f - beta1) * gi;
    const float vi = beta2 * v[i] + (1.0f - beta2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) * inv_sqrt_bc2 + eps;
    pi -= step * mi / denom;
    p[i] = __float2bfloat16_rn(pi);
  }
}

}  // namespace omni

using namespace omni;

extern "C" int omni_ce_fwd(const void* logits, const int64_t* targets, float* loss, float* lse, int64_t rows, int32_t V,
                           int64_t ld, int64_t ignore_index, void* stream) {
  OMNI_CHECK_ARG(logits && targets && loss && lse && rows >= 0 && V > 0 && (ld % 8) == 0 && ld >= V);
  if (rows == 0) return OMNI_OK;
  ce_fwd_kernel<<<(unsigned)rows, CE_THREADS, 0, (cudaStream_t)stream>>>((const bf16*)logits, targets, loss, lse, V, ld,
                                                                         ignore_index);
  OMNI_LAUNCH_CHECK();
  return OMNI_OK;
}

extern "C" int omni_ce_bwd(void* logits, const int64_t* targets, const float* lse, const float* scale, int64_t rows,
                           int32_t V, int64_t ld, int64_t ignore_index, void* stream) {
  OMNI_CHECK_ARG(logits && targets && lse && scale && rows >= 0 && V > 0 && (ld % 8) == 0 && ld >= V);
  if (rows == 0) return OMNI_OK;
  ce_bwd_kernel<<<(unsigned)rows, CE_THREADS, 0, (cudaStream_t)stream>>>((bf16*)logits, targets, lse, scale, V, ld,
                                                                         ignore_index);
  OMNI_LAUNCH_CHECK();
  return OMNI_OK;
}

extern "C" int omni_argmax(const void* logits, int64_t* out, int64_t rows, int32_t V, int64_t ld, void* stream) {
  OMNI_CHECK_ARG(logits && out && rows >= 0 && V > 0 && ld >= V);
  if (rows == 0) return OMNI_OK;
  argmax_kernel<<<(unsigned)rows, CE_THREADS, 0, (cudaStream_t)stream>>>((const bf16*)logits, out, V, ld);
  OMNI_LAUNCH_CHECK();
  return OMNI_OK;
}

extern "C" int omni_decode_pick(const void* logits, int32_t V, int64_t ld, int64_t* unfinished, const int64_t* eos,
                                const int64_t* pad, const int64_t* step_idx, int64_t* out, int32_t B, int64_t* alive,
                                const void* embed, int64_t ld_embed, void* x_next, int64_t ld_x, int32_t H, void* stream) {
  OMNI_CHECK_ARG(logits && unfinished && eos && pad && step_idx && out && alive && embed && x_next && B > 0 && V > 0 && ld >= V);
  OMNI_CHECK_ARG(H > 0 && (H % 8) == 0 && (ld_embed % 8) == 0 && (ld_x % 8) == 0 &&
                 (reinterpret_cast<uintptr_t>(embed) & 15) == 0 && (reinterpret_cast<uintptr_t>(x_next) & 15) == 0);
  decode_pick_kernel<<<(unsigned)B, CE_THREADS, 0, (cudaStream_t)stream>>>(
      (const bf16*)logits, V, ld, (long long*)unfinished, (const long long*)eos, (const long long*)pad,
      (const long long*)step_idx, (long long*)out, B, (long long*)alive, (const bf16*)embed, ld_embed, (bf16*)x_next, ld_x, H / 8);
  OMNI_LAUNCH_CHECK();
  return OMNI_OK;
}

extern "C" int omni_decode_advance(int64_t* step_idx, int64_t* len_idx, int32_t* pos, int32_t n_pos, void* stream) {
  OMNI_CHECK_ARG(step_idx && len_idx && pos && n_pos >= 0);
  decode_advance_kernel<<<(n_pos + 255) / 256 + (n_pos == 0), 256, 0, (cudaStream_t)stream>>>(
      (long long*)step_idx, (long long*)len_idx, pos, n_pos);
  OMNI_LAUNCH_CHECK();
  return OMNI_OK;
}

extern "C" int omni_sumsq(const void* g, int64_t n, float* acc, void* stream) {
  OMNI_CHECK_ARG(g && acc && n >= 0 && (reinterpret_cast<uintptr_t>(g) & 15) == 0);
  if (n == 0) return OMNI_OK;
  long long blocks = ceil_div_ll(n / 8 + 1, 256);
  if (blocks > kNumSMs * 8LL) blocks = kNumSMs * 8LL;
  sumsq_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>((const bf16*)g, n / 8, n, acc);
  OMNI_LAUNCH_CHECK();
  return OMNI_OK;
}

extern "C" int omni_adamw(void* p, const void* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
                          float eps, float weight_decay, int32_t step, float grad_scale, float max_norm,
                          const float* sumsq, void* stream) {
  OMNI_CHECK_ARG(p && g && m && v && n >= 0 && step >= 1);
  if (n == 0) return OMNI_OK;
  const float bc1 = 1.0f - powf(beta1, (float)step);
  const float bc2 = 1.0f - powf(beta2, (float)step);
  long long blocks = ceil_div_ll(n, 256);
  if (blocks > kNumSMs * 8LL) blocks = kNumSMs * 8LL;
  adamw_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>((bf16*)p, (const bf16*)g, m, v, n, lr, beta1, beta2, eps,
                                                              weight_decay, bc1, bc2, grad_scale, max_norm, sumsq);
  OMNI_LAUNCH_CHECK();
  return OMNI_OK;
}
