// HBM-bound kernels of the Matryoshka path: token compression (avg-pool / stack) of encoder features,
// and the splice of media tokens + marker / prompt / text embeddings into the LLM input rows (+ labels).
// All loads/stores are 16-byte vectors, consecutive threads touch consecutive 16-byte chunks of one row,
// several independent loads are kept in flight per thread; grids are sized from the SM count.
//
// Reference semantics: Omni_AVSR/modeling_OmniAVSR.py:537-588 (audio), :465-514 (video) for compression,
// :270-299, :337-395 (train) and :406-458 (infer) for the splice and the labels.
#include "splice_common.cuh"

namespace omni {

constexpr int CS_THREADS = 256;

// ---------------------------------------------------------------------------------------------
// avg-pool: one thread = 8 channels of one output token
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(CS_THREADS)
compress_avg_kernel(const bf16* __restrict__ x, bf16* __restrict__ out, int n_out, int D8, long long x_bs,
                    int rate, long long total) {
  const float frate = static_cast<float>(rate);
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c8 = static_cast<int>(idx % D8);
    const long long t = idx / D8;
    const int j = static_cast<int>(t % n_out);
    const int b = static_cast<int>(t / n_out);
    const uint4* src = reinterpret_cast<const uint4*>(x + b * x_bs) + (static_cast<long long>(j) * rate) * D8 + c8;
    float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    int i = 0;
    for (; i + 4 <= rate; i += 4) {
      // 4 independent 16B loads in flight, accumulated in window order (matches the sequential fp32 sum)
      const uint4 u0 = ld_nc_u4(src + static_cast<long long>(i + 0) * D8);
      const uint4 u1 = ld_nc_u4(src + static_cast<long long>(i + 1) * D8);
      const uint4 u2 = ld_nc_u4(src + static_cast<long long>(i + 2) * D8);
      const uint4 u3 = ld_nc_u4(src + static_cast<long long>(i + 3) * D8);
      acc8(a, u0); acc8(a, u1); acc8(a, u2); acc8(a, u3);
    }
    for (; i < rate; ++i) acc8(a, ld_nc_u4(src + static_cast<long long>(i) * D8));
    uint4 o;
    o.x = f2_to_bf2(a[0] / frate, a[1] / frate);
    o.y = f2_to_bf2(a[2] / frate, a[3] / frate);
    o.z = f2_to_bf2(a[4] / frate, a[5] / frate);
    o.w = f2_to_bf2(a[6] / frate, a[7] / frate);
    st_na_u4(reinterpret_cast<uint4*>(out) + idx, o);
  }
}

// Same, with the window length known at compile time: all RATE 16-byte loads of a window are issued before the first
// add (one DRAM round trip per output instead of RATE/4), the sum still runs in window order (bit-exact with ATen).
template <int RATE>
__global__ void __launch_bounds__(CS_THREADS)
compress_avg_fixed_kernel(const bf16* __restrict__ x, bf16* __restrict__ out, int n_out, int D8, long long x_bs,
                          long long total) {
  constexpr float frate = static_cast<float>(RATE);
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c8 = static_cast<int>(idx % D8);
    const long long t = idx / D8;
    const int j = static_cast<int>(t % n_out);
    const int b = static_cast<int>(t / n_out);
    const uint4* src = reinterpret_cast<const uint4*>(x + b * x_bs) + (static_cast<long long>(j) * RATE) * D8 + c8;
    uint4 u[RATE];
#pragma unroll
    for (int i = 0; i < RATE; ++i) u[i] = ld_nc_u4(src + static_cast<long long>(i) * D8);
    float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < RATE; ++i) acc8(a, u[i]);
    uint4 o;
    o.x = f2_to_bf2(a[0] / frate, a[1] / frate);
    o.y = f2_to_bf2(a[2] / frate, a[3] / frate);
    o.z = f2_to_bf2(a[4] / frate, a[5] / frate);
    o.w = f2_to_bf2(a[6] / frate, a[7] / frate);
    st_na_u4(reinterpret_cast<uint4*>(out) + idx, o);
  }
}

// stack = per-clip contiguous copy of the first n_out*rate rows (row-major [n_out, rate*D] == [n_out*rate, D])
__global__ void __launch_bounds__(CS_THREADS)
compress_stack_kernel(const bf16* __restrict__ x, bf16* __restrict__ out, long long per_clip8, long long x_bs,
                      long long total) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  for (; idx + 3 * stride < total; idx += 4 * stride) {
    uint4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long k = idx + u * stride;
      const long long b = k / per_clip8;
      const long long o = k - b * per_clip8;
      v[u] = ld_nc_u4(reinterpret_cast<const uint4*>(x + b * x_bs) + o);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) st_na_u4(reinterpret_cast<uint4*>(out) + idx + u * stride, v[u]);
  }
  for (; idx < total; idx += stride) {
    const long long b = idx / per_clip8;
    const long long o = idx - b * per_clip8;
    st_na_u4(reinterpret_cast<uint4*>(out) + idx, ld_nc_u4(reinterpret_cast<const uint4*>(x + b * x_bs) + o));
  }
}

// backward: thread = 8 channels of one input row t < n_tok
__global__ void __launch_bounds__(CS_THREADS)
compress_bwd_kernel(const bf16* __restrict__ dout, bf16* __restrict__ dx, int n_tok, int n_out, int D8,
                    long long dx_bs, int rate, int mode, long long total) {
  const float frate = static_cast<float>(rate);
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c8 = static_cast<int>(idx % D8);
    const long long t2 = idx / D8;
    const int t = static_cast<int>(t2 % n_tok);
    const int b = static_cast<int>(t2 / n_tok);
    const int j = t / rate;
    uint4 o = make_uint4(0u, 0u, 0u, 0u);
    if (j < n_out) {
      if (mode == OMNI_COMPRESS_AVG) {
        const uint4 g = ld_nc_u4(reinterpret_cast<const uint4*>(dout) + (static_cast<long long>(b) * n_out + j) * D8 + c8);
        float2 f;
        f = bf2_to_f2(g.x); o.x = f2_to_bf2(f.x / frate, f.y / frate);
        f = bf2_to_f2(g.y); o.y = f2_to_bf2(f.x / frate, f.y / frate);
        f = bf2_to_f2(g.z); o.z = f2_to_bf2(f.x / frate, f.y / frate);
        f = bf2_to_f2(g.w); o.w = f2_to_bf2(f.x / frate, f.y / frate);
      } else {
        const int i = t - j * rate;
        o = ld_nc_u4(reinterpret_cast<const uint4*>(dout) +
                     ((static_cast<long long>(b) * n_out + j) * rate + i) * D8 + c8);
      }
    }
    st_na_u4(reinterpret_cast<uint4*>(dx + b * dx_bs) + static_cast<long long>(t) * D8 + c8, o);
  }
}

// ---------------------------------------------------------------------------------------------
// splice
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(CS_THREADS)
splice_kernel(const SpliceK k, long long total_rows) {
  // one warp per destination row; each lane moves 16B chunks lane, lane+32, ... (4 loads in flight)
  const int lane = threadIdx.x & 31;
  const long long warp_global = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
  const long long n_warps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  for (long long g = warp_global; g < total_rows; g += n_warps) {
    int t = 0;
    long long base = 0;
    if (g >= k.row_end[0]) { t = 1; base = k.row_end[0]; }
    if (g >= k.row_end[1]) { t = 2; base = k.row_end[1]; }
    const long long rr = g - base;
    const int b = static_cast<int>(rr / k.S[t]);
    const int pos = static_cast<int>(rr - static_cast<long long>(b) * k.S[t]);
    const RowSrc src = splice_resolve(k, t, b, pos);
    if (lane == 0 && k.out_labels[t]) k.out_labels[t][rr] = src.label;
    if (!k.out[t]) continue;
    uint4* dst = reinterpret_cast<uint4*>(k.out[t]) + rr * k.H8;
    if (src.ptr) {
      const uint4* s = reinterpret_cast<const uint4*>(src.ptr);
      int c = lane;
      for (; c + 96 < k.H8; c += 128) {
        const uint4 v0 = ld_nc_u4(s + c);
        const uint4 v1 = ld_nc_u4(s + c + 32);
        const uint4 v2 = ld_nc_u4(s + c + 64);
        const uint4 v3 = ld_nc_u4(s + c + 96);
        st_na_u4(dst + c, v0);
        st_na_u4(dst + c + 32, v1);
        st_na_u4(dst + c + 64, v2);
        st_na_u4(dst + c + 96, v3);
      }
      for (; c < k.H8; c += 32) st_na_u4(dst + c, ld_nc_u4(s + c));
    } else {
      for (int c = lane; c < k.H8; c += 32) st_na_u4(dst + c, make_uint4(0u, 0u, 0u, 0u));
    }
  }
}

// backward wrt the media tokens: one warp per media token row (audio rows first, then video rows)
__global__ void __launch_bounds__(CS_THREADS)
splice_bwd_kernel(const SpliceK k, const bf16* d0, const bf16* d1, const bf16* d2, bf16* d_audio, bf16* d_video,
                  long long rows_a, long long rows_total) {
  const int lane = threadIdx.x & 31;
  const long long warp_global = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
  const long long n_warps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  for (long long g = warp_global; g < rows_total; g += n_warps) {
    const bool is_audio = g < rows_a;
    const long long rr = is_audio ? g : g - rows_a;
    const int n = is_audio ? k.n_a : k.n_v;
    const int b = static_cast<int>(rr / n);
    const int i = static_cast<int>(rr - static_cast<long long>(b) * n);
    // own-task sequence (task 0 for audio, 1 for video) and the AVSR sequence (task 2)
    const int t_own = is_audio ? 0 : 1;
    const bf16* d_own = is_audio ? d0 : d1;
    const int pos_own = k.has_bos + 1 + i;
    const int pos_av = k.has_bos + ((is_audio || !k.has_a[2]) ? 0 : (k.n_a + 2)) + 1 + i;
    const uint4* s0 = (d_own && k.S[t_own]) ? reinterpret_cast<const uint4*>(d_own) +
                                                  (static_cast<long long>(b) * k.S[t_own] + pos_own) * k.H8
                                            : nullptr;
    const uint4* s1 = (d2 && k.S[2] && (is_audio ? k.has_a[2] : k.has_v[2]))
                          ? reinterpret_cast<const uint4*>(d2) + (static_cast<long long>(b) * k.S[2] + pos_av) * k.H8
                          : nullptr;
    uint4* dst = reinterpret_cast<uint4*>(is_audio ? d_audio : d_video) + rr * k.H8;
    for (int c = lane; c < k.H8; c += 32) {
      float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (s0) acc8(a, ld_nc_u4(s0 + c));
      if (s1) acc8(a, ld_nc_u4(s1 + c));
      uint4 o;
      o.x = f2_to_bf2(a[0], a[1]); o.y = f2_to_bf2(a[2], a[3]);
      o.z = f2_to_bf2(a[4], a[5]); o.w = f2_to_bf2(a[6], a[7]);
      st_na_u4(dst + c, o);
    }
  }
}

static int grid_for(long long work_items, int per_block) {
  long long blocks = ceil_div_ll(work_items, per_block);
  const long long cap = static_cast<long long>(kNumSMs) * 8;  // 8 resident 256-thread CTAs per SM
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return static_cast<int>(blocks);
}

}  // namespace omni

extern "C" int32_t omni_splice_seq_len(const omni_splice_args* a, int32_t t) {
  if (!a || t < 0 || t > 2) return -1;
  const int has_a = ((t == 0 || t == 2) && a->audio_tok != nullptr) ? 1 : 0;
  const int has_v = ((t == 1 || t == 2) && a->video_tok != nullptr) ? 1 : 0;
  return (a->has_bos ? 1 : 0) + has_a * (a->n_a + 2) + has_v * (a->n_v + 2) + a->prompt_len[t] +
         (a->L - (a->has_bos ? 1 : 0));
}

extern "C" int omni_splice_prompt(const omni_splice_args* a, void* stream) {
  using namespace omni;
  OMNI_CHECK_ARG(a != nullptr);
  SpliceK k;
  int rc = fill_splice(a, &k);
  if (rc) return rc;
  const long long total_rows = k.row_end[2];
  if (total_rows == 0) return OMNI_OK;
  const int grid = grid_for(total_rows, CS_THREADS / 32);
  splice_kernel<<<grid, CS_THREADS, 0, reinterpret_cast<cudaStream_t>(stream)>>>(k, total_rows);
  OMNI_LAUNCH_CHECK();
  return OMNI_OK;
}

extern "C" int omni_splice_prompt_bwd(const omni_splice_args* a, const void* const dout[3], void* d_audio_tok,
                                      void* d_video_tok, void* stream) {
  using namespace omni;
  OMNI_CHECK_ARG(a != nullptr && dout != nullptr);
  SpliceK k;
  int rc = fill_splice(a, &k);
  if (rc) return rc;
  const long long rows_a = d_audio_tok ? static_cast<long long>(a->B) * a->n_a : 0;
  const long long rows_v = d_video_tok ? static_cast<long long>(a->B) * a->n_v : 0;
  if (d_audio_tok) OMNI_CHECK_ARG(a->audio_tok != nullptr);
  if (d_video_tok) OMNI_CHECK_ARG(a->video_tok != nullptr);
  if (rows_a + rows_v == 0) return OMNI_OK;
  const int grid = grid_for(rows_a + rows_v, CS_THREADS / 32);
  splice_bwd_kernel<<<grid, CS_THREADS, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      k, reinterpret_cast<const bf16*>(dout[0]), reinterpret_cast<const bf16*>(dout[1]),
      reinterpret_cast<const bf16*>(dout[2]), reinterpret_cast<bf16*>(d_audio_tok),
      reinterpret_cast<bf16*>(d_video_tok), rows_a, rows_a + rows_v);
  OMNI_LAUNCH_CHECK();
  return OMNI_OK;
}

extern "C" int omni_matryoshka_compress(const void* x, void* out, int32_t B, int32_t n_tok, int32_t D, int64_t x_bs,
                                        int32_t rate, int32_t mode, void* stream) {
  using namespace omni;
  OMNI_CHECK_ARG(x && out && B > 0 && n_tok >= 0 && D > 0 && (D % 8) == 0 && rate >= 1);
  OMNI_CHECK_ARG((x_bs % 8) == 0 && x_bs >= static_cast<int64_t>(n_tok) * D);
  OMNI_CHECK_ARG(mode == OMNI_COMPRESS_AVG || mode == OMNI_COMPRESS_STACK);
  const int n_out = n_tok / rate;
  if (n_out == 0) return OMNI_OK;
  const int D8 = D / 8;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (mode == OMNI_COMPRESS_AVG) {
    const long long total = static_cast<long long>(B) * n_out * D8;
    const bf16* xp = reinterpret_cast<const bf16*>(x);
    bf16* op = reinterpret_cast<bf16*>(out);
    const int grid = grid_for(total, CS_THREADS);
    switch (rate) {   // the rates of the reference's recipes (README: audio 4/16, video 2/5) + neighbours
      case 2: compress_avg_fixed_kernel<2><<<grid, CS_THREADS, 0, st>>>(xp, op, n_out, D8, x_bs, total); break;
      case 3: compress_avg_fixed_kernel<3><<<grid, CS_THREADS, 0, st>>>(xp, op, n_out, D8, x_bs, total); break;
      case 4: compress_avg_fixed_kernel<4><<<grid, CS_THREADS, 0, st>>>(xp, op, n_out, D8, x_bs, total); break;
      case 5: compress_avg_fixed_kernel<5><<<grid, CS_THREADS, 0, st>>>(xp, op, n_out, D8, x_bs, total); break;
      case 8: compress_avg_fixed_kernel<8><<<grid, CS_THREADS, 0, st>>>(xp, op, n_out, D8, x_bs, total); break;
      case 16: compress_avg_fixed_kernel<16><<<grid, CS_THREADS, 0, st>>>(xp, op, n_out, D8, x_bs, total); break;
      default: compress_avg_kernel<<<grid, CS_THREADS, 0, st>>>(xp, op, n_out, D8, x_bs, rate, total);
    }
  } else {
    const long long per_clip8 = static_cast<long long>(n_out) * rate * D8;
    const long long total = per_clip8 * B;
    compress_stack_kernel<<<grid_for(total, CS_THREADS * 4), CS_THREADS, 0, st>>>(
        reinterpret_cast<const bf16*>(x), reinterpret_cast<bf16*>(out), per_clip8, x_bs, total);
  }
  OMNI_LAUNCH_CHECK();
  return OMNI_OK;
}

extern "C" int omni_matryoshka_compress_bwd(const void* dout, void* dx, int32_t B, int32_t n_tok, int32_t D,
                                            int64_t dx_bs, int32_t rate, int32_t mode, void* stream) {
  using namespace omni;
  OMNI_CHECK_ARG(dout && dx && B > 0 && n_tok > 0 && D > 0 && (D % 8) == 0 && rate >= 1);
  OMNI_CHECK_ARG((dx_bs % 8) == 0 && dx_bs >= static_cast<int64_t>(n_tok) * D);
  OMNI_CHECK_ARG(mode == OMNI_COMPRESS_AVG || mode == OMNI_COMPRESS_STACK);
  const int D8 = D / 8;
  const long long total = static_cast<long long>(B) * n_tok * D8;
  compress_bwd_kernel<<<grid_for(total, CS_THREADS), CS_THREADS, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const bf16*>(dout), reinterpret_cast<bf16*>(dx), n_tok, n_tok / rate, D8, dx_bs, rate, mode,
      total);
  OMNI_LAUNCH_CHECK();
  return OMNI_OK;
}
