// Single-token (decode-step) attention over the static KV cache, HBM-bound: one CTA per (clip, KV head) streams that
// head's K and V rows exactly once for all G query heads of the GQA group, and writes the new token's K / V into the
// cache on the way (replaces the index_copy_ + key-mask + F.scaled_dot_product_attention sequence of the eager step;
// reference call sites: Llama_LoRA.py:284-300 / Qwen_LoRA.py:590-606 with past_key_value during HF generate).
//
// Layout: HD/8 lanes cover one key row (one 16-byte load each, so a warp instruction reads 32/(HD/8) whole rows);
// the query fragments of all G heads live in registers.  Pass 1: scores -> shared memory; softmax per head by one
// warp; pass 2: P.V with per-thread partial sums over its key slots, reduced across the slots / warps at the end.
// The position of the new token is read from device memory (the decode step is replayed from a CUDA graph).
#include "common.cuh"
#include "../../include/omni_avsr.h"

namespace omni {

constexpr int DA_THREADS = 128;
constexpr int DA_WARPS = DA_THREADS / 32;
constexpr int DA_UNROLL = 8;       // key rows in flight per lane

__device__ __forceinline__ void unpack_u4(const uint4& u, float (&f)[8]) {
  float2 t;
  t = bf2_to_f2(u.x); f[0] = t.x; f[1] = t.y;
  t = bf2_to_f2(u.y); f[2] = t.x; f[3] = t.y;
  t = bf2_to_f2(u.z); f[4] = t.x; f[5] = t.y;
  t = bf2_to_f2(u.w); f[6] = t.x; f[7] = t.y;
}

__device__ __forceinline__ float rbf16(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

// RoPE of 8 consecutive dims of one head held by a lane (partner dims +- HD/2 live in lane l ^ (LPK / 2)):
//   out = bf16( bf16(x * cos) + bf16(rotate_half(x) * sin) )  -- the rounding points of apply_rotary_pos_emb in bf16
// (transformers modeling_llama.py, called at Llama_LoRA.py:277), identical to omni_rope.
template <int LPK>
__device__ __forceinline__ void rope8(float (&x)[8], const float (&cs)[8], const float (&sn)[8], bool second_half) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float xp = __shfl_xor_sync(0xffffffffu, x[i], LPK / 2);
    const float rot = second_half ? xp : -xp;
    x[i] = rbf16(rbf16(x[i] * cs[i]) + rbf16(rot * sn[i]));
  }
}

template <int HD, int G>
__global__ void __launch_bounds__(DA_THREADS, 4)
decode_attn_kernel(const bf16* __restrict__ qkv, long long ld, bf16* __restrict__ kc, bf16* __restrict__ vc,
                   const long long* __restrict__ len_idx, bf16* __restrict__ out, long long out_ld, int n_kv_heads,
                   int max_len, float scale_log2, const bf16* __restrict__ cos_t, const bf16* __restrict__ sin_t) {
  pdl_launch_dependents();
  pdl_wait();                          // q|k|v row and the cache position come from predecessors
  constexpr int LPK = HD / 8;          // lanes per key row
  constexpr int KPW = 32 / LPK;        // key rows per warp instruction
  extern __shared__ float sm[];
  float* sc = sm;                                   // [G][max_len] scores, then unnormalised probabilities
  float* red = sm + G * max_len;                    // [DA_WARPS][G][HD] partial outputs
  float* inv_sum = red + DA_WARPS * G * HD;         // [G]

  const int b = blockIdx.x / n_kv_heads, kvh = blockIdx.x - b * n_kv_heads;
  const int n_heads = n_kv_heads * G;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sub = lane / LPK, l = lane - sub * LPK;
  const int pos = static_cast<int>(*len_idx);       // position of the new token == number of cached keys
  if (pos >= max_len) return;                       // cache full: the host sized max_len = prefill + max_new_tokens
  const int n_keys = pos + 1;

  const bf16* row = qkv + static_cast<long long>(b) * ld;
  // fused RoPE (cos_t != null): the new token's position is `pos`; q heads and the new key are rotated on the fly, so the
  // separate in-place rotation of the packed q|k|v row (one more launch per layer and step) disappears
  float cs[8], sn[8];
  const bool rope = cos_t != nullptr;
  const bool second_half = l >= LPK / 2;
  if (rope) {
    unpack_u4(__ldg(reinterpret_cast<const uint4*>(cos_t + static_cast<long long>(pos) * HD) + l), cs);
    unpack_u4(__ldg(reinterpret_cast<const uint4*>(sin_t + static_cast<long long>(pos) * HD) + l), sn);
  }
  float q[G][8];
#pragma unroll
  for (int g = 0; g < G; ++g) {
    unpack_u4(*reinterpret_cast<const uint4*>(row + (kvh * G + g) * HD + l * 8), q[g]);
    if (rope) rope8<LPK>(q[g], cs, sn, second_half);
#pragma unroll
    for (int i = 0; i < 8; ++i) q[g][i] *= scale_log2;
  }
  uint4 k_new = *reinterpret_cast<const uint4*>(row + (n_heads + kvh) * HD + l * 8);
  if (rope) {
    float kf[8];
    unpack_u4(k_new, kf);
    rope8<LPK>(kf, cs, sn, second_half);
    k_new.x = f2_to_bf2(kf[0], kf[1]); k_new.y = f2_to_bf2(kf[2], kf[3]);
    k_new.z = f2_to_bf2(kf[4], kf[5]); k_new.w = f2_to_bf2(kf[6], kf[7]);
  }
  const uint4 v_new = *reinterpret_cast<const uint4*>(row + (n_heads + n_kv_heads + kvh) * HD + l * 8);
  bf16* krow = kc + (static_cast<long long>(b) * n_kv_heads + kvh) * max_len * HD;
  bf16* vrow = vc + (static_cast<long long>(b) * n_kv_heads + kvh) * max_len * HD;
  if (threadIdx.x < LPK) {                          // append the new token to the cache (read back by later steps only)
    *reinterpret_cast<uint4*>(krow + static_cast<long long>(pos) * HD + l * 8) = k_new;
    *reinterpret_cast<uint4*>(vrow + static_cast<long long>(pos) * HD + l * 8) = v_new;
  }

  // ---- pass 1: scores (DA_UNROLL independent 16-byte loads in flight per lane: the kernel is latency-bound otherwise) ----
  for (int p0 = warp * KPW; p0 < n_keys; p0 += DA_UNROLL * DA_WARPS * KPW) {
    uint4 kk[DA_UNROLL];
#pragma unroll
    for (int u = 0; u < DA_UNROLL; ++u) {
      const int p = p0 + u * DA_WARPS * KPW + sub;
      kk[u] = k_new;
      if (p < n_keys && p != pos) {
        kk[u] = ld_nc_u4(krow + static_cast<long long>(p) * HD + l * 8);
        // pull the matching V row towards L2 now: pass 2 (which can only start after the softmax) then runs out of L2, and
        // the HBM streams of K and V overlap instead of being serialised by the two barriers in between
        if ((l & 7) == 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(vrow + static_cast<long long>(p) * HD + l * 8));
      }
    }
#pragma unroll
    for (int u = 0; u < DA_UNROLL; ++u) {
      const int p = p0 + u * DA_WARPS * KPW + sub;
      // packed fp32 (FFMA2): the kernel is issue-bound (ncu: 53 % issue slots at 13 warps / SM), the dot products are its
      // dominant instruction class -- two lanes of a register pair per slot
      const float2 kf2[4] = {bf2_to_f2(kk[u].x), bf2_to_f2(kk[u].y), bf2_to_f2(kk[u].z), bf2_to_f2(kk[u].w)};
#pragma unroll
      for (int g = 0; g < G; ++g) {
        float2 s2 = make_float2(0.f, 0.f);
#pragma unroll
        for (int i = 0; i < 4; ++i) s2 = ffma2(make_float2(q[g][2 * i], q[g][2 * i + 1]), kf2[i], s2);
        float s = s2.x + s2.y;
#pragma unroll
        for (int o = LPK >> 1; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (p < n_keys && l == 0) sc[g * max_len + p] = s;
      }
    }
  }
  __syncthreads();

  // ---- softmax: warp w handles heads w, w + DA_WARPS, ... ----
  for (int g = warp; g < G; g += DA_WARPS) {
    float* s = sc + g * max_len;
    float mx = -INFINITY;
    for (int p = lane; p < n_keys; p += 32) mx = fmaxf(mx, s[p]);
    mx = warp_max(mx);
    float sum = 0.f;
    for (int p = lane; p < n_keys; p += 32) {
      const float e = ex2_approx(s[p] - mx);
      s[p] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    if (lane == 0) inv_sum[g] = 1.0f / sum;
  }
  __syncthreads();

  // ---- pass 2: P.V ----
  float acc[G][8];
#pragma unroll
  for (int g = 0; g < G; ++g)
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[g][i] = 0.f;
  for (int p0 = warp * KPW; p0 < n_keys; p0 += DA_UNROLL * DA_WARPS * KPW) {
    uint4 vv[DA_UNROLL];
#pragma unroll
    for (int u = 0; u < DA_UNROLL; ++u) {
      const int p = p0 + u * DA_WARPS * KPW + sub;
      vv[u] = v_new;
      if (p < n_keys && p != pos) vv[u] = ld_nc_u4(vrow + static_cast<long long>(p) * HD + l * 8);
    }
#pragma unroll
    for (int u = 0; u < DA_UNROLL; ++u) {
      const int p = p0 + u * DA_WARPS * KPW + sub;
      if (p < n_keys) {
        const float2 vf2[4] = {bf2_to_f2(vv[u].x), bf2_to_f2(vv[u].y), bf2_to_f2(vv[u].z), bf2_to_f2(vv[u].w)};
#pragma unroll
        for (int g = 0; g < G; ++g) {
          const float pg = sc[g * max_len + p];
          const float2 pg2 = make_float2(pg, pg);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float2 a = ffma2(pg2, vf2[i], make_float2(acc[g][2 * i], acc[g][2 * i + 1]));
            acc[g][2 * i] = a.x;
            acc[g][2 * i + 1] = a.y;
          }
        }
      }
    }
  }
  // sum the key slots of a warp (lanes with equal l), then the warps through shared memory
#pragma unroll
  for (int g = 0; g < G; ++g)
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float a = acc[g][i];
#pragma unroll
      for (int o = LPK; o < 32; o <<= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
      acc[g][i] = a;
    }
  if (sub == 0) {
#pragma unroll
    for (int g = 0; g < G; ++g)
#pragma unroll
      for (int i = 0; i < 8; ++i) red[(warp * G + g) * HD + l * 8 + i] = acc[g][i];
  }
  __syncthreads();
  for (int e = threadIdx.x; e < G * HD; e += DA_THREADS) {
    const int g = e / HD;
    float a = 0.f;
#pragma unroll
    for (int w = 0; w < DA_WARPS; ++w) a += red[w * G * HD + e];
    out[static_cast<long long>(b) * out_ld + (kvh * G + g) * HD + (e - g * HD)] = __float2bfloat16_rn(a * inv_sum[g]);
  }
}


// ---------------------------------------------------------------------------------------------------------------------
// Tensor-core variant (default).  The FFMA kernel above is issue-bound (ncu: 53 % issue slots at 13 warps / SM, 22 us per
// layer against 9 us of K / V traffic at B = 64): per key and query head it spends 4 FFMA2 + 3 shuffle / add pairs on an
// 8-lane dot product.  Here both contractions run as warp-level mma.sync.m16n8k16 (bf16 in, fp32 accumulate) on fragments
// built DIRECTLY from the 16-byte global loads, no shared-memory staging of K or V:
//   pass 1  S^T[16 keys x 8 heads] += K[16 keys x 16 d] . Q^T[16 d x 8 heads]: the d index of a contraction may be permuted
//           freely, so lane (r = lane / 4, c = lane % 4) loads chunk c + 4q (8 consecutive d) of key rows r and r + 8 and
//           feeds the four registers of each uint4 as the (k = 2c, 2c + 1) / (k = 2c + 8, 2c + 9) halves of two MMAs; the
//           Q fragment of head r is the same chunk of the query row.  Heads >= G are zero fragments.
//   pass 2  O^T[16 d x 8 heads] += V^T[16 d x 16 keys] . P^T[16 keys x 8 heads]: lane (r, c) loads chunk r (+ 8h) of keys
//           4c .. 4c + 3, two PRMTs per register pair interleave two keys of one d, P = exp2(s - max) is computed from the
//           fp32 scores in shared memory on the fly (no separate softmax stage: the maximum comes out of pass 1's
//           accumulators) and packed to bf16 like every tensor-core attention does.
// One barrier between the passes, one before the cross-warp reduction of the outputs.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int DM_THREADS = 256;
constexpr int DM_WARPS = DM_THREADS / 32;

__device__ __forceinline__ void mma_16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                          uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

__device__ __forceinline__ uint4 ld_plain_u4(const void* p) {      // coherent load (rows this CTA wrote before a barrier)
  uint4 v;
  asm volatile("ld.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}

// RoPE of one 8-dim chunk given its partner chunk (+- HD/2); rounding points of omni_rope / rope8 above.
__device__ __forceinline__ uint4 rope_chunk(const uint4& x, const uint4& xp, const uint4& cs, const uint4& sn,
                                            bool second_half) {
  float xf[8], pf[8], cf[8], sf[8];
  unpack_u4(x, xf); unpack_u4(xp, pf); unpack_u4(cs, cf); unpack_u4(sn, sf);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float rot = second_half ? pf[i] : -pf[i];
    xf[i] = rbf16(rbf16(xf[i] * cf[i]) + rbf16(rot * sf[i]));
  }
  uint4 o;
  o.x = f2_to_bf2(xf[0], xf[1]); o.y = f2_to_bf2(xf[2], xf[3]);
  o.z = f2_to_bf2(xf[4], xf[5]); o.w = f2_to_bf2(xf[6], xf[7]);
  return o;
}

constexpr int DM_MAX_NEW = 128;      // generated positions a beam-search indirection row can describe

// BEAM: the B rows are utterances x K beams and the cache row of a key position is looked up (see omni_decode_attention_beam
// in the header: prompt rows once per utterance, generated positions through the indirection table of omni_beam_select).
template <int HD, int G, bool BEAM>
__global__ void __launch_bounds__(DM_THREADS, (HD == 64 ? 4 : 2))
decode_attn_mma_kernel(const bf16* __restrict__ qkv, long long ld, bf16* __restrict__ kc, bf16* __restrict__ vc,
                       const long long* __restrict__ len_idx, bf16* __restrict__ out, long long out_ld, int n_kv_heads,
                       int max_len, int sc_stride, float scale_log2, const bf16* __restrict__ cos_t,
                       const bf16* __restrict__ sin_t, const int* __restrict__ beam_ind, int ind_ld, long long ind_plane,
                       const long long* __restrict__ prefill_len, int beam_k) {
  pdl_launch_dependents();
  constexpr int NCH = HD / 8;          // 16-byte chunks per row
  constexpr int NQ = HD / 32;          // chunks per lane in pass 1 (c + 4q)
  constexpr int NH = HD / 64;          // chunks per lane in pass 2 (r + 8h)
  constexpr int U1 = (HD == 64) ? 2 : 1;   // 16-key blocks in flight per warp (8 x 16-byte loads per lane either way)
  constexpr int U2 = 1;
  extern __shared__ float sm[];
  float* sc = sm;                                   // [G][sc_stride] scaled scores (log2 units)
  float* red = sm + G * sc_stride + 16;             // [DM_WARPS][G][HD] partial outputs
  float* wmax = red + DM_WARPS * G * HD;            // [DM_WARPS][8]
  float* wsum = wmax + DM_WARPS * 8;                // [DM_WARPS][8]

  const int b = blockIdx.x / n_kv_heads, kvh = blockIdx.x - b * n_kv_heads;
  const int n_heads = n_kv_heads * G;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = lane >> 2, c = lane & 3;
  const int pos = static_cast<int>(*len_idx);       // position of the new token == number of cached keys
  if (pos >= max_len) return;                       // cache full: the host sized max_len = prefill + max_new_tokens
  const int n_keys = pos + 1;
  const int n_blocks = (n_keys + 15) >> 4;
  const bool rope = cos_t != nullptr;

  const bf16* row = qkv + static_cast<long long>(b) * ld;
  const long long row_elems = static_cast<long long>(max_len) * HD;
  bf16* krow = kc + (static_cast<long long>(b) * n_kv_heads + kvh) * row_elems;
  bf16* vrow = vc + (static_cast<long long>(b) * n_kv_heads + kvh) * row_elems;
  // Everything above and the cached rows [0, pos) were written by EARLIER steps (the position counter by the previous step's
  // omni_decode_advance, the rows by this layer's launches of the previous steps): under programmatic dependent launch they
  // are complete while the producer of the q|k|v row (the immediate predecessor) is still running, so this CTA pulls its K
  // and V rows towards L2 before griddepcontrol.wait -- the HBM stream of the cache overlaps the q|k|v GEMM instead of
  // starting after this kernel's own prologue.  (BEAM: the prompt part, which is all but <= max_new positions.)
  {
    const int n_old = BEAM ? min(pos, static_cast<int>(*prefill_len)) : pos;
    const long long poff = BEAM ? (static_cast<long long>(b / beam_k * beam_k) - b) * n_kv_heads * row_elems : 0;
    const int n_lines = (n_old * HD * 2) >> 7;                   // 128-byte lines of one row range
    for (int i = threadIdx.x; i < n_lines; i += DM_THREADS) {
      asm volatile("prefetch.global.L2 [%0];" ::"l"(krow + poff + i * 64));
      asm volatile("prefetch.global.L2 [%0];" ::"l"(vrow + poff + i * 64));
    }
  }
  pdl_wait();                          // the q|k|v row comes from the predecessor
  __shared__ int s_ind[BEAM ? DM_MAX_NEW : 1];
  int s0 = 0;
  long long prompt_off = 0;            // element offset of (prompt row, kvh) relative to (row b, kvh)
  if (BEAM) {
    s0 = static_cast<int>(*prefill_len);
    const int t_new = pos - s0;        // generated position of the new token
    const int* ind = beam_ind + ((t_new + 1) & 1) * ind_plane + static_cast<long long>(b) * ind_ld;
    for (int t = threadIdx.x; t < t_new && t < DM_MAX_NEW; t += DM_THREADS) s_ind[t] = ind[t];
    prompt_off = (static_cast<long long>(b / beam_k * beam_k) - b) * n_kv_heads * row_elems;
  }
  // element offset (relative to this CTA's own cache row) of the cache row holding key position p
  auto key_off = [&](int p) -> long long {
    if (!BEAM || p >= pos) return 0;
    if (p < s0) return prompt_off;
    return (static_cast<long long>(s_ind[p - s0]) - b) * n_kv_heads * row_elems;
  };

  // ---- append the new token (RoPE on the key), read back below with coherent loads after the barrier ----
  if (threadIdx.x < NCH) {
    const int ch = threadIdx.x;
    const bf16* ksrc = row + (n_heads + kvh) * HD;
    uint4 x = *reinterpret_cast<const uint4*>(ksrc + ch * 8);
    if (rope) {
      const uint4 xp = *reinterpret_cast<const uint4*>(ksrc + ((ch + NCH / 2) % NCH) * 8);
      const uint4 cs = __ldg(reinterpret_cast<const uint4*>(cos_t + static_cast<long long>(pos) * HD) + ch);
      const uint4 sn = __ldg(reinterpret_cast<const uint4*>(sin_t + static_cast<long long>(pos) * HD) + ch);
      x = rope_chunk(x, xp, cs, sn, ch >= NCH / 2);
    }
    *reinterpret_cast<uint4*>(krow + static_cast<long long>(pos) * HD + ch * 8) = x;
  } else if (threadIdx.x < 2 * NCH) {
    const int ch = threadIdx.x - NCH;
    *reinterpret_cast<uint4*>(vrow + static_cast<long long>(pos) * HD + ch * 8) =
        *reinterpret_cast<const uint4*>(row + (n_heads + n_kv_heads + kvh) * HD + ch * 8);
  }

  // ---- Q fragments of head r (zero for r >= G): chunks c + 4q ----
  uint4 qf[NQ];
#pragma unroll
  for (int q = 0; q < NQ; ++q) qf[q] = make_uint4(0u, 0u, 0u, 0u);
  if (r < G) {
    const bf16* qsrc = row + (kvh * G + r) * HD;
#pragma unroll
    for (int q = 0; q < NQ; ++q) qf[q] = *reinterpret_cast<const uint4*>(qsrc + (c + 4 * q) * 8);
    if (rope) {
      uint4 rot[NQ];
#pragma unroll
      for (int q = 0; q < NQ; ++q) {
        const int ch = c + 4 * q;
        const uint4 cs = __ldg(reinterpret_cast<const uint4*>(cos_t + static_cast<long long>(pos) * HD) + ch);
        const uint4 sn = __ldg(reinterpret_cast<const uint4*>(sin_t + static_cast<long long>(pos) * HD) + ch);
        rot[q] = rope_chunk(qf[q], qf[(q + NQ / 2) % NQ], cs, sn, ch >= NCH / 2);
      }
#pragma unroll
      for (int q = 0; q < NQ; ++q) qf[q] = rot[q];
    }
  }
  __syncthreads();                                   // the appended K / V row is visible to the whole CTA

  // ---- pass 1: scores ----
  float m0 = -INFINITY, m1 = -INFINITY;              // running maxima of heads 2c, 2c + 1 over this lane's keys
  for (int blk0 = warp; blk0 < n_blocks; blk0 += U1 * DM_WARPS) {
    uint4 ka[U1][NQ], kb[U1][NQ];
#pragma unroll
    for (int u = 0; u < U1; ++u) {
      const int blk = blk0 + u * DM_WARPS;
      const int pa = blk * 16 + r, pb = pa + 8;
      const bool last = blk == n_blocks - 1;         // the block holding the row this CTA just wrote
#pragma unroll
      for (int q = 0; q < NQ; ++q) {
        ka[u][q] = make_uint4(0u, 0u, 0u, 0u);
        kb[u][q] = make_uint4(0u, 0u, 0u, 0u);
        const bf16* pka = krow + (pa < n_keys ? key_off(pa) : 0) + static_cast<long long>(pa) * HD + (c + 4 * q) * 8;
        const bf16* pkb = krow + (pb < n_keys ? key_off(pb) : 0) + static_cast<long long>(pb) * HD + (c + 4 * q) * 8;
        if (blk < n_blocks) {
          if (last) {
            if (pa < n_keys) ka[u][q] = ld_plain_u4(pka);
            if (pb < n_keys) kb[u][q] = ld_plain_u4(pkb);
          } else {
            ka[u][q] = ld_nc_u4(pka);
            kb[u][q] = ld_nc_u4(pkb);
          }
        }
      }
      // BEAM: the generated positions live in other beams' rows and were not covered by the prefetch above
      if (BEAM && blk < n_blocks && c < 2 * NH) {
        const int pv = blk * 16 + r + 8 * (c / NH);
        if (pv < n_keys)
          asm volatile("prefetch.global.L2 [%0];" ::"l"(vrow + key_off(pv) + static_cast<long long>(pv) * HD + (c % NH) * 64));
      }
    }
#pragma unroll
    for (int u = 0; u < U1; ++u) {
      const int blk = blk0 + u * DM_WARPS;
      if (blk < n_blocks) {
        float s[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
          mma_16816(s, ka[u][q].x, kb[u][q].x, ka[u][q].y, kb[u][q].y, qf[q].x, qf[q].y);
          mma_16816(s, ka[u][q].z, kb[u][q].z, ka[u][q].w, kb[u][q].w, qf[q].z, qf[q].w);
        }
        const int pa = blk * 16 + r, pb = pa + 8;
        if (2 * c < G) {
          float* s0 = sc + (2 * c) * sc_stride;
          float* s1 = s0 + sc_stride;
          const bool h1 = 2 * c + 1 < G;              // odd G: the group's last column pair has one head only
          if (pa < n_keys) {
            const float a0 = s[0] * scale_log2, a1 = s[1] * scale_log2;
            s0[pa] = a0; m0 = fmaxf(m0, a0);
            if (h1) { s1[pa] = a1; m1 = fmaxf(m1, a1); }
          }
          if (pb < n_keys) {
            const float a2 = s[2] * scale_log2, a3 = s[3] * scale_log2;
            s0[pb] = a2; m0 = fmaxf(m0, a2);
            if (h1) { s1[pb] = a3; m1 = fmaxf(m1, a3); }
          }
        }
      }
    }
  }
  // maxima over the 8 key rows of the lane group (lanes with equal c), then per warp into shared memory
#pragma unroll
  for (int o = 4; o < 32; o <<= 1) {
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, o));
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, o));
  }
  if (r == 0) {
    wmax[warp * 8 + 2 * c] = m0;
    wmax[warp * 8 + 2 * c + 1] = m1;
  }
  __syncthreads();

  // ---- pass 2: O^T += V^T . P^T with P = exp2(s - max) formed on the fly ----
  float mx = -INFINITY;
#pragma unroll
  for (int w = 0; w < DM_WARPS; ++w) mx = fmaxf(mx, wmax[w * 8 + r]);   // head r (unused for r >= G)
  float acc[NH][4][4];
#pragma unroll
  for (int h = 0; h < NH; ++h)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[h][j][i] = 0.f;
  float psum = 0.f;
  const float* srow = sc + (r < G ? r : 0) * sc_stride;
  for (int blk0 = warp; blk0 < n_blocks; blk0 += U2 * DM_WARPS) {
    uint4 vv[U2][NH][4];
#pragma unroll
    for (int u = 0; u < U2; ++u) {
      const int blk = blk0 + u * DM_WARPS;
      const bool last = blk == n_blocks - 1;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int p = blk * 16 + 4 * c + i;
#pragma unroll
        for (int h = 0; h < NH; ++h) {
          vv[u][h][i] = make_uint4(0u, 0u, 0u, 0u);
          if (blk < n_blocks && p < n_keys) {
            const bf16* pv = vrow + key_off(p) + static_cast<long long>(p) * HD + (r + 8 * h) * 8;
            vv[u][h][i] = last ? ld_plain_u4(pv) : ld_nc_u4(pv);
          }
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U2; ++u) {
      const int blk = blk0 + u * DM_WARPS;
      if (blk < n_blocks) {
        const int p = blk * 16 + 4 * c;
        float e[4] = {0.f, 0.f, 0.f, 0.f};
        if (r < G) {
          const float4 s4 = *reinterpret_cast<const float4*>(srow + p);
          if (p + 0 < n_keys) e[0] = ex2_approx(s4.x - mx);
          if (p + 1 < n_keys) e[1] = ex2_approx(s4.y - mx);
          if (p + 2 < n_keys) e[2] = ex2_approx(s4.z - mx);
          if (p + 3 < n_keys) e[3] = ex2_approx(s4.w - mx);
          psum += (e[0] + e[1]) + (e[2] + e[3]);
        }
        const uint32_t b0 = f2_to_bf2(e[0], e[1]), b1 = f2_to_bf2(e[2], e[3]);
#pragma unroll
        for (int h = 0; h < NH; ++h) {
          const uint4 &v0 = vv[u][h][0], &v1 = vv[u][h][1], &v2 = vv[u][h][2], &v3 = vv[u][h][3];
          mma_16816(acc[h][0], __byte_perm(v0.x, v1.x, 0x5410), __byte_perm(v0.x, v1.x, 0x7632),
                    __byte_perm(v2.x, v3.x, 0x5410), __byte_perm(v2.x, v3.x, 0x7632), b0, b1);
          mma_16816(acc[h][1], __byte_perm(v0.y, v1.y, 0x5410), __byte_perm(v0.y, v1.y, 0x7632),
                    __byte_perm(v2.y, v3.y, 0x5410), __byte_perm(v2.y, v3.y, 0x7632), b0, b1);
          mma_16816(acc[h][2], __byte_perm(v0.z, v1.z, 0x5410), __byte_perm(v0.z, v1.z, 0x7632),
                    __byte_perm(v2.z, v3.z, 0x5410), __byte_perm(v2.z, v3.z, 0x7632), b0, b1);
          mma_16816(acc[h][3], __byte_perm(v0.w, v1.w, 0x5410), __byte_perm(v0.w, v1.w, 0x7632),
                    __byte_perm(v2.w, v3.w, 0x5410), __byte_perm(v2.w, v3.w, 0x7632), b0, b1);
        }
      }
    }
  }
  // row sums: over the 4 key groups of a head (lanes r, c = 0..3), then per warp
  psum += __shfl_xor_sync(0xffffffffu, psum, 1);
  psum += __shfl_xor_sync(0xffffffffu, psum, 2);
  if (c == 0) wsum[warp * 8 + r] = psum;
  // accumulators: acc[h][j] = {O[d][head 2c], O[d][head 2c+1], O[d+1][head 2c], O[d+1][head 2c+1]}, d = 8 (r + 8h) + 2j
  if (2 * c < G) {
#pragma unroll
    for (int h = 0; h < NH; ++h)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int d = 8 * (r + 8 * h) + 2 * j;
        float* r0 = red + (warp * G + 2 * c) * HD + d;
        *reinterpret_cast<float2*>(r0) = make_float2(acc[h][j][0], acc[h][j][2]);
        if (2 * c + 1 < G) *reinterpret_cast<float2*>(r0 + HD) = make_float2(acc[h][j][1], acc[h][j][3]);
      }
  }
  __syncthreads();
  for (int e = threadIdx.x; e < G * HD; e += DM_THREADS) {
    const int g = e / HD;
    float a = 0.f, sum = 0.f;
#pragma unroll
    for (int w = 0; w < DM_WARPS; ++w) {
      a += red[w * G * HD + e];
      sum += wsum[w * 8 + g];
    }
    out[static_cast<long long>(b) * out_ld + (kvh * G + g) * HD + (e - g * HD)] = __float2bfloat16_rn(a / sum);
  }
}

static bool da_use_ffma() {
  static const bool v = [] { const char* e = getenv("OMNI_DA_FFMA"); return e && e[0] == '1'; }();
  return v;
}

struct BeamView {
  const int* ind = nullptr;
  int ind_ld = 0;
  long long ind_plane = 0;
  const long long* prefill_len = nullptr;
  int K = 1;
};

template <int HD, int G>
static int launch_decode_attn(const bf16* qkv, long long ld, bf16* kc, bf16* vc, const long long* len_idx, bf16* out,
                              long long out_ld, int B, int n_kv_heads, int max_len, float scale, cudaStream_t st,
                              const bf16* cos_t, const bf16* sin_t, const BeamView& bv) {
  if (da_use_ffma() && !bv.ind) {      // measurement switch: the FFMA formulation (see the comment above the MMA kernel)
    auto kfn = decode_attn_kernel<HD, G>;
    const int smem = (G * max_len + DA_WARPS * G * HD + G) * 4;
    if (smem > 200 * 1024) return OMNI_ERR_UNSUPPORTED;
    if (smem > 48 * 1024 &&
        cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess)
      return OMNI_ERR_CUDA;
    if (omni_launch_pdl(kfn, dim3(B * n_kv_heads), dim3(DA_THREADS), smem, st, qkv, ld, kc, vc, len_idx, out, out_ld,
                        n_kv_heads, max_len, scale * 1.4426950408889634f, cos_t, sin_t) != cudaSuccess)
      return OMNI_ERR_CUDA;
    return OMNI_OK;
  }
  const int sc_stride = (max_len + 15) / 16 * 16 + 8;     // + 8: heads 2c / 2c + 1 of a store land in different banks
  const int smem = (G * sc_stride + 16 + DM_WARPS * G * HD + 2 * DM_WARPS * 8) * 4;
  if (smem > 200 * 1024) return OMNI_ERR_UNSUPPORTED;
  auto go = [&](auto kfn) -> int {
    if (smem > 48 * 1024 &&
        cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess)
      return OMNI_ERR_CUDA;
    if (omni_launch_pdl(kfn, dim3(B * n_kv_heads), dim3(DM_THREADS), smem, st, qkv, ld, kc, vc, len_idx, out, out_ld,
                        n_kv_heads, max_len, sc_stride, scale * 1.4426950408889634f, cos_t, sin_t, bv.ind, bv.ind_ld,
                        bv.ind_plane, bv.prefill_len, bv.K) != cudaSuccess)
      return OMNI_ERR_CUDA;
    return OMNI_OK;
  };
  return bv.ind ? go(decode_attn_mma_kernel<HD, G, true>) : go(decode_attn_mma_kernel<HD, G, false>);
}

}  // namespace omni

extern "C" int omni_decode_attention(const void* qkv, int64_t ld, void* k_cache, void* v_cache, const int64_t* len_idx,
                                     void* out, int64_t out_ld, int32_t B, int32_t n_heads, int32_t n_kv_heads,
                                     int32_t head_dim, int32_t max_len, float scale, void* stream) {
  return omni_decode_attention_rope(qkv, ld, k_cache, v_cache, len_idx, out, out_ld, B, n_heads, n_kv_heads, head_dim, max_len,
                                    scale, nullptr, nullptr, 0, stream);
}

static int decode_attention_impl(const void* qkv, int64_t ld, void* k_cache, void* v_cache, const int64_t* len_idx,
                                 void* out, int64_t out_ld, int32_t B, int32_t n_heads, int32_t n_kv_heads,
                                 int32_t head_dim, int32_t max_len, float scale, const void* cos_t, const void* sin_t,
                                 int32_t table_rows, const omni::BeamView& bv, void* stream) {
  using namespace omni;
  OMNI_CHECK_ARG((cos_t == nullptr) == (sin_t == nullptr));
  if (cos_t) OMNI_CHECK_ARG(table_rows >= max_len && (reinterpret_cast<uintptr_t>(cos_t) & 15) == 0 &&
                            (reinterpret_cast<uintptr_t>(sin_t) & 15) == 0);
  const bf16* ct = reinterpret_cast<const bf16*>(cos_t);
  const bf16* stb = reinterpret_cast<const bf16*>(sin_t);
  OMNI_CHECK_ARG(qkv && k_cache && v_cache && len_idx && out && B > 0 && n_heads > 0 && n_kv_heads > 0 && max_len > 0);
  OMNI_CHECK_ARG(n_heads % n_kv_heads == 0 && (ld % 8) == 0 && (out_ld % 8) == 0);
  OMNI_CHECK_ARG(ld >= static_cast<int64_t>(n_heads + 2 * n_kv_heads) * head_dim);
  const int G = n_heads / n_kv_heads;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const bf16* q = reinterpret_cast<const bf16*>(qkv);
  bf16* kc = reinterpret_cast<bf16*>(k_cache);
  bf16* vc = reinterpret_cast<bf16*>(v_cache);
  const long long* li = reinterpret_cast<const long long*>(len_idx);
  bf16* o = reinterpret_cast<bf16*>(out);
#define OMNI_DA(HD_, G_) \
  if (head_dim == HD_ && G == G_) \
    return launch_decode_attn<HD_, G_>(q, ld, kc, vc, li, o, out_ld, B, n_kv_heads, max_len, scale, st, ct, stb, bv);
  // every GQA group size of the reference's model table: 4 (Llama-3.2-1B / 3.1-8B), 3 (Llama-3.2-3B), 7 (Qwen2.5-0.5B / 7B),
  // 6 (1.5B), 8 (3B), 5 (14B / 32B), + 1 / 2 for MHA-like test geometries
  OMNI_DA(64, 1) OMNI_DA(64, 2) OMNI_DA(64, 3) OMNI_DA(64, 4) OMNI_DA(64, 5) OMNI_DA(64, 6) OMNI_DA(64, 7) OMNI_DA(64, 8)
  OMNI_DA(128, 1) OMNI_DA(128, 2) OMNI_DA(128, 3) OMNI_DA(128, 4) OMNI_DA(128, 5) OMNI_DA(128, 6) OMNI_DA(128, 7)
  OMNI_DA(128, 8)
#undef OMNI_DA
  return OMNI_ERR_UNSUPPORTED;
}

extern "C" int omni_decode_attention_rope(const void* qkv, int64_t ld, void* k_cache, void* v_cache, const int64_t* len_idx,
                                          void* out, int64_t out_ld, int32_t B, int32_t n_heads, int32_t n_kv_heads,
                                          int32_t head_dim, int32_t max_len, float scale, const void* cos_t,
                                          const void* sin_t, int32_t table_rows, void* stream) {
  return decode_attention_impl(qkv, ld, k_cache, v_cache, len_idx, out, out_ld, B, n_heads, n_kv_heads, head_dim, max_len,
                               scale, cos_t, sin_t, table_rows, omni::BeamView(), stream);
}

extern "C" int omni_decode_attention_beam(const void* qkv, int64_t ld, void* k_cache, void* v_cache, const int64_t* len_idx,
                                          void* out, int64_t out_ld, int32_t B, int32_t n_heads, int32_t n_kv_heads,
                                          int32_t head_dim, int32_t max_len, float scale, const void* cos_t,
                                          const void* sin_t, int32_t table_rows, const int32_t* beam_ind, int32_t ind_ld,
                                          const int64_t* prefill_len, int32_t K, void* stream) {
  OMNI_CHECK_ARG(beam_ind && prefill_len && K > 0 && B % K == 0 && ind_ld > 0 && ind_ld <= omni::DM_MAX_NEW);
  omni::BeamView bv;
  bv.ind = beam_ind;
  bv.ind_ld = ind_ld;
  bv.ind_plane = static_cast<long long>(B) * ind_ld;
  bv.prefill_len = reinterpret_cast<const long long*>(prefill_len);
  bv.K = K;
  return decode_attention_impl(qkv, ld, k_cache, v_cache, len_idx, out, out_ld, B, n_heads, n_kv_heads, head_dim, max_len,
                               scale, cos_t, sin_t, table_rows, bv, stream);
}
