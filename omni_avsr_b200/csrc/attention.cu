// Flash-attention forward on the sm_100a tensor cores (head_dim 64 or 128): S = Q.K^T and O += P.V are tcgen05.mma with the
// accumulators in TMEM; the online softmax runs on 128 threads (one per query row) between the two MMAs.
//
// Replaces F.scaled_dot_product_attention of: Llama_LoRA.py:300 / Qwen_LoRA.py:606 (causal, GQA, no mask on this path),
// HF WhisperEncoderLayer self-attention (non-causal, 1500 keys incl. the zero padding, no mask) and fairseq
// multihead_attention.py:619-654 (non-causal; q pre-scaling by head_dim^-0.5 == the softmax scale used here).
//
// Layout: packed q|k|v rows [M, ld] bf16 (q heads, then k heads, then v heads in every row); a launch covers one
// segment of B clips x S tokens starting at row `row0`.  grid = (ceil(S/128), n_heads, B); 192 threads:
//   warp 0: TMA producer (Q once; K and V tile per step)      warp 1: MMA issuer (+ TMEM alloc)
//   warps 2..5: softmax / accumulation, thread = query row.
// Per KV tile of 128 keys:  S(TMEM) = Q K^T  ->  rows: m, l, P = exp2(s - m) (bf16, written to smem in the K-major
// 128B-swizzled operand layout)  ->  O_tile(TMEM) = P V (V consumed as an MN-major operand straight from its
// [keys, head_dim] TMA tile)  ->  rows: O = O * alpha + O_tile.   With head_dim 64 two CTAs fit per SM (80 KB smem, 256
// TMEM columns), so one CTA's softmax overlaps the other's MMAs.
#include "common.cuh"
#include "../../include/omni_avsr.h"

namespace omni {

constexpr int AT_BQ = 128;     // query rows per CTA
constexpr int AT_BK = 128;     // keys per step
constexpr int AT_THREADS = 192;

template <int AT_HD>           // head dim: 64 or 128 (one or two 128-byte swizzle blocks per row)
struct AttnSmem {
  static constexpr int Q_BYTES = AT_BQ * AT_HD * 2;        // 16 / 32 KB
  static constexpr int K_BYTES = AT_BK * AT_HD * 2;
  static constexpr int V_BYTES = AT_BK * AT_HD * 2;
  static constexpr int P_BYTES = AT_BQ * AT_BK * 2;        // 32 KB (two [128 x 64] swizzled blocks)
  static constexpr int STAGE_BYTES = K_BYTES + V_BYTES;    // K/V tiles are double-buffered: the TMA load of tile j+1
  static constexpr int OFF_Q = 0, OFF_KV = Q_BYTES;        // runs under the softmax / MMAs of tile j
  static constexpr int OFF_P = OFF_KV + 2 * STAGE_BYTES;
  static constexpr int BAR_OFFSET = OFF_P + P_BYTES;
  // 2 CTAs / SM need <= 113 KB each: the 1024-byte alignment slack is trimmed to 768 (the kernel traps if the dynamic
  // shared-memory base is less aligned than that allows; in practice it is 1024-aligned)
  static constexpr int ALIGN_SLACK = AT_HD == 64 ? 768 : 1024;
  static constexpr int TOTAL = BAR_OFFSET + 8 * 8 + 16 + ALIGN_SLACK;
};

struct AttnParams {
  bf16* out;             // [M, out_ld]
  float* lse;            // optional [n_heads, M] (natural-log-sum-exp of the scaled scores) or nullptr
  long long out_ld;
  long long M;           // rows of the packed buffer (lse stride)
  int row0, S, n_heads, n_kv_heads, causal;
  float scale_log2;      // softmax scale * log2(e)
};

template <int AT_HD>
__global__ void __launch_bounds__(AT_THREADS, AT_HD == 64 ? 2 : 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tm, const AttnParams p) {
  using SM = AttnSmem<AT_HD>;
  constexpr int NB = AT_HD / 64;     // 64-column TMA boxes per Q / K / V tile
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SM::BAR_OFFSET);
  uint64_t* q_full = bars + 0;
  uint64_t* kv_full = bars + 1;      // [2]
  uint64_t* kv_empty = bars + 3;     // [2]
  uint64_t* s_full = bars + 5;
  uint64_t* p_full = bars + 6;
  uint64_t* o_full = bars + 7;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 8);
  if (static_cast<int>(smem - smem_raw) > SM::ALIGN_SLACK) __trap();

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int qt = blockIdx.x, head = blockIdx.y, clip = blockIdx.z;
  const int kvh = head / (p.n_heads / p.n_kv_heads);
  const int q0 = qt * AT_BQ;
  const int clip_row0 = p.row0 + clip * p.S;
  const int n_kv = p.causal ? (qt + 1) : (p.S + AT_BK - 1) / AT_BK;
  const int col_q = head * AT_HD;
  const int col_k = (p.n_heads + kvh) * AT_HD;
  const int col_v = (p.n_heads + p.n_kv_heads + kvh) * AT_HD;

  if (warp == 0 && lane == 0) tma_prefetch_desc(&tm);
  if (warp == 1) {
    if (lane == 0) {
      mbar_init(q_full, 1);
      mbar_init(kv_full, 1);
      mbar_init(kv_full + 1, 1);
      mbar_init(kv_empty, 1);
      mbar_init(kv_empty + 1, 1);
      mbar_init(s_full, 1);
      mbar_init(p_full, 4);      // one arrive per softmax warp
      mbar_init(o_full, 1);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(tmem_ptr_smem, 256);    // S: columns [0,128), O tile: columns [128,192)
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const uint32_t tmem_s = tmem_base;
  const uint32_t tmem_o = tmem_base + 128;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(q_full, SM::Q_BYTES);
#pragma unroll
      for (int b = 0; b < NB; ++b) tma_load_2d(&tm, q_full, smem + SM::OFF_Q + b * 16384, col_q + b * 64, clip_row0 + q0);
      for (int j = 0; j < n_kv; ++j) {
        const int st = j & 1;
        uint8_t* kv = smem + SM::OFF_KV + st * SM::STAGE_BYTES;
        mbar_wait(kv_empty + st, ((j >> 1) & 1) ^ 1);
        mbar_expect_tx(kv_full + st, SM::STAGE_BYTES);
#pragma unroll
        for (int b = 0; b < NB; ++b) {
          tma_load_2d(&tm, kv_full + st, kv + b * 16384, col_k + b * 64, clip_row0 + j * AT_BK);
          tma_load_2d(&tm, kv_full + st, kv + SM::K_BYTES + b * 16384, col_v + b * 64, clip_row0 + j * AT_BK);
        }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc_s = make_idesc_bf16(AT_BQ, AT_BK, 0, 0);   // Q (K-major) x K (K-major)
    constexpr uint32_t idesc_o = make_idesc_bf16(AT_BQ, AT_HD, 0, 1);   // P (K-major) x V (MN-major: [keys, hd] tile)
    const uint32_t sQ = smem_u32(smem + SM::OFF_Q);
    const uint32_t sP = smem_u32(smem + SM::OFF_P);
    mbar_wait(q_full, 0);
    for (int j = 0; j < n_kv; ++j) {
      const uint32_t ph = j & 1;
      const int st = j & 1;
      const uint32_t sK = smem_u32(smem + SM::OFF_KV + st * SM::STAGE_BYTES);
      const uint32_t sV = sK + SM::K_BYTES;
      mbar_wait(kv_full + st, (j >> 1) & 1);
      tc_fence_after();
      if (lane == 0) {
#pragma unroll
        for (int k = 0; k < AT_HD / 16; ++k) {
          const uint64_t qd = make_smem_desc_sw128(sQ + (k >> 2) * 16384, 16, 1024) + 2 * (k & 3);
          const uint64_t kd = make_smem_desc_sw128(sK + (k >> 2) * 16384, 16, 1024) + 2 * (k & 3);
          umma_bf16(tmem_s, qd, kd, idesc_s, k > 0 ? 1u : 0u);
        }
        umma_commit(s_full);
      }
      __syncwarp();
      mbar_wait(p_full, ph);      // softmax consumed S and wrote P
      tc_fence_after();
      if (lane == 0) {
        // MN-major V: 8 key rows per 1 KB group (SBO), next 64 head-dim columns one 16 KB box further (LBO)
        const uint64_t vd = make_smem_desc_sw128(sV, 16384, 1024);
#pragma unroll
        for (int k = 0; k < AT_BK / 16; ++k) {
          const uint64_t pd = make_smem_desc_sw128(sP + (k >> 2) * 16384, 16, 1024) + 2 * (k & 3);
          umma_bf16(tmem_o, pd, vd + 128 * k, idesc_o, k > 0 ? 1u : 0u);
        }
        umma_commit(o_full);
        umma_commit(kv_empty + st);
      }
      __syncwarp();
    }
  } else {
    // ===== softmax / accumulation: thread = query row =====
    const int q = warp & 3;
    const int r = q * 32 + lane;                 // row inside the tile == TMEM lane
    const int qpos = q0 + r;                     // position inside the clip
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    uint8_t* sP = smem + SM::OFF_P;
    float m = -INFINITY, l = 0.f;
    float o[AT_HD];
#pragma unroll
    for (int i = 0; i < AT_HD; ++i) o[i] = 0.f;

    for (int j = 0; j < n_kv; ++j) {
      const uint32_t ph = j & 1;
      const int k0 = j * AT_BK;
      mbar_wait(s_full, ph);
      tc_fence_after();
      // pass 1: row maximum of the (masked) scores.  Interior tiles (every key visible) take the compare-free path.
      float mx = -INFINITY;
      const int kmax = p.causal ? min(qpos, p.S - 1) : (p.S - 1);      // last visible key position
      const bool unmasked = (k0 + AT_BK - 1 <= kmax);
#pragma unroll 1
      for (int c = 0; c < AT_BK / 32; ++c) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_s + lane_addr + c * 32, v);
        tmem_ld_wait();
        if (unmasked) {
#pragma unroll
          for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(v[i]));
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (k0 + c * 32 + i <= kmax) mx = fmaxf(mx, __uint_as_float(v[i]));
        }
      }
      const float m_new = fmaxf(m, mx * p.scale_log2);
      const float m_use = (m_new == -INFINITY) ? 0.f : m_new;          // fully masked row so far
      const float alpha = ex2_approx(m - m_use);                       // m = -inf -> 0
      // pass 2: P = exp2(s * scale - m), row sum, bf16 P into the swizzled operand tile
      float sum = 0.f;
      const float sc = p.scale_log2;
#pragma unroll 1
      for (int c = 0; c < AT_BK / 32; ++c) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_s + lane_addr + c * 32, v);
        tmem_ld_wait();
        uint32_t packed[16];
        if (unmasked) {
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const float p0 = ex2_approx(fmaf(__uint_as_float(v[i]), sc, -m_use));
            const float p1 = ex2_approx(fmaf(__uint_as_float(v[i + 1]), sc, -m_use));
            sum += p0 + p1;
            packed[i >> 1] = f2_to_bf2(p0, p1);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            float p0 = 0.f, p1 = 0.f;
            if (k0 + c * 32 + i <= kmax) p0 = ex2_approx(fmaf(__uint_as_float(v[i]), sc, -m_use));
            if (k0 + c * 32 + i + 1 <= kmax) p1 = ex2_approx(fmaf(__uint_as_float(v[i + 1]), sc, -m_use));
            sum += p0 + p1;
            packed[i >> 1] = f2_to_bf2(p0, p1);
          }
        }
        // 32 keys = 4 chunks of 16 bytes; key block (64 keys) = c >> 1, chunk index inside the 128-byte row = (c & 1) * 4 + t
        uint8_t* blk = sP + (c >> 1) * 16384 + r * 128;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const int chunk = (c & 1) * 4 + t;
          *reinterpret_cast<uint4*>(blk + ((chunk ^ (r & 7)) << 4)) =
              make_uint4(packed[4 * t], packed[4 * t + 1], packed[4 * t + 2], packed[4 * t + 3]);
        }
      }
      l = l * alpha + sum;
      m = m_new;
      // make the generic-proxy smem writes visible to the tensor core (async proxy), then release S / publish P
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
      // O = O * alpha + P V
      mbar_wait(o_full, ph);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < AT_HD / 32; ++c) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_o + lane_addr + c * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) o[c * 32 + i] = o[c * 32 + i] * alpha + __uint_as_float(v[i]);
      }
      tc_fence_before();
    }
    if (qpos < p.S) {
      const float inv = l > 0.f ? 1.0f / l : 0.f;
      const long long row = static_cast<long long>(clip_row0) + qpos;
      bf16* op = p.out + row * p.out_ld + head * AT_HD;
#pragma unroll
      for (int i = 0; i < AT_HD; i += 8) {
        uint4 u;
        u.x = f2_to_bf2(o[i] * inv, o[i + 1] * inv);
        u.y = f2_to_bf2(o[i + 2] * inv, o[i + 3] * inv);
        u.z = f2_to_bf2(o[i + 4] * inv, o[i + 5] * inv);
        u.w = f2_to_bf2(o[i + 6] * inv, o[i + 7] * inv);
        *reinterpret_cast<uint4*>(op + i) = u;
      }
      if (p.lse) p.lse[static_cast<long long>(head) * p.M + row] = (m + log2f(l)) * 0.6931471805599453f;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 256);
}

template <int HD>
static int launch_attn(const CUtensorMap& tm, const AttnParams& p, dim3 grid, cudaStream_t st) {
  auto kfn = attn_fwd_kernel<HD>;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, AttnSmem<HD>::TOTAL) != cudaSuccess)
      return OMNI_ERR_CUDA;
    attr_set = true;
  }
  kfn<<<grid, AT_THREADS, AttnSmem<HD>::TOTAL, st>>>(tm, p);
  OMNI_LAUNCH_CHECK();
  return OMNI_OK;
}

}  // namespace omni

extern "C" int omni_attention_fwd(const void* qkv, int64_t M, int64_t ld, void* out, int64_t out_ld, float* lse,
                                  int32_t row0, int32_t B, int32_t S, int32_t n_heads, int32_t n_kv_heads,
                                  int32_t head_dim, int32_t causal, float scale, void* stream) {
  using namespace omni;
  OMNI_CHECK_ARG(qkv && out && M > 0 && B > 0 && S > 0 && n_heads > 0 && n_kv_heads > 0);
  OMNI_CHECK_ARG(n_heads % n_kv_heads == 0 && row0 >= 0 && static_cast<int64_t>(row0) + static_cast<int64_t>(B) * S <= M);
  OMNI_CHECK_ARG((ld % 8) == 0 && (out_ld % 8) == 0 && ld >= static_cast<int64_t>(n_heads + 2 * n_kv_heads) * head_dim);
  if (head_dim != 64 && head_dim != 128) return OMNI_ERR_UNSUPPORTED;
  CUtensorMap tm;
  int rc = omni_make_tmap_2d_bf16(&tm, qkv, (uint64_t)M, (uint64_t)(n_heads + 2 * n_kv_heads) * head_dim, (uint64_t)ld,
                                  AT_BQ, 64, 1);
  if (rc) return rc;
  AttnParams p;
  p.out = reinterpret_cast<bf16*>(out);
  p.lse = lse;
  p.out_ld = out_ld;
  p.M = M;
  p.row0 = row0; p.S = S; p.n_heads = n_heads; p.n_kv_heads = n_kv_heads; p.causal = causal ? 1 : 0;
  p.scale_log2 = scale * 1.4426950408889634f;
  dim3 grid(ceil_div(S, AT_BQ), n_heads, B);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  return head_dim == 64 ? launch_attn<64>(tm, p, grid, st) : launch_attn<128>(tm, p, grid, st);
}
