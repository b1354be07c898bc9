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
  // grid (head, clip, tile): blocks are dispatched x-fastest, so all CTAs of the heaviest tile (the last one under the causal
  // mask: it sees every key) start first and the light ones fill the tail of the launch
  const int head = blockIdx.x, clip = blockIdx.y;
  const int qt = p.causal ? static_cast<int>(gridDim.z - 1 - blockIdx.z) : static_cast<int>(blockIdx.z);
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

// ---------------------------------------------------------------------------------------------------------------
// Software-pipelined variant (used for head_dim 128, where only one CTA fits per SM anyway): 8 softmax warps (two per
// TMEM lane quarter, 64 keys each, row maximum exchanged through the idle P tile), two S buffers and two O tiles in TMEM
// (512 columns), two P tiles in shared memory: S(j+1) is computed while the softmax warps convert S(j), and the O update of
// tile j-1 is applied after the conversion of tile j.  Measured (profiles/attention_microbench_r1.jsonl): 1.2x over the
// unpipelined kernel at head_dim 128; at head_dim 64 two unpipelined CTAs per SM are faster (the softmax warps of ONE CTA
// run in lockstep and contend for the MUFU pipe, two CTAs de-phase naturally), so head_dim 64 keeps the kernel above.
// ---------------------------------------------------------------------------------------------------------------
constexpr int AT_PIPE_THREADS = 320;   // warp 0 TMA, warp 1 MMA, warps 2..9 softmax

template <int AT_HD>           // head dim: 64 or 128 (one or two 128-byte swizzle blocks per row)
struct AttnPipeSmem {
  static constexpr int Q_BYTES = AT_BQ * AT_HD * 2;        // 16 / 32 KB
  static constexpr int K_BYTES = AT_BK * AT_HD * 2;
  static constexpr int V_BYTES = AT_BK * AT_HD * 2;
  static constexpr int P_BYTES = AT_BQ * AT_BK * 2;        // 32 KB (two [128 x 64] swizzled blocks), double-buffered
  static constexpr int STAGE_BYTES = K_BYTES + V_BYTES;
  static constexpr int NS = AT_HD == 64 ? 3 : 2;           // K/V stages (TMA runs NS - 1 tiles ahead of the MMAs)
  static constexpr int OFF_Q = 0, OFF_KV = Q_BYTES;
  static constexpr int OFF_P = OFF_KV + NS * STAGE_BYTES;
  static constexpr int BAR_OFFSET = OFF_P + 2 * P_BYTES;
  static constexpr int TOTAL = BAR_OFFSET + 16 * 8 + 16 + 1024;
};

template <int AT_HD>
__global__ void __launch_bounds__(AT_PIPE_THREADS, 1)
attn_fwd_pipe_kernel(const __grid_constant__ CUtensorMap tm, const AttnParams p) {
  using SM = AttnPipeSmem<AT_HD>;
  constexpr int NB = AT_HD / 64;     // 64-column TMA boxes per Q / K / V tile
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SM::BAR_OFFSET);
  constexpr int NS = SM::NS;
  uint64_t* q_full = bars + 0;
  uint64_t* kv_full = bars + 1;            // [NS]
  uint64_t* kv_empty = bars + 1 + NS;      // [NS]
  uint64_t* s_full = bars + 1 + 2 * NS;    // [2]  S(j) landed in TMEM buffer j & 1
  uint64_t* p_full = s_full + 2;           // [2]  softmax consumed S(j) and wrote P(j) into P buffer j & 1
  uint64_t* o_full = p_full + 2;           // [2]  P(j).V landed in O buffer j & 1
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 16);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // grid (head, clip, tile): blocks are dispatched x-fastest, so all CTAs of the heaviest tile (the last one under the causal
  // mask: it sees every key) start first and the light ones fill the tail of the launch
  const int head = blockIdx.x, clip = blockIdx.y;
  const int qt = p.causal ? static_cast<int>(gridDim.z - 1 - blockIdx.z) : static_cast<int>(blockIdx.z);
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
      for (int i = 0; i < NS; ++i) {
        mbar_init(kv_full + i, 1);
        mbar_init(kv_empty + i, 1);
      }
      for (int i = 0; i < 2; ++i) {
        mbar_init(s_full + i, 1);
        mbar_init(p_full + i, 8);      // one arrive per softmax warp
        mbar_init(o_full + i, 1);
      }
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(tmem_ptr_smem, 512);    // S buffers: columns [0,128) [128,256); O tiles: [256,256+HD) [256+HD,256+2HD)
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const uint32_t tmem_s = tmem_base;           // + 128 * (j & 1)
  const uint32_t tmem_o = tmem_base + 256;     // + AT_HD * (j & 1)

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(q_full, SM::Q_BYTES);
#pragma unroll
      for (int b = 0; b < NB; ++b) tma_load_2d(&tm, q_full, smem + SM::OFF_Q + b * 16384, col_q + b * 64, clip_row0 + q0);
      for (int j = 0; j < n_kv; ++j) {
        const int st = j % NS;
        uint8_t* kv = smem + SM::OFF_KV + st * SM::STAGE_BYTES;
        mbar_wait(kv_empty + st, ((j / NS) & 1) ^ 1);
        mbar_expect_tx(kv_full + st, SM::STAGE_BYTES);
#pragma unroll
        for (int b = 0; b < NB; ++b) {
          tma_load_2d(&tm, kv_full + st, kv + b * 16384, col_k + b * 64, clip_row0 + j * AT_BK);
          tma_load_2d(&tm, kv_full + st, kv + SM::K_BYTES + b * 16384, col_v + b * 64, clip_row0 + j * AT_BK);
        }
      }
    }
  } else if (warp == 1) {
    // Software pipeline: S(j+1) = Q.K(j+1)^T is issued BEFORE the wait on P(j), so the tensor core computes the next
    // score tile while the softmax warps convert the current one (two S buffers, two P buffers, two O tiles).
    constexpr uint32_t idesc_s = make_idesc_bf16(AT_BQ, AT_BK, 0, 0);   // Q (K-major) x K (K-major)
    constexpr uint32_t idesc_o = make_idesc_bf16(AT_BQ, AT_HD, 0, 1);   // P (K-major) x V (MN-major: [keys, hd] tile)
    const uint32_t sQ = smem_u32(smem + SM::OFF_Q);
    auto issue_s = [&](int j) {
      const int st = j % NS;
      const uint32_t sK = smem_u32(smem + SM::OFF_KV + st * SM::STAGE_BYTES);
      mbar_wait(kv_full + st, (j / NS) & 1);
      tc_fence_after();
      if (lane == 0) {
#pragma unroll
        for (int k = 0; k < AT_HD / 16; ++k) {
          const uint64_t qd = make_smem_desc_sw128(sQ + (k >> 2) * 16384, 16, 1024) + 2 * (k & 3);
          const uint64_t kd = make_smem_desc_sw128(sK + (k >> 2) * 16384, 16, 1024) + 2 * (k & 3);
          umma_bf16(tmem_s + 128 * (j & 1), qd, kd, idesc_s, k > 0 ? 1u : 0u);
        }
        umma_commit(s_full + (j & 1));
      }
      __syncwarp();
    };
    mbar_wait(q_full, 0);
    issue_s(0);
    for (int j = 0; j < n_kv; ++j) {
      const int b = j & 1;
      // S buffer (j+1)&1 is free: P(j-1) arrived one iteration ago, i.e. the softmax warps are done with S(j-1)
      if (j + 1 < n_kv) issue_s(j + 1);
      mbar_wait(p_full + b, (j >> 1) & 1);      // softmax consumed S(j) and wrote P(j); O tile b was read two tiles ago
      tc_fence_after();
      if (lane == 0) {
        const int st = j % NS;
        const uint32_t sV = smem_u32(smem + SM::OFF_KV + st * SM::STAGE_BYTES) + SM::K_BYTES;
        const uint32_t sP = smem_u32(smem + SM::OFF_P + b * SM::P_BYTES);
        // MN-major V: 8 key rows per 1 KB group (SBO), next 64 head-dim columns one 16 KB box further (LBO)
        const uint64_t vd = make_smem_desc_sw128(sV, 16384, 1024);
#pragma unroll
        for (int k = 0; k < AT_BK / 16; ++k) {
          const uint64_t pd = make_smem_desc_sw128(sP + (k >> 2) * 16384, 16, 1024) + 2 * (k & 3);
          umma_bf16(tmem_o + AT_HD * b, pd, vd + 128 * k, idesc_o, k > 0 ? 1u : 0u);
        }
        umma_commit(o_full + b);
        umma_commit(kv_empty + st);
      }
      __syncwarp();
    }
  } else {
    // ===== softmax / accumulation: thread = (query row, half of the tile's keys) =====
    // Warps w and w + 4 read the same TMEM lane quarter (hardware rule: lanes 32*(w%4)..); each converts 64 of the 128
    // keys of its row and accumulates half of the head_dim columns of O.  The only per-tile exchange is the row maximum:
    // it goes through the first bytes of the thread's own (not yet written) row of the P tile, under two 64-thread
    // named barriers; the row sums stay partial until the end.
    const int q = warp & 3;
    const int hf = (warp - 2) >> 2;              // key half / O-column half
    const int r = q * 32 + lane;                 // row inside the tile == TMEM lane
    const int qpos = q0 + r;                     // position inside the clip
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    const int bar_id = 1 + q;
    constexpr int OC = AT_HD / 2;                // O columns owned by this thread
    float m = -INFINITY, l = 0.f, alpha_prev = 0.f;
    float o[OC];
#pragma unroll
    for (int i = 0; i < OC; ++i) o[i] = 0.f;
    // O = O * alpha + P(j).V for the tile whose MMA was issued one iteration ago (deferred by one tile, so the wait on
    // the P.V result never sits between two softmax passes)
    auto accumulate_o = [&](int j, float alpha) {
      mbar_wait(o_full + (j & 1), (j >> 1) & 1);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < OC / 32; ++c) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_o + AT_HD * (j & 1) + lane_addr + hf * OC + c * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) o[c * 32 + i] = o[c * 32 + i] * alpha + __uint_as_float(v[i]);
      }
      tc_fence_before();
    };

    for (int j = 0; j < n_kv; ++j) {
      const int b = j & 1;
      const uint32_t tmem_sj = tmem_s + 128 * b;
      uint8_t* sP = smem + SM::OFF_P + b * SM::P_BYTES;      // idle since P(j-2).V completed (waited last iteration)
      float* xch_mine = reinterpret_cast<float*>(sP + hf * 16384 + r * 128);
      const float* xch_peer = reinterpret_cast<const float*>(sP + (hf ^ 1) * 16384 + r * 128);
      const int k0 = j * AT_BK + hf * 64;        // first key of this thread's half
      mbar_wait(s_full + b, (j >> 1) & 1);
      tc_fence_after();
      // pass 1: maximum of the (masked) scores of this half.  Interior halves take the compare-free path.
      float mx = -INFINITY;
      const int kmax = p.causal ? min(qpos, p.S - 1) : (p.S - 1);      // last visible key position
      const bool unmasked = (k0 + 63 <= kmax);
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_sj + lane_addr + hf * 64 + c * 32, v);
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
      *xch_mine = mx;
      asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");
      mx = fmaxf(mx, *xch_peer);
      asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");      // peer has read before P overwrites the slot
      const float m_new = fmaxf(m, mx * p.scale_log2);
      const float m_use = (m_new == -INFINITY) ? 0.f : m_new;          // fully masked row so far
      const float alpha = ex2_approx(m - m_use);                       // m = -inf -> 0
      // pass 2: P = exp2(s * scale - m), partial row sum, bf16 P into this half's swizzled [128 x 64] operand block
      float sum = 0.f;
      const float sc = p.scale_log2;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_sj + lane_addr + hf * 64 + c * 32, v);
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
        // 32 keys = 4 chunks of 16 bytes at chunk index c * 4 + t of the 128-byte row
        uint8_t* blk = sP + hf * 16384 + r * 128;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const int chunk = c * 4 + t;
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
      if (lane == 0) mbar_arrive(p_full + b);
      if (j > 0) accumulate_o(j - 1, alpha_prev);
      alpha_prev = alpha;
    }
    accumulate_o(n_kv - 1, alpha_prev);
    // total row sum = own half + peer's half (both P tiles are idle after the last o_full)
    {
      uint8_t* sP = smem + SM::OFF_P;
      *reinterpret_cast<float*>(sP + hf * 16384 + r * 128) = l;
      asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");
      l += *reinterpret_cast<const float*>(sP + (hf ^ 1) * 16384 + r * 128);
    }
    if (qpos < p.S) {
      const float inv = l > 0.f ? 1.0f / l : 0.f;
      const long long row = static_cast<long long>(clip_row0) + qpos;
      bf16* op = p.out + row * p.out_ld + head * AT_HD + hf * OC;
#pragma unroll
      for (int i = 0; i < OC; i += 8) {
        uint4 u;
        u.x = f2_to_bf2(o[i] * inv, o[i + 1] * inv);
        u.y = f2_to_bf2(o[i + 2] * inv, o[i + 3] * inv);
        u.z = f2_to_bf2(o[i + 4] * inv, o[i + 5] * inv);
        u.w = f2_to_bf2(o[i + 6] * inv, o[i + 7] * inv);
        *reinterpret_cast<uint4*>(op + i) = u;
      }
      if (p.lse && hf == 0) p.lse[static_cast<long long>(head) * p.M + row] = (m + log2f(l)) * 0.6931471805599453f;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
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
  OMNI_CHECK_ARG(B <= 65535);   // clip index rides in gridDim.y
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
  dim3 grid(n_heads, B, ceil_div(S, AT_BQ));
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (head_dim == 64) return launch_attn<64>(tm, p, grid, st);
  auto kfn = attn_fwd_pipe_kernel<128>;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, AttnPipeSmem<128>::TOTAL) != cudaSuccess)
      return OMNI_ERR_CUDA;
    attr_set = true;
  }
  kfn<<<grid, AT_PIPE_THREADS, AttnPipeSmem<128>::TOTAL, st>>>(tm, p);
  OMNI_LAUNCH_CHECK();
  return OMNI_OK;
}
