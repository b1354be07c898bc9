// Flash-attention forward on the sm_100a tensor cores (head_dim 64 or 128): S = Q.K^T and O += P.V are tcgen05.mma with the
// accumulators in TMEM; the online softmax runs on CUDA cores (one thread per query row) between the two MMAs.
//
// Replaces F.scaled_dot_product_attention of: Llama_LoRA.py:300 / Qwen_LoRA.py:606 (causal, GQA, no mask on this path),
// HF WhisperEncoderLayer self-attention (non-causal, 1500 keys incl. the zero padding, no mask) and fairseq
// multihead_attention.py:619-654 (non-causal; q pre-scaling by head_dim^-0.5 == the softmax scale used here).
//
// Layout: packed q|k|v rows [M, ld] bf16 (q heads, then k heads, then v heads in every row); a launch covers one
// segment of B clips x S tokens starting at row `row0`.  grid = (n_heads, B, ceil(S/128)); warp 0: TMA producer (Q once;
// K and V tile per step), warp 1: MMA issuer (+ TMEM alloc), remaining warps: softmax.  P = exp2(s - m) is written as
// bf16 into shared memory in the K-major 128B-swizzled operand layout; V is consumed as an MN-major operand straight from
// its [keys, head_dim] TMA tile.  Two kernels:
//   attn_fwd_pipe_kernel<128>  head_dim 128 (Qwen2.5-3B, Llama-3.1-8B): 8 softmax warps, two S / P / O buffers
//   attn_fwd64_kernel          head_dim 64 (Whisper, AV-HuBERT, Llama-3.2-1B): two independent softmax streams per CTA,
//                              O resident in TMEM with lazy rescaling, 2 CTAs / SM
// Every tcgen05 / TMA instruction is issued from an `elect.sync` branch: with `if (lane == 0)` ptxas wraps each
// UTCHMMA / UTCBAR / UTMALDG in an ELECT + R2UR.BROADCAST + BRA.U.ANY loop (~100 cycles of issue latency per MMA, more
// than the 32..64 cycles these small MMAs execute for).
#include <stdlib.h>
#include "common.cuh"
#include "../../include/omni_avsr.h"

namespace omni {

constexpr int AT_BQ = 128;     // query rows per CTA
constexpr int AT_BK = 128;     // keys per step

struct AttnParams {
  bf16* out;             // [M, out_ld]
  float* lse;            // optional [n_heads, M] (natural-log-sum-exp of the scaled scores) or nullptr
  long long out_ld;
  long long M;           // rows of the packed buffer (lse stride)
  int row0, S, n_heads, n_kv_heads, causal;
  float scale_log2;      // softmax scale * log2(e)
};

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
    if (elect_one()) {
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
      if (elect_one()) {
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
      if (elect_one()) {
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

// ---------------------------------------------------------------------------------------------------------------
// head_dim 64 forward (the kernel dispatched for head_dim 64): two independent online-softmax streams per CTA over
// alternating 64-key steps, O resident in TMEM.  320 threads: warp 0 TMA, warp 1 MMA, warps 2..5 softmax group 0 (even
// steps), warps 6..9 softmax group 1 (odd steps); two CTAs per SM (256 TMEM columns, 113 KB shared memory each), i.e.
// four softmax warps per SM sub-partition to keep the MUFU (ex2) pipe fed -- that pipe bounds head_dim 64 attention
// (1 ex2 per score against 256 tensor-core flop).  Against a one-stream kernel that keeps O in registers (the first version of this file, 398 TFLOP/s on Whisper's shape):
//   * S is double-buffered in TMEM (128 x 64 fp32 per group) and P in shared memory (16 KB K-major tile per group): the
//     MMA warp runs S(t+1) = Q.K_half^T while group t&1 converts S(t), and P(t).V_half while the other group converts
//     S(t+1) -- a softmax warp never waits for a tensor-core round trip;
//   * the 64 scores of a row are read from TMEM ONCE (two 32-column loads in flight, one wait) and stay in registers for
//     the maximum and the exponentials;
//   * each group owns its running (m, l) and its own O accumulator in TMEM (split-KV inside the CTA; merged once in the
//     epilogue), so the groups never exchange row statistics;
//   * O is never read back per step: P.V accumulates in TMEM, and a group raises its reference maximum -- with a TMEM
//     read-modify-write of the row's 64 O columns -- only when it grew by more than 2^8 (lazy rescaling: l and O stay
//     consistent for ANY reference value m, so the result is exact; P <= 256 in between).  Rarely taken after the
//     first steps.
// MMA issue order: S(0) S(1) | PV(0) S(2) | PV(1) S(3) | ...  tcgen05 MMAs complete in issue order, so s_full(t) also means
// P.V(t-2) finished: group t&1's P tile is free and its O accumulator is stable.
// ---------------------------------------------------------------------------------------------------------------
constexpr int AT64_THREADS = 320;
// AT64_POLY (template parameter of the kernel): 1 of every AT64_POLY exponentials runs on the FMA pipe (0: none)

struct AttnSmem64 {
  static constexpr int Q_BYTES = AT_BQ * 64 * 2;           // 16 KB
  static constexpr int K_BYTES = AT_BK * 64 * 2;           // 16 KB (128 keys per TMA tile, consumed as two halves)
  static constexpr int STAGE_BYTES = 2 * K_BYTES;          // K tile + V tile
  static constexpr int P_BYTES = AT_BQ * 64 * 2;           // 16 KB per 64-key P tile, one per group
  static constexpr int OFF_Q = 0, OFF_KV = Q_BYTES;
  static constexpr int OFF_P = OFF_KV + 2 * STAGE_BYTES;
  static constexpr int BAR_OFFSET = OFF_P + 2 * P_BYTES;
  static constexpr int ALIGN_SLACK = 768;
  static constexpr int TOTAL = BAR_OFFSET + 12 * 8 + 16 + ALIGN_SLACK;
};

template <int AT64_POLY>
__global__ void __launch_bounds__(AT64_THREADS, 2)
attn_fwd64_kernel(const __grid_constant__ CUtensorMap tm, const AttnParams p) {
  constexpr int AT_HD = 64;
  constexpr int HK = 64;             // keys per pipeline step
  using SM = AttnSmem64;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SM::BAR_OFFSET);
  uint64_t* q_full = bars + 0;
  uint64_t* kv_full = bars + 1;      // [2]
  uint64_t* kv_empty = bars + 3;     // [2]
  uint64_t* s_full = bars + 5;       // [2]  S(t) landed in TMEM buffer t & 1
  uint64_t* p_full = bars + 7;       // [2]  group t & 1 consumed S(t), wrote P(t) (and rescaled its O)
  uint64_t* o_full = bars + 9;       //      last P.V finished
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 12);
  if (static_cast<int>(smem - smem_raw) > SM::ALIGN_SLACK) __trap();

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int head = blockIdx.x, clip = blockIdx.y;
  const int qt = p.causal ? static_cast<int>(gridDim.z - 1 - blockIdx.z) : static_cast<int>(blockIdx.z);
  const int kvh = head / (p.n_heads / p.n_kv_heads);
  const int q0 = qt * AT_BQ;
  const int clip_row0 = p.row0 + clip * p.S;
  const int n_kv = p.causal ? (qt + 1) : (p.S + AT_BK - 1) / AT_BK;
  // 64-key steps; the second half of the last tile is skipped when it holds no key of the clip
  const int last_keys = (p.causal ? min(p.S, q0 + AT_BQ) : p.S) - (n_kv - 1) * AT_BK;
  const int T = 2 * n_kv - (last_keys <= HK ? 1 : 0);
  const int col_q = head * AT_HD;
  const int col_k = (p.n_heads + kvh) * AT_HD;
  const int col_v = (p.n_heads + p.n_kv_heads + kvh) * AT_HD;

  if (warp == 0 && lane == 0) tma_prefetch_desc(&tm);
  if (warp == 1) {
    if (lane == 0) {
      mbar_init(q_full, 1);
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        mbar_init(kv_full + i, 1);
        mbar_init(kv_empty + i, 1);
        mbar_init(s_full + i, 1);
        mbar_init(p_full + i, 4);      // one arrive per softmax warp of the group
      }
      mbar_init(o_full, 1);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(tmem_ptr_smem, 256);    // S buffers: columns [0,64), [64,128); O of group 0 / 1: [128,192), [192,256)
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const uint32_t tmem_o = tmem_base + 128;

  if (warp == 0) {
    if (elect_one()) {
      mbar_expect_tx(q_full, SM::Q_BYTES);
      tma_load_2d(&tm, q_full, smem + SM::OFF_Q, col_q, clip_row0 + q0);
      for (int j = 0; j < n_kv; ++j) {
        const int st = j & 1;
        uint8_t* kv = smem + SM::OFF_KV + st * SM::STAGE_BYTES;
        mbar_wait(kv_empty + st, ((j >> 1) & 1) ^ 1);
        mbar_expect_tx(kv_full + st, SM::STAGE_BYTES);
        tma_load_2d(&tm, kv_full + st, kv, col_k, clip_row0 + j * AT_BK);
        tma_load_2d(&tm, kv_full + st, kv + SM::K_BYTES, col_v, clip_row0 + j * AT_BK);
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc_s = make_idesc_bf16(AT_BQ, HK, 0, 0);      // Q (K-major) x K_half (K-major)
    constexpr uint32_t idesc_o = make_idesc_bf16(AT_BQ, AT_HD, 0, 1);   // P (K-major) x V_half (MN-major)
    const uint32_t sQ = smem_u32(smem + SM::OFF_Q);
    const uint32_t sP = smem_u32(smem + SM::OFF_P);
    auto issue_s = [&](int t) {
      const int j = t >> 1, h = t & 1, st = j & 1;
      if (h == 0) {
        mbar_wait(kv_full + st, (j >> 1) & 1);
        tc_fence_after();
      }
      if (elect_one()) {
        const uint32_t sK = smem_u32(smem + SM::OFF_KV + st * SM::STAGE_BYTES) + h * (HK * 128);
#pragma unroll
        for (int k = 0; k < AT_HD / 16; ++k) {
          const uint64_t qd = make_smem_desc_sw128(sQ, 16, 1024) + 2 * k;
          const uint64_t kd = make_smem_desc_sw128(sK, 16, 1024) + 2 * k;
          umma_bf16(tmem_base + (t & 1) * HK, qd, kd, idesc_s, k > 0 ? 1u : 0u);
        }
        umma_commit(s_full + (t & 1));
      }
      __syncwarp();
    };
    mbar_wait(q_full, 0);
    issue_s(0);
    if (T > 1) issue_s(1);
    for (int t = 0; t < T; ++t) {
      const int j = t >> 1, h = t & 1, st = j & 1;
      mbar_wait(p_full + (t & 1), (t >> 1) & 1);      // group t&1 consumed S(t), wrote P(t), rescaled its O if it had to
      tc_fence_after();
      if (elect_one()) {
        // MN-major V: 8 key rows per 1 KB group (SBO); this half starts 64 key rows (8 KB) into the tile
        const uint32_t sV = smem_u32(smem + SM::OFF_KV + st * SM::STAGE_BYTES + SM::K_BYTES) + h * (HK * 128);
        const uint64_t vd = make_smem_desc_sw128(sV, 16384, 1024);
#pragma unroll
        for (int k = 0; k < HK / 16; ++k) {
          const uint64_t pd = make_smem_desc_sw128(sP + (t & 1) * SM::P_BYTES, 16, 1024) + 2 * k;
          umma_bf16(tmem_o + (t & 1) * AT_HD, pd, vd + 128 * k, idesc_o, (t > 1 || k > 0) ? 1u : 0u);
        }
        if (h == 1 || t == T - 1) umma_commit(kv_empty + st);
        if (t == T - 1) umma_commit(o_full);
      }
      __syncwarp();
      if (t + 2 < T) issue_s(t + 2);
    }
  } else {
    // ===== softmax: thread = query row; group g owns the steps t with t & 1 == g =====
    const int g = (warp - 2) >> 2;
    const int q = warp & 3;
    const int r = q * 32 + lane;                 // row inside the tile == TMEM lane
    const int qpos = q0 + r;                     // position inside the clip
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    const uint32_t tmem_sg = tmem_base + lane_addr + g * HK;
    const uint32_t tmem_og = tmem_o + lane_addr + g * AT_HD;
    uint8_t* prow = smem + SM::OFF_P + g * SM::P_BYTES + r * 128;
    float m = -INFINITY, l = 0.f;
    const float sc = p.scale_log2;
    const int kmax = p.causal ? min(qpos, p.S - 1) : (p.S - 1);      // last visible key position

    for (int t = g; t < T; t += 2) {
      const int k0 = t * HK;
      mbar_wait(s_full + g, (t >> 1) & 1);
      tc_fence_after();
      uint32_t v0[32], v1[32];
      tmem_ld_32x32(tmem_sg, v0);
      tmem_ld_32x32(tmem_sg + 32, v1);
      tmem_ld_wait();
      if (k0 + HK - 1 > kmax) {      // boundary step: masked scores become -inf once, the passes below stay branch-free
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          if (k0 + i > kmax) v0[i] = 0xff800000u;
          if (k0 + 32 + i > kmax) v1[i] = 0xff800000u;
        }
      }
      float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
      for (int i = 0; i < 32; i += 2) {
        mx0 = fmaxf(mx0, __uint_as_float(v0[i]));
        mx1 = fmaxf(mx1, __uint_as_float(v0[i + 1]));
        mx2 = fmaxf(mx2, __uint_as_float(v1[i]));
        mx3 = fmaxf(mx3, __uint_as_float(v1[i + 1]));
      }
      const float m_tile = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * sc;
      // lazy rescaling: raise the reference maximum only when it grew by more than 8 (log2 units)
      const bool raise = m_tile > m + 8.0f;        // true on the group's first step (m = -inf) unless fully masked
      float alpha = 1.0f;
      if (raise) {
        alpha = ex2_approx(m - m_tile);            // m = -inf -> 0
        m = m_tile;
      }
      const float neg_m = (m == -INFINITY) ? 0.f : -m;
      float2 s01 = make_float2(0.f, 0.f), s23 = make_float2(0.f, 0.f);
      const float2 sc2 = make_float2(sc, sc), nm2 = make_float2(neg_m, neg_m);
      auto emit = [&](uint32_t (&v)[32], int c, float2& sab) {
        uint32_t packed[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          // packed fp32 (FFMA2 / FADD2): one issue slot per pair of scores for the scaling and for the row sum;
          // every AT64_POLY-th exponential runs on the FMA pipe (ex2_poly) instead of the MUFU
          const float2 x = ffma2(make_float2(__uint_as_float(v[i]), __uint_as_float(v[i + 1])), sc2, nm2);
          const float p0 = ex2_approx(x.x);                                           // -inf -> 0
          const float p1 = (AT64_POLY > 0 && ((i + 1) % AT64_POLY) == AT64_POLY - 1) ? ex2_poly(x.y) : ex2_approx(x.y);
          sab = fadd2(sab, make_float2(p0, p1));
          packed[i >> 1] = f2_to_bf2(p0, p1);
        }
        // 32 keys = 4 chunks of 16 bytes; chunk index inside the 128-byte row = c * 4 + t4, XOR-swizzled with the row
#pragma unroll
        for (int t4 = 0; t4 < 4; ++t4) {
          const int chunk = c * 4 + t4;
          *reinterpret_cast<uint4*>(prow + ((chunk ^ (r & 7)) << 4)) =
              make_uint4(packed[4 * t4], packed[4 * t4 + 1], packed[4 * t4 + 2], packed[4 * t4 + 3]);
        }
      };
      emit(v0, 0, s01);
      emit(v1, 1, s23);
      l = l * alpha + ((s01.x + s01.y) + (s23.x + s23.y));
      // O_g(TMEM) *= alpha for the rows of this warp, only when some row raised its maximum (never on the group's first
      // step: its first P.V overwrites O_g).  The group's previous P.V(t-2) has finished (s_full(t) was committed after it).
      if (t >= 2 && __any_sync(0xffffffffu, raise)) {
#pragma unroll
        for (int c = 0; c < AT_HD / 32; ++c) {
          uint32_t o[32];
          tmem_ld_32x32(tmem_og + c * 32, o);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
          tmem_st_32x32(tmem_og + c * 32, o);
        }
        tmem_st_wait();
      }
      // make the generic-proxy smem writes visible to the tensor core (async proxy), then release S / publish P
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full + g);
    }
    // ---- merge the two streams: row r of group g handles output columns [32 g, 32 g + 32) ----
    mbar_wait(o_full, 0);
    tc_fence_after();
    float2* stat = reinterpret_cast<float2*>(smem + SM::OFF_P);       // P tiles are idle now: [2][128] (m, l)
    stat[g * AT_BQ + r] = make_float2(m, l);
    asm volatile("bar.sync 1, 256;" ::: "memory");
    const float2 other = stat[(g ^ 1) * AT_BQ + r];
    const float m_all = fmaxf(m, other.x);
    const float w_own = (m == -INFINITY) ? 0.f : ex2_approx(m - m_all);
    const float w_oth = (other.x == -INFINITY) ? 0.f : ex2_approx(other.x - m_all);
    const float l_all = l * w_own + other.y * w_oth;
    const float w0 = g == 0 ? w_own : w_oth, w1 = g == 0 ? w_oth : w_own;     // weights of O_0 / O_1
    uint32_t oa[32], ob[32];
    tmem_ld_32x32(tmem_o + lane_addr + g * 32, oa);                           // O_0[:, 32 g ..]
    if (T > 1) {
      tmem_ld_32x32(tmem_o + lane_addr + AT_HD + g * 32, ob);                 // O_1[:, 32 g ..]
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i) ob[i] = 0u;                                // group 1 never ran: its O is unwritten
    }
    tmem_ld_wait();
    if (qpos < p.S) {
      const float inv = l_all > 0.f ? 1.0f / l_all : 0.f;
      const float a0 = w0 * inv, a1 = (T > 1) ? w1 * inv : 0.f;
      const long long row = static_cast<long long>(clip_row0) + qpos;
      bf16* op = p.out + row * p.out_ld + head * AT_HD + g * 32;
#pragma unroll
      for (int i = 0; i < 32; i += 8) {
        float f[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(oa[i + e]) * a0 + __uint_as_float(ob[i + e]) * a1;
        uint4 u;
        u.x = f2_to_bf2(f[0], f[1]);
        u.y = f2_to_bf2(f[2], f[3]);
        u.z = f2_to_bf2(f[4], f[5]);
        u.w = f2_to_bf2(f[6], f[7]);
        *reinterpret_cast<uint4*>(op + i) = u;
      }
      if (p.lse && g == 0) p.lse[static_cast<long long>(head) * p.M + row] = (m_all + log2f(l_all)) * 0.6931471805599453f;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 256);
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
  if (head_dim == 64) {
    // share of the exponentials computed on the FMA pipe: OMNI_AT64_POLY = 0 (none, default), 2 (1/2), 3 (1/6), 4 (1/4).
    // Measured at Whisper's shape (B=16, S=1500, 16 x 64; profiles/README.md): 0.273 / 0.286 / 0.280 / 0.297 ms for
    // 0 / 3 / 4 / 2 -- the kernel is bound by issue slots and latency (MUFU pipe at 58 %), so trading one MUFU op for
    // eight FMA-pipe instructions loses; kept as a switch for the experiment's record.
    static const int poly = getenv("OMNI_AT64_POLY") ? atoi(getenv("OMNI_AT64_POLY")) : 0;
    auto k64 = poly == 0 ? attn_fwd64_kernel<0> : poly == 2 ? attn_fwd64_kernel<2> : poly == 3 ? attn_fwd64_kernel<3>
                                                                                               : attn_fwd64_kernel<4>;
    static bool attr64 = false;
    if (!attr64) {
      if (cudaFuncSetAttribute(k64, cudaFuncAttributeMaxDynamicSharedMemorySize, AttnSmem64::TOTAL) != cudaSuccess)
        return OMNI_ERR_CUDA;
      attr64 = true;
    }
    k64<<<grid, AT64_THREADS, AttnSmem64::TOTAL, st>>>(tm, p);
    OMNI_LAUNCH_CHECK();
    return OMNI_OK;
  }
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
