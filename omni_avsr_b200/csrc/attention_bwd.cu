// Flash-attention backward on the sm_100a tensor cores (head_dim 64 or 128), two kernels, no atomics:
//   attn_bwd_dq_kernel   : CTA = one 128-query tile of one head;  loops over 64-key steps;  dQ accumulates in TMEM
//   attn_bwd_dkv_kernel  : CTA = one 128-key tile of one KV head; loops over (query head of the group, 64-query step);
//                          dK and dV accumulate in TMEM (the GQA group sum happens in the accumulator)
// The 64-wide steps keep a CTA at 256 TMEM columns and < 113 KB of shared memory for head_dim 64, so two CTAs share
// an SM and one CTA's exp2 / dS arithmetic runs under the other's MMAs; the streamed operand tiles are double-buffered.
// Both recompute S and dP on the tensor cores from the packed q|k|v rows and dOut, rebuild P = exp2(s*scale - lse)
// from the log-sum-exp saved by the forward kernel, form dS = P o (dP - delta) per thread (thread = TMEM lane) and
// feed bf16 P / dS back to tcgen05.mma as K-major operands through a 128B-swizzled shared-memory tile, exactly like
// the forward kernel's P.  K / V / Q / dO tiles are used twice from the same TMA tile: K-major for the score GEMMs,
// MN-major for the gradient GEMMs.  delta[h, row] = sum_d dO*O comes from a small HBM-bound pre-pass.
//
// Replaces what autograd derives for F.scaled_dot_product_attention at Llama_LoRA.py:300 / Qwen_LoRA.py:606 (causal GQA)
// and fairseq multihead_attention.py:619-654 (non-causal, AV-HuBERT LoRA fine-tuning).
#include "common.cuh"
#include "../../include/omni_avsr.h"

namespace omni {

constexpr int AB_T = 128;          // queries per tile == keys per tile
constexpr int AB_THREADS = 320;    // warp 0 TMA, warp 1 MMA, warps 2..9 compute: thread = (TMEM lane, 32-column half of the step);
                                   // warp w may only read TMEM lanes 32*(w%4).., so warps w and w+4 share a lane quarter

struct AttnBwdParams {
  bf16* dqkv;            // [M, dqkv_ld] packed gradient rows, same column layout as qkv
  long long dqkv_ld;
  const float* stats;    // [n_heads][B][n_qs][2][64]: per 64-query step, -lse * log2(e) (-inf for the queries past S, which makes
                         // their P = exp2(s - inf) = 0 without any test) then -rowsum(dO o O); written by attn_delta_kernel
  long long M;
  int row0, S, n_heads, n_kv_heads, causal, B, n_qs;
  float scale, scale_log2;
};

// ---------------------------------------------------------------------------------------------------------------
// statistics pre-pass: one warp per (clip, padded query position), HD/8 lanes per head (16-byte loads), fp32 sum.
// Writes the per-step statistics blocks the two kernels read, NEGATED (they are the addends of packed FFMA2 / FADD2
// instructions): stats[h][clip][pos / 64][0][pos % 64] = -lse * log2(e), [1][pos % 64] = -delta = -rowsum(dO o O);
// positions in [S, 64 n_qs) get (-inf, 0).  The blocks are 512 bytes, 512-byte
// aligned: the dK/dV kernel fetches one per step with a single bulk copy next to its Q / dO tiles.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
attn_delta_kernel(const bf16* __restrict__ dout, long long do_ld, const bf16* __restrict__ out, long long o_ld,
                  const float* __restrict__ lse, float* __restrict__ stats, long long M, int row0, int B, int S, int n_qs,
                  int n_heads, int hd) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int s_pad = n_qs * 64;
  if (w >= B * s_pad) return;
  const int clip = w / s_pad, pos = w - clip * s_pad;
  const bool valid = pos < S;
  const long long row = static_cast<long long>(row0) + static_cast<long long>(clip) * S + pos;
  const int lph = hd >> 3;                 // lanes per head
  const int hpp = 32 / lph;                // heads per pass
  const int sub = lane / lph, l = lane % lph;
  for (int h0 = 0; h0 < n_heads; h0 += hpp) {
    const int h = h0 + sub;
    float acc = 0.f;
    if (h < n_heads && valid) {
      const uint4 a = ld_nc_u4(dout + row * do_ld + h * hd + l * 8);
      const uint4 b = ld_nc_u4(out + row * o_ld + h * hd + l * 8);
      const uint32_t av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 x = bf2_to_f2(av[i]), y = bf2_to_f2(bv[i]);
        acc = fmaf(x.x, y.x, acc);
        acc = fmaf(x.y, y.y, acc);
      }
    }
    for (int o = lph >> 1; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (h < n_heads && l == 0) {
      float* blk = stats + ((static_cast<long long>(h) * B + clip) * n_qs + (pos >> 6)) * 128 + (pos & 63);
      blk[0] = valid ? -lse[static_cast<long long>(h) * M + row] * 1.4426950408889634f : -INFINITY;
      blk[64] = -acc;
    }
  }
}

// write 32 bf16 (16 packed words) of row r, column block c (32 columns, c = 0 or 1) into a [128 x 64] K-major
// 128B-swizzled operand tile
__device__ __forceinline__ void store_operand_chunk(uint8_t* tile, int r, int c, const uint32_t (&w)[16]) {
  uint8_t* blk = tile + r * 128;
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const int chunk = c * 4 + t;
    *reinterpret_cast<uint4*>(blk + ((chunk ^ (r & 7)) << 4)) = make_uint4(w[4 * t], w[4 * t + 1], w[4 * t + 2], w[4 * t + 3]);
  }
}

constexpr int AB_STEP = 64;        // streamed rows (keys in the dQ kernel, queries in the dK/dV kernel) per step

// K-major descriptor of k-step k (16 elements) inside a tile made of 64-column 128B-swizzled blocks of `blk_bytes`
__device__ __forceinline__ uint64_t kmajor_desc(uint32_t base, int k, uint32_t blk_bytes) {
  return make_smem_desc_sw128(base + (k >> 2) * blk_bytes, 16, 1024) + 2 * (k & 3);
}

// ---------------------------------------------------------------------------------------------------------------
// dQ
// ---------------------------------------------------------------------------------------------------------------
template <int HD>
struct BwdQSmem {
  static constexpr int TILE = AB_T * HD * 2;               // Q, dO: [128 x HD]
  static constexpr int STEP = AB_STEP * HD * 2;            // K, V step tiles: [64 x HD]
  static constexpr int OFF_Q = 0, OFF_DO = TILE, OFF_KV = 2 * TILE;     // K/V: 2 stages of (K | V)
  static constexpr int OFF_DS = OFF_KV + 4 * STEP;         // dS [128 x 64] bf16
  static constexpr int BAR_OFFSET = OFF_DS + 16384;
  static constexpr int ALIGN_SLACK = HD == 64 ? 768 : 1024;
  static constexpr int TOTAL = BAR_OFFSET + 128 + ALIGN_SLACK;
  static constexpr int TMEM_COLS = 256;                    // S [0,64) | dP [64,128) | dQ [128,128+HD)
};

template <int HD>
__global__ void __launch_bounds__(AB_THREADS, HD == 64 ? 2 : 1)
attn_bwd_dq_kernel(const __grid_constant__ CUtensorMap tm, const __grid_constant__ CUtensorMap tm64,
                   const __grid_constant__ CUtensorMap tmdo, const AttnBwdParams p) {
  using SM = BwdQSmem<HD>;
  constexpr int NB = HD / 64;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  if (static_cast<int>(smem - smem_raw) > SM::ALIGN_SLACK) __trap();
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SM::BAR_OFFSET);
  uint64_t* q_full = bars + 0;
  uint64_t* kv_full = bars + 1;      // [2]
  uint64_t* kv_empty = bars + 3;     // [2]
  uint64_t* s_full = bars + 5;
  uint64_t* ds_full = bars + 6;
  uint64_t* dq_full = bars + 7;
  uint64_t* sd_read = bars + 8;      // compute warps hold S / dP of the step in registers: the TMEM buffers are free
  uint64_t* ds_free = bars + 9;      // dQ MMAs of the step finished reading the dS tile
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 10);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // grid (head, clip, tile), heaviest query tile (the last one under the causal mask) dispatched first
  const int head = blockIdx.x, clip = blockIdx.y;
  const int qt = p.causal ? static_cast<int>(gridDim.z - 1 - blockIdx.z) : static_cast<int>(blockIdx.z);
  const int kvh = head / (p.n_heads / p.n_kv_heads);
  const int q0 = qt * AB_T;
  const int clip_row0 = p.row0 + clip * p.S;
  const int n_all = (p.S + AB_STEP - 1) / AB_STEP;
  const int n_kv = p.causal ? min(2 * qt + 2, n_all) : n_all;
  const int col_q = head * HD;
  const int col_k = (p.n_heads + kvh) * HD;
  const int col_v = (p.n_heads + p.n_kv_heads + kvh) * HD;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm);
    tma_prefetch_desc(&tm64);
    tma_prefetch_desc(&tmdo);
  }
  if (warp == 1) {
    if (lane == 0) {
      mbar_init(q_full, 1);
      mbar_init(kv_full, 1);
      mbar_init(kv_full + 1, 1);
      mbar_init(kv_empty, 1);
      mbar_init(kv_empty + 1, 1);
      mbar_init(s_full, 1);
      mbar_init(ds_full, 8);
      mbar_init(dq_full, 1);
      mbar_init(sd_read, 8);
      mbar_init(ds_free, 1);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(tmem_ptr_smem, SM::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const uint32_t tmem_s = tmem_base, tmem_dp = tmem_base + 64, tmem_dq = tmem_base + 128;

  if (warp == 0) {
    if (elect_one()) {
      mbar_expect_tx(q_full, 2 * SM::TILE);
#pragma unroll
      for (int b = 0; b < NB; ++b) {
        tma_load_2d(&tm, q_full, smem + SM::OFF_Q + b * 16384, col_q + b * 64, clip_row0 + q0);
        tma_load_2d(&tmdo, q_full, smem + SM::OFF_DO + b * 16384, col_q + b * 64, clip_row0 + q0);
      }
      for (int j = 0; j < n_kv; ++j) {
        const int st = j & 1;
        uint8_t* kv = smem + SM::OFF_KV + st * 2 * SM::STEP;
        mbar_wait(kv_empty + st, ((j >> 1) & 1) ^ 1);
        mbar_expect_tx(kv_full + st, 2 * SM::STEP);
#pragma unroll
        for (int b = 0; b < NB; ++b) {
          tma_load_2d(&tm64, kv_full + st, kv + b * 8192, col_k + b * 64, clip_row0 + j * AB_STEP);
          tma_load_2d(&tm64, kv_full + st, kv + SM::STEP + b * 8192, col_v + b * 64, clip_row0 + j * AB_STEP);
        }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc_s = make_idesc_bf16(AB_T, AB_STEP, 0, 0);
    constexpr uint32_t idesc_g = make_idesc_bf16(AB_T, HD, 0, 1);     // dS (K-major) x K step tile (MN-major)
    const uint32_t sQ = smem_u32(smem + SM::OFF_Q), sDO = smem_u32(smem + SM::OFF_DO);
    const uint32_t sDS = smem_u32(smem + SM::OFF_DS);
    // Issue order: S,dP(0) | S,dP(1) dQ(0) | S,dP(2) dQ(1) | ...  The score MMAs of step j+1 go out as soon as the compute
    // warps hold step j's S / dP in registers (sd_read), i.e. they run under step j's exp2 / dS arithmetic.
    auto issue_scores = [&](int j) {
      const int st = j & 1;
      const uint32_t sK = smem_u32(smem + SM::OFF_KV + st * 2 * SM::STEP), sV = sK + SM::STEP;
      mbar_wait(kv_full + st, (j >> 1) & 1);
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < HD / 16; ++k)
          umma_bf16(tmem_s, kmajor_desc(sQ, k, 16384), kmajor_desc(sK, k, 8192), idesc_s, k > 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < HD / 16; ++k)
          umma_bf16(tmem_dp, kmajor_desc(sDO, k, 16384), kmajor_desc(sV, k, 8192), idesc_s, k > 0 ? 1u : 0u);
        umma_commit(s_full);
      }
      __syncwarp();
    };
    mbar_wait(q_full, 0);
    issue_scores(0);
    for (int j = 0; j < n_kv; ++j) {
      const uint32_t ph = j & 1;
      const int st = j & 1;
      const uint32_t sK = smem_u32(smem + SM::OFF_KV + st * 2 * SM::STEP);
      if (j + 1 < n_kv) {
        mbar_wait(sd_read, ph);
        tc_fence_after();
        issue_scores(j + 1);
      }
      mbar_wait(ds_full, ph);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t kd = make_smem_desc_sw128(sK, 8192, 1024);       // MN-major [64 keys, HD]
#pragma unroll
        for (int k = 0; k < AB_STEP / 16; ++k)
          umma_bf16(tmem_dq, kmajor_desc(sDS, k, 0), kd + 128 * k, idesc_g, (j > 0 || k > 0) ? 1u : 0u);
        umma_commit(kv_empty + st);
        umma_commit(ds_free);
        if (j == n_kv - 1) umma_commit(dq_full);
      }
      __syncwarp();
    }
  } else {
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;            // which 32 of the step's 64 columns this thread converts
    const int r = q * 32 + lane;
    const int qpos = q0 + r;
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    const long long row = static_cast<long long>(clip_row0) + qpos;
    const bool row_ok = qpos < p.S;
    // -lse2 = -inf for the rows past the clip: their P (and dS) come out as exactly 0 on the compare-free path
    const float* sblk = p.stats + ((static_cast<long long>(head) * p.B + clip) * p.n_qs + (qpos >> 6)) * 128 + (qpos & 63);
    const float nlse2 = row_ok ? sblk[0] : -INFINITY;      // -lse * log2(e)
    const float ndl = row_ok ? sblk[64] : 0.f;             // -delta
    const int kmax = p.causal ? min(qpos, p.S - 1) : (p.S - 1);      // last visible key position
    const float sc = p.scale_log2;
    uint8_t* sDS = smem + SM::OFF_DS;
    for (int j = 0; j < n_kv; ++j) {
      const uint32_t ph = j & 1;
      const int k0 = j * AB_STEP;
      mbar_wait(s_full, ph);
      tc_fence_after();
      const bool full = (k0 + AB_STEP - 1 <= kmax);
      {
        const int c = half;
        uint32_t s[32], d[32], w[16];
        tmem_ld_32x32(tmem_s + lane_addr + c * 32, s);
        tmem_ld_32x32(tmem_dp + lane_addr + c * 32, d);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(sd_read);     // S / dP are in registers: the next step's score MMAs may overwrite TMEM
        if (!full) {       // boundary step: masked scores become -inf once (exp2 -> 0), the pass below stays branch-free
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (k0 + c * 32 + i > kmax) s[i] = 0xff800000u;
        }
        const float2 sc2 = make_float2(sc, sc), nl2 = make_float2(nlse2, nlse2), ndl2 = make_float2(ndl, ndl);
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          // packed fp32 (FFMA2 / FADD2 / FMUL2): three issue slots per PAIR of elements around the two exponentials
          const float2 x = ffma2(make_float2(__uint_as_float(s[i]), __uint_as_float(s[i + 1])), sc2, nl2);
          const float2 t = fadd2(make_float2(__uint_as_float(d[i]), __uint_as_float(d[i + 1])), ndl2);
          const float2 g2 = fmul2(make_float2(ex2_approx(x.x), ex2_approx(x.y)), t);
          w[i >> 1] = f2_to_bf2(g2.x, g2.y);
        }
        if (j > 0) mbar_wait(ds_free, (j - 1) & 1);      // the previous step's dQ MMAs are done with the dS tile
        store_operand_chunk(sDS, r, c, w);
      }
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(ds_full);
    }
    mbar_wait(dq_full, 0);
    tc_fence_after();
    bf16* op = p.dqkv + row * p.dqkv_ld + col_q;
#pragma unroll
    for (int cc = 0; cc < HD / 64; ++cc) {
      const int c = half * (HD / 64) + cc;
      uint32_t v[32];
      tmem_ld_32x32(tmem_dq + lane_addr + c * 32, v);
      tmem_ld_wait();
      if (row_ok) {
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          uint4 u;
          u.x = f2_to_bf2(__uint_as_float(v[i]) * p.scale, __uint_as_float(v[i + 1]) * p.scale);
          u.y = f2_to_bf2(__uint_as_float(v[i + 2]) * p.scale, __uint_as_float(v[i + 3]) * p.scale);
          u.z = f2_to_bf2(__uint_as_float(v[i + 4]) * p.scale, __uint_as_float(v[i + 5]) * p.scale);
          u.w = f2_to_bf2(__uint_as_float(v[i + 6]) * p.scale, __uint_as_float(v[i + 7]) * p.scale);
          *reinterpret_cast<uint4*>(op + c * 32 + i) = u;
        }
      }
    }
    tc_fence_before();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, SM::TMEM_COLS);
}

// ---------------------------------------------------------------------------------------------------------------
// dK, dV
// ---------------------------------------------------------------------------------------------------------------
template <int HD>
struct BwdKVSmem {
  static constexpr int TILE = AB_T * HD * 2;               // K, V: [128 x HD]
  static constexpr int STEP = AB_STEP * HD * 2;            // Q, dO step tiles: [64 x HD]
  static constexpr int OFF_K = 0, OFF_V = TILE, OFF_QDO = 2 * TILE;     // 2 stages of (Q | dO)
  static constexpr int OFF_P = OFF_QDO + 4 * STEP;         // P^T [128 x 64] bf16
  static constexpr int OFF_DS = OFF_P + 16384;             // dS^T [128 x 64] bf16
  static constexpr int OFF_STAT = OFF_DS + 16384;          // [2 stages][lse2 | delta][64] fp32 (bulk copy per step)
  static constexpr int BAR_OFFSET = OFF_STAT + 2 * 2 * AB_STEP * 4;
  static constexpr int ALIGN_SLACK = HD == 64 ? 768 : 1024;
  static constexpr int TOTAL = BAR_OFFSET + 128 + ALIGN_SLACK;
  static constexpr int TMEM_COLS = HD == 64 ? 256 : 512;   // S^T [0,64) | dP^T [64,128) | dV [128,128+HD) | dK behind
};

template <int HD>
__global__ void __launch_bounds__(AB_THREADS, HD == 64 ? 2 : 1)
attn_bwd_dkv_kernel(const __grid_constant__ CUtensorMap tm, const __grid_constant__ CUtensorMap tm64,
                    const __grid_constant__ CUtensorMap tmdo64, const AttnBwdParams p) {
  using SM = BwdKVSmem<HD>;
  constexpr int NB = HD / 64;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  if (static_cast<int>(smem - smem_raw) > SM::ALIGN_SLACK) __trap();
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SM::BAR_OFFSET);
  uint64_t* kv_full = bars + 0;
  uint64_t* q_full = bars + 1;       // [2]
  uint64_t* q_empty = bars + 3;      // [2]
  uint64_t* s_full = bars + 5;
  uint64_t* pds_full = bars + 6;
  uint64_t* acc_full = bars + 7;
  uint64_t* sd_read = bars + 8;      // compute warps hold S^T / dP^T of the step in registers
  uint64_t* pds_free = bars + 9;     // dV / dK MMAs of the step finished reading the P^T / dS^T tiles
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 10);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // grid (KV head, clip, tile): key tile 0 is the heaviest under the causal mask (every query sees it) and goes first
  const int kvh = blockIdx.x, clip = blockIdx.y, jt = blockIdx.z;
  const int G = p.n_heads / p.n_kv_heads;
  const int k0 = jt * AB_T;
  const int clip_row0 = p.row0 + clip * p.S;
  const int n_qs = (p.S + AB_STEP - 1) / AB_STEP;          // 64-query steps in the clip
  const int i0 = p.causal ? 2 * jt : 0;                    // first step that sees any key of this tile
  const int per_head = n_qs - i0;
  const int n_it = G * per_head;
  const int col_k = (p.n_heads + kvh) * HD;
  const int col_v = (p.n_heads + p.n_kv_heads + kvh) * HD;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm);
    tma_prefetch_desc(&tm64);
    tma_prefetch_desc(&tmdo64);
  }
  if (warp == 1) {
    if (lane == 0) {
      mbar_init(kv_full, 1);
      mbar_init(q_full, 1);
      mbar_init(q_full + 1, 1);
      mbar_init(q_empty, 1);
      mbar_init(q_empty + 1, 1);
      mbar_init(s_full, 1);
      mbar_init(pds_full, 8);
      mbar_init(acc_full, 1);
      mbar_init(sd_read, 8);
      mbar_init(pds_free, 1);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(tmem_ptr_smem, SM::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const uint32_t tmem_s = tmem_base, tmem_dp = tmem_base + 64, tmem_dv = tmem_base + 128, tmem_dk = tmem_base + 128 + HD;

  if (warp == 0) {
    if (elect_one()) {
      mbar_expect_tx(kv_full, 2 * SM::TILE);
#pragma unroll
      for (int b = 0; b < NB; ++b) {
        tma_load_2d(&tm, kv_full, smem + SM::OFF_K + b * 16384, col_k + b * 64, clip_row0 + k0);
        tma_load_2d(&tm, kv_full, smem + SM::OFF_V + b * 16384, col_v + b * 64, clip_row0 + k0);
      }
      for (int it = 0; it < n_it; ++it) {
        const int head = kvh * G + it / per_head;
        const int q0 = (i0 + it % per_head) * AB_STEP;
        const int st = it & 1;
        uint8_t* qd = smem + SM::OFF_QDO + st * 2 * SM::STEP;
        mbar_wait(q_empty + st, ((it >> 1) & 1) ^ 1);
        mbar_expect_tx(q_full + st, 2 * SM::STEP + 512);
#pragma unroll
        for (int b = 0; b < NB; ++b) {
          tma_load_2d(&tm64, q_full + st, qd + b * 8192, head * HD + b * 64, clip_row0 + q0);
          tma_load_2d(&tmdo64, q_full + st, qd + SM::STEP + b * 8192, head * HD + b * 64, clip_row0 + q0);
        }
        // the step's 64 (lse2, delta) pairs ride on the same barrier
        bulk_load_1d(smem + SM::OFF_STAT + st * 512,
                     p.stats + ((static_cast<long long>(head) * p.B + clip) * p.n_qs + (i0 + it % per_head)) * 128, 512,
                     q_full + st);
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc_s = make_idesc_bf16(AB_T, AB_STEP, 0, 0);
    constexpr uint32_t idesc_g = make_idesc_bf16(AB_T, HD, 0, 1);
    const uint32_t sK = smem_u32(smem + SM::OFF_K), sV = smem_u32(smem + SM::OFF_V);
    const uint32_t sP = smem_u32(smem + SM::OFF_P), sDS = smem_u32(smem + SM::OFF_DS);
    // Issue order: S,dP(0) | S,dP(1) dV,dK(0) | S,dP(2) dV,dK(1) | ...  (see the dQ kernel)
    auto issue_scores = [&](int it) {
      const int st = it & 1;
      const uint32_t sQ = smem_u32(smem + SM::OFF_QDO + st * 2 * SM::STEP), sDO = sQ + SM::STEP;
      mbar_wait(q_full + st, (it >> 1) & 1);
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < HD / 16; ++k)        // S^T = K Q^T
          umma_bf16(tmem_s, kmajor_desc(sK, k, 16384), kmajor_desc(sQ, k, 8192), idesc_s, k > 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < HD / 16; ++k)        // dP^T = V dO^T
          umma_bf16(tmem_dp, kmajor_desc(sV, k, 16384), kmajor_desc(sDO, k, 8192), idesc_s, k > 0 ? 1u : 0u);
        umma_commit(s_full);
      }
      __syncwarp();
    };
    mbar_wait(kv_full, 0);
    if (n_it > 0) issue_scores(0);
    for (int it = 0; it < n_it; ++it) {
      const uint32_t ph = it & 1;
      const int st = it & 1;
      const uint32_t sQ = smem_u32(smem + SM::OFF_QDO + st * 2 * SM::STEP), sDO = sQ + SM::STEP;
      if (it + 1 < n_it) {
        mbar_wait(sd_read, ph);
        tc_fence_after();
        issue_scores(it + 1);
      }
      mbar_wait(pds_full, ph);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t dod = make_smem_desc_sw128(sDO, 8192, 1024);     // MN-major [64 queries, HD]
        const uint64_t qd = make_smem_desc_sw128(sQ, 8192, 1024);
#pragma unroll
        for (int k = 0; k < AB_STEP / 16; ++k)   // dV += P^T dO
          umma_bf16(tmem_dv, kmajor_desc(sP, k, 0), dod + 128 * k, idesc_g, (it > 0 || k > 0) ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < AB_STEP / 16; ++k)   // dK += dS^T Q
          umma_bf16(tmem_dk, kmajor_desc(sDS, k, 0), qd + 128 * k, idesc_g, (it > 0 || k > 0) ? 1u : 0u);
        umma_commit(q_empty + st);
        umma_commit(pds_free);
        if (it == n_it - 1) umma_commit(acc_full);
      }
      __syncwarp();
    }
  } else {
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int r = q * 32 + lane;                 // key row inside the tile == TMEM lane
    const int kpos = k0 + r;
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    const bool row_ok = kpos < p.S;
    const int qlo = row_ok ? (p.causal ? kpos : 0) : 0x7fffffff;     // first query position that sees this key
    const float sc = p.scale_log2;
    uint8_t* sP = smem + SM::OFF_P;
    uint8_t* sDS = smem + SM::OFF_DS;
    const float* stat = reinterpret_cast<const float*>(smem + SM::OFF_STAT);
    int step = 0;                                // 64-query step inside the current head (no integer division in the loop)
    for (int it = 0; it < n_it; ++it) {
      const uint32_t ph = it & 1;
      const int q0 = (i0 + step) * AB_STEP;
      if (++step == per_head) step = 0;
      const float* st = stat + (it & 1) * 128;
      mbar_wait(q_full + (it & 1), (it >> 1) & 1);         // the step's statistics block landed (TMA bulk copy)
      mbar_wait(s_full, ph);
      tc_fence_after();
      // every query of the step sees this key (queries past S carry -lse2 = -inf: P = 0 without a test)
      const bool full = qlo <= q0;
      {
        const int c = half;
        uint32_t s[32], d[32], wp[16], wd[16];
        tmem_ld_32x32(tmem_s + lane_addr + c * 32, s);
        tmem_ld_32x32(tmem_dp + lane_addr + c * 32, d);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(sd_read);     // the next step's score MMAs may overwrite S^T / dP^T
        if (!full) {       // diagonal step (or a key past S): masked scores become -inf once, the pass below is branch-free
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (q0 + c * 32 + i < qlo) s[i] = 0xff800000u;
        }
        const float2 sc2 = make_float2(sc, sc);
        const float4* L4 = reinterpret_cast<const float4*>(st + c * 32);
        const float4* D4 = reinterpret_cast<const float4*>(st + AB_STEP + c * 32);
#pragma unroll
        for (int i4 = 0; i4 < 8; ++i4) {
          const float4 l = L4[i4], dd = D4[i4];
          // packed fp32 (FFMA2 / FADD2 / FMUL2): s * scale - lse2, dP - delta, P * (dP - delta) for two queries per issue slot
          // (the statistics block holds -lse2 and -delta: plain addends)
          const float2 xa = ffma2(make_float2(__uint_as_float(s[i4 * 4]), __uint_as_float(s[i4 * 4 + 1])), sc2, make_float2(l.x, l.y));
          const float2 xb = ffma2(make_float2(__uint_as_float(s[i4 * 4 + 2]), __uint_as_float(s[i4 * 4 + 3])), sc2, make_float2(l.z, l.w));
          const float p0 = ex2_approx(xa.x), p1 = ex2_approx(xa.y), p2 = ex2_approx(xb.x), p3 = ex2_approx(xb.y);
          const float2 ta = fadd2(make_float2(__uint_as_float(d[i4 * 4]), __uint_as_float(d[i4 * 4 + 1])), make_float2(dd.x, dd.y));
          const float2 tb = fadd2(make_float2(__uint_as_float(d[i4 * 4 + 2]), __uint_as_float(d[i4 * 4 + 3])), make_float2(dd.z, dd.w));
          const float2 ga = fmul2(make_float2(p0, p1), ta), gb = fmul2(make_float2(p2, p3), tb);
          wp[i4 * 2] = f2_to_bf2(p0, p1);
          wp[i4 * 2 + 1] = f2_to_bf2(p2, p3);
          wd[i4 * 2] = f2_to_bf2(ga.x, ga.y);
          wd[i4 * 2 + 1] = f2_to_bf2(gb.x, gb.y);
        }
        if (it > 0) mbar_wait(pds_free, (it - 1) & 1);   // the previous step's dV / dK MMAs are done with the tiles
        store_operand_chunk(sP, r, c, wp);
        store_operand_chunk(sDS, r, c, wd);
      }
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(pds_full);
    }
    mbar_wait(acc_full, 0);
    tc_fence_after();
    const long long row = static_cast<long long>(clip_row0) + kpos;
    bf16* ov = p.dqkv + row * p.dqkv_ld + col_v;
    bf16* ok_ = p.dqkv + row * p.dqkv_ld + col_k;
#pragma unroll
    for (int cc = 0; cc < HD / 32; ++cc) {
      const int c = half * (HD / 32) + cc;                 // half 0 stores dV, half 1 stores dK
      uint32_t v[32];
      tmem_ld_32x32(tmem_dv + lane_addr + c * 32, v);      // dV columns first, dK right behind
      tmem_ld_wait();
      const bool is_k = c >= HD / 32;
      const float mul = is_k ? p.scale : 1.0f;
      bf16* op = is_k ? (ok_ + (c - HD / 32) * 32) : (ov + c * 32);
      if (row_ok) {
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          uint4 u;
          u.x = f2_to_bf2(__uint_as_float(v[i]) * mul, __uint_as_float(v[i + 1]) * mul);
          u.y = f2_to_bf2(__uint_as_float(v[i + 2]) * mul, __uint_as_float(v[i + 3]) * mul);
          u.z = f2_to_bf2(__uint_as_float(v[i + 4]) * mul, __uint_as_float(v[i + 5]) * mul);
          u.w = f2_to_bf2(__uint_as_float(v[i + 6]) * mul, __uint_as_float(v[i + 7]) * mul);
          *reinterpret_cast<uint4*>(op + i) = u;
        }
      }
    }
    tc_fence_before();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, SM::TMEM_COLS);
}

template <int HD>
static int launch_attn_bwd(const CUtensorMap& tm, const CUtensorMap& tm64, const CUtensorMap& tmdo,
                           const CUtensorMap& tmdo64, const AttnBwdParams& p, int B, cudaStream_t st) {
  auto kq = attn_bwd_dq_kernel<HD>;
  auto kkv = attn_bwd_dkv_kernel<HD>;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(kq, cudaFuncAttributeMaxDynamicSharedMemorySize, BwdQSmem<HD>::TOTAL) != cudaSuccess)
      return OMNI_ERR_CUDA;
    if (cudaFuncSetAttribute(kkv, cudaFuncAttributeMaxDynamicSharedMemorySize, BwdKVSmem<HD>::TOTAL) != cudaSuccess)
      return OMNI_ERR_CUDA;
    attr_set = true;
  }
  const int nt = ceil_div(p.S, AB_T);
  kkv<<<dim3(p.n_kv_heads, B, nt), AB_THREADS, BwdKVSmem<HD>::TOTAL, st>>>(tm, tm64, tmdo64, p);
  OMNI_LAUNCH_CHECK();
  kq<<<dim3(p.n_heads, B, nt), AB_THREADS, BwdQSmem<HD>::TOTAL, st>>>(tm, tm64, tmdo, p);
  OMNI_LAUNCH_CHECK();
  return OMNI_OK;
}

}  // namespace omni

extern "C" int64_t omni_attention_bwd_scratch_floats(int32_t B, int32_t S, int32_t n_heads) {
  if (B <= 0 || S <= 0 || n_heads <= 0) return 0;
  return static_cast<int64_t>(n_heads) * B * ((S + omni::AB_STEP - 1) / omni::AB_STEP) * 2 * omni::AB_STEP;
}

extern "C" int omni_attention_bwd(const void* qkv, int64_t M, int64_t ld, const void* out, int64_t out_ld,
                                  const void* dout, int64_t dout_ld, const float* lse, float* delta, void* dqkv,
                                  int64_t dqkv_ld, int32_t row0, int32_t B, int32_t S, int32_t n_heads,
                                  int32_t n_kv_heads, int32_t head_dim, int32_t causal, float scale, void* stream) {
  using namespace omni;
  OMNI_CHECK_ARG(qkv && out && dout && lse && delta && dqkv && M > 0 && B > 0 && S > 0 && n_heads > 0 && n_kv_heads > 0);
  OMNI_CHECK_ARG(B <= 65535);   // clip index rides in gridDim.y
  OMNI_CHECK_ARG(n_heads % n_kv_heads == 0 && row0 >= 0 && static_cast<int64_t>(row0) + static_cast<int64_t>(B) * S <= M);
  const int64_t width = static_cast<int64_t>(n_heads + 2 * n_kv_heads) * head_dim;
  OMNI_CHECK_ARG((ld % 8) == 0 && (out_ld % 8) == 0 && (dout_ld % 8) == 0 && (dqkv_ld % 8) == 0 && ld >= width &&
                 dqkv_ld >= width);
  if (head_dim != 64 && head_dim != 128) return OMNI_ERR_UNSUPPORTED;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  CUtensorMap tm, tm64, tmdo, tmdo64;
  int rc = omni_make_tmap_2d_bf16(&tm, qkv, (uint64_t)M, (uint64_t)width, (uint64_t)ld, AB_T, 64, 1);
  if (rc) return rc;
  rc = omni_make_tmap_2d_bf16(&tm64, qkv, (uint64_t)M, (uint64_t)width, (uint64_t)ld, AB_STEP, 64, 1);
  if (rc) return rc;
  rc = omni_make_tmap_2d_bf16(&tmdo, dout, (uint64_t)M, (uint64_t)n_heads * head_dim, (uint64_t)dout_ld, AB_T, 64, 1);
  if (rc) return rc;
  rc = omni_make_tmap_2d_bf16(&tmdo64, dout, (uint64_t)M, (uint64_t)n_heads * head_dim, (uint64_t)dout_ld, AB_STEP, 64, 1);
  if (rc) return rc;
  const int n_qs = ceil_div(S, AB_STEP);
  const long long rows = static_cast<long long>(B) * n_qs * AB_STEP;
  OMNI_CHECK_ARG((reinterpret_cast<uintptr_t>(delta) & 15) == 0 && rows / 8 < 0x7fffffffLL);
  attn_delta_kernel<<<static_cast<unsigned>(ceil_div_ll(rows, 8)), 256, 0, st>>>(
      reinterpret_cast<const bf16*>(dout), dout_ld, reinterpret_cast<const bf16*>(out), out_ld, lse, delta, M, row0, B, S,
      n_qs, n_heads, head_dim);
  OMNI_LAUNCH_CHECK();
  AttnBwdParams p;
  p.dqkv = reinterpret_cast<bf16*>(dqkv);
  p.dqkv_ld = dqkv_ld;
  p.stats = delta;
  p.M = M;
  p.B = B; p.n_qs = n_qs;
  p.row0 = row0; p.S = S; p.n_heads = n_heads; p.n_kv_heads = n_kv_heads; p.causal = causal ? 1 : 0;
  p.scale = scale;
  p.scale_log2 = scale * 1.4426950408889634f;
  return head_dim == 64 ? launch_attn_bwd<64>(tm, tm64, tmdo, tmdo64, p, B, st)
                        : launch_attn_bwd<128>(tm, tm64, tmdo, tmdo64, p, B, st);
}
