// Weight-streaming GEMM of the decode step (M <= 128 token rows): out[t, n] = epi(alpha * sum_k x[t, k] * W[n, k]).
//
// A decode step is HBM-bound: every weight byte is read once per step for a handful of token rows (reference: the HF
// generate loop over LlamaDecoderLayer_lora, Llama_LoRA.py:580-655, one token per sequence and step).  Measured on B200,
// one SM pulls at most ~45 GB/s of HBM-missing TMA traffic, so a weight matrix only streams at HBM speed when (a) ~all SMs
// work on it and (b) each keeps ~100 KB of WEIGHT bytes in flight.  The general GEMM kernels (gemm_tcgen05.cu) put the
// tokens on the 128-row M side of the MMA: two thirds of every pipeline stage are re-fetched activations, a 2048-wide
// output yields 32 CTAs, and the step's GEMMs ran at 0.4-2.5 TB/s.  This kernel swaps the operands:
//   * the WEIGHT tile is the M operand of tcgen05.mma (128 output features x 64 k, 16 KB per stage, all of it payload), the
//     tokens are the N operand (64 or 128 rows, 8 / 16 KB per stage, L2-resident), accumulator in TMEM = [feature lane,
//     token column];
//   * split-K over SPLIT co-resident CTAs (same 128 features, ~K / SPLIT each) brings every GEMM of the step to 120-148
//     CTAs.  The fp32 partials are exchanged through an L2-resident workspace: every rank stores its [token][feature] tile,
//     releases a per-tile counter, waits until all SPLIT ranks have arrived (the grid never exceeds the SM count, so all
//     CTAs are resident and the spin cannot deadlock), and then finishes the token slice [s * TSL, (s + 1) * TSL) itself --
//     summed in rank order (deterministic), every rank running 1 / SPLIT of the epilogue.  (A hardware cluster with a DSMEM
//     exchange was measured first: clusters of 8 CTAs with ~215 KB of shared memory each cost ~12 us of scheduling per
//     launch inside the back-to-back decode graph, clusters of 4 leave half the SMs idle on the 2048-wide outputs.)
//   * epilogue: alpha, bias, rounding points of the unfused bf16 op sequence, residual add, or the SwiGLU of LlamaMLP on
//     [gate 64 | up 64] interleaved weight rows; the [token, feature] transposition goes through a shared-memory tile so
//     that global stores are full 128 / 256-byte rows;
//   * the Omni-LoRA up-projection rides as extra K blocks (x2 = s * h A^T, W2 = B rows of the task's adapter) on the last
//     rank, the adapter rows picked by task id (tile_group[0]) exactly as in omni_gemm_bf16.
#include <stdlib.h>
#include <string.h>
#include "gemm_epilogue.cuh"

namespace omni {

constexpr int SK_THREADS = 192;     // warp 0 TMA, warp 1 MMA (+ TMEM alloc), warps 2..5 epilogue
constexpr int SK_FEATS = 128;       // output features per CTA (the M of the MMA)
constexpr int SK_WS_COUNTER_BYTES = 4096;  // [<= 512 tiles][2] ints at the start of the split-K workspace
constexpr int SK_TPITCH = SK_FEATS + 8;   // bf16 pitch of the [token][feature] transposition tile (conflict-free 16 B reads)

struct SkinnyParams {
  int M, N, K;
  int kb_total;             // 64-wide k-blocks of the reduction (rank r takes [kb * r / SPLIT, kb * (r + 1) / SPLIT))
  float* ws_part;           // [n_tiles][SPLIT][NTOK][128] fp32 partial tiles (SPLIT > 1)
  int* ws_cnt;              // [n_tiles][2] arrive / done counters, zero between launches (self-resetting)
  int n_ext;                // K-extension slots per (group, tile)
  const int* tile_group;    // [1] task id of the token tile or null
  const int* b_row_table;   // [groups][ceil(N / 64)] row of W per 64-feature block, or null (row = feature)
  const int4* ext_table;    // [groups][n_tiles][n_ext] {x2 column, W2 row, W2 column, -}; W2 row < 0: skip
  int n_tiles;              // ceil(N / 128)
  int n_blocks64;           // ceil(N / 64)
  const bf16* bias;
  const bf16* residual;
  bf16* out;
  long long ldo, ldr;
  int act;                  // OMNI_ACT_NONE / OMNI_ACT_RELU / OMNI_ACT_GELU / OMNI_ACT_SWIGLU64
  float alpha;
  bf16* out2;               // SWIGLU64: [M, N / 2]
  long long ldo2;
  // fused RMSNorm of the finished rows (split-K launches only): see omni_gemm_args.norm_out
  const bf16* norm_w;
  bf16* norm_out;
  long long norm_ld;
  float norm_eps;
  int* norm_cnt;            // [SPLIT] arrivals per token slice, zero between launches (self-resetting)
};

template <int NTOK, int SPLIT, int STAGES>
struct SkinnySmem {
  static constexpr int W_BYTES = SK_FEATS * BK * 2;        // 16 KB
  static constexpr int X_BYTES = NTOK * BK * 2;            // 8 / 16 KB
  static constexpr int STAGE_BYTES = W_BYTES + X_BYTES;
  static constexpr int TSL = ((NTOK + SPLIT - 1) / SPLIT + 3) / 4 * 4;   // tokens finished by one rank (multiple of 4)
  // the [token][feature] transposition tile of the epilogue ALIASES the pipeline stages: it is written only after the
  // accumulator barrier, i.e. when every MMA has consumed its stage and no TMA load is pending.  (Measured: shrinking the
  // pipeline to 4 stages so that the CTAs of two consecutive GEMMs are co-resident -- the next GEMM's weight prefetch
  // would then run under this GEMM's split-K exchange -- made the step 5 % SLOWER: 64 KB in flight per CTA no longer
  // covers the HBM latency at ~40 GB/s per SM.  One CTA per SM with 8 stages it is.)
  static constexpr int T_OFFSET = 0;
  static constexpr int T_BYTES = TSL * SK_TPITCH * 2;
  static_assert(T_BYTES <= STAGES * STAGE_BYTES, "transposition tile must fit in the stage area");
  static constexpr int BAR_OFFSET = STAGES * STAGE_BYTES;
  static constexpr int EXT_OFFSET = BAR_OFFSET + ((2 * STAGES + 1) * 8 + 16 + 15) / 16 * 16;
  static constexpr int TOTAL = EXT_OFFSET + 32 * 16 + 1024;
};

template <int NTOK, int SPLIT, int STAGES>
__global__ void __launch_bounds__(SK_THREADS, 1)
gemm_skinny_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmX,
                   const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmX2, const SkinnyParams p) {
  using S = SkinnySmem<NTOK, SPLIT, STAGES>;
  constexpr int TSL = S::TSL;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::BAR_OFFSET);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);
  int4* ext_smem = reinterpret_cast<int4*>(smem + S::EXT_OFFSET);       // [32] the tile's K-extension entries
  bf16* ttile = reinterpret_cast<bf16*>(smem + S::T_OFFSET);            // [TSL][SK_TPITCH]

  pdl_launch_dependents();      // the next kernel of the step may start its own prologue / weight prefetch right away
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_tile = blockIdx.x / SPLIT;
  const int rank = blockIdx.x - n_tile * SPLIT;
  const int kb0 = (p.kb_total * rank) / SPLIT;
  const int kb_mine = (p.kb_total * (rank + 1)) / SPLIT - kb0;
  const int n0 = n_tile * SK_FEATS;
  const int group = p.tile_group ? p.tile_group[0] : 0;
  const int4* ext = p.ext_table ? p.ext_table + static_cast<long long>(group * p.n_tiles + n_tile) * p.n_ext : nullptr;
  // K-extension list of the tile (static data: read before the predecessor finishes).  One PARALLEL load per warp -- lane j
  // fetches entry j -- instead of a dependent global load per entry in the producer loop: on the last rank those ~0.7 us
  // round trips sat in front of the split-K exchange every other rank of the tile waits in (qkv: 20 us -> see profiles/).
  int4 ext_mine = make_int4(0, -1, 0, 0);
  if (ext && rank == SPLIT - 1 && lane < p.n_ext) ext_mine = ext[lane];
  const unsigned ext_mask = __ballot_sync(0xffffffffu, ext_mine.y >= 0);
  const int n_ext_valid = __popc(ext_mask);
  const int total_iters = kb_mine + n_ext_valid;

  if (warp == 0) ext_smem[lane] = ext_mine;      // visible to the producer lane after the __syncthreads below
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmW);
    tma_prefetch_desc(&tmX);
    if (ext) {
      tma_prefetch_desc(&tmW2);
      tma_prefetch_desc(&tmX2);
    }
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < STAGES; ++s) {
        mbar_init(&full_bar[s], 1);
        mbar_init(&empty_bar[s], 1);
      }
      mbar_init(tmem_full_bar, 1);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(tmem_ptr_smem, NTOK);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ===== TMA producer: weight boxes first (the HBM stream), then the L2-resident token rows =====
    if (elect_one()) {
      int row_a = n0, row_b = n0 + 64;
      if (p.b_row_table) {
        const int* tb = p.b_row_table + group * p.n_blocks64 + 2 * n_tile;
        row_a = tb[0];
        row_b = (2 * n_tile + 1 < p.n_blocks64) ? tb[1] : (1 << 30);       // beyond the tensor: zero-filled
      }
      int stage = 0;
      uint32_t phase = 0;
      // programmatic dependent launch: the WEIGHT boxes of the first pipeline fill (up to STAGES x 16 KB per CTA) go out
      // before the predecessor has finished -- weights are static -- and only the token rows wait for it
      const int pre = kb_mine < STAGES ? kb_mine : STAGES;
      for (int it = 0; it < pre; ++it) {
        uint8_t* sW = smem + it * S::STAGE_BYTES;
        mbar_expect_tx(&full_bar[it], S::STAGE_BYTES);
        tma_load_2d(&tmW, &full_bar[it], sW, (kb0 + it) * BK, row_a);
        tma_load_2d(&tmW, &full_bar[it], sW + S::W_BYTES / 2, (kb0 + it) * BK, row_b);
      }
      pdl_wait();
      for (int it = 0; it < pre; ++it)
        tma_load_2d(&tmX, &full_bar[it], smem + it * S::STAGE_BYTES + S::W_BYTES, (kb0 + it) * BK, 0);
      stage = pre == STAGES ? 0 : pre;
      phase = pre == STAGES ? 1 : 0;
      for (int it = pre; it < kb_mine; ++it) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sW = smem + stage * S::STAGE_BYTES;
        uint8_t* sX = sW + S::W_BYTES;
        mbar_expect_tx(&full_bar[stage], S::STAGE_BYTES);
        tma_load_2d(&tmW, &full_bar[stage], sW, (kb0 + it) * BK, row_a);
        tma_load_2d(&tmW, &full_bar[stage], sW + S::W_BYTES / 2, (kb0 + it) * BK, row_b);
        tma_load_2d(&tmX, &full_bar[stage], sX, (kb0 + it) * BK, 0);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      if (n_ext_valid) {
        for (int j = 0; j < p.n_ext; ++j) {
          const int4 e = ext_smem[j];
          if (e.y < 0) continue;
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sW = smem + stage * S::STAGE_BYTES;
          uint8_t* sX = sW + S::W_BYTES;
          mbar_expect_tx(&full_bar[stage], S::STAGE_BYTES);
          tma_load_2d(&tmW2, &full_bar[stage], sW, e.z, e.y);
          tma_load_2d(&tmW2, &full_bar[stage], sW + S::W_BYTES / 2, e.z, e.y + 64);
          tma_load_2d(&tmX2, &full_bar[stage], sX, e.x, 0);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: D[feature, token] += W[feature, k] * x[token, k] =====
    constexpr uint32_t idesc = make_idesc_bf16(SK_FEATS, NTOK, 0, 0);
    int stage = 0;
    uint32_t phase = 0;
    for (int it = 0; it < total_iters; ++it) {
      mbar_wait(&full_bar[stage], phase);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t sW = smem_u32(smem + stage * S::STAGE_BYTES);
        const uint32_t sX = sW + S::W_BYTES;
        const uint64_t adesc = make_smem_desc_sw128(sW, 16, 1024);
        const uint64_t bdesc = make_smem_desc_sw128(sX, 16, 1024);
#pragma unroll
        for (int k = 0; k < BK / UMMA_K; ++k)
          umma_bf16(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc, (it > 0 || k > 0) ? 1u : 0u);
        umma_commit(&empty_bar[stage]);
        if (it == total_iters - 1) umma_commit(tmem_full_bar);
      }
      __syncwarp();
      if (++stage == STAGES) { stage = 0; phase ^= 1; }
    }
  }

  // ===== epilogue part 1 (warps 2..5): TMEM -> (split-K: partial tile -> workspace -> this rank's token slice) =====
  const int q = warp & 3;
  const int f = q * 32 + lane;                 // feature row of this thread inside the tile (epilogue warps only)
  float v[TSL];
  if (warp >= 2) {
    pdl_wait();                                // residual / workspace / output buffers belong to the predecessor until here
    if (total_iters > 0) {
      mbar_wait(tmem_full_bar, 0);
      tc_fence_after();
    }
    if constexpr (SPLIT > 1) {
      float* mine = p.ws_part + (static_cast<long long>(n_tile) * SPLIT + rank) * NTOK * SK_FEATS + f;
#pragma unroll
      for (int c = 0; c < NTOK / 32; ++c) {
        uint32_t r[32];
        if (total_iters > 0) {
          tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(c * 32), r);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) r[i] = 0u;
        }
#pragma unroll
        for (int i = 0; i < 32; ++i)            // one token per instruction: 32 lanes = 128 contiguous bytes
          __stcg(mine + (c * 32 + i) * SK_FEATS, __uint_as_float(r[i]));
      }
      tc_fence_before();
      // release this rank's partial tile, wait for the other ranks of the tile
      asm volatile("bar.sync 1, 128;" ::: "memory");
      int* cnt = p.ws_cnt + 2 * n_tile;
      if (warp == 2 && lane == 0) {
        __threadfence();
        atomicAdd(cnt, 1);
        uint32_t spins = 0;
        while (true) {
          int seen;
          asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(seen) : "l"(cnt) : "memory");
          if (seen >= SPLIT) break;
          if (++spins > (1u << 26)) __trap();
        }
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      // this rank finishes tokens [rank * TSL, rank * TSL + TSL): sum the partials in rank order (deterministic)
#pragma unroll
      for (int t = 0; t < TSL; ++t) v[t] = 0.f;
      const float* base = p.ws_part + static_cast<long long>(n_tile) * SPLIT * NTOK * SK_FEATS + f;
#pragma unroll
      for (int src = 0; src < SPLIT; ++src) {
#pragma unroll
        for (int t = 0; t < TSL; ++t) {
          const int tok = rank * TSL + t;
          if (tok < NTOK) v[t] += __ldcg(base + (static_cast<long long>(src) * NTOK + tok) * SK_FEATS);
        }
      }
      // every rank has read what it needs before the counters are cleared for the next launch
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (warp == 2 && lane == 0) {
        if (atomicAdd(cnt + 1, 1) == SPLIT - 1) {
          cnt[0] = 0;
          cnt[1] = 0;
        }
      }
    } else {
#pragma unroll
      for (int c = 0; c < NTOK / 32; ++c) {
        uint32_t r[32];
        tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(c * 32), r);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) v[c * 32 + i] = __uint_as_float(r[i]);
      }
      tc_fence_before();
    }
  }

  // ===== epilogue part 2: alpha, bias, rounding -> [token][feature] tile -> coalesced global rows =====
  if (warp >= 2) {
    const int feat = n0 + f;
    const float bias = (p.bias && feat < p.N) ? __bfloat162float(p.bias[feat]) : 0.f;
    const bool round_first = p.act != OMNI_ACT_NONE || p.residual != nullptr;
#pragma unroll
    for (int t = 0; t < TSL; ++t) {
      float x = v[t] * p.alpha + bias;
      if (round_first) x = __bfloat162float(__float2bfloat16_rn(x));     // the linear's bf16 output feeds act / residual
      if (p.act == OMNI_ACT_RELU) x = fmaxf(x, 0.f);
      else if (p.act == OMNI_ACT_GELU) x = gelu_fast(x);
      ttile[t * SK_TPITCH + f] = __float2bfloat16_rn(x);
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");
    const int e = (warp - 2) * 32 + lane;
    const int tok0 = rank * TSL;                   // (tokens >= M are never stored: tg < p.M below, and M <= NTOK)
    if (p.act == OMNI_ACT_SWIGLU64) {
      // tile rows = [gate 64 | up 64] of 64 intermediate channels: act = bf16(bf16(silu(gate)) * up), 128 B per token
      const int chunk = e & 7;
      for (int t = e >> 3; t < TSL; t += 16) {
        const int tg = tok0 + t;
        if (tg >= p.M) continue;
        const uint4 g4 = *reinterpret_cast<const uint4*>(ttile + t * SK_TPITCH + chunk * 8);
        const uint4 u4 = *reinterpret_cast<const uint4*>(ttile + t * SK_TPITCH + 64 + chunk * 8);
        const uint32_t gs[4] = {g4.x, g4.y, g4.z, g4.w};
        const uint32_t us[4] = {u4.x, u4.y, u4.z, u4.w};
        uint32_t o[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 g = bf2_to_f2(gs[i]);
          const float2 u = bf2_to_f2(us[i]);
          o[i] = f2_to_bf2(__bfloat162float(__float2bfloat16_rn(silu(g.x))) * u.x,
                           __bfloat162float(__float2bfloat16_rn(silu(g.y))) * u.y);
        }
        const int col = n_tile * 64 + chunk * 8;
        if (col + 8 <= p.N / 2)
          *reinterpret_cast<uint4*>(p.out2 + static_cast<long long>(tg) * p.ldo2 + col) = make_uint4(o[0], o[1], o[2], o[3]);
      }
      if (p.out) {                                   // the gate | up tile itself (not needed by the decode step)
        const int chunk16 = e & 15;
        for (int t = e >> 4; t < TSL; t += 8) {
          const int tg = tok0 + t;
          const int col = n0 + chunk16 * 8;
          if (tg < p.M && col + 8 <= p.N)
            *reinterpret_cast<uint4*>(p.out + static_cast<long long>(tg) * p.ldo + col) =
                *reinterpret_cast<const uint4*>(ttile + t * SK_TPITCH + chunk16 * 8);
        }
      }
    } else {
      const int chunk = e & 15;                      // 16 lanes x 16 B = the tile's 128 features of one token
      for (int t = e >> 4; t < TSL; t += 8) {
        const int tg = tok0 + t;
        if (tg >= p.M) continue;
        const int col = n0 + chunk * 8;
        if (col >= p.N) continue;
        uint4 val = *reinterpret_cast<const uint4*>(ttile + t * SK_TPITCH + chunk * 8);
        bf16* op = p.out + static_cast<long long>(tg) * p.ldo + col;
        if (col + 8 <= p.N) {
          if (p.residual) {
            const uint4 r4 = ld_nc_u4(p.residual + static_cast<long long>(tg) * p.ldr + col);
            float2 a, b;
            a = bf2_to_f2(val.x); b = bf2_to_f2(r4.x); val.x = f2_to_bf2(a.x + b.x, a.y + b.y);
            a = bf2_to_f2(val.y); b = bf2_to_f2(r4.y); val.y = f2_to_bf2(a.x + b.x, a.y + b.y);
            a = bf2_to_f2(val.z); b = bf2_to_f2(r4.z); val.z = f2_to_bf2(a.x + b.x, a.y + b.y);
            a = bf2_to_f2(val.w); b = bf2_to_f2(r4.w); val.w = f2_to_bf2(a.x + b.x, a.y + b.y);
          }
          *reinterpret_cast<uint4*>(op) = val;
        } else {
          const bf16* tv = reinterpret_cast<const bf16*>(&val);
          for (int i = 0; i < 8 && col + i < p.N; ++i) {
            float x = __bfloat162float(tv[i]);
            if (p.residual) x += __bfloat162float(p.residual[static_cast<long long>(tg) * p.ldr + col + i]);
            op[i] = __float2bfloat16_rn(x);
          }
        }
      }
      if constexpr (SPLIT > 1) {
        if (p.norm_out) {
          // Fused RMSNorm: this CTA has written features [n0, n0 + 128) of the token slice [rank * TSL, rank * TSL + TSL).  The
          // CTA that completes the slice's LAST feature tile normalises its rows (re-read from L2, where every CTA's stores
          // landed) -- one warp per row, chunks lane, lane + 32, ... and the butterfly sum of rmsnorm_fwd_kernel: same bits.
          __shared__ int s_last;
          asm volatile("bar.sync 1, 128;" ::: "memory");
          if (warp == 2 && lane == 0) {
            __threadfence();
            const int old = atomicAdd(p.norm_cnt + rank, 1);
            s_last = (old == p.n_tiles - 1) ? 1 : 0;
            if (s_last) p.norm_cnt[rank] = 0;          // every CTA of the slice has arrived: ready for the next launch
          }
          asm volatile("bar.sync 1, 128;" ::: "memory");
          if (s_last) {
            __threadfence();
            const int H8 = p.N >> 3;
            for (int t = warp - 2; t < TSL; t += 4) {
              const int tg = tok0 + t;
              if (tg >= p.M) continue;
              const uint4* xr = reinterpret_cast<const uint4*>(p.out + static_cast<long long>(tg) * p.ldo);
              float ss = 0.f;
              for (int c = lane; c < H8; c += 32) {
                const uint4 u = __ldcg(xr + c);
                const uint32_t w4[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const float2 f = bf2_to_f2(w4[i]);
                  ss += f.x * f.x;
                  ss += f.y * f.y;
                }
              }
              ss = warp_sum(ss);
              const float rstd = rsqrtf(ss / static_cast<float>(H8 * 8) + p.norm_eps);
              uint4* yr = reinterpret_cast<uint4*>(p.norm_out + static_cast<long long>(tg) * p.norm_ld);
              for (int c = lane; c < H8; c += 32) {
                const uint4 u = __ldcg(xr + c);
                const uint4 g = __ldg(reinterpret_cast<const uint4*>(p.norm_w) + c);
                const uint32_t x4[4] = {u.x, u.y, u.z, u.w}, g4[4] = {g.x, g.y, g.z, g.w};
                uint32_t o4[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const float2 f = bf2_to_f2(x4[i]), gw = bf2_to_f2(g4[i]);
                  o4[i] = f2_to_bf2(gw.x * __bfloat162float(__float2bfloat16_rn(f.x * rstd)),
                                    gw.y * __bfloat162float(__float2bfloat16_rn(f.y * rstd)));
                }
                yr[c] = make_uint4(o4[0], o4[1], o4[2], o4[3]);
              }
            }
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, NTOK);
}

static int skinny_sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0, v = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return kNumSMs;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) return kNumSMs;
    n = v;
  }
  return n;
}

template <int NTOK, int SPLIT, int STAGES>
static int launch_skinny(const CUtensorMap& tmW, const CUtensorMap& tmX, const CUtensorMap& tmW2, const CUtensorMap& tmX2,
                         const SkinnyParams& p, cudaStream_t st) {
  using S = SkinnySmem<NTOK, SPLIT, STAGES>;
  static_assert(S::TOTAL <= 232448, "shared memory budget");
  auto kfn = gemm_skinny_kernel<NTOK, SPLIT, STAGES>;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL) != cudaSuccess) return OMNI_ERR_CUDA;
    attr_set = true;
  }
  if (omni_launch_pdl(kfn, dim3(p.n_tiles * SPLIT), dim3(SK_THREADS), S::TOTAL, st, tmW, tmX, tmW2, tmX2, p) != cudaSuccess)
    return OMNI_ERR_CUDA;
  return OMNI_OK;
}

}  // namespace omni

// Entry point shared with omni_gemm_bf16: same argument block, A = token rows [M <= 128, K], B = weights.
extern "C" int omni_gemm_skinny_bf16(const omni_gemm_args* a, void* stream) {
  using namespace omni;
  OMNI_CHECK_ARG(a != nullptr);
  OMNI_CHECK_ARG(a->M > 0 && a->M <= 128 && a->N > 0 && a->K > 0 && (a->K % BK) == 0);
  OMNI_CHECK_ARG(a->A && a->B && (a->out || a->act == OMNI_ACT_SWIGLU64));
  OMNI_CHECK_ARG((a->lda % 8) == 0 && (a->ldb % 8) == 0);
  OMNI_CHECK_ARG((reinterpret_cast<uintptr_t>(a->A) & 15) == 0 && (reinterpret_cast<uintptr_t>(a->B) & 15) == 0);
  OMNI_CHECK_ARG(!a->out || ((a->ldo % 8) == 0 && a->ldo >= a->N && (reinterpret_cast<uintptr_t>(a->out) & 15) == 0));
  if (a->residual) OMNI_CHECK_ARG((a->ldr % 8) == 0 && (reinterpret_cast<uintptr_t>(a->residual) & 15) == 0);
  if (a->out_fp32 || a->act == OMNI_ACT_GELU_KEEP) return OMNI_ERR_UNSUPPORTED;
  if (a->act == OMNI_ACT_SWIGLU64) {
    OMNI_CHECK_ARG(a->out2 && (a->ldo2 % 8) == 0 && a->ldo2 >= a->N / 2 && (reinterpret_cast<uintptr_t>(a->out2) & 15) == 0);
    if ((a->N % 128) != 0 || a->residual || a->bias || a->ext_table || a->b_row_table) return OMNI_ERR_UNSUPPORTED;
  }
  if (a->ext_table) {
    OMNI_CHECK_ARG(a->A2 && a->B2 && a->n_ext > 0 && a->n_ext <= 32 && (a->lda2 % 8) == 0 && (a->ldb2 % 8) == 0);
    OMNI_CHECK_ARG(a->block_n == 128);           // the table is indexed with this kernel's 128-feature tiles
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int ntok = a->M <= 64 ? 64 : 128;

  SkinnyParams p;
  memset(&p, 0, sizeof(p));
  p.M = a->M; p.N = a->N; p.K = a->K;
  p.n_tiles = ceil_div(a->N, SK_FEATS);
  p.n_blocks64 = ceil_div(a->N, 64);
  p.n_ext = a->ext_table ? a->n_ext : 0;
  p.tile_group = a->tile_group;
  p.b_row_table = a->b_row_table;
  p.ext_table = reinterpret_cast<const int4*>(a->ext_table);
  p.bias = reinterpret_cast<const bf16*>(a->bias);
  p.residual = reinterpret_cast<const bf16*>(a->residual);
  p.out = reinterpret_cast<bf16*>(a->out);
  p.ldo = a->ldo; p.ldr = a->ldr;
  p.act = a->act; p.alpha = a->alpha;
  p.out2 = reinterpret_cast<bf16*>(a->out2); p.ldo2 = a->ldo2;
  p.norm_w = reinterpret_cast<const bf16*>(a->norm_weight);
  p.norm_out = reinterpret_cast<bf16*>(a->norm_out);
  p.norm_ld = a->norm_ld; p.norm_eps = a->norm_eps;
  if (a->norm_out) {
    OMNI_CHECK_ARG(a->norm_weight && a->out && (a->N % 8) == 0 && (a->norm_ld % 8) == 0 && a->norm_ld >= a->N &&
                   (reinterpret_cast<uintptr_t>(a->norm_out) & 15) == 0 && (reinterpret_cast<uintptr_t>(a->norm_weight) & 15) == 0);
    if (a->act != OMNI_ACT_NONE || !a->workspace) return OMNI_ERR_UNSUPPORTED;
  }

  // split K until ~all SMs stream weights; the ranks of a tile wait for each other, so the whole grid must be resident:
  // n_tiles * split <= SM count (one CTA per SM at this shared-memory footprint)
  const int kb = a->K / BK;
  const int sms = skinny_sm_count();
  // Powers of two only.  Measured in isolation (tools/skinny_probe.py, 64 tokens, L2 flushed, ~5 us of event overhead
  // included): qkv 64 x 3072 x 2048 -- split 4: 15.4 us, 5: 25.6, 6: 17.4, 3: 19.5; o_proj 2048 x 2048 -- split 8: 13.3, 6: 23.6,
  // 4: 13.3; Qwen2.5-3B qkv 2560 x 2048 -- split 4: 15.3, 6: 23.6.  The uneven K slices of 3 / 5 / 6 cost more than the extra
  // CTAs bring.
  static const int choices[] = {8, 4, 2};
  // measured inside the decode graph: 144 CTAs (24 tiles x 6) ran 2x slower than 96 (x 4) -- near the SM count the last
  // CTAs of the grid only become resident when the predecessor's stragglers have left, and their peers spin meanwhile
  const int cta_cap = sms < 132 ? sms : 132;
  int split = 1;
  static const char* force = getenv("OMNI_SKINNY_SPLIT");
  if (force) {
    split = atoi(force);
    if (split != 1 && (p.n_tiles * split > cta_cap || split > kb)) split = 1;
  } else if (a->workspace) {
    for (int c : choices) {
      if (p.n_tiles * c <= cta_cap && kb >= 2 * c) { split = c; break; }
    }
  }
  p.kb_total = kb;
  if (split > 1) {
    const int64_t part = static_cast<int64_t>(p.n_tiles) * split * ntok * SK_FEATS * 4;
    // fixed-size counter region in front (the partial tiles of one launch must never land on another launch's counters)
    const int64_t cnt = SK_WS_COUNTER_BYTES;
    if (p.n_tiles * 8 > cnt) return OMNI_ERR_UNSUPPORTED;
    if (!a->workspace || a->workspace_bytes < part + cnt || (reinterpret_cast<uintptr_t>(a->workspace) & 255) != 0)
      return OMNI_ERR_WORKSPACE;
    p.ws_cnt = reinterpret_cast<int*>(a->workspace);                          // counters first (they stay zero between launches)
    p.ws_part = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(a->workspace) + cnt);
    // the fused norm's per-slice arrival counters sit in the last 32 bytes of the counter region
    if (a->norm_out && p.n_tiles * 8 > cnt - 32) return OMNI_ERR_UNSUPPORTED;
    p.norm_cnt = reinterpret_cast<int*>(reinterpret_cast<uint8_t*>(a->workspace) + cnt - 32);
  } else if (a->norm_out) {
    return OMNI_ERR_UNSUPPORTED;               // no split-K: one CTA would have to normalise every row
  }

  CUtensorMap tmW, tmX, tmW2, tmX2;
  int rc = omni_make_tmap_2d_bf16(&tmW, a->B, (uint64_t)a->b_rows, (uint64_t)a->K, (uint64_t)a->ldb, 64, BK, 1);
  if (rc) return rc;
  rc = omni_make_tmap_2d_bf16(&tmX, a->A, (uint64_t)a->M, (uint64_t)a->K, (uint64_t)a->lda, ntok, BK, 1);
  if (rc) return rc;
  if (a->ext_table) {
    rc = omni_make_tmap_2d_bf16(&tmW2, a->B2, (uint64_t)a->b2_rows, (uint64_t)a->b2_cols, (uint64_t)a->ldb2, 64, BK, 1);
    if (rc) return rc;
    rc = omni_make_tmap_2d_bf16(&tmX2, a->A2, (uint64_t)a->M, (uint64_t)a->a2_cols, (uint64_t)a->lda2, ntok, BK, 1);
    if (rc) return rc;
  } else {
    tmW2 = tmW;
    tmX2 = tmX;
  }
#define OMNI_SK(NT, SP, STG) return launch_skinny<NT, SP, STG>(tmW, tmX, tmW2, tmX2, p, st);
  if (ntok == 64) {
    switch (split) {
      case 1: OMNI_SK(64, 1, 8)
      case 2: OMNI_SK(64, 2, 8)
      case 3: OMNI_SK(64, 3, 8)
      case 4: OMNI_SK(64, 4, 8)
      case 5: OMNI_SK(64, 5, 8)
      case 6: OMNI_SK(64, 6, 8)
      case 8: OMNI_SK(64, 8, 8)
    }
  } else {
    switch (split) {
      case 1: OMNI_SK(128, 1, 6)
      case 2: OMNI_SK(128, 2, 6)
      case 3: OMNI_SK(128, 3, 6)
      case 4: OMNI_SK(128, 4, 6)
      case 5: OMNI_SK(128, 5, 6)
      case 6: OMNI_SK(128, 6, 6)
      case 8: OMNI_SK(128, 8, 6)
    }
  }
#undef OMNI_SK
  return OMNI_ERR_BAD_ARG;
}

extern "C" int64_t omni_gemm_skinny_workspace_bytes(void) {
  const int64_t sms = omni::skinny_sm_count();
  return sms * 128 * omni::SK_FEATS * 4 + omni::SK_WS_COUNTER_BYTES;
}
