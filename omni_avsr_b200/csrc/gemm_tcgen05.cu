// bf16 GEMM on the sm_100a tensor cores: tcgen05.mma with TMEM accumulators, TMA-fed 128B-swizzled
// shared-memory pipeline, warp-specialised (TMA producer / MMA issuer / 4 epilogue warps).
//
//   out[M,N] = epilogue( alpha * ( A[M,K] . B[rowsel(N),K]^T  +  sum_j A2[M, a2_col_j : +64] . B2[b2_row_j.., b2_col_j : +64]^T ) )
//
// * A, B, A2, B2 are row-major with the reduction dimension contiguous ("K-major"), i.e. exactly the
//   nn.Linear layout (x [tokens, in], W [out, in]).
// * tile_group[m_tile] (optional) is the task id of a 128-row tile of packed tokens; b_row_table /
//   ext_table are indexed by (group, n_tile) and give the row of B (grouped weights) and the K-extension
//   blocks (LoRA up-projection riding in the same TMEM accumulator) for that tile.
//   This is how the Omni-LoRA adapter is "chosen per sample by task id"
//   (reference semantics: Omni_AVSR/Llama_LoRA.py:246-259, Qwen_LoRA.py:557-570).
// * epilogue: (+bias) -> round bf16 -> act -> (+residual) -> bf16 / fp32 store, mirroring the rounding
//   points of the reference's unfused bf16 op sequence.
#include <stdlib.h>
#include "gemm_epilogue.cuh"

namespace omni {

template <int BN, int STAGES>
struct GemmSmem {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BAR_OFFSET = STAGES * STAGE_BYTES;
  static constexpr int TOTAL = BAR_OFFSET + (2 * STAGES + 1) * 8 + 16 + 1024;  // +1024 for manual alignment
};

template <int BN, int STAGES>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_tn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmB2,
                    const GemmKParams p) {
  using S = GemmSmem<BN, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::BAR_OFFSET);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_tile = blockIdx.x;
  const int m_tile = blockIdx.y;
  const int m0 = m_tile * BM;
  const int n0 = n_tile * BN;

  const int group = p.tile_group ? p.tile_group[m_tile] : 0;
  const int tbl = group * p.n_tiles + n_tile;
  const int b_row = p.b_row_table ? p.b_row_table[tbl] : n0;
  const int4* ext = p.ext_table ? (p.ext_table + static_cast<long long>(tbl) * p.n_ext) : nullptr;
  int n_ext_valid = 0;
  if (ext) {
    for (int j = 0; j < p.n_ext; ++j) n_ext_valid += (ext[j].y >= 0) ? 1 : 0;
  }
  const int total_iters = p.num_k_blocks + n_ext_valid;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (ext) {
      tma_prefetch_desc(&tmA2);
      tma_prefetch_desc(&tmB2);
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
    tmem_alloc(tmem_ptr_smem, BN);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ===== TMA producer =====
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0; it < p.num_k_blocks; ++it) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sA = smem + stage * S::STAGE_BYTES;
        uint8_t* sB = sA + S::A_BYTES;
        mbar_expect_tx(&full_bar[stage], S::STAGE_BYTES);
        tma_load_2d(&tmA, &full_bar[stage], sA, it * BK, m0);
        tma_load_2d(&tmB, &full_bar[stage], sB, it * BK, b_row);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      for (int j = 0; j < p.n_ext; ++j) {
        if (!ext) break;
        const int4 e = ext[j];
        if (e.y < 0) continue;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sA = smem + stage * S::STAGE_BYTES;
        uint8_t* sB = sA + S::A_BYTES;
        mbar_expect_tx(&full_bar[stage], S::STAGE_BYTES);
        tma_load_2d(&tmA2, &full_bar[stage], sA, e.x, m0);
        tma_load_2d(&tmB2, &full_bar[stage], sB, e.z, e.y);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: lane 0 issues every tcgen05.mma and every commit =====
    constexpr uint32_t idesc = make_idesc_bf16(BM, BN, 0, 0);
    int stage = 0;
    uint32_t phase = 0;
    for (int it = 0; it < total_iters; ++it) {
      mbar_wait(&full_bar[stage], phase);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t sA = smem_u32(smem + stage * S::STAGE_BYTES);
        const uint32_t sB = sA + S::A_BYTES;
        const uint64_t adesc = make_smem_desc_sw128(sA, 16, 1024);
        const uint64_t bdesc = make_smem_desc_sw128(sB, 16, 1024);
#pragma unroll
        for (int k = 0; k < BK / UMMA_K; ++k) {
          // advance 32 bytes (16 bf16) inside the 128B swizzle atom: +2 in the (addr >> 4) field
          umma_bf16(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc, (it > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(&empty_bar[stage]);
        if (it == total_iters - 1) umma_commit(tmem_full_bar);
      }
      __syncwarp();
      if (++stage == STAGES) { stage = 0; phase ^= 1; }
    }
  } else {
    // ===== epilogue: warps 2..5, TMEM lane quarter = warp % 4 =====
    const int q = warp & 3;
    const int row = m0 + q * 32 + lane;
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    const bool row_ok = row < p.M;
    const float alpha = p.alpha;
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      uint32_t r[32];
      tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(c * 32), r);
      tmem_ld_wait();
      const int col0 = n0 + c * 32;
      if (!row_ok || col0 >= p.N) continue;
      float v[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]) * alpha;
      epilogue_store_32(p, v, row, col0);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tmem_dealloc(tmem_base, BN);
  }
}

// ---------------------------------------------------------------------------------------------------
// Persistent variant: one CTA per SM loops over output tiles; the TMEM accumulator is double-buffered
// (2 x BN columns) so the epilogue of tile i overlaps the main loop of tile i+1, and the TMA pipeline
// never drains between tiles.
// ---------------------------------------------------------------------------------------------------
template <int BN, int STAGES>
struct GemmSmemP {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BAR_OFFSET = STAGES * STAGE_BYTES;
  static constexpr int TOTAL = BAR_OFFSET + (2 * STAGES + 4) * 8 + 16 + 1024;
};

template <int BN, int STAGES>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_tn_persistent(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                        const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmB2,
                        const GemmKParams p, const int m_fast) {
  using S = GemmSmemP<BN, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::BAR_OFFSET);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;   // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;   // [2]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_tiles = p.m_tiles * p.n_tiles;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (p.ext_table) {
      tma_prefetch_desc(&tmA2);
      tma_prefetch_desc(&tmB2);
    }
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < STAGES; ++s) {
        mbar_init(&full_bar[s], 1);
        mbar_init(&empty_bar[s], 1);
      }
      for (int s = 0; s < 2; ++s) {
        mbar_init(&tmem_full_bar[s], 1);
        mbar_init(&tmem_empty_bar[s], 4);   // one arrive per epilogue warp
      }
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(tmem_ptr_smem, 2 * BN);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ===== TMA producer =====
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        int m_tile, n_tile;
        tile_coords(t, p.m_tiles, p.n_tiles, m_fast, m_tile, n_tile);
        const int m0 = m_tile * BM;
        const int group = p.tile_group ? p.tile_group[m_tile] : 0;
        const int tbl = group * p.n_tiles + n_tile;
        const int b_row = p.b_row_table ? p.b_row_table[tbl] : n_tile * BN;
        for (int it = 0; it < p.num_k_blocks; ++it) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sA = smem + stage * S::STAGE_BYTES;
          uint8_t* sB = sA + S::A_BYTES;
          mbar_expect_tx(&full_bar[stage], S::STAGE_BYTES);
          tma_load_2d(&tmA, &full_bar[stage], sA, it * BK, m0);
          tma_load_2d(&tmB, &full_bar[stage], sB, it * BK, b_row);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (p.ext_table) {
          const int4* ext = p.ext_table + static_cast<long long>(tbl) * p.n_ext;
          for (int j = 0; j < p.n_ext; ++j) {
            const int4 e = ext[j];
            if (e.y < 0) continue;
            mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* sA = smem + stage * S::STAGE_BYTES;
            uint8_t* sB = sA + S::A_BYTES;
            mbar_expect_tx(&full_bar[stage], S::STAGE_BYTES);
            tma_load_2d(&tmA2, &full_bar[stage], sA, e.x, m0);
            tma_load_2d(&tmB2, &full_bar[stage], sB, e.z, e.y);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    constexpr uint32_t idesc = make_idesc_bf16(BM, BN, 0, 0);
    int stage = 0;
    uint32_t phase = 0;
    int as = 0;
    uint32_t aphase = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      int m_tile, n_tile;
      tile_coords(t, p.m_tiles, p.n_tiles, m_fast, m_tile, n_tile);
      int total_iters = p.num_k_blocks;
      if (p.ext_table) {
        const int group = p.tile_group ? p.tile_group[m_tile] : 0;
        const int4* ext = p.ext_table + static_cast<long long>(group * p.n_tiles + n_tile) * p.n_ext;
        for (int j = 0; j < p.n_ext; ++j) total_iters += (ext[j].y >= 0) ? 1 : 0;
      }
      mbar_wait(&tmem_empty_bar[as], aphase ^ 1);   // epilogue has drained this accumulator stage
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(as * BN);
      for (int it = 0; it < total_iters; ++it) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sA = smem_u32(smem + stage * S::STAGE_BYTES);
          const uint32_t sB = sA + S::A_BYTES;
          const uint64_t adesc = make_smem_desc_sw128(sA, 16, 1024);
          const uint64_t bdesc = make_smem_desc_sw128(sB, 16, 1024);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k)
            umma_bf16(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (it > 0 || k > 0) ? 1u : 0u);
          umma_commit(&empty_bar[stage]);
          if (it == total_iters - 1) umma_commit(&tmem_full_bar[as]);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      as ^= 1;
      if (as == 0) aphase ^= 1;
    }
  } else {
    // ===== epilogue warps 2..5 =====
    const int q = warp & 3;
    int as = 0;
    uint32_t aphase = 0;
    const float alpha = p.alpha;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      int m_tile, n_tile;
      tile_coords(t, p.m_tiles, p.n_tiles, m_fast, m_tile, n_tile);
      const int row = m_tile * BM + q * 32 + lane;
      const int n0 = n_tile * BN;
      const bool row_ok = row < p.M;
      ResPrefetch res_cur, res_nxt;
      res_prefetch(p, row, n0, row_ok, res_nxt);          // in flight while the main loop of this tile finishes
      mbar_wait(&tmem_full_bar[as], aphase);
      tc_fence_after();
      const uint32_t tmem_acc = tmem_base + static_cast<uint32_t>(as * BN) + (static_cast<uint32_t>(q * 32) << 16);
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        res_cur = res_nxt;
        if (c + 1 < BN / 32) res_prefetch(p, row, n0 + (c + 1) * 32, row_ok, res_nxt);
        uint32_t r[32];
        tmem_ld_32x32(tmem_acc + static_cast<uint32_t>(c * 32), r);
        tmem_ld_wait();
        if (c == BN / 32 - 1) {
          // all TMEM reads of this warp for this tile are done: hand the accumulator stage back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tmem_empty_bar[as]);
        }
        const int col0 = n0 + c * 32;
        if (!row_ok || col0 >= p.N) continue;
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]) * alpha;
        epilogue_store_32(p, v, row, col0, &res_cur);
      }
      as ^= 1;
      if (as == 0) aphase ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tmem_dealloc(tmem_base, 2 * BN);
  }
}

// ---------------------------------------------------------------------------------------------------
// Split-K variant for the decode step (M <= 128: one tile of activation rows, the weight matrix streamed once).
// Measured (profiles/decode_gemm_r1): one SM pulls only ~25 B/clk of HBM-missing TMA traffic, so a [2048 x 2048] weight
// spread over 8..32 CTAs streams at 0.4-1.5 TB/s.  Here a cluster of SPLIT CTAs shares one 128 x 64 output tile, each
// CTA accumulating a K slice in its own TMEM; ranks 1.. push their fp32 partials into rank 0's shared memory over DSMEM
// (column-major: 32 lanes = 32 consecutive floats), one cluster barrier, and rank 0 runs the usual epilogue on the sum.
// ---------------------------------------------------------------------------------------------------
template <int SPLIT, int STAGES>
struct GemmSmemSK {
  static constexpr int BN = 64;
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int PART_OFFSET = STAGES * STAGE_BYTES;                 // [SPLIT-1][64 cols][128 rows] fp32
  static constexpr int PART_BYTES = (SPLIT - 1) * BN * BM * 4;
  static constexpr int BAR_OFFSET = PART_OFFSET + PART_BYTES;
  static constexpr int TOTAL = BAR_OFFSET + (2 * STAGES + 1) * 8 + 16 + 1024;
};

template <int SPLIT, int STAGES>
__global__ void __cluster_dims__(SPLIT, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_tn_splitk(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmKParams p) {
  using S = GemmSmemSK<SPLIT, STAGES>;
  constexpr int BN = S::BN;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::BAR_OFFSET);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);
  float* part = reinterpret_cast<float*>(smem + S::PART_OFFSET);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int rank = static_cast<int>(cluster_ctarank());
  const int n_tile = blockIdx.x / SPLIT;
  const int kb_per = p.num_k_blocks / SPLIT;
  const int kb0 = rank * kb_per;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
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
    tmem_alloc(tmem_ptr_smem, BN);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0; it < kb_per; ++it) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sA = smem + stage * S::STAGE_BYTES;
        uint8_t* sB = sA + S::A_BYTES;
        mbar_expect_tx(&full_bar[stage], S::STAGE_BYTES);
        tma_load_2d(&tmB, &full_bar[stage], sB, (kb0 + it) * BK, n_tile * BN);     // the HBM stream first
        tma_load_2d(&tmA, &full_bar[stage], sA, (kb0 + it) * BK, 0);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = make_idesc_bf16(BM, BN, 0, 0);
    int stage = 0;
    uint32_t phase = 0;
    for (int it = 0; it < kb_per; ++it) {
      mbar_wait(&full_bar[stage], phase);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t sA = smem_u32(smem + stage * S::STAGE_BYTES);
        const uint32_t sB = sA + S::A_BYTES;
        const uint64_t adesc = make_smem_desc_sw128(sA, 16, 1024);
        const uint64_t bdesc = make_smem_desc_sw128(sB, 16, 1024);
#pragma unroll
        for (int k = 0; k < BK / UMMA_K; ++k)
          umma_bf16(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc, (it > 0 || k > 0) ? 1u : 0u);
        umma_commit(&empty_bar[stage]);
        if (it == kb_per - 1) umma_commit(tmem_full_bar);
      }
      __syncwarp();
      if (++stage == STAGES) { stage = 0; phase ^= 1; }
    }
  } else if (rank != 0) {
    // ===== partial producer: TMEM -> rank 0's shared memory =====
    const int q = warp & 3;
    const int r = q * 32 + lane;
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(part)), "r"(0));
    remote += static_cast<uint32_t>(((rank - 1) * BN * BM + r) * 4);
#pragma unroll
    for (int c = 0; c < BN / 32; ++c) {
      uint32_t v[32];
      tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(c * 32), v);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i)
        asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(remote + static_cast<uint32_t>((c * 32 + i) * BM * 4)), "r"(v[i])
                     : "memory");
    }
    tc_fence_before();
  }

  // partials published (release) / visible to rank 0 (acquire)
  cluster_sync_all();

  if (rank == 0 && warp >= 2) {
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const int row = r;
    const int n0 = n_tile * BN;
    const bool row_ok = row < p.M;
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
#pragma unroll
    for (int c = 0; c < BN / 32; ++c) {
      uint32_t acc[32];
      tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(c * 32), acc);
      tmem_ld_wait();
      const int col0 = n0 + c * 32;
      if (!row_ok || col0 >= p.N) continue;
      float v[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        float a = __uint_as_float(acc[i]);
#pragma unroll
        for (int s2 = 0; s2 < SPLIT - 1; ++s2) a += part[(s2 * BN + c * 32 + i) * BM + r];
        v[i] = a * p.alpha;
      }
      epilogue_store_32(p, v, row, col0);
    }
    tc_fence_before();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, BN);
}

// ---------------------------------------------------------------------------------------------------
// Cluster variant: two CTAs (one per SM of a TPC pair) work on two vertically adjacent M tiles of the same N tile.
// Each CTA loads its own A tile and HALF of the shared B tile; the TMA unit multicasts that half into both CTAs'
// shared memory, so the L2 -> SM traffic per MAC drops by a third (48 KB -> 32 KB per 128x256x64 block).  A smem
// slot may only be refilled when BOTH MMA warps are done with it: tcgen05.commit arrives on the empty barrier of
// both CTAs (count 2).  Everything else (TMEM double buffering, epilogue) is the persistent kernel.
// ---------------------------------------------------------------------------------------------------
template <int BN, int STAGES>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_tn_cluster(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmBh,
                     const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmB2,
                     const GemmKParams p, const int m_fast) {
  using S = GemmSmemP<BN, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::BAR_OFFSET);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;   // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;   // [2]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int rank = static_cast<int>(cluster_ctarank());
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;
  const int m_pairs = (p.m_tiles + 1) >> 1;
  const int num_pairs = m_pairs * p.n_tiles;
  constexpr int B_HALF_BYTES = S::B_BYTES / 2;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmBh);
    if (p.ext_table) {
      tma_prefetch_desc(&tmA2);
      tma_prefetch_desc(&tmB2);
    }
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < STAGES; ++s) {
        mbar_init(&full_bar[s], 1);
        mbar_init(&empty_bar[s], 2);      // both CTAs' MMA warps release a slot
      }
      for (int s = 0; s < 2; ++s) {
        mbar_init(&tmem_full_bar[s], 1);
        mbar_init(&tmem_empty_bar[s], 4);
      }
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(tmem_ptr_smem, 2 * BN);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                     // partner's barriers are initialised before any remote arrive / multicast
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = cluster_id; t < num_pairs; t += num_clusters) {
        int m_pair, n_tile;
        tile_coords(t, m_pairs, p.n_tiles, m_fast, m_pair, n_tile);
        const int m_tile = 2 * m_pair + rank;
        const int m0 = m_tile * BM;                  // may be >= M for the phantom tile of an odd M: TMA zero-fills
        const int n0 = n_tile * BN;
        for (int it = 0; it < p.num_k_blocks; ++it) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sA = smem + stage * S::STAGE_BYTES;
          uint8_t* sB = sA + S::A_BYTES;
          mbar_expect_tx(&full_bar[stage], S::STAGE_BYTES);
          tma_load_2d(&tmA, &full_bar[stage], sA, it * BK, m0);
          tma_load_2d_multicast(&tmBh, &full_bar[stage], sB + rank * B_HALF_BYTES, it * BK, n0 + rank * (BN / 2),
                                static_cast<uint16_t>(3));
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (p.ext_table) {
          const int mt = m_tile < p.m_tiles ? m_tile : p.m_tiles - 1;
          const int group = p.tile_group ? p.tile_group[mt] : 0;
          const int4* ext = p.ext_table + static_cast<long long>(group * p.n_tiles + n_tile) * p.n_ext;
          for (int j = 0; j < p.n_ext; ++j) {
            const int4 e = ext[j];
            if (e.y < 0) continue;
            mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* sA = smem + stage * S::STAGE_BYTES;
            uint8_t* sB = sA + S::A_BYTES;
            mbar_expect_tx(&full_bar[stage], S::STAGE_BYTES);
            tma_load_2d(&tmA2, &full_bar[stage], sA, e.x, m0);
            // the adapter (hence B2) may differ between the two CTAs: private, non-multicast load of both halves
            tma_load_2d(&tmB2, &full_bar[stage], sB, e.z, e.y);
            tma_load_2d(&tmB2, &full_bar[stage], sB + B_HALF_BYTES, e.z, e.y + BN / 2);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = make_idesc_bf16(BM, BN, 0, 0);
    int stage = 0;
    uint32_t phase = 0;
    int as = 0;
    uint32_t aphase = 0;
    for (int t = cluster_id; t < num_pairs; t += num_clusters) {
      int m_pair, n_tile;
      tile_coords(t, m_pairs, p.n_tiles, m_fast, m_pair, n_tile);
      int total_iters = p.num_k_blocks;
      if (p.ext_table) {
        // the number of extension blocks depends only on the N tile (Q / K / V column class), not on the task
        const int4* ext = p.ext_table + static_cast<long long>(n_tile) * p.n_ext;
        for (int j = 0; j < p.n_ext; ++j) total_iters += (ext[j].y >= 0) ? 1 : 0;
      }
      mbar_wait(&tmem_empty_bar[as], aphase ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(as * BN);
      for (int it = 0; it < total_iters; ++it) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sA = smem_u32(smem + stage * S::STAGE_BYTES);
          const uint32_t sB = sA + S::A_BYTES;
          const uint64_t adesc = make_smem_desc_sw128(sA, 16, 1024);
          const uint64_t bdesc = make_smem_desc_sw128(sB, 16, 1024);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k)
            umma_bf16(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (it > 0 || k > 0) ? 1u : 0u);
          umma_commit_multicast(&empty_bar[stage], static_cast<uint16_t>(3));
          if (it == total_iters - 1) umma_commit(&tmem_full_bar[as]);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      as ^= 1;
      if (as == 0) aphase ^= 1;
    }
  } else {
    const int q = warp & 3;
    int as = 0;
    uint32_t aphase = 0;
    const float alpha = p.alpha;
    for (int t = cluster_id; t < num_pairs; t += num_clusters) {
      int m_pair, n_tile;
      tile_coords(t, m_pairs, p.n_tiles, m_fast, m_pair, n_tile);
      const int row = (2 * m_pair + rank) * BM + q * 32 + lane;
      const int n0 = n_tile * BN;
      const bool row_ok = row < p.M;
      ResPrefetch res_cur, res_nxt;
      res_prefetch(p, row, n0, row_ok, res_nxt);          // in flight while the main loop of this tile finishes
      mbar_wait(&tmem_full_bar[as], aphase);
      tc_fence_after();
      const uint32_t tmem_acc = tmem_base + static_cast<uint32_t>(as * BN) + (static_cast<uint32_t>(q * 32) << 16);
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        res_cur = res_nxt;
        if (c + 1 < BN / 32) res_prefetch(p, row, n0 + (c + 1) * 32, row_ok, res_nxt);
        uint32_t r[32];
        tmem_ld_32x32(tmem_acc + static_cast<uint32_t>(c * 32), r);
        tmem_ld_wait();
        if (c == BN / 32 - 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tmem_empty_bar[as]);
        }
        const int col0 = n0 + c * 32;
        if (!row_ok || col0 >= p.N) continue;
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]) * alpha;
        epilogue_store_32(p, v, row, col0, &res_cur);
      }
      as ^= 1;
      if (as == 0) aphase ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                     // nobody exits while the partner can still multicast into / arrive on it
  if (warp == 1) {
    tmem_dealloc(tmem_base, 2 * BN);
  }
}

// ---------------------------------------------------------------------------------------------------
// CTA-pair variant (tcgen05.mma.cta_group::2): the two SMs of a TPC compute one 256 x 256 output tile together.
// Each CTA stages its own 128 rows of A and HALF (128 rows) of the B tile; the pair's tensor cores read both halves,
// so shared-memory fill and operand-read traffic per MAC are halved w.r.t. the single-CTA kernels -- the single-CTA
// 128x256 tile needs 96 B/clk of TMA fill + 96 B/clk of MMA operand reads against a 128 B/clk shared memory.
// Protocol: both CTAs' TMA loads credit the LEADER's full barrier; the leader's MMA warp issues every MMA and
// multicasts its commits to both CTAs' empty / tmem_full barriers; both epilogues arrive on the leader's tmem_empty.
// Plain GEMMs only (no grouping tables / K-extension): those stay on the kernels above.
// ---------------------------------------------------------------------------------------------------
template <int STAGES>
struct GemmSmem2 {
  static constexpr int BN2 = 256;
  static constexpr int A_BYTES = BM * BK * 2;            // own 128 rows of A
  static constexpr int B_BYTES = (BN2 / 2) * BK * 2;     // own half of the B tile
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;  // 32 KB
  static constexpr int EPI_OFFSET = STAGES * STAGE_BYTES;                    // 8 per-warp transposition tiles
  static constexpr int BAR_OFFSET = EPI_OFFSET + 8 * EPI_WARP_BYTES;
  static constexpr int TOTAL = BAR_OFFSET + (2 * STAGES + 4) * 8 + 16 + 1024;
};

// EPI: 0 plain epilogue, 1 OMNI_ACT_SWIGLU64, 2 OMNI_ACT_GELU_KEEP (own instantiations: their extra registers stay out of
// the plain kernel)
template <int STAGES, int EPI = 0>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM2_THREADS, 1)
gemm_bf16_tn_2cta(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmBh,
                  const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmB2h, const GemmKParams p,
                  const int m_fast) {
  using S = GemmSmem2<STAGES>;
  constexpr int BN2 = S::BN2;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::BAR_OFFSET);   // used in the leader only
  uint64_t* empty_bar = full_bar + STAGES;                                   // one per CTA
  uint64_t* tmem_full_bar = empty_bar + STAGES;                              // [2], one set per CTA
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;                              // [2], used in the leader only (count 8)
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int rank = static_cast<int>(cluster_ctarank());
  const bool leader = rank == 0;
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;
  const int m_pairs = (p.m_tiles + 1) >> 1;
  const int n_tiles2 = (p.N + BN2 - 1) / BN2;
  const int num_pairs = m_pairs * n_tiles2;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmBh);
    if (p.ext_table) {
      tma_prefetch_desc(&tmA2);
      tma_prefetch_desc(&tmB2h);
    }
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < STAGES; ++s) {
        mbar_init(&full_bar[s], 1);
        mbar_init(&empty_bar[s], 1);
      }
      for (int s = 0; s < 2; ++s) {
        mbar_init(&tmem_full_bar[s], 1);
        mbar_init(&tmem_empty_bar[s], 16);  // 8 epilogue warps x 2 CTAs
      }
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc_pair(tmem_ptr_smem, 2 * BN2);
    tmem_relinquish_pair();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ===== TMA producer (both CTAs) =====
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = cluster_id; t < num_pairs; t += num_clusters) {
        int m_pair, n_tile;
        tile_coords(t, m_pairs, n_tiles2, m_fast, m_pair, n_tile);
        const int m0 = (2 * m_pair + rank) * BM;
        const int nrow = n_tile * BN2 + rank * (BN2 / 2);
        for (int it = 0; it < p.num_k_blocks; ++it) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sA = smem + stage * S::STAGE_BYTES;
          uint8_t* sB = sA + S::A_BYTES;
          if (leader) mbar_expect_tx(&full_bar[stage], 2 * S::STAGE_BYTES);   // bytes of BOTH CTAs land on this barrier
          tma_load_2d_pair(&tmA, &full_bar[stage], sA, it * BK, m0);
          tma_load_2d_pair(&tmBh, &full_bar[stage], sB, it * BK, nrow);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (p.ext_table) {
          // K-extension (LoRA up-projection): both M tiles of a pair belong to the same task (256-row aligned
          // segments), so the pair shares the adapter rows of B2; each CTA stages its half
          const int mt = (2 * m_pair < p.m_tiles) ? 2 * m_pair : p.m_tiles - 1;
          const int group = p.tile_group ? p.tile_group[mt] : 0;
          const int4* ext = p.ext_table + static_cast<long long>(group * n_tiles2 + n_tile) * p.n_ext;
          for (int j = 0; j < p.n_ext; ++j) {
            const int4 e = ext[j];
            if (e.y < 0) continue;
            mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* sA = smem + stage * S::STAGE_BYTES;
            uint8_t* sB = sA + S::A_BYTES;
            if (leader) mbar_expect_tx(&full_bar[stage], 2 * S::STAGE_BYTES);
            tma_load_2d_pair(&tmA2, &full_bar[stage], sA, e.x, m0);
            tma_load_2d_pair(&tmB2h, &full_bar[stage], sB, e.z, e.y + rank * (BN2 / 2));
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (leader CTA only) =====
    if (leader) {
      constexpr uint32_t idesc = make_idesc_bf16(2 * BM, BN2, 0, 0);   // M = 256 across the pair
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int t = cluster_id; t < num_pairs; t += num_clusters) {
        int total_iters = p.num_k_blocks;
        if (p.ext_table) {
          int m_pair, n_tile;
          tile_coords(t, m_pairs, n_tiles2, m_fast, m_pair, n_tile);
          const int4* ext = p.ext_table + static_cast<long long>(n_tile) * p.n_ext;   // count depends on the N tile only
          for (int j = 0; j < p.n_ext; ++j) total_iters += (ext[j].y >= 0) ? 1 : 0;
        }
        mbar_wait(&tmem_empty_bar[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(as * BN2);
        for (int it = 0; it < total_iters; ++it) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t sA = smem_u32(smem + stage * S::STAGE_BYTES);
            const uint32_t sB = sA + S::A_BYTES;
            const uint64_t adesc = make_smem_desc_sw128(sA, 16, 1024);
            const uint64_t bdesc = make_smem_desc_sw128(sB, 16, 1024);
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k)
              umma_bf16_pair(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (it > 0 || k > 0) ? 1u : 0u);
            umma_commit_pair_multicast(&empty_bar[stage], static_cast<uint16_t>(3));
            if (it == total_iters - 1) umma_commit_pair_multicast(&tmem_full_bar[as], static_cast<uint16_t>(3));
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        as ^= 1;
        if (as == 0) aphase ^= 1;
      }
    }
  } else {
    // ===== epilogue (both CTAs: each drains its own 128 TMEM lanes) =====
    // Eight warps: warp w reads TMEM lane quarter w % 4 (hardware rule) and the column half (w - 2) / 4, so a thread
    // owns 128 of the 256 accumulator columns of its row.  All four residual chunks of the tile are requested before
    // the wait on the accumulator: the (strided, 64 bytes per thread) loads fly under the tile's main loop.
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    constexpr int CH = BN2 / 64;            // 32-column chunks per thread
    uint8_t* stage_tile = smem + S::EPI_OFFSET + (warp - 2) * EPI_WARP_BYTES;
    int as = 0;
    uint32_t aphase = 0;
    const float alpha = p.alpha;
    for (int t = cluster_id; t < num_pairs; t += num_clusters) {
      int m_pair, n_tile;
      tile_coords(t, m_pairs, n_tiles2, m_fast, m_pair, n_tile);
      const int row0 = (2 * m_pair + rank) * BM + q * 32;        // first of this warp's 32 rows
      const int row = row0 + lane;
      const int n0 = n_tile * BN2 + half * (BN2 / 2);
      const bool row_ok = row < p.M;
      // fast path: bf16 output and both 64-column blocks inside N -> coalesced through the transposition tile
      const bool fast = !p.out_fp32 && (n0 + BN2 / 2 <= p.N);
      // EPI 4 / 5 read `residual` as the saved forward tensor of the fused backward, not as an addend
      const bool use_res = EPI < 4 && fast && p.residual != nullptr && row0 < p.M;
      ResTile rt;
      if (use_res) res_tile_prefetch(p, row0, n0, lane, rt);       // in flight under the rest of this tile's main loop
      ResTile rt2;
      if constexpr (EPI == 4) {
        // gate | up blocks of this thread's first 64 intermediate channels (interleaved layout: block b at columns 128 b)
        if (fast && row0 < p.M) {
          tile_prefetch(p.residual, p.ldr, p.M, row0, 2 * n0, lane, rt);
          tile_prefetch(p.residual, p.ldr, p.M, row0, 2 * n0 + 64, lane, rt2);
        }
      }
      if constexpr (EPI == 5) {
        if (fast && row0 < p.M) tile_prefetch(p.residual, p.ldr, p.M, row0, n0, lane, rt);
      }
      mbar_wait(&tmem_full_bar[as], aphase);
      tc_fence_after();
      const uint32_t tmem_acc = tmem_base + static_cast<uint32_t>(as * BN2 + half * (BN2 / 2)) +
                                (static_cast<uint32_t>(q * 32) << 16);
      if constexpr (EPI == 4) {
        // fused SwiGLU backward (the dgrad GEMM of down_proj): this thread's d(act) row, 2 blocks of 64 intermediate channels,
        // becomes the d(gate) | d(up) blocks of d(gate_up); the accumulator is read in 32-column pieces to stay in registers
        const bool active = row0 < p.M;       // warp-uniform
        uint8_t* my_row = stage_tile + lane * EPI_PITCH;
        bf16* o = reinterpret_cast<bf16*>(p.out);
#pragma unroll
        for (int c2 = 0; c2 < 2; ++c2) {
          uint32_t g[32], du[32];
          if (active) {
            tile_to_row(stage_tile, lane, rt, g);
            res_tile_publish(stage_tile, lane, rt2);      // the up block stays in the tile: consumed and overwritten in place
          }
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            uint32_t r[32];
            tmem_ld_32x32(tmem_acc + static_cast<uint32_t>(c2 * 64 + h * 32), r);
            tmem_ld_wait();
            if (c2 == 1 && h == 1) {
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive_cluster(&tmem_empty_bar[as], 0);   // leader's barrier
            }
            if (active) swiglu_bwd_half(r, alpha, g, h, my_row, du);
          }
          if (!active) continue;
          if (c2 == 0) {     // next block's gate | up: in flight under this block's two tile stores
            tile_prefetch(p.residual, p.ldr, p.M, row0, 2 * (n0 + 64), lane, rt);
            tile_prefetch(p.residual, p.ldr, p.M, row0, 2 * (n0 + 64) + 64, lane, rt2);
          }
          stage_to_global(o, p.ldo, p.M, stage_tile, lane, row0, 2 * (n0 + c2 * 64));
          store_tile64_packed(o, p.ldo, p.M, stage_tile, du, lane, row0, 2 * (n0 + c2 * 64) + 64);
        }
      } else if (fast) {
        uint32_t gate[EPI == 1 ? 32 : 1];     // the rounded gate block, packed bf16 pairs
#pragma unroll
        for (int c2 = 0; c2 < 2; ++c2) {
          uint32_t r0[32], r1[32];
          tmem_ld_32x32(tmem_acc + static_cast<uint32_t>(c2 * 64), r0);
          tmem_ld_32x32(tmem_acc + static_cast<uint32_t>(c2 * 64 + 32), r1);
          tmem_ld_wait();
          if (c2 == 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(&tmem_empty_bar[as], 0);   // leader's barrier
          }
          if (row0 >= p.M) continue;                                      // warp-uniform
          float v[64];
          const float2 al2 = make_float2(alpha, alpha);
#pragma unroll
          for (int i = 0; i < 32; i += 2) {          // packed scaling (FMUL2)
            const float2 a = fmul2(make_float2(__uint_as_float(r0[i]), __uint_as_float(r0[i + 1])), al2);
            const float2 b = fmul2(make_float2(__uint_as_float(r1[i]), __uint_as_float(r1[i + 1])), al2);
            v[i] = a.x; v[i + 1] = a.y;
            v[32 + i] = b.x; v[32 + i + 1] = b.y;
          }
          if constexpr (EPI == 5) {
            // fused GELU backward (the dgrad GEMM of fc2): d(act) block x gelu'(pre-activation block)
            uint32_t x[32], dx[32];
            tile_to_row(stage_tile, lane, rt, x);
            if (c2 == 0) tile_prefetch(p.residual, p.ldr, p.M, row0, n0 + 64, lane, rt);
            gelu_bwd_row64(v, x, dx);
            store_tile64_packed(reinterpret_cast<bf16*>(p.out), p.ldo, p.M, stage_tile, dx, lane, row0, n0 + c2 * 64);
            continue;
          }
          if (use_res) {
            res_tile_publish(stage_tile, lane, rt);
            if (c2 == 0) res_tile_prefetch(p, row0, n0 + 64, lane, rt);   // lands while block 0 is converted and stored
          }
          if constexpr (EPI == 3) {
            epilogue_tile64_prelu_ring(p, stage_tile, v, lane, row0, n0 + c2 * 64, use_res);
          } else {
            // rounds v to bf16 when act != NONE; SwiGLU without a gate|up output (inference): only the activation is stored
            epilogue_tile64(p, stage_tile, v, lane, row0, n0 + c2 * 64, use_res, EPI != 1 || p.out != nullptr);
          }
          if constexpr (EPI == 2) {
            // v holds the rounded pre-activation that was just stored: second output = gelu of it
#pragma unroll
            for (int i = 0; i < 64; i += 2) {
              const float2 g = gelu_fast2(make_float2(v[i], v[i + 1]));
              v[i] = g.x;
              v[i + 1] = g.y;
            }
            store_tile64(p.out2, p.ldo2, p.M, stage_tile, v, lane, row0, n0 + c2 * 64);
          }
          if constexpr (EPI == 1) {
            // this thread's 128 columns are [gate 64 | up 64] of the same 64 intermediate channels
            if (c2 == 0) {
#pragma unroll
              for (int i = 0; i < 32; ++i) gate[i] = f2_to_bf2(v[2 * i], v[2 * i + 1]);      // exact: v is already rounded
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                // bf16(silu(gate)) * up on pairs, packed fp32 arithmetic (silu = g * sigmoid(g): the multiplication by the
                // reciprocal that __fdividef(g, 1 + e) performs)
                const float2 g = bf2_to_f2(gate[i]);
                const float2 sl = bf2_to_f2(f2_to_bf2_pair(fmul2(g, sigmoid2(g))));
                const float2 r = fmul2(sl, make_float2(v[2 * i], v[2 * i + 1]));
                v[2 * i] = r.x;
                v[2 * i + 1] = r.y;
              }
              store_tile64(p.out2, p.ldo2, p.M, stage_tile, v, lane, row0, n0 >> 1);
            }
          }
        }
      } else {
#pragma unroll
        for (int c = 0; c < CH; ++c) {
          uint32_t r[32];
          tmem_ld_32x32(tmem_acc + static_cast<uint32_t>(c * 32), r);
          tmem_ld_wait();
          if (c == CH - 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(&tmem_empty_bar[as], 0);   // leader's barrier
          }
          const int col0 = n0 + c * 32;
          if (!row_ok || col0 >= p.N) continue;
          float v[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]) * alpha;
          epilogue_store_32(p, v, row, col0);      // edge tiles / fp32 output: direct row-per-thread path
        }
      }
      as ^= 1;
      if (as == 0) aphase ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tmem_dealloc_pair(tmem_base, 2 * BN2);
  }
}

static int omni_sm_count() {
  static int n = 0;   // immutable once resolved
  if (n == 0) {
    int dev = 0, v = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return kNumSMs;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) return kNumSMs;
    n = v;
  }
  return n;
}

template <int BN, int STAGES>
static int launch_gemm(const omni_gemm_args* a, cudaStream_t stream) {
  using S = GemmSmem<BN, STAGES>;
  CUtensorMap tmA, tmB, tmA2, tmB2;
  int rc;
  // K == 0: the whole reduction is the K-extension list (table-driven convolution, ops.conv_frames); CTA-pair kernel only
  const bool ext_only = a->K == 0;
  if (!ext_only) {
    rc = omni_make_tmap_2d_bf16(&tmA, a->A, (uint64_t)a->M, (uint64_t)a->K, (uint64_t)a->lda, BM, BK, 1);
    if (rc) return rc;
    rc = omni_make_tmap_2d_bf16(&tmB, a->B, (uint64_t)a->b_rows, (uint64_t)a->K, (uint64_t)a->ldb, BN, BK, 1);
    if (rc) return rc;
  }
  if (a->ext_table) {
    rc = omni_make_tmap_2d_bf16(&tmA2, a->A2, (uint64_t)a->M, (uint64_t)a->a2_cols, (uint64_t)a->lda2, BM, BK, 1);
    if (rc) return rc;
    rc = omni_make_tmap_2d_bf16(&tmB2, a->B2, (uint64_t)a->b2_rows, (uint64_t)a->b2_cols, (uint64_t)a->ldb2, BN, BK, 1);
    if (rc) return rc;
  } else {
    tmA2 = tmA;
    tmB2 = tmB;
  }
  if (ext_only) {
    tmA = tmA2;
    tmB = tmB2;
  }
  GemmKParams p;
  p.M = a->M; p.N = a->N; p.K = a->K;
  p.num_k_blocks = ceil_div(a->K, BK);
  p.n_tiles = ceil_div(a->N, BN);
  p.m_tiles = ceil_div(a->M, BM);
  p.n_ext = a->ext_table ? a->n_ext : 0;
  p.tile_group = a->tile_group;
  p.b_row_table = a->b_row_table;
  p.ext_table = reinterpret_cast<const int4*>(a->ext_table);
  p.bias = reinterpret_cast<const bf16*>(a->bias);
  p.residual = reinterpret_cast<const bf16*>(a->residual);
  p.out = a->out;
  p.ldo = a->ldo; p.ldr = a->ldr;
  p.act = a->act; p.out_fp32 = a->out_fp32; p.alpha = a->alpha;
  p.out2 = reinterpret_cast<bf16*>(a->out2); p.ldo2 = a->ldo2;
  p.slope = reinterpret_cast<const bf16*>(a->slope);
  p.res_bias = reinterpret_cast<const bf16*>(a->res_bias);
  p.ring_hp = a->ring_h + 2; p.ring_wp = a->ring_w + 2; p.ring_g = a->ring_group; p.ring_c = a->ring_c;

  if constexpr (BN == 64) {
    // decode step: one tile of rows, few N tiles -> split K over a 4-CTA cluster so that enough SMs pull on HBM
    constexpr int SPLIT = 4, SKST = 4;
    static const bool no_splitk = (getenv("OMNI_GEMM_NO_SPLITK") != nullptr);
    if (!no_splitk && p.m_tiles == 1 && !a->ext_table && !a->b_row_table && !a->tile_group && p.n_tiles * SPLIT <= 160 &&
        p.num_k_blocks % SPLIT == 0 && p.num_k_blocks >= 4 * SPLIT && (a->K % BK) == 0) {
      using SK = GemmSmemSK<SPLIT, SKST>;
      auto ks = gemm_bf16_tn_splitk<SPLIT, SKST>;
      static bool attr_set_sk = false;
      if (!attr_set_sk) {
        if (cudaFuncSetAttribute(ks, cudaFuncAttributeMaxDynamicSharedMemorySize, SK::TOTAL) != cudaSuccess)
          return OMNI_ERR_CUDA;
        attr_set_sk = true;
      }
      ks<<<p.n_tiles * SPLIT, GEMM_THREADS, SK::TOTAL, stream>>>(tmA, tmB, p);
      OMNI_LAUNCH_CHECK();
      return OMNI_OK;
    }
  }
  static const bool use_v1 = (getenv("OMNI_GEMM_V1") != nullptr);   // debugging switch: one tile per CTA
  if (use_v1 && ext_only) return OMNI_ERR_UNSUPPORTED;
  if (!use_v1) {
    using SP = GemmSmemP<BN, STAGES>;
    auto kp = gemm_bf16_tn_persistent<BN, STAGES>;
    static bool attr_set_p = false;
    if (!attr_set_p) {
      if (cudaFuncSetAttribute(kp, cudaFuncAttributeMaxDynamicSharedMemorySize, SP::TOTAL) != cudaSuccess)
        return OMNI_ERR_CUDA;
      attr_set_p = true;
    }
    const int tiles = p.m_tiles * p.n_tiles;
    const int sms = omni_sm_count();
    // raster: walk M fastest when the A panel is small enough to live in L2 and there are more N tiles than M tiles
    // (lm_head-like shapes: the big B operand is then streamed from HBM exactly once)
    const long long a_bytes = static_cast<long long>(a->M) * a->K * 2;
    const int m_fast = (!a->b_row_table && !a->ext_table && p.m_tiles < p.n_tiles && a_bytes <= (40ll << 20)) ? 1 : 0;
    static const bool no_2cta = (getenv("OMNI_GEMM_NO_2CTA") != nullptr);
    const bool pair_ok = !no_2cta && BN == 256 && !a->b_row_table && (!a->ext_table || a->pair_aligned) &&
                         ((p.m_tiles >= 2 && tiles >= sms / 2) || ext_only);
    if ((a->act == OMNI_ACT_SWIGLU64 || a->act == OMNI_ACT_GELU_KEEP || a->act == OMNI_ACT_PRELU_RING ||
         a->act == OMNI_ACT_SWIGLU_BWD64 || a->act == OMNI_ACT_GELU_BWD || ext_only) && !pair_ok)
      return OMNI_ERR_UNSUPPORTED;
    if (pair_ok) {
      // CTA pairs (tcgen05.mma.cta_group::2): 256 x 256 tile per pair, half the shared-memory traffic per MAC
      constexpr int ST2 = 5;     // 5 x 32 KB operand stages + 36 KB of epilogue transposition tiles
      using S2 = GemmSmem2<ST2>;
      const int epi = a->act == OMNI_ACT_SWIGLU64 ? 1 : a->act == OMNI_ACT_GELU_KEEP ? 2 : a->act == OMNI_ACT_PRELU_RING ? 3
                      : a->act == OMNI_ACT_SWIGLU_BWD64 ? 4 : a->act == OMNI_ACT_GELU_BWD ? 5 : 0;
      auto k2 = epi == 1 ? gemm_bf16_tn_2cta<ST2, 1> : epi == 2 ? gemm_bf16_tn_2cta<ST2, 2>
                : epi == 3 ? gemm_bf16_tn_2cta<ST2, 3> : epi == 4 ? gemm_bf16_tn_2cta<ST2, 4>
                : epi == 5 ? gemm_bf16_tn_2cta<ST2, 5> : gemm_bf16_tn_2cta<ST2, 0>;
      static bool attr_set_2[6] = {false, false, false, false, false, false};
      if (!attr_set_2[epi]) {
        if (cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, S2::TOTAL) != cudaSuccess)
          return OMNI_ERR_CUDA;
        attr_set_2[epi] = true;
      }
      CUtensorMap tmBh, tmB2h2;
      if (!ext_only) {
        rc = omni_make_tmap_2d_bf16(&tmBh, a->B, (uint64_t)a->b_rows, (uint64_t)a->K, (uint64_t)a->ldb, 128, BK, 1);
        if (rc) return rc;
      }
      if (a->ext_table) {
        rc = omni_make_tmap_2d_bf16(&tmB2h2, a->B2, (uint64_t)a->b2_rows, (uint64_t)a->b2_cols, (uint64_t)a->ldb2, 128,
                                    BK, 1);
        if (rc) return rc;
      } else {
        tmB2h2 = tmBh;
      }
      if (ext_only) tmBh = tmB2h2;
      const int pairs = ((p.m_tiles + 1) / 2) * ceil_div(a->N, 256);
      int clusters = sms / 2;
      if (pairs < clusters) clusters = pairs;
      int m_fast2 = (!ext_only && ((p.m_tiles + 1) / 2) < ceil_div(a->N, 256) && a_bytes <= (40ll << 20)) ? 1 : 0;
      // L2-friendly raster for a B operand that does not stay L2-resident under the output stream (gate_up: 67 MB of weights
      // against 1 GB of output per launch): groups of N tiles whose B panel is <= 32 MB, all M tiles per group.  Costs one more
      // pass over A per extra group, so only when A is not the bigger operand.
      static const bool no_group = (getenv("OMNI_GEMM_NO_NGROUP") != nullptr);
      const long long b_bytes = static_cast<long long>(a->N) * a->K * 2;
      if (!no_group && !m_fast2 && !a->ext_table && !ext_only && b_bytes > (48ll << 20) && a_bytes <= 2 * b_bytes) {
        const long long tile_bytes = 256ll * a->K * 2;
        const int ng = static_cast<int>((32ll << 20) / tile_bytes);
        if (ng >= 2 && ng < ceil_div(a->N, 256)) m_fast2 = ng;
      }
      k2<<<2 * clusters, GEMM2_THREADS, S2::TOTAL, stream>>>(tmA, tmBh, tmA2, tmB2h2, p, m_fast2);
      OMNI_LAUNCH_CHECK();
      return OMNI_OK;
    }
    static const bool no_cluster = (getenv("OMNI_GEMM_NO_CLUSTER") != nullptr);
    if (!no_cluster && BN >= 128 && !a->b_row_table && p.m_tiles >= 2 && tiles >= sms / 2) {
      // 2-CTA clusters with TMA multicast of the shared B tile
      auto kc = gemm_bf16_tn_cluster<BN, STAGES>;
      static bool attr_set_c = false;
      if (!attr_set_c) {
        if (cudaFuncSetAttribute(kc, cudaFuncAttributeMaxDynamicSharedMemorySize, SP::TOTAL) != cudaSuccess)
          return OMNI_ERR_CUDA;
        attr_set_c = true;
      }
      CUtensorMap tmBh, tmB2h;
      rc = omni_make_tmap_2d_bf16(&tmBh, a->B, (uint64_t)a->b_rows, (uint64_t)a->K, (uint64_t)a->ldb, BN / 2, BK, 1);
      if (rc) return rc;
      if (a->ext_table) {
        rc = omni_make_tmap_2d_bf16(&tmB2h, a->B2, (uint64_t)a->b2_rows, (uint64_t)a->b2_cols, (uint64_t)a->ldb2,
                                    BN / 2, BK, 1);
        if (rc) return rc;
      } else {
        tmB2h = tmBh;
      }
      const int pairs = ((p.m_tiles + 1) / 2) * p.n_tiles;
      int clusters = sms / 2;
      if (pairs < clusters) clusters = pairs;
      kc<<<2 * clusters, GEMM_THREADS, SP::TOTAL, stream>>>(tmA, tmBh, tmA2, tmB2h, p, m_fast);
      OMNI_LAUNCH_CHECK();
      return OMNI_OK;
    }
    kp<<<tiles < sms ? tiles : sms, GEMM_THREADS, SP::TOTAL, stream>>>(tmA, tmB, tmA2, tmB2, p, m_fast);
    OMNI_LAUNCH_CHECK();
    return OMNI_OK;
  }
  auto kfn = gemm_bf16_tn_kernel<BN, STAGES>;
  static bool attr_set = false;  // idempotent attribute; benign if set twice
  if (!attr_set) {
    if (cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL) != cudaSuccess)
      return OMNI_ERR_CUDA;
    attr_set = true;
  }
  dim3 grid(p.n_tiles, p.m_tiles, 1);
  kfn<<<grid, GEMM_THREADS, S::TOTAL, stream>>>(tmA, tmB, tmA2, tmB2, p);
  OMNI_LAUNCH_CHECK();
  return OMNI_OK;
}

}  // namespace omni

extern "C" int omni_gemm_bf16(const omni_gemm_args* a, void* stream) {
  using namespace omni;
  OMNI_CHECK_ARG(a != nullptr);
  OMNI_CHECK_ARG(a->M > 0 && a->N > 0 && (a->K > 0 || (a->K == 0 && a->ext_table)));
  OMNI_CHECK_ARG((a->K == 0 || (a->A && a->B)) && (a->out || (a->act == OMNI_ACT_SWIGLU64 && a->out2)));
  OMNI_CHECK_ARG((a->lda % 8) == 0 && (a->ldb % 8) == 0);
  OMNI_CHECK_ARG((reinterpret_cast<uintptr_t>(a->A) & 15) == 0 && (reinterpret_cast<uintptr_t>(a->B) & 15) == 0);
  if (a->out) {
    OMNI_CHECK_ARG(a->ldo >= a->N);
    OMNI_CHECK_ARG((a->ldo % 8) == 0 && (reinterpret_cast<uintptr_t>(a->out) & 15) == 0);
  }
  if (a->residual) OMNI_CHECK_ARG((a->ldr % 8) == 0 && (reinterpret_cast<uintptr_t>(a->residual) & 15) == 0);
  if (a->bias) OMNI_CHECK_ARG((reinterpret_cast<uintptr_t>(a->bias) & 15) == 0);
  if (a->ext_table) {
    OMNI_CHECK_ARG(a->A2 && a->B2 && a->n_ext > 0);
    OMNI_CHECK_ARG((a->lda2 % 8) == 0 && (a->ldb2 % 8) == 0);
  }
  if (a->b_row_table || a->ext_table) OMNI_CHECK_ARG(a->block_n == 64 || a->block_n == 128 || a->block_n == 256);
  if (a->act == OMNI_ACT_GELU_KEEP) {
    OMNI_CHECK_ARG(a->out2 && (a->ldo2 % 8) == 0 && a->ldo2 >= a->N && (reinterpret_cast<uintptr_t>(a->out2) & 15) == 0);
    if (a->block_n != 256 || (a->N % 256) != 0 || a->out_fp32 || a->residual || a->ext_table || a->b_row_table)
      return OMNI_ERR_UNSUPPORTED;
  }
  if (a->act == OMNI_ACT_PRELU_RING) {
    OMNI_CHECK_ARG(a->slope && a->bias && a->ring_h >= 0 && a->ring_w >= 0 && a->ring_group > 0 && a->ring_c > 0);
    OMNI_CHECK_ARG((a->ring_c % 64) == 0 && a->ring_group * a->ring_c == a->N);
    OMNI_CHECK_ARG((reinterpret_cast<uintptr_t>(a->slope) & 15) == 0 &&
                   (!a->res_bias || (reinterpret_cast<uintptr_t>(a->res_bias) & 15) == 0));
    if (a->block_n != 256 || (a->N % 128) != 0 || a->out_fp32 || a->b_row_table) return OMNI_ERR_UNSUPPORTED;
  }
  if (a->act == OMNI_ACT_SWIGLU_BWD64 || a->act == OMNI_ACT_GELU_BWD) {
    // `residual` = the saved forward tensor ([M, 2N] gate|up blocks, resp. [M, N] pre-activation); out = d of it
    const int64_t width = a->act == OMNI_ACT_SWIGLU_BWD64 ? 2 * static_cast<int64_t>(a->N) : a->N;
    OMNI_CHECK_ARG(a->residual && a->ldr >= width && a->ldo >= width);
    if (a->block_n != 256 || (a->N % 256) != 0 || a->out_fp32 || a->bias || a->ext_table || a->b_row_table)
      return OMNI_ERR_UNSUPPORTED;
  }
  if (a->act == OMNI_ACT_SWIGLU64) {
    OMNI_CHECK_ARG(a->out2 && (a->ldo2 % 8) == 0 && a->ldo2 >= a->N / 2 && (reinterpret_cast<uintptr_t>(a->out2) & 15) == 0);
    if (a->block_n != 256 || (a->N % 256) != 0 || a->out_fp32 || a->residual || a->bias || a->ext_table || a->b_row_table)
      return OMNI_ERR_UNSUPPORTED;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  int bn = a->block_n;
  if (bn == 0) bn = (a->N <= 64) ? 64 : 128;
  switch (bn) {
    case 64: return launch_gemm<64, 8>(a, st);
    case 128: return launch_gemm<128, 6>(a, st);
    case 256: return launch_gemm<256, 4>(a, st);
    default: return OMNI_ERR_BAD_ARG;
  }
}
