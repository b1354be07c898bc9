// Fused Matryoshka path (north_star subsystem 1): token compression -> projector MLP -> splice into the LLM rows,
// both modalities, ONE persistent launch.
//
// Reference op chain replaced (Omni_AVSR/modeling_OmniAVSR.py): avg-pool / stack :544-546,562-568 (audio),
// :469-471,487-493 (video); projector Linear+ReLU+Linear :366 / :353; marker cat :347-368; task cat + labels :270-299,
// :373-387 (train) and :406-458 (infer).
//
// Why this shape.  At the named sizes (D = 1024, I = H = 2048) the projector is tensor-bound (5.03 GFLOP per utterance
// against 7.14 MB: 300-1000 FLOP/B, ridge 217), and the [128, I] intermediate of one row tile (512 KB in bf16) fits neither
// the 227 KB of shared memory nor the 256 KB of tensor memory of an SM, so "GEMM-1 stays on chip" would need a K-split of
// GEMM-2 over a cluster with 1 MB of fp32 partials per CTA crossing DSMEM per row tile (21 B/clk) -- slower than the L2.
// What B200 does offer is a 126 MB L2: `pooled` (13 MB per modality at B = 32) and `hidden` (26 MB) are written and read
// back by other SMs microseconds later without touching HBM.  So the kernel is a persistent CTA-pair GEMM
// (tcgen05.mma cta_group::2, 256x256 tiles, TMEM double-buffered accumulators, 4-stage TMA pipeline) that walks the
// static work list  [audio GEMM-1 | audio GEMM-2 | video GEMM-1 | video GEMM-2]  in order (the audio GEMMs run while the
// pool warps are still compressing the video rows), plus four "pool" warps per
// CTA that run concurrently with the tensor pipe:
//   * pool warps: TMA-stage the r encoder rows of one output token (3-D tensor map over [B, T, D], box [r, 256]) into a
//     private 3-slot ring, sum them in window order in fp32, divide by r, round to bf16 (the arithmetic of AvgPool1d, bit
//     for bit) or copy them (stack), write the `pooled` row, bump the row block's counter.  When the compression work is
//     done the same warps copy the marker / prompt / text embedding rows into the LLM rows and write every label.
//   * TMA producer: before loading the A tile of a work item it waits for the counter of that 256-row block (`pooled`
//     for GEMM-1, `hidden` for GEMM-2; acquire load + generic->async proxy fence).  Every dependency points to an earlier
//     item of the list or to the pool warps, and all CTAs are co-resident (grid <= SM count, 1 CTA / SM), so it cannot
//     deadlock.
//   * epilogue warps (8): GEMM-1: +bias, round, ReLU -> `hidden`, then release the block counter.  GEMM-2: +bias, round,
//     and each row goes straight to its TWO destinations in the packed LLM buffer (own-task sequence + AVSR sequence)
//     through the warp's transposition tile, so every store instruction writes full 128-byte row segments.
#include <stdlib.h>
#include <string.h>
#include "gemm_epilogue.cuh"
#include "splice_common.cuh"

namespace omni {

constexpr int PPS_STAGES = 4;
constexpr int PPS_EPI_WARPS = 8;
constexpr int PPS_POOL_WARPS = 6;        // warps 2, 3 and 12..15 (the epilogue warps 4..11 keep TMEM lane quarter = warp % 4)
constexpr int PPS_THREADS = (2 + PPS_EPI_WARPS + PPS_POOL_WARPS) * 32;   // 512 = 128 registers per thread
constexpr int PPS_RING_BYTES = 10240;    // TMA landing ring of one pool warp, cut into slots of one unit each
constexpr int PPS_MAX_SLOTS = 4;         // slots per ring: 4 x 2560 B or 2 x 5120 B (power of two: index = count & (n - 1))
constexpr int PPS_SLOT_BYTES = 4096;     // largest unit
constexpr int PPS_CHUNK = 8;             // pool units per dependency release; divides the units of a full 256-row block
constexpr int PPS_BN = 256;

struct PpsSmem {
  static constexpr int A_BYTES = BM * BK * 2;             // own 128 rows of A
  static constexpr int B_BYTES = (PPS_BN / 2) * BK * 2;   // own half of the B tile
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;   // 32 KB
  static constexpr int EPI_OFFSET = PPS_STAGES * STAGE_BYTES;
  static constexpr int POOL_OFFSET = EPI_OFFSET + PPS_EPI_WARPS * EPI_WARP_BYTES;
  static constexpr int BAR_OFFSET = POOL_OFFSET + PPS_POOL_WARPS * PPS_RING_BYTES;
  static constexpr int N_BARS = 2 * PPS_STAGES + 4 + PPS_POOL_WARPS * PPS_MAX_SLOTS;
  static constexpr int TOTAL = BAR_OFFSET + N_BARS * 8 + 16 + 1024;
};
static_assert(PpsSmem::TOTAL <= 232448, "shared memory budget");

struct PpsGemmDesc {
  int M, N;
  int num_k_blocks;
  int m_pairs, n_tiles;      // 256-row blocks x 256-column tiles
  int item_begin;            // first work item of this GEMM in the list
  int relu;                  // 1: GEMM-1 (bias + ReLU -> hidden);  0: GEMM-2 (bias -> scatter into the LLM rows)
  int mod;                   // 0 audio, 1 video
  const bf16* bias;
  bf16* out;                 // GEMM-1: hidden [M, N];  GEMM-2: optional dense copy of the projected tokens (may be null)
  long long ldo;
  const int* wait_cnt;       // [m_pairs] counters the A operand of a row block depends on
  int wait_per_row;          // > 0: target = rows_in_block * wait_per_row (pooled units);  0: target = wait_fixed
  int wait_fixed;
  int* signal_cnt;           // [m_pairs] bumped once per CTA per tile (GEMM-1) or null
};

struct PpsScatter {          // where projected token (clip b, index j) of a modality goes
  bf16* dst[2];              // own-task sequence rows, AVSR sequence rows (null: not written)
  int S[2];                  // sequence length of that task
  int pos0[2];               // position of token 0 inside a sequence
  int n;                     // tokens per clip
};

struct PpsPool {
  int M, n, r, D, K1;
  int cw;                    // columns per TMA box
  int lane_cols;             // columns per lane: 8 (16-byte accesses) or 4 (8-byte accesses, cw <= 128 so all lanes work)
  int pow2;                  // rate is a power of two: x / r == x * rcp exactly
  float rcp, fr;
  int n_boxes_row;           // boxes per output row
  int boxes_per_unit;
  int units_per_row;
  long long units;           // M * units_per_row
  bf16* pooled;
  int* ready;                // [m_pairs]
};

struct PpsParams {
  PpsGemmDesc g[4];          // audio GEMM-1, audio GEMM-2, video GEMM-1, video GEMM-2 (absent modality: M = 0)
  PpsScatter sc[2];
  PpsPool pool[2];
  SpliceK splice;
  long long splice_rows;
  int n_items;
  int mode;
  int H;
  int slot_bytes, n_slots;   // geometry of a pool warp's TMA ring (one unit per slot), common to both modalities
};

struct PpsMaps {
  CUtensorMap a[4], b[4], x[2];
};

__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }

__device__ __forceinline__ void tma_load_3d(const CUtensorMap* m, uint64_t* bar, void* smem_dst, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// Bounded acquire-spin on a dependency counter (another CTA's pool / epilogue warps release it).
__device__ __forceinline__ void wait_counter(const int* p, int target) {
  uint32_t spins = 0;
  while (true) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    if (v >= target) break;
    __nanosleep(64);
    if (++spins > (1u << 25)) __trap();
  }
}
// all lanes' global stores of this warp -> visible to whoever acquires the counter (also to its TMA loads)
__device__ __forceinline__ void release_counter(int* p, int lane, int count = 1) {
  __syncwarp();
  if (lane == 0) {
    __threadfence();
    fence_proxy_async_global();
    atomicAdd(p, count);
  }
}

__device__ __forceinline__ int item_gemm(const PpsParams& P, int it) {
  int gi = 0;
  if (it >= P.g[1].item_begin) gi = 1;
  if (it >= P.g[2].item_begin) gi = 2;
  if (it >= P.g[3].item_begin) gi = 3;
  return gi;
}

// 32 rows x 64 bf16 columns held one row per thread -> the one or two destination rows recorded in the pad of the
// warp's transposition tile (16 bytes after the 128 data bytes of every row) and, optionally, a dense copy.
__device__ __forceinline__ void scatter_tile64(uint8_t* stage, const float (&v)[64], int lane, int col0, bf16* dense,
                                               long long ldd, int row0, int M) {
  uint8_t* my_row = stage + lane * EPI_PITCH;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    uint4 o;
    o.x = f2_to_bf2(v[8 * i + 0], v[8 * i + 1]);
    o.y = f2_to_bf2(v[8 * i + 2], v[8 * i + 3]);
    o.z = f2_to_bf2(v[8 * i + 4], v[8 * i + 5]);
    o.w = f2_to_bf2(v[8 * i + 6], v[8 * i + 7]);
    *reinterpret_cast<uint4*>(my_row + 16 * i) = o;
  }
  __syncwarp();
  const int sub = lane >> 3;
  const int cofs = col0 + (lane & 7) * 8;
#pragma unroll
  for (int pass = 0; pass < 8; ++pass) {
    const int rr = sub + 4 * pass;
    const uint8_t* src = stage + rr * EPI_PITCH;
    const uint4 val = *reinterpret_cast<const uint4*>(src + (lane & 7) * 16);
    const ulonglong2 d = *reinterpret_cast<const ulonglong2*>(src + 128);
    if (d.x) *reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(d.x) + cofs) = val;
    if (d.y) *reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(d.y) + cofs) = val;
    if (dense && row0 + rr < M) *reinterpret_cast<uint4*>(dense + static_cast<long long>(row0 + rr) * ldd + cofs) = val;
  }
  __syncwarp();
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(PPS_THREADS, 1)
pool_project_splice_kernel(const __grid_constant__ PpsMaps maps, const __grid_constant__ PpsParams P) {
  using S = PpsSmem;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::BAR_OFFSET);   // used in the leader only
  uint64_t* empty_bar = full_bar + PPS_STAGES;                               // one set per CTA
  uint64_t* tmem_full_bar = empty_bar + PPS_STAGES;                          // [2], one set per CTA
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;                              // [2], used in the leader only (count 16)
  uint64_t* pool_bar = tmem_empty_bar + 2;                                   // [pool warps][slots]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(pool_bar + PPS_POOL_WARPS * PPS_MAX_SLOTS);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int rank = static_cast<int>(cluster_ctarank());
  const bool leader = rank == 0;
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < 4; ++i) {
      if (P.g[i].M > 0) {
        tma_prefetch_desc(&maps.a[i]);
        tma_prefetch_desc(&maps.b[i]);
      }
    }
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < PPS_STAGES; ++s) {
        mbar_init(&full_bar[s], 1);
        mbar_init(&empty_bar[s], 1);
      }
      for (int s = 0; s < 2; ++s) {
        mbar_init(&tmem_full_bar[s], 1);
        mbar_init(&tmem_empty_bar[s], 2 * PPS_EPI_WARPS);
      }
      for (int s = 0; s < PPS_POOL_WARPS * PPS_MAX_SLOTS; ++s) mbar_init(&pool_bar[s], 1);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc_pair(tmem_ptr_smem, 2 * PPS_BN);
    tmem_relinquish_pair();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ===== TMA producer (both CTAs of the pair) =====
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int it = cluster_id; it < P.n_items; it += num_clusters) {
        const int gi = item_gemm(P, it);
        const PpsGemmDesc& g = P.g[gi];
        const int local = it - g.item_begin;
        const int mp = local / g.n_tiles;
        const int nt = local - mp * g.n_tiles;
        const int rows_blk = min(2 * BM, g.M - mp * 2 * BM);
        wait_counter(g.wait_cnt + mp, g.wait_per_row > 0 ? rows_blk * g.wait_per_row : g.wait_fixed);
        fence_proxy_async_global();
        const int m0 = (2 * mp + rank) * BM;
        const int nrow = nt * PPS_BN + rank * (PPS_BN / 2);
        const CUtensorMap* tA = &maps.a[gi];
        const CUtensorMap* tB = &maps.b[gi];
        for (int kb = 0; kb < g.num_k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sA = smem + stage * S::STAGE_BYTES;
          uint8_t* sB = sA + S::A_BYTES;
          if (leader) mbar_expect_tx(&full_bar[stage], 2 * S::STAGE_BYTES);   // bytes of BOTH CTAs land on this barrier
          tma_load_2d_pair(tA, &full_bar[stage], sA, kb * BK, m0);
          tma_load_2d_pair(tB, &full_bar[stage], sB, kb * BK, nrow);
          if (++stage == PPS_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (leader CTA only) =====
    if (leader) {
      constexpr uint32_t idesc = make_idesc_bf16(2 * BM, PPS_BN, 0, 0);   // M = 256 across the pair
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int it = cluster_id; it < P.n_items; it += num_clusters) {
        const int nkb = P.g[item_gemm(P, it)].num_k_blocks;
        mbar_wait(&tmem_empty_bar[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(as * PPS_BN);
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t sA = smem_u32(smem + stage * S::STAGE_BYTES);
            const uint32_t sB = sA + S::A_BYTES;
            const uint64_t adesc = make_smem_desc_sw128(sA, 16, 1024);
            const uint64_t bdesc = make_smem_desc_sw128(sB, 16, 1024);
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k)
              umma_bf16_pair(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
            umma_commit_pair_multicast(&empty_bar[stage], static_cast<uint16_t>(3));
            if (kb == nkb - 1) umma_commit_pair_multicast(&tmem_full_bar[as], static_cast<uint16_t>(3));
          }
          __syncwarp();
          if (++stage == PPS_STAGES) { stage = 0; phase ^= 1; }
        }
        as ^= 1;
        if (as == 0) aphase ^= 1;
      }
    }
  } else if (warp >= 4 && warp < 4 + PPS_EPI_WARPS) {
    // ===== epilogue (both CTAs: each drains its own 128 TMEM lanes; warp w: lane quarter w % 4, column half (w-2)/4) =====
    const int q = warp & 3;
    const int half = (warp - 4) >> 2;
    uint8_t* stage_tile = smem + S::EPI_OFFSET + (warp - 4) * EPI_WARP_BYTES;
    int as = 0;
    uint32_t aphase = 0;
    for (int it = cluster_id; it < P.n_items; it += num_clusters) {
      const int gi = item_gemm(P, it);
      const PpsGemmDesc& g = P.g[gi];
      const int local = it - g.item_begin;
      const int mp = local / g.n_tiles;
      const int nt = local - mp * g.n_tiles;
      const int row0 = (2 * mp + rank) * BM + q * 32;
      const int row = row0 + lane;
      const int n0 = nt * PPS_BN + half * (PPS_BN / 2);
      const bool row_ok = row < g.M;
      const bool fast = (n0 + PPS_BN / 2 <= g.N);
      GemmKParams gp;
      gp.M = g.M; gp.N = g.N; gp.bias = g.bias; gp.residual = nullptr; gp.out = g.out; gp.ldo = g.ldo; gp.ldr = 0;
      gp.act = g.relu ? OMNI_ACT_RELU : OMNI_ACT_NONE; gp.out_fp32 = 0; gp.alpha = 1.0f; gp.out2 = nullptr; gp.ldo2 = 0;
      unsigned long long d0 = 0, d1 = 0;
      if (!g.relu) {
        // destinations of this thread's row: own-task sequence and AVSR sequence
        const PpsScatter& sc = P.sc[g.mod];
        if (row_ok) {
          const int b = row / sc.n;
          const int j = row - b * sc.n;
          if (sc.dst[0])
            d0 = reinterpret_cast<unsigned long long>(sc.dst[0] + (static_cast<long long>(b) * sc.S[0] + sc.pos0[0] + j) * P.H);
          if (sc.dst[1])
            d1 = reinterpret_cast<unsigned long long>(sc.dst[1] + (static_cast<long long>(b) * sc.S[1] + sc.pos0[1] + j) * P.H);
        }
        *reinterpret_cast<ulonglong2*>(stage_tile + lane * EPI_PITCH + 128) = make_ulonglong2(d0, d1);
        __syncwarp();
      }
      mbar_wait(&tmem_full_bar[as], aphase);
      tc_fence_after();
      const uint32_t tmem_acc = tmem_base + static_cast<uint32_t>(as * PPS_BN + half * (PPS_BN / 2)) +
                                (static_cast<uint32_t>(q * 32) << 16);
      if (fast) {
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
          if (row0 >= g.M) continue;                                      // warp-uniform
          float v[64];
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            v[i] = __uint_as_float(r0[i]);
            v[32 + i] = __uint_as_float(r1[i]);
          }
          const int col0 = n0 + c2 * 64;
          if (g.relu) {
            epilogue_tile64(gp, stage_tile, v, lane, row0, col0, false);
          } else {
            const uint4* bp = reinterpret_cast<const uint4*>(g.bias + col0);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const uint4 bb = __ldg(bp + i);
              float2 f;
              f = bf2_to_f2(bb.x); v[8 * i + 0] += f.x; v[8 * i + 1] += f.y;
              f = bf2_to_f2(bb.y); v[8 * i + 2] += f.x; v[8 * i + 3] += f.y;
              f = bf2_to_f2(bb.z); v[8 * i + 4] += f.x; v[8 * i + 5] += f.y;
              f = bf2_to_f2(bb.w); v[8 * i + 6] += f.x; v[8 * i + 7] += f.y;
            }
            scatter_tile64(stage_tile, v, lane, col0, g.out, g.ldo, row0, g.M);
          }
        }
      } else {
        // edge tiles (N not a multiple of 128 inside this half): direct row-per-thread path
#pragma unroll 1
        for (int c = 0; c < PPS_BN / 64; ++c) {
          uint32_t r[32];
          tmem_ld_32x32(tmem_acc + static_cast<uint32_t>(c * 32), r);
          tmem_ld_wait();
          if (c == PPS_BN / 64 - 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(&tmem_empty_bar[as], 0);
          }
          const int col0 = n0 + c * 32;
          if (!row_ok || col0 >= g.N) continue;
          float v[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
          if (g.relu) {
            epilogue_store_32(gp, v, row, col0);
          } else {
            for (int i = 0; i < 32; ++i) {
              if (col0 + i >= g.N) break;
              const bf16 o = __float2bfloat16_rn(v[i] + __bfloat162float(g.bias[col0 + i]));
              if (d0) reinterpret_cast<bf16*>(d0)[col0 + i] = o;
              if (d1) reinterpret_cast<bf16*>(d1)[col0 + i] = o;
              if (g.out) g.out[static_cast<long long>(row) * g.ldo + col0 + i] = o;
            }
          }
        }
      }
      if (g.signal_cnt) {
        // all eight epilogue warps have stored their part of the hidden tile: one gpu-scope release per CTA and tile
        asm volatile("bar.sync 1, %0;" ::"n"(PPS_EPI_WARPS * 32) : "memory");
        if (warp == 4 && lane == 0) {
          __threadfence();
          fence_proxy_async_global();
          atomicAdd(g.signal_cnt + mp, 1);
        }
      }
      as ^= 1;
      if (as == 0) aphase ^= 1;
    }
  } else {
    // ===== pool warps: Matryoshka compression, then the marker / prompt / text rows and the labels =====
    const int pw_local = warp < 4 ? warp - 2 : warp - (2 + PPS_EPI_WARPS);
    const long long pw = static_cast<long long>(blockIdx.x) * PPS_POOL_WARPS + pw_local;
    const long long PW = static_cast<long long>(gridDim.x) * PPS_POOL_WARPS;
    uint8_t* ring = smem + S::POOL_OFFSET + pw_local * PPS_RING_BYTES;
    uint64_t* bars = pool_bar + pw_local * PPS_MAX_SLOTS;
    const long long units_a = P.pool[0].units;

    // Work is handed out in CHUNKS of PPS_CHUNK consecutive units (a unit = one ring slot of TMA boxes = part of one output
    // row); a chunk never straddles a 256-row block, and the block counter is released ONCE per chunk -- a gpu-scope fence
    // per 4 KB unit capped the whole kernel at ~0.8 TB/s of encoder rows.
    const long long ch_a = ceil_div_ll(units_a, PPS_CHUNK);
    const long long ch_total = ch_a + ceil_div_ll(P.pool[1].units, PPS_CHUNK);
    struct Cur {                 // position in this warp's unit sequence; (m, uq, b, j) advance incrementally: one set of
      long long chunk;           // divisions per chunk, none per unit (lane 0 used to spend ~300 instructions per issue)
      int idx, len, mod;
      int m, uq, b, j;           // output row, unit within the row, clip, token within the clip
    };
    auto load_chunk = [&](Cur& c) {
      c.idx = 0;
      if (c.chunk >= ch_total) { c.len = 0; return; }
      c.mod = c.chunk >= ch_a ? 1 : 0;
      const PpsPool& pl = P.pool[c.mod];
      const long long u0 = (c.chunk - (c.mod ? ch_a : 0)) * PPS_CHUNK;
      const long long left = pl.units - u0;
      c.len = left < PPS_CHUNK ? static_cast<int>(left) : PPS_CHUNK;
      c.m = static_cast<int>(u0 / pl.units_per_row);
      c.uq = static_cast<int>(u0 - static_cast<long long>(c.m) * pl.units_per_row);
      c.b = c.m / pl.n;
      c.j = c.m - c.b * pl.n;
    };
    auto advance = [&](Cur& c) {
      if (++c.idx >= c.len) {
        c.chunk += PW;
        load_chunk(c);
        return;
      }
      const PpsPool& pl = P.pool[c.mod];
      if (++c.uq == pl.units_per_row) {
        c.uq = 0;
        ++c.m;
        if (++c.j == pl.n) { c.j = 0; ++c.b; }
      }
    };
    // issue the TMA boxes of the unit under the cursor into ring slot s (lane 0)
    auto issue = [&](const Cur& c, int s) {
      const PpsPool& pl = P.pool[c.mod];
      const int box0 = c.uq * pl.boxes_per_unit;
      const int nb = min(pl.boxes_per_unit, pl.n_boxes_row - box0);
      const uint32_t box_bytes = static_cast<uint32_t>(pl.r) * pl.cw * 2;
      mbar_expect_tx(&bars[s], box_bytes * nb);
      for (int k = 0; k < nb; ++k)
        tma_load_3d(&maps.x[c.mod], &bars[s], ring + s * P.slot_bytes + k * box_bytes, (box0 + k) * pl.cw, c.j * pl.r, c.b);
    };

    Cur cc, ci;
    cc.chunk = ci.chunk = pw;
    load_chunk(cc);
    load_chunk(ci);
    int issued = 0, consumed = 0;
    const int slot_mask = P.n_slots - 1;                               // n_slots is 2 or 4
    const int slot_shift = P.n_slots == 4 ? 2 : 1;
    for (int s = 0; s < slot_mask && ci.len > 0; ++s) {                // prologue: n_slots - 1 units in flight
      if (lane == 0) issue(ci, issued & slot_mask);
      ++issued;
      advance(ci);
    }
    while (cc.len > 0) {
      if (ci.len > 0) {
        // this slot held the unit consumed in the previous iteration (fully read before its closing __syncwarp)
        if (lane == 0) issue(ci, issued & slot_mask);
        ++issued;
        advance(ci);
      }
      const int s = consumed & slot_mask;
      mbar_wait(&bars[s], static_cast<uint32_t>((consumed >> slot_shift) & 1));
      const PpsPool& pl = P.pool[cc.mod];
      const int m = cc.m;
      const int box0 = cc.uq * pl.boxes_per_unit;
      const int nb = min(pl.boxes_per_unit, pl.n_boxes_row - box0);
      const int r = pl.r;
      const int cw = pl.cw;
      const uint8_t* slot = ring + s * P.slot_bytes;
      if (pl.lane_cols == 8) {
        for (int k = 0; k < nb; ++k) {
          const int col = (box0 + k) * cw + lane * 8;
          if (lane * 8 < cw && col < pl.D) {
            const uint8_t* src = slot + (k * r * cw + lane * 8) * 2;
            if (P.mode == OMNI_COMPRESS_AVG) {
              float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
              for (int i = 0; i < r; ++i) acc8(a, *reinterpret_cast<const uint4*>(src + i * cw * 2));   // window order
              if (pl.pow2) {                              // x / 2^k == x * 2^-k exactly: no IEEE division sequence
#pragma unroll
                for (int i = 0; i < 8; ++i) a[i] *= pl.rcp;
              } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) a[i] = a[i] / pl.fr;
              }
              uint4 o;
              o.x = f2_to_bf2(a[0], a[1]);
              o.y = f2_to_bf2(a[2], a[3]);
              o.z = f2_to_bf2(a[4], a[5]);
              o.w = f2_to_bf2(a[6], a[7]);
              *reinterpret_cast<uint4*>(pl.pooled + static_cast<long long>(m) * pl.K1 + col) = o;
            } else {
              bf16* dst = pl.pooled + static_cast<long long>(m) * pl.K1 + col;
#pragma unroll 4
              for (int i = 0; i < r; ++i)
                *reinterpret_cast<uint4*>(dst + static_cast<long long>(i) * pl.D) = *reinterpret_cast<const uint4*>(src + i * cw * 2);
            }
          }
        }
      } else {
        // narrow boxes (large rates: r * cw * 2 bytes must fit a ring slot): 4 columns per lane keep all 32 lanes busy
        for (int k = 0; k < nb; ++k) {
          const int col = (box0 + k) * cw + lane * 4;
          if (lane * 4 < cw && col < pl.D) {
            const uint8_t* src = slot + (k * r * cw + lane * 4) * 2;
            if (P.mode == OMNI_COMPRESS_AVG) {
              float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll 4
              for (int i = 0; i < r; ++i) {
                const uint2 u = *reinterpret_cast<const uint2*>(src + i * cw * 2);
                float2 f;
                f = bf2_to_f2(u.x); a0 += f.x; a1 += f.y;
                f = bf2_to_f2(u.y); a2 += f.x; a3 += f.y;
              }
              if (pl.pow2) { a0 *= pl.rcp; a1 *= pl.rcp; a2 *= pl.rcp; a3 *= pl.rcp; }
              else { a0 = a0 / pl.fr; a1 = a1 / pl.fr; a2 = a2 / pl.fr; a3 = a3 / pl.fr; }
              uint2 o;
              o.x = f2_to_bf2(a0, a1);
              o.y = f2_to_bf2(a2, a3);
              *reinterpret_cast<uint2*>(pl.pooled + static_cast<long long>(m) * pl.K1 + col) = o;
            } else {
              bf16* dst = pl.pooled + static_cast<long long>(m) * pl.K1 + col;
#pragma unroll 4
              for (int i = 0; i < r; ++i)
                *reinterpret_cast<uint2*>(dst + static_cast<long long>(i) * pl.D) = *reinterpret_cast<const uint2*>(src + i * cw * 2);
            }
          }
        }
      }
      ++consumed;
      __syncwarp();                                   // slot s may be refilled from here on
      if (cc.idx == cc.len - 1) release_counter(pl.ready + (m >> 8), lane, cc.len);
      advance(cc);
    }

    // marker / prompt / text embedding rows + labels (media rows are written by the GEMM-2 epilogue; only their label)
    const SpliceK& k = P.splice;
    for (long long gr = pw; gr < P.splice_rows; gr += PW) {
      int t = 0;
      long long base = 0;
      if (gr >= k.row_end[0]) { t = 1; base = k.row_end[0]; }
      if (gr >= k.row_end[1]) { t = 2; base = k.row_end[1]; }
      const long long rr = gr - base;
      const int b = static_cast<int>(static_cast<unsigned>(rr) / static_cast<unsigned>(k.S[t]));   // rows < 2^31 (checked)
      const int pos = static_cast<int>(rr - static_cast<long long>(b) * k.S[t]);
      const RowSrc src = splice_resolve(k, t, b, pos);
      if (lane == 0 && k.out_labels[t]) k.out_labels[t][rr] = src.label;
      if (!k.out[t] || src.media) continue;
      uint4* dst = reinterpret_cast<uint4*>(k.out[t]) + rr * k.H8;
      if (src.ptr) {
        const uint4* sp = reinterpret_cast<const uint4*>(src.ptr);
        int c = lane;
        for (; c + 96 < k.H8; c += 128) {
          const uint4 v0 = ld_nc_u4(sp + c);
          const uint4 v1 = ld_nc_u4(sp + c + 32);
          const uint4 v2 = ld_nc_u4(sp + c + 64);
          const uint4 v3 = ld_nc_u4(sp + c + 96);
          st_na_u4(dst + c, v0);
          st_na_u4(dst + c + 32, v1);
          st_na_u4(dst + c + 64, v2);
          st_na_u4(dst + c + 96, v3);
        }
        for (; c < k.H8; c += 32) st_na_u4(dst + c, ld_nc_u4(sp + c));
      } else {
        for (int c = lane; c < k.H8; c += 32) st_na_u4(dst + c, make_uint4(0u, 0u, 0u, 0u));
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) tmem_dealloc_pair(tmem_base, 2 * PPS_BN);
}

// 3-D bf16 tensor map over the encoder output [B, T, D] (batch stride x_bs elements), box [1, r, cw], no swizzle.
static int make_tmap_pool(CUtensorMap* out, const void* x, uint64_t B, uint64_t rows, uint64_t D, uint64_t x_bs, uint32_t r,
                          uint32_t cw) {
  omni_cuTensorMapEncodeTiled_t enc = omni_get_tmap_encoder();
  if (!enc) return OMNI_ERR_NO_DRIVER;
  cuuint64_t gdim[3] = {D, rows, B};
  cuuint64_t gstride[2] = {D * sizeof(bf16), x_bs * sizeof(bf16)};
  cuuint32_t box[3] = {cw, r, 1};
  cuuint32_t estride[3] = {1, 1, 1};
  CUresult rc = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(x), gdim, gstride, box, estride,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return rc == CUDA_SUCCESS ? OMNI_OK : OMNI_ERR_CUDA;
}

static int pps_sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0, v = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return kNumSMs;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) return kNumSMs;
    n = v;
  }
  return n;
}

struct PpsCounts {
  int pairs[2];       // 256-row blocks per modality
  int total_ints;
};
static PpsCounts pps_counts(const omni_pps_args* a) {
  PpsCounts c;
  const omni_pps_modality* mods[2] = {&a->audio, &a->video};
  for (int i = 0; i < 2; ++i) {
    const omni_pps_modality* m = mods[i];
    const long long M = (m->x && m->rate > 0) ? static_cast<long long>(a->splice.B) * (m->n_tok / m->rate) : 0;
    c.pairs[i] = static_cast<int>((M + 2 * BM - 1) / (2 * BM));
  }
  c.total_ints = 2 * (c.pairs[0] + c.pairs[1]);
  return c;
}

}  // namespace omni

extern "C" int64_t omni_pps_workspace_bytes(const omni_pps_args* a) {
  if (!a) return -1;
  const omni::PpsCounts c = omni::pps_counts(a);
  return static_cast<int64_t>(c.total_ints) * 4 + 256;
}

extern "C" int omni_pool_project_splice(const omni_pps_args* a, void* stream) {
  using namespace omni;
  OMNI_CHECK_ARG(a != nullptr);
  OMNI_CHECK_ARG(a->I > 0 && (a->I % 8) == 0);
  OMNI_CHECK_ARG(a->mode == OMNI_COMPRESS_AVG || a->mode == OMNI_COMPRESS_STACK);
  const omni_splice_args* sp = &a->splice;
  const omni_pps_modality* mods[2] = {&a->audio, &a->video};
  const bool present[2] = {a->audio.x != nullptr, a->video.x != nullptr};
  PpsParams P;
  PpsMaps maps;
  memset(&P, 0, sizeof(P));
  memset(&maps, 0, sizeof(maps));
  int rc = fill_splice(sp, &P.splice, present[0] ? 1 : 0, present[1] ? 1 : 0);
  if (rc) return rc;
  P.splice_rows = P.splice.row_end[2];
  OMNI_CHECK_ARG(P.splice_rows < (1ll << 31));
  P.mode = a->mode;
  P.H = sp->H;
  const int H = sp->H;
  const PpsCounts cnt = pps_counts(a);
  OMNI_CHECK_ARG(a->workspace != nullptr && a->workspace_bytes >= omni_pps_workspace_bytes(a));
  OMNI_CHECK_ARG((reinterpret_cast<uintptr_t>(a->workspace) & 15) == 0);
  int* counters = reinterpret_cast<int*>(a->workspace);
  int* pooled_cnt[2] = {counters, counters + cnt.pairs[0]};
  int* hidden_cnt[2] = {counters + cnt.pairs[0] + cnt.pairs[1], counters + 2 * cnt.pairs[0] + cnt.pairs[1]};
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);

  int items = 0;
  for (int i = 0; i < 2; ++i) {
    for (int layer = 0; layer < 2; ++layer) {
      const int gidx = 2 * i + layer;
      PpsGemmDesc& g = P.g[gidx];
      g.item_begin = items;
      g.mod = i;
      if (!present[i]) continue;
      const omni_pps_modality* m = mods[i];
      OMNI_CHECK_ARG(m->rate >= 1 && m->rate <= 256 && m->D > 0 && (m->D % 8) == 0 && m->n_tok >= 0);
      OMNI_CHECK_ARG((m->x_bs % 8) == 0 && m->x_bs >= static_cast<int64_t>(m->n_tok) * m->D);
      OMNI_CHECK_ARG(m->w1 && m->b1 && m->w2 && m->b2 && m->pooled && m->hidden);
      const int n = m->n_tok / m->rate;
      OMNI_CHECK_ARG(n == (i == 0 ? sp->n_a : sp->n_v));
      const long long Mll = static_cast<long long>(sp->B) * n;
      OMNI_CHECK_ARG(Mll < (1ll << 30));
      const int M = static_cast<int>(Mll);
      if (M == 0) continue;
      const int K1 = a->mode == OMNI_COMPRESS_AVG ? m->D : m->D * m->rate;
      g.M = M;
      g.m_pairs = cnt.pairs[i];
      if (layer == 0) {
        g.N = a->I;
        g.num_k_blocks = ceil_div(K1, BK);
        g.relu = 1;
        g.bias = reinterpret_cast<const bf16*>(m->b1);
        g.out = reinterpret_cast<bf16*>(m->hidden);
        g.ldo = a->I;
        g.wait_cnt = pooled_cnt[i];
        g.signal_cnt = hidden_cnt[i];
        rc = omni_make_tmap_2d_bf16(&maps.a[gidx], m->pooled, (uint64_t)M, (uint64_t)K1, (uint64_t)K1, BM, BK, 1);
        if (rc) return rc;
        rc = omni_make_tmap_2d_bf16(&maps.b[gidx], m->w1, (uint64_t)a->I, (uint64_t)K1, (uint64_t)K1, BM, BK, 1);
        if (rc) return rc;
      } else {
        g.N = H;
        g.num_k_blocks = ceil_div(a->I, BK);
        g.relu = 0;
        g.bias = reinterpret_cast<const bf16*>(m->b2);
        g.out = reinterpret_cast<bf16*>(m->tok);
        g.ldo = H;
        g.wait_cnt = hidden_cnt[i];
        g.wait_per_row = 0;
        g.wait_fixed = 2 * ceil_div(a->I, PPS_BN);                     // both CTAs of every GEMM-1 tile of the block
        g.signal_cnt = nullptr;
        rc = omni_make_tmap_2d_bf16(&maps.a[gidx], m->hidden, (uint64_t)M, (uint64_t)a->I, (uint64_t)a->I, BM, BK, 1);
        if (rc) return rc;
        rc = omni_make_tmap_2d_bf16(&maps.b[gidx], m->w2, (uint64_t)H, (uint64_t)a->I, (uint64_t)a->I, BM, BK, 1);
        if (rc) return rc;
      }
      g.n_tiles = ceil_div(g.N, PPS_BN);
      items += g.m_pairs * g.n_tiles;
    }
  }
  P.n_items = items;

  // compression plan per modality
  for (int i = 0; i < 2; ++i) {
    if (!present[i] || P.g[2 * i].M == 0) continue;
    const omni_pps_modality* m = mods[i];
    PpsPool& pl = P.pool[i];
    pl.M = P.g[2 * i].M;
    pl.n = m->n_tok / m->rate;
    pl.r = m->rate;
    pl.D = m->D;
    pl.K1 = a->mode == OMNI_COMPRESS_AVG ? m->D : m->D * m->rate;
    int cw = m->D < 256 ? m->D : 256;
    while (cw > 8 && static_cast<long long>(m->rate) * cw * 2 > PPS_SLOT_BYTES) cw = (cw / 2 + 7) / 8 * 8;
    if (static_cast<long long>(m->rate) * cw * 2 > PPS_SLOT_BYTES) return OMNI_ERR_UNSUPPORTED;
    pl.cw = cw;
    pl.lane_cols = (cw <= 128 && cw % 4 == 0 && m->D > 128) ? 4 : 8;
    pl.pow2 = (m->rate & (m->rate - 1)) == 0 ? 1 : 0;
    pl.fr = static_cast<float>(m->rate);
    pl.rcp = 1.0f / pl.fr;
    pl.n_boxes_row = ceil_div(m->D, cw);
    int bpu = 2048 / (m->rate * cw * 2);          // ~2 KB units: four ring slots per warp, enough bytes to amortise a unit
    if (bpu < 1) bpu = 1;
    if (bpu > pl.n_boxes_row) bpu = pl.n_boxes_row;
    {
      const int unit_bytes = (bpu * m->rate * cw * 2 + 127) / 128 * 128;
      if (unit_bytes > P.slot_bytes) P.slot_bytes = unit_bytes;
    }
    pl.boxes_per_unit = bpu;
    pl.units_per_row = ceil_div(pl.n_boxes_row, bpu);
    pl.units = static_cast<long long>(pl.M) * pl.units_per_row;
    pl.pooled = reinterpret_cast<bf16*>(m->pooled);
    pl.ready = pooled_cnt[i];
    P.g[2 * i].wait_per_row = pl.units_per_row;
    rc = make_tmap_pool(&maps.x[i], m->x, (uint64_t)sp->B, (uint64_t)m->n_tok, (uint64_t)m->D, (uint64_t)m->x_bs,
                        (uint32_t)m->rate, (uint32_t)cw);
    if (rc) return rc;
    // scatter destinations: own-task sequence (task i) and the AVSR sequence (task 2)
    PpsScatter& sc = P.sc[i];
    sc.n = pl.n;
    const int hb = P.splice.has_bos;
    sc.dst[0] = (P.splice.S[i] > 0) ? P.splice.out[i] : nullptr;
    sc.S[0] = P.splice.S[i];
    sc.pos0[0] = hb + 1;
    sc.dst[1] = (P.splice.S[2] > 0) ? P.splice.out[2] : nullptr;
    sc.S[1] = P.splice.S[2];
    sc.pos0[1] = hb + ((i == 1 && P.splice.has_a[2]) ? (P.splice.n_a + 2) : 0) + 1;
  }

  {
    // profiling switches (results are wrong with either): 1 = compression + row copies only, 2 = GEMMs only
    static const char* dbg = getenv("OMNI_PPS_ONLY");
    if (dbg && (dbg[0] == '1' || dbg[0] == '5')) P.n_items = items = 0;
    if (dbg && dbg[0] == '5') P.splice_rows = 0;                   // 5: compression only
    if (dbg && dbg[0] == '2') {
      P.pool[0].units = P.pool[1].units = 0;
      P.splice_rows = 0;
      for (int i = 0; i < 4; ++i) { P.g[i].wait_per_row = 0; P.g[i].wait_fixed = 0; }
    }
  }
  if (P.slot_bytes < 128) P.slot_bytes = 128;
  P.n_slots = (PPS_RING_BYTES / P.slot_bytes) >= 4 ? 4 : 2;
  if (P.n_slots * P.slot_bytes > PPS_RING_BYTES) return OMNI_ERR_UNSUPPORTED;
  if (cudaMemsetAsync(a->workspace, 0, static_cast<size_t>(cnt.total_ints) * 4, st) != cudaSuccess) return OMNI_ERR_CUDA;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(pool_project_splice_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PpsSmem::TOTAL) !=
        cudaSuccess)
      return OMNI_ERR_CUDA;
    attr_set = true;
  }
  const int sms = pps_sm_count();
  int clusters = sms / 2;
  // at least one CTA pair even when there is no GEMM work (pure splice); never more pairs than the SMs hold at once
  long long want = items;
  {
    const long long pool_work = (P.pool[0].units + P.pool[1].units + PPS_CHUNK - 1) / PPS_CHUNK + P.splice_rows / 8;
    const long long w2 = (pool_work + 2 * PPS_POOL_WARPS - 1) / (2 * PPS_POOL_WARPS);
    if (w2 > want) want = w2;
  }
  if (want < 1) want = 1;
  if (want < clusters) clusters = static_cast<int>(want);
  if (clusters < 1) clusters = 1;
  pool_project_splice_kernel<<<2 * clusters, PPS_THREADS, PpsSmem::TOTAL, st>>>(maps, P);
  OMNI_LAUNCH_CHECK();
  return OMNI_OK;
}
