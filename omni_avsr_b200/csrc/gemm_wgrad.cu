// Weight-gradient GEMM on the tcgen05 tensor cores:  out[z][i, j] = alpha * sum_{k in [k0_z, k1_z)} A[k, a_col0+i] * B[k, b_col0+j]
//
// Both operands are token-major activations ([tokens, features] row-major), i.e. the reduction runs over the SLOW
// dimension: the tiles are TMA-loaded as [64 tokens x 64 features] 128B-swizzled boxes and consumed by tcgen05.mma as
// MN-major operands (instruction-descriptor a_major = b_major = 1), so no transposed copy of dY or X is ever made.
// grid.z enumerates independent token ranges (the task runs of the packed rows): the per-task LoRA gradients
// (reference: autograd of Llama_LoRA.py:246-259) and the projector gradients come out of one launch each.
#include "common.cuh"
#include "../../include/omni_avsr.h"

namespace omni {

constexpr int WG_BM = 128;
constexpr int WG_BK = 64;          // tokens per pipeline stage
constexpr int WG_THREADS = 192;

struct WgradKParams {
  int Mo, No;
  int k0[OMNI_WGRAD_MAX_RANGES], k1[OMNI_WGRAD_MAX_RANGES];
  int a_col0, b_col0;
  void* out;
  long long ldo, out_zstride;
  int out_fp32;
  int accumulate;
  float alpha;
};

template <int BN, int STAGES>
struct WgradSmem {
  static constexpr int A_BYTES = WG_BM * WG_BK * 2;   // 2 boxes of [64 k][64 mn]
  static constexpr int B_BYTES = BN * WG_BK * 2;      // BN/64 boxes
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BAR_OFFSET = STAGES * STAGE_BYTES;
  static constexpr int TOTAL = BAR_OFFSET + (2 * STAGES + 1) * 8 + 16 + 1024;
};

template <int BN, int STAGES>
__global__ void __launch_bounds__(WG_THREADS, 1)
gemm_bf16_wgrad_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                       const WgradKParams p) {
  using S = WgradSmem<BN, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::BAR_OFFSET);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * BN;
  const int m0 = blockIdx.y * WG_BM;
  const int z = blockIdx.z;
  const int kbeg = p.k0[z];
  const int num_k_blocks = (p.k1[z] - kbeg + WG_BK - 1) / WG_BK;

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
    tmem_alloc(tmem_ptr_smem, BN < 32 ? 32 : BN);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (num_k_blocks > 0) {
    if (warp == 0) {
      if (elect_one()) {
        int stage = 0;
        uint32_t phase = 0;
        for (int it = 0; it < num_k_blocks; ++it) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sA = smem + stage * S::STAGE_BYTES;
          uint8_t* sB = sA + S::A_BYTES;
          mbar_expect_tx(&full_bar[stage], S::STAGE_BYTES);
          const int krow = kbeg + it * WG_BK;
#pragma unroll
          for (int c = 0; c < WG_BM / 64; ++c)
            tma_load_2d(&tmA, &full_bar[stage], sA + c * 8192, p.a_col0 + m0 + c * 64, krow);
#pragma unroll
          for (int c = 0; c < BN / 64; ++c)
            tma_load_2d(&tmB, &full_bar[stage], sB + c * 8192, p.b_col0 + n0 + c * 64, krow);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    } else if (warp == 1) {
      constexpr uint32_t idesc = make_idesc_bf16(WG_BM, BN, 1, 1);   // both operands MN-major
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0; it < num_k_blocks; ++it) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sA = smem_u32(smem + stage * S::STAGE_BYTES);
          const uint32_t sB = sA + S::A_BYTES;
          // MN-major, 128B swizzle: LBO = stride between 64-wide MN chunks (one TMA box, 8 KB),
          //                         SBO = stride between groups of 8 K rows (1 KB)
          const uint64_t adesc = make_smem_desc_sw128(sA, 8192, 1024);
          const uint64_t bdesc = make_smem_desc_sw128(sB, 8192, 1024);
#pragma unroll
          for (int k = 0; k < WG_BK / 16; ++k) {
            // 16 K rows = 2048 bytes further down the box: +128 in the (addr >> 4) field
            umma_bf16(tmem_base, adesc + 128 * k, bdesc + 128 * k, idesc, (it > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);
          if (it == num_k_blocks - 1) umma_commit(tmem_full_bar);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  }
  if (warp >= 2) {
    const int q = warp & 3;
    const int row = m0 + q * 32 + lane;
    if (num_k_blocks > 0) {
      mbar_wait(tmem_full_bar, 0);
      tc_fence_after();
    }
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      uint32_t r[32];
      if (num_k_blocks > 0) {
        tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(c * 32), r);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) r[i] = 0u;
      }
      const int col0 = n0 + c * 32;
      if (row >= p.Mo || col0 >= p.No) continue;
      const long long off = static_cast<long long>(z) * p.out_zstride + static_cast<long long>(row) * p.ldo + col0;
      if (p.out_fp32) {
        float* op = reinterpret_cast<float*>(p.out) + off;
        for (int i = 0; i < 32; ++i) {
          if (col0 + i < p.No) {
            const float v = __uint_as_float(r[i]) * p.alpha;
            op[i] = p.accumulate ? op[i] + v : v;
          }
        }
      } else {
        bf16* op = reinterpret_cast<bf16*>(p.out) + off;
        for (int i = 0; i < 32; ++i) {
          if (col0 + i < p.No) {
            float v = __uint_as_float(r[i]) * p.alpha;
            if (p.accumulate) v += __bfloat162float(op[i]);
            op[i] = __float2bfloat16_rn(v);
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, BN < 32 ? 32 : BN);
}

template <int BN, int STAGES>
static int launch_wgrad(const omni_wgrad_args* a, cudaStream_t stream) {
  using S = WgradSmem<BN, STAGES>;
  CUtensorMap tmA, tmB;
  int rc = omni_make_tmap_2d_bf16(&tmA, a->A, (uint64_t)a->K, (uint64_t)a->a_cols, (uint64_t)a->lda, 64, 64, 1);
  if (rc) return rc;
  rc = omni_make_tmap_2d_bf16(&tmB, a->B, (uint64_t)a->K, (uint64_t)a->b_cols, (uint64_t)a->ldb, 64, 64, 1);
  if (rc) return rc;
  WgradKParams p;
  p.Mo = a->Mo; p.No = a->No;
  for (int i = 0; i < OMNI_WGRAD_MAX_RANGES; ++i) { p.k0[i] = a->k0[i]; p.k1[i] = a->k1[i]; }
  p.a_col0 = a->a_col0; p.b_col0 = a->b_col0;
  p.out = a->out; p.ldo = a->ldo; p.out_zstride = a->out_zstride;
  p.out_fp32 = a->out_fp32; p.accumulate = a->accumulate; p.alpha = a->alpha;
  auto kfn = gemm_bf16_wgrad_kernel<BN, STAGES>;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL) != cudaSuccess)
      return OMNI_ERR_CUDA;
    attr_set = true;
  }
  dim3 grid(ceil_div(a->No, BN), ceil_div(a->Mo, WG_BM), a->n_ranges);
  kfn<<<grid, WG_THREADS, S::TOTAL, stream>>>(tmA, tmB, p);
  OMNI_LAUNCH_CHECK();
  return OMNI_OK;
}

// column sums: out[j] = sum_k x[k, j]  (bias gradients), fp32 accumulate, bf16 result
__global__ void __launch_bounds__(256)
colsum_kernel(const bf16* __restrict__ x, bf16* __restrict__ out, long long rows, int cols, long long ld) {
  // block = 32 columns x 8 row-lanes
  __shared__ float sh[8][33];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int col = blockIdx.x * 32 + cx;
  float s = 0.f;
  if (col < cols)
    for (long long r = ry; r < rows; r += 8) s += __bfloat162float(x[r * ld + col]);
  sh[ry][cx] = s;
  __syncthreads();
  if (ry == 0 && col < cols) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += sh[i][cx];
    out[col] = __float2bfloat16_rn(t);
  }
}

}  // namespace omni

extern "C" int omni_gemm_wgrad_bf16(const omni_wgrad_args* a, void* stream) {
  using namespace omni;
  OMNI_CHECK_ARG(a && a->A && a->B && a->out);
  OMNI_CHECK_ARG(a->Mo > 0 && a->No > 0 && a->K > 0 && a->n_ranges >= 1 && a->n_ranges <= OMNI_WGRAD_MAX_RANGES);
  OMNI_CHECK_ARG((a->lda % 8) == 0 && (a->ldb % 8) == 0);
  OMNI_CHECK_ARG((reinterpret_cast<uintptr_t>(a->A) & 15) == 0 && (reinterpret_cast<uintptr_t>(a->B) & 15) == 0);
  OMNI_CHECK_ARG(a->a_col0 >= 0 && a->b_col0 >= 0 && a->a_col0 + a->Mo <= a->a_cols && a->b_col0 + a->No <= a->b_cols);
  for (int i = 0; i < a->n_ranges; ++i) OMNI_CHECK_ARG(a->k0[i] >= 0 && a->k1[i] >= a->k0[i] && a->k1[i] <= a->K);
  // a range that does not end on a 64-token boundary must be the last tokens of the tensor (TMA zero-fill), otherwise
  // the tail box would pull in rows of the next range
  for (int i = 0; i < a->n_ranges; ++i)
    OMNI_CHECK_ARG(((a->k1[i] - a->k0[i]) % WG_BK) == 0 || a->k1[i] == a->K);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (a->No <= 64) return launch_wgrad<64, 6>(a, st);
  if (a->No <= 128) return launch_wgrad<128, 6>(a, st);
  return launch_wgrad<256, 4>(a, st);
}

extern "C" int omni_colsum_bf16(const void* x, void* out, int64_t rows, int32_t cols, int64_t ld, void* stream) {
  using namespace omni;
  OMNI_CHECK_ARG(x && out && rows >= 0 && cols > 0 && ld >= cols);
  colsum_kernel<<<ceil_div(cols, 32), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const bf16*>(x), reinterpret_cast<bf16*>(out), rows, cols, ld);
  OMNI_LAUNCH_CHECK();
  return OMNI_OK;
}
