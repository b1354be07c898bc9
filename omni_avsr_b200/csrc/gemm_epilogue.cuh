// Pieces of the tcgen05 GEMM shared by gemm_tcgen05.cu and pool_project_splice.cu: tile constants, the kernel
// parameter block, and the epilogue helpers (bias -> rounding point -> activation -> residual -> store; the coalesced
// 32x64 transposition tile of the CTA-pair kernels).
#pragma once
#include "common.cuh"
#include "../../include/omni_avsr.h"

namespace omni {

constexpr int BM = 128;
constexpr int BK = 64;           // 64 bf16 = 128 bytes = one swizzle-128B row
constexpr int UMMA_K = 16;
constexpr int GEMM_THREADS = 192;  // warp0 TMA, warp1 MMA(+TMEM alloc), warps 2..5 epilogue
constexpr int GEMM2_THREADS = 320; // CTA-pair kernel: warps 2..9 epilogue (two warps per TMEM lane quarter, 128 columns each)

struct GemmKParams {
  int M, N, K;
  int num_k_blocks;
  int n_tiles, m_tiles;
  int n_ext;
  const int* tile_group;
  const int* b_row_table;
  const int4* ext_table;  // {a2_col, b2_row, b2_col, unused}; b2_row < 0 => skip
  const bf16* bias;
  const bf16* residual;
  void* out;
  long long ldo, ldr;
  int act;
  int out_fp32;
  float alpha;
  bf16* out2;        // OMNI_ACT_SWIGLU64: [M, N/2]
  long long ldo2;
  // OMNI_ACT_PRELU_RING (ResNet BasicBlock epilogue on ring-padded frames)
  const bf16* slope;
  const bf16* res_bias;
  int ring_hp, ring_wp, ring_g, ring_c;
};

// Residual values of one 32-column chunk, fetched one chunk AHEAD of the TMEM read so that the (strided, 64 bytes per
// thread) loads are in flight while the previous chunk is converted and stored -- without this the epilogue is
// latency-bound on short-K GEMMs (8 serialized ~1 us round trips per 128x256 tile).
struct ResPrefetch {
  uint4 r[4];
  bool valid;
};
__device__ __forceinline__ void res_prefetch(const GemmKParams& p, int row, int col0, bool row_ok, ResPrefetch& o) {
  o.valid = p.residual != nullptr && row_ok && (col0 + 32 <= p.N);
  if (o.valid) {
    const uint4* rp4 = reinterpret_cast<const uint4*>(p.residual + static_cast<long long>(row) * p.ldr + col0);
#pragma unroll
    for (int i = 0; i < 4; ++i) o.r[i] = ld_nc_u4(rp4 + i);
  }
}

__device__ __forceinline__ float gelu_fast(float x) {
  // exact-erf GELU, x * Phi(x), with ONE special-function op: Phi(-|x|) = 0.5 * erfc(|x|/sqrt2) = 2^(q(z) - 1), where q is
  // a degree-7 fit of log2(erfc(z)) on z = |x|/sqrt2 in [0, 4.5] (relative error of the GELU value <= 7e-6 -- 300x below
  // bf16 resolution, also in the negative tail where Phi is tiny; the bf16-rounded result differs from the erff-based one
  // in 0.03 % of inputs, by one ulp).  The Abramowitz-Stegun form used before needed an exp AND a reciprocal: with 256
  // values per thread per tile the epilogue was MUFU-bound on the K = 1024 encoder GEMMs.
  const float z = fminf(fabsf(x) * 0.70710678118654752440f, 4.5f);
  float q = -2.045480869e-05f;
  q = fmaf(q, z, 4.882984795e-04f);
  q = fmaf(q, z, -5.237891804e-03f);
  q = fmaf(q, z, 3.395747021e-02f);
  q = fmaf(q, z, -1.525140703e-01f);
  q = fmaf(q, z, -9.170033932e-01f);
  q = fmaf(q, z, -1.628095627e+00f);
  q = fmaf(q, z, 3.904249297e-06f - 1.0f);
  const float h = ex2_approx(q);                 // Phi(-|x|)
  return x * (x < 0.f ? h : 1.0f - h);
}

// gelu_fast on two values with packed fp32 arithmetic (FMUL2 / FFMA2): the same operations per lane, hence the same bits, at
// ~9.5 instead of ~15 issue slots per element -- the GELU epilogues of the K = 1024 encoder GEMMs are issue-bound (two
// epilogue warps per sub-partition against a 4096-clk main loop).
__device__ __forceinline__ float2 gelu_fast2(float2 x) {
  float2 z = fmul2(make_float2(fabsf(x.x), fabsf(x.y)), make_float2(0.70710678118654752440f, 0.70710678118654752440f));
  z.x = fminf(z.x, 4.5f);
  z.y = fminf(z.y, 4.5f);
  auto c2 = [](float c) { return make_float2(c, c); };
  float2 q = c2(-2.045480869e-05f);
  q = ffma2(q, z, c2(4.882984795e-04f));
  q = ffma2(q, z, c2(-5.237891804e-03f));
  q = ffma2(q, z, c2(3.395747021e-02f));
  q = ffma2(q, z, c2(-1.525140703e-01f));
  q = ffma2(q, z, c2(-9.170033932e-01f));
  q = ffma2(q, z, c2(-1.628095627e+00f));
  q = ffma2(q, z, c2(3.904249297e-06f - 1.0f));
  const float h0 = ex2_approx(q.x), h1 = ex2_approx(q.y);
  return fmul2(x, make_float2(x.x < 0.f ? h0 : 1.0f - h0, x.y < 0.f ? h1 : 1.0f - h1));
}

// d gelu(x) / dx = Phi(x) + x * phi(x) with the same one-MUFU Phi as gelu_fast (plus one ex2 for the density): ~20 instructions
// against ~45 for erff + __expf -- the stand-alone backward kernel was compute-bound at 2x its HBM time.  Relative error of
// Phi <= 7e-6 (see gelu_fast), of phi the ex2.approx error (2^-22).
__device__ __forceinline__ float gelu_grad_fast(float x) {
  const float z = fminf(fabsf(x) * 0.70710678118654752440f, 4.5f);
  float q = -2.045480869e-05f;
  q = fmaf(q, z, 4.882984795e-04f);
  q = fmaf(q, z, -5.237891804e-03f);
  q = fmaf(q, z, 3.395747021e-02f);
  q = fmaf(q, z, -1.525140703e-01f);
  q = fmaf(q, z, -9.170033932e-01f);
  q = fmaf(q, z, -1.628095627e+00f);
  q = fmaf(q, z, 3.904249297e-06f - 1.0f);
  const float h = ex2_approx(q);                                   // Phi(-|x|)
  const float cdf = x < 0.f ? h : 1.0f - h;
  const float pdf = 0.39894228040143267794f * ex2_approx(-0.72134752044448170368f * x * x);     // exp(-x^2 / 2) / sqrt(2 pi)
  return fmaf(x, pdf, cdf);
}

// Epilogue of 32 accumulator columns of one row: alpha, bias, rounding point, activation, residual, store.
__device__ __forceinline__ void epilogue_store_32(const GemmKParams& p, float (&v)[32], int row, int col0,
                                                  const ResPrefetch* pre = nullptr) {
  const bool full = (col0 + 32 <= p.N);
  if (p.bias) {
    if (full) {
      const uint4* bp = reinterpret_cast<const uint4*>(p.bias + col0);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint4 b = __ldg(bp + i);
        const uint32_t bw[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {               // packed fp32 adds (FADD2)
          const float2 r = fadd2(make_float2(v[8 * i + 2 * j], v[8 * i + 2 * j + 1]), bf2_to_f2(bw[j]));
          v[8 * i + 2 * j] = r.x;
          v[8 * i + 2 * j + 1] = r.y;
        }
      }
    } else {
      for (int i = 0; i < 32; ++i)
        if (col0 + i < p.N) v[i] += __bfloat162float(p.bias[col0 + i]);
    }
  }
  if (p.act != OMNI_ACT_NONE || p.residual) {
    // reference rounding point: the linear's bf16 output feeds the activation / the residual add
#pragma unroll
    for (int i = 0; i < 32; i += 2) {
      const float2 f = bf2_to_f2(f2_to_bf2(v[i], v[i + 1]));
      v[i] = f.x;
      v[i + 1] = f.y;
    }
    if (p.act == OMNI_ACT_RELU) {
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.0f);
    } else if (p.act == OMNI_ACT_GELU) {
#pragma unroll
      for (int i = 0; i < 32; i += 2) {
        const float2 g = gelu_fast2(make_float2(v[i], v[i + 1]));
        const float2 f = bf2_to_f2(f2_to_bf2(g.x, g.y));
        v[i] = f.x;
        v[i + 1] = f.y;
      }
    }
  }
  if (p.residual) {
    const bf16* rp = p.residual + static_cast<long long>(row) * p.ldr + col0;
    if (full) {
      const uint4* rp4 = reinterpret_cast<const uint4*>(rp);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        uint4 b = (pre && pre->valid) ? pre->r[i] : ld_nc_u4(rp4 + i);
        float2 f;
        f = bf2_to_f2(b.x); v[8 * i + 0] += f.x; v[8 * i + 1] += f.y;
        f = bf2_to_f2(b.y); v[8 * i + 2] += f.x; v[8 * i + 3] += f.y;
        f = bf2_to_f2(b.z); v[8 * i + 4] += f.x; v[8 * i + 5] += f.y;
        f = bf2_to_f2(b.w); v[8 * i + 6] += f.x; v[8 * i + 7] += f.y;
      }
    } else {
      for (int i = 0; i < 32; ++i)
        if (col0 + i < p.N) v[i] += __bfloat162float(rp[i]);
    }
  }
  if (p.out_fp32) {
    float* op = reinterpret_cast<float*>(p.out) + static_cast<long long>(row) * p.ldo + col0;
    if (full) {
#pragma unroll
      for (int i = 0; i < 8; ++i)
        *reinterpret_cast<float4*>(op + 4 * i) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
    } else {
      for (int i = 0; i < 32; ++i)
        if (col0 + i < p.N) op[i] = v[i];
    }
  } else {
    bf16* op = reinterpret_cast<bf16*>(p.out) + static_cast<long long>(row) * p.ldo + col0;
    if (full) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        uint4 o;
        o.x = f2_to_bf2(v[8 * i + 0], v[8 * i + 1]);
        o.y = f2_to_bf2(v[8 * i + 2], v[8 * i + 3]);
        o.z = f2_to_bf2(v[8 * i + 4], v[8 * i + 5]);
        o.w = f2_to_bf2(v[8 * i + 6], v[8 * i + 7]);
        *reinterpret_cast<uint4*>(op + 8 * i) = o;
      }
    } else {
      for (int i = 0; i < 32; ++i)
        if (col0 + i < p.N) op[i] = __float2bfloat16_rn(v[i]);
    }
  }
}


// ---- coalesced epilogue of the CTA-pair kernel ------------------------------------------------------------------
// A thread owns one accumulator row (TMEM lane), so direct global accesses are 16 bytes per thread at the row pitch:
// 32 different 128-byte lines per warp instruction, and the LSU retires about one line per clock -- on short-K GEMMs
// (K = 1024: 8192 clk of MMA per 256x256 tile) the residual loads + output stores alone cost as much as the main loop.
// Each epilogue warp therefore transposes 32 rows x 64 columns through a private padded shared-memory tile (pitch
// 144 B: conflict-free for the 16-byte row writes and for the 8-lanes-per-row reads), so that global loads / stores are
// full 128-byte row segments (4 lines per warp instruction).
constexpr int EPI_PITCH = 144;
constexpr int EPI_WARP_BYTES = 32 * EPI_PITCH;

struct ResTile {            // residual of a 32-row x 64-column block in the coalesced layout: lane -> (row l/8 + 4*pass, 16 B l%8)
  uint4 r[8];
};
__device__ __forceinline__ void res_tile_prefetch(const GemmKParams& p, int row0, int col0, int lane, ResTile& o) {
  const bf16* rp = p.residual + static_cast<long long>(row0 + (lane >> 3)) * p.ldr + col0 + (lane & 7) * 8;
#pragma unroll
  for (int pass = 0; pass < 8; ++pass) {
    const bool ok = (row0 + (lane >> 3) + 4 * pass) < p.M;
    o.r[pass] = ok ? ld_nc_u4(rp + static_cast<long long>(4 * pass) * p.ldr) : make_uint4(0, 0, 0, 0);
  }
}

// publish a prefetched residual block into the warp's transposition tile (frees its registers for the next prefetch)
__device__ __forceinline__ void res_tile_publish(uint8_t* stage, int lane, const ResTile& res) {
  uint8_t* co_ptr = stage + (lane >> 3) * EPI_PITCH + (lane & 7) * 16;
#pragma unroll
  for (int pass = 0; pass < 8; ++pass) *reinterpret_cast<uint4*>(co_ptr + 4 * pass * EPI_PITCH) = res.r[pass];
  __syncwarp();
}

// 32 rows x 64 bf16 columns held one row per thread -> global memory as full 128-byte rows through the warp's padded
// transposition tile.
__device__ __forceinline__ void store_tile64(bf16* out, long long ldo, int M, uint8_t* stage, const float (&v)[64], int lane,
                                             int row0, int col0) {
  uint8_t* my_row = stage + lane * EPI_PITCH;
  const uint8_t* co_ptr = stage + (lane >> 3) * EPI_PITCH + (lane & 7) * 16;
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
  bf16* op = out + static_cast<long long>(row0 + (lane >> 3)) * ldo + col0 + (lane & 7) * 8;
#pragma unroll
  for (int pass = 0; pass < 8; ++pass) {
    if (row0 + (lane >> 3) + 4 * pass < M)
      *reinterpret_cast<uint4*>(op + static_cast<long long>(4 * pass) * ldo) =
          *reinterpret_cast<const uint4*>(co_ptr + 4 * pass * EPI_PITCH);
  }
  __syncwarp();
}

__device__ __forceinline__ void epilogue_tile64(const GemmKParams& p, uint8_t* stage, float (&v)[64], int lane, int row0,
                                                int col0, bool res_staged, bool store = true) {
  if (p.bias) {
    const uint4* bp = reinterpret_cast<const uint4*>(p.bias + col0);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const uint4 b = __ldg(bp + i);
      const uint32_t bw[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {                 // packed fp32 adds (FADD2): same sums, half the issue slots
        const float2 r = fadd2(make_float2(v[8 * i + 2 * j], v[8 * i + 2 * j + 1]), bf2_to_f2(bw[j]));
        v[8 * i + 2 * j] = r.x;
        v[8 * i + 2 * j + 1] = r.y;
      }
    }
  }
  if (p.act != OMNI_ACT_NONE || p.residual) {
#pragma unroll
    for (int i = 0; i < 64; i += 2) {               // one packed conversion per pair (same round-to-nearest-even)
      const float2 f = bf2_to_f2(f2_to_bf2(v[i], v[i + 1]));
      v[i] = f.x;
      v[i + 1] = f.y;
    }
    if (p.act == OMNI_ACT_RELU) {
#pragma unroll
      for (int i = 0; i < 64; ++i) v[i] = fmaxf(v[i], 0.0f);
    } else if (p.act == OMNI_ACT_GELU) {
#pragma unroll
      for (int i = 0; i < 64; i += 2) {
        const float2 g = gelu_fast2(make_float2(v[i], v[i + 1]));
        const float2 f = bf2_to_f2(f2_to_bf2(g.x, g.y));
        v[i] = f.x;
        v[i + 1] = f.y;
      }
    }
  }
  uint8_t* my_row = stage + lane * EPI_PITCH;
  uint8_t* co_ptr = stage + (lane >> 3) * EPI_PITCH + (lane & 7) * 16;
  if (res_staged) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const uint4 b = *reinterpret_cast<const uint4*>(my_row + 16 * i);
      const uint32_t bw[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 r = fadd2(make_float2(v[8 * i + 2 * j], v[8 * i + 2 * j + 1]), bf2_to_f2(bw[j]));
        v[8 * i + 2 * j] = r.x;
        v[8 * i + 2 * j + 1] = r.y;
      }
    }
    __syncwarp();
  }
  if (store) store_tile64(reinterpret_cast<bf16*>(p.out), p.ldo, p.M, stage, v, lane, row0, col0);
}


// 32 rows x 64 bf16 columns already packed (32 words per thread = one row) -> global memory, through the transposition tile
// the warp's transposition tile (every thread has written its own 128-byte row) -> global memory as full 128-byte rows
__device__ __forceinline__ void stage_to_global(bf16* out, long long ldo, int M, const uint8_t* stage, int lane, int row0,
                                                int col0) {
  const uint8_t* co_ptr = stage + (lane >> 3) * EPI_PITCH + (lane & 7) * 16;
  __syncwarp();
  bf16* op = out + static_cast<long long>(row0 + (lane >> 3)) * ldo + col0 + (lane & 7) * 8;
#pragma unroll
  for (int pass = 0; pass < 8; ++pass) {
    if (row0 + (lane >> 3) + 4 * pass < M)
      *reinterpret_cast<uint4*>(op + static_cast<long long>(4 * pass) * ldo) =
          *reinterpret_cast<const uint4*>(co_ptr + 4 * pass * EPI_PITCH);
  }
  __syncwarp();
}
__device__ __forceinline__ void store_tile64_packed(bf16* out, long long ldo, int M, uint8_t* stage, const uint32_t (&w)[32],
                                                    int lane, int row0, int col0) {
  uint8_t* my_row = stage + lane * EPI_PITCH;
#pragma unroll
  for (int i = 0; i < 8; ++i)
    *reinterpret_cast<uint4*>(my_row + 16 * i) = make_uint4(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
  stage_to_global(out, ldo, M, stage, lane, row0, col0);
}

// a 32-row x 64-column bf16 block of `src` (row pitch ld) in the coalesced lane layout of ResTile
__device__ __forceinline__ void tile_prefetch(const bf16* src, long long ld, int M, int row0, int col0, int lane, ResTile& o) {
  const bf16* rp = src + static_cast<long long>(row0 + (lane >> 3)) * ld + col0 + (lane & 7) * 8;
#pragma unroll
  for (int pass = 0; pass < 8; ++pass) {
    const bool ok = (row0 + (lane >> 3) + 4 * pass) < M;
    o.r[pass] = ok ? ld_nc_u4(rp + static_cast<long long>(4 * pass) * ld) : make_uint4(0, 0, 0, 0);
  }
}

// publish a prefetched block and read back this thread's own row (64 bf16 = 32 packed words)
__device__ __forceinline__ void tile_to_row(uint8_t* stage, int lane, const ResTile& t, uint32_t (&w)[32]) {
  res_tile_publish(stage, lane, t);
  const uint8_t* my_row = stage + lane * EPI_PITCH;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const uint4 b = *reinterpret_cast<const uint4*>(my_row + 16 * i);
    w[4 * i] = b.x; w[4 * i + 1] = b.y; w[4 * i + 2] = b.z; w[4 * i + 3] = b.w;
  }
  __syncwarp();
}

// OMNI_ACT_SWIGLU_BWD64, 32 of a block's 64 intermediate channels (h = 0 / 1): r = the accumulator columns (fp32 bits), g / u
// = the saved gate block of the row (packed bf16; the up block is read from the transposition tile) -> du words 16 h .. 16 h + 15
// stay in registers, the d(gate) words go straight into the thread's row of the transposition tile.  Same arithmetic as swiglu_bwd_kernel on the GEMM's bf16-rounded
// output, with the hardware reciprocal for the sigmoid (<= 1 ulp before the bf16 rounding).
__device__ __forceinline__ void swiglu_bwd_half(const uint32_t (&r)[32], float alpha, const uint32_t (&g)[32], int h,
                                                uint8_t* my_row, uint32_t (&du)[32]) {
  const float2 al2 = make_float2(alpha, alpha), one2 = make_float2(1.0f, 1.0f), m1 = make_float2(-1.0f, -1.0f);
#pragma unroll
  for (int i4 = 0; i4 < 4; ++i4) {
    // the up block of the row sits in the thread's row of the transposition tile; each 16-byte piece is replaced in place
    // by the d(gate) words of the same channels
    const uint4 ub = *reinterpret_cast<const uint4*>(my_row + 64 * h + 16 * i4);
    const uint32_t u[4] = {ub.x, ub.y, ub.z, ub.w};
    uint32_t dgw[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      // two channels per step, packed fp32 (FMUL2 / FFMA2): the epilogue is issue-bound, this halves its arithmetic slots
      const int i = i4 * 4 + e, w = h * 16 + i;
      const float2 gf = bf2_to_f2(g[w]), uf = bf2_to_f2(u[e]);
      const float2 d = bf2_to_f2(f2_to_bf2_pair(fmul2(make_float2(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1])), al2)));
      const float2 s = sigmoid2(gf);
      du[w] = f2_to_bf2_pair(fmul2(d, fmul2(gf, s)));                                        // d * silu(g)
      const float2 t = fmul2(s, ffma2(gf, ffma2(s, m1, one2), one2));                        // s * (1 + g * (1 - s))
      dgw[e] = f2_to_bf2_pair(fmul2(fmul2(d, uf), t));                                       // d * u * silu'(g)
    }
    *reinterpret_cast<uint4*>(my_row + 64 * h + 16 * i4) = make_uint4(dgw[0], dgw[1], dgw[2], dgw[3]);
  }
}

// OMNI_ACT_GELU_BWD: v = d(act) (one row per thread, 64 columns), x = the saved pre-activation -> dx packed bf16 (same
// arithmetic as gelu_bwd_kernel on the GEMM's bf16-rounded output)
__device__ __forceinline__ void gelu_bwd_row64(const float (&v)[64], const uint32_t (&x)[32], uint32_t (&dx)[32]) {
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    const float2 xf = bf2_to_f2(x[i]);
    const float d0 = __bfloat162float(__float2bfloat16_rn(v[2 * i]));
    const float d1 = __bfloat162float(__float2bfloat16_rn(v[2 * i + 1]));
    dx[i] = f2_to_bf2(d0 * gelu_grad_fast(xf.x), d1 * gelu_grad_fast(xf.y));
  }
}

// OMNI_ACT_PRELU_RING on a 32-row x 64-column block held one row per thread: f = bf16(bf16(acc) + bias); with a residual
// (staged in the warp's transposition tile) f = bf16(f + bf16(res + res_bias)); PReLU; ring pixels -> 0 (same arithmetic and
// rounding points as prelu_res_ring_kernel applied to the GEMM's bf16 output).
__device__ __forceinline__ void epilogue_tile64_prelu_ring(const GemmKParams& p, uint8_t* stage, float (&v)[64], int lane,
                                                           int row0, int col0, bool res_staged) {
  const int ch0 = col0 % p.ring_c;                        // 64-column blocks never straddle a pixel (ring_c % 64 == 0)
  const long long pix = static_cast<long long>(row0 + lane) * p.ring_g + col0 / p.ring_c;
  const int px = static_cast<int>(pix % p.ring_wp);
  const int py = static_cast<int>((pix / p.ring_wp) % p.ring_hp);
  // ring_h == 0 (ring_hp == 2): plain frames without a ring (ops.conv_frames), nothing is zeroed
  const bool ring = p.ring_hp > 2 && (px == 0 || py == 0 || px == p.ring_wp - 1 || py == p.ring_hp - 1);
  uint8_t* my_row = stage + lane * EPI_PITCH;
  const uint4* bp = reinterpret_cast<const uint4*>(p.bias + ch0);
  const uint4* sp = reinterpret_cast<const uint4*>(p.slope + ch0);
  const uint4* rbp = p.res_bias ? reinterpret_cast<const uint4*>(p.res_bias + ch0) : nullptr;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const uint4 b = __ldg(bp + i);
    const uint4 s = __ldg(sp + i);
    const uint32_t bs[4] = {b.x, b.y, b.z, b.w};
    const uint32_t ss[4] = {s.x, s.y, s.z, s.w};
    uint32_t rs[4] = {0u, 0u, 0u, 0u}, rbs[4] = {0u, 0u, 0u, 0u};
    if (res_staged) {
      const uint4 r = *reinterpret_cast<const uint4*>(my_row + 16 * i);
      rs[0] = r.x; rs[1] = r.y; rs[2] = r.z; rs[3] = r.w;
      if (rbp) {
        const uint4 rb = __ldg(rbp + i);
        rbs[0] = rb.x; rbs[1] = rb.y; rbs[2] = rb.z; rbs[3] = rb.w;
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      // two channels per step: packed fp32 adds / multiply and one packed bf16 conversion per rounding point
      auto rb2 = [](float2 x) { return bf2_to_f2(f2_to_bf2_pair(x)); };
      const float2 sf = bf2_to_f2(ss[j]);
      float2 f = rb2(make_float2(v[8 * i + 2 * j], v[8 * i + 2 * j + 1]));
      f = rb2(fadd2(f, bf2_to_f2(bs[j])));
      if (res_staged) {
        float2 rf = bf2_to_f2(rs[j]);
        if (rbp) rf = rb2(fadd2(rf, bf2_to_f2(rbs[j])));
        f = rb2(fadd2(f, rf));
      }
      const float2 neg = fmul2(f, sf);
      v[8 * i + 2 * j] = ring ? 0.f : (f.x > 0.f ? f.x : neg.x);
      v[8 * i + 2 * j + 1] = ring ? 0.f : (f.y > 0.f ? f.y : neg.y);
    }
  }
  if (res_staged) __syncwarp();
  store_tile64(reinterpret_cast<bf16*>(p.out), p.ldo, p.M, stage, v, lane, row0, col0);
}

// Tile order of the persistent kernels.  m_fast = 0: N fastest (an M tile's row of output tiles is contiguous in time);
// 1: M fastest (small A, huge B: lm_head); >= 2: N fastest inside GROUPS of m_fast N tiles, all M tiles per group -- the
// group's B panel (m_fast * BN * K * 2 bytes, sized to stay L2-resident under the output stream) is read from HBM once
// instead of once per M wave.
__device__ __forceinline__ void tile_coords(int t, int m_tiles, int n_tiles, int m_fast, int& m_tile, int& n_tile) {
  if (m_fast >= 2) {
    const int per_group = m_fast * m_tiles;
    const int g = t / per_group;
    const int r = t - g * per_group;
    const int n0 = g * m_fast;
    const int w = min(m_fast, n_tiles - n0);
    m_tile = r / w;
    n_tile = n0 + r - m_tile * w;
  } else if (m_fast) {
    n_tile = t / m_tiles;
    m_tile = t - n_tile * m_tiles;
  } else {
    m_tile = t / n_tiles;
    n_tile = t - m_tile * n_tiles;
  }
}

}  // namespace omni
