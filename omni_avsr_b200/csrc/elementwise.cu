// HBM-bound row kernels of the LLM / encoder blocks: RMSNorm, LayerNorm, RoPE, SwiGLU, GELU, row gather.
// One warp owns one row (16-byte vector accesses, lanes interleaved over 16-byte chunks), reductions by
// shuffles; every kernel mirrors the bf16 rounding points of the reference's unfused torch op sequence.
//
// Reference semantics (third-party transformers==4.43.1 classes used by Omni_AVSR/Llama_LoRA.py:12, Qwen_LoRA.py:7):
//   LlamaRMSNorm / Qwen2RMSNorm, apply_rotary_pos_emb (Llama_LoRA.py:277), LlamaMLP (SwiGLU),
//   fairseq LayerNorm + gelu (av_hubert/fairseq/fairseq/modules/{layer_norm,gelu}.py), WhisperEncoderLayer LN/GELU.
#include "gemm_epilogue.cuh"      // gelu_fast / gelu_grad_fast (shared with the GEMM epilogues: same bits)

namespace omni {

constexpr int EW_THREADS = 256;
constexpr int EW_WARPS = EW_THREADS / 32;

static int rows_grid(long long rows) {
  long long blocks = ceil_div_ll(rows, EW_WARPS);
  const long long cap = static_cast<long long>(kNumSMs) * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return static_cast<int>(blocks);
}

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  float2 t;
  t = bf2_to_f2(u.x); f[0] = t.x; f[1] = t.y;
  t = bf2_to_f2(u.y); f[2] = t.x; f[3] = t.y;
  t = bf2_to_f2(u.z); f[4] = t.x; f[5] = t.y;
  t = bf2_to_f2(u.w); f[6] = t.x; f[7] = t.y;
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 o;
  o.x = f2_to_bf2(f[0], f[1]); o.y = f2_to_bf2(f[2], f[3]);
  o.z = f2_to_bf2(f[4], f[5]); o.w = f2_to_bf2(f[6], f[7]);
  return o;
}
__device__ __forceinline__ float rbf(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

// ------------------------------------------------------------------------------------------------
// Norm kernels.  NPL = 16-byte chunks per lane (row width = 256 * NPL elements: 4 -> 1024, 8 -> 2048, 16 -> 4096): the
// row (and dy) is loaded ONCE with all NPL (2 NPL) loads of a lane in flight and kept in registers for the statistics
// and the output pass.  NPL = 0 is the generic width (runtime loops, the row is re-read from L1/L2 for each pass).  Both
// forms accumulate in the same order (chunk c = lane, lane + 32, ...), so they produce identical bits.
// ------------------------------------------------------------------------------------------------
template <int NPL>
struct RowRegs {
  uint4 v[NPL > 0 ? NPL : 1];
  __device__ __forceinline__ void load(const uint4* row, int lane) {
#pragma unroll
    for (int k = 0; k < NPL; ++k) v[k] = __ldg(row + lane + 32 * k);
  }
  // Opaque to the optimiser: the next pass has to unpack the bf16 pairs again instead of keeping 8 fp32 registers per
  // chunk alive across the warp reduction (which would not fit in the register file).
  __device__ __forceinline__ void pin() {
#pragma unroll
    for (int k = 0; k < NPL; ++k) asm volatile("" : "+r"(v[k].x), "+r"(v[k].y), "+r"(v[k].z), "+r"(v[k].w));
  }
};
#define OMNI_ROW_LOOP(k, c) \
  _Pragma("unroll") for (int k = 0, c = lane; NPL > 0 ? k < NPL : c < H8; ++k, c += 32)

// RMSNorm:  y = w * bf16(x * rsqrt(mean(x^2) + eps))     (fp32 statistics)
template <int NPL>
__global__ void __launch_bounds__(EW_THREADS, 1)
rmsnorm_fwd_kernel(const bf16* __restrict__ x, const bf16* __restrict__ w, bf16* __restrict__ y,
                   float* __restrict__ rstd_out, long long rows, int H8, long long ldx, long long ldy, float eps) {
  pdl_launch_dependents();
  pdl_wait();                                   // x is the predecessor's output (no-op without the launch attribute)
  const int lane = threadIdx.x & 31;
  const long long wg = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
  const long long nw = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  for (long long r = wg; r < rows; r += nw) {
    const uint4* xr = reinterpret_cast<const uint4*>(x + r * ldx);
    RowRegs<NPL> X;
    X.load(xr, lane);
    float ss = 0.f;
    OMNI_ROW_LOOP(k, c) {
      float f[8];
      unpack8(NPL > 0 ? X.v[NPL > 0 ? k : 0] : __ldg(xr + c), f);
#pragma unroll
      for (int i = 0; i < 8; ++i) ss += f[i] * f[i];
    }
    ss = warp_sum(ss);
    X.pin();
    const float rstd = rsqrtf(ss / static_cast<float>(H8 * 8) + eps);
    if (lane == 0 && rstd_out) rstd_out[r] = rstd;
    uint4* yr = reinterpret_cast<uint4*>(y + r * ldy);
    OMNI_ROW_LOOP(k, c) {
      float f[8], g[8];
      unpack8(NPL > 0 ? X.v[NPL > 0 ? k : 0] : __ldg(xr + c), f);
      unpack8(__ldg(reinterpret_cast<const uint4*>(w) + c), g);
#pragma unroll
      for (int i = 0; i < 8; ++i) f[i] = g[i] * rbf(f[i] * rstd);
      yr[c] = pack8(f);
    }
  }
}

// dx = rstd * (g - xhat * mean(g * xhat)),  g = dy * w, xhat = x * rstd   (weights are frozen: no dw)
template <int NPL>
__global__ void __launch_bounds__(EW_THREADS, 1)
rmsnorm_bwd_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ x, const bf16* __restrict__ w,
                   const float* __restrict__ rstd_in, bf16* __restrict__ dx, const bf16* __restrict__ dx_add,
                   long long rows, int H8) {
  const int lane = threadIdx.x & 31;
  const long long wg = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
  const long long nw = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  const float invH = 1.0f / static_cast<float>(H8 * 8);
  for (long long r = wg; r < rows; r += nw) {
    const uint4* xr = reinterpret_cast<const uint4*>(x) + r * H8;
    const uint4* dr = reinterpret_cast<const uint4*>(dy) + r * H8;
    const uint4* ar = dx_add ? reinterpret_cast<const uint4*>(dx_add) + r * H8 : nullptr;
    RowRegs<NPL> X, D, A;
    X.load(xr, lane);
    D.load(dr, lane);
    if (ar) A.load(ar, lane);
    const float rstd = rstd_in[r];
    float dot = 0.f;
    OMNI_ROW_LOOP(k, c) {
      float f[8], d[8], g[8];
      unpack8(NPL > 0 ? X.v[NPL > 0 ? k : 0] : __ldg(xr + c), f);
      unpack8(NPL > 0 ? D.v[NPL > 0 ? k : 0] : __ldg(dr + c), d);
      unpack8(__ldg(reinterpret_cast<const uint4*>(w) + c), g);
#pragma unroll
      for (int i = 0; i < 8; ++i) dot += d[i] * g[i] * f[i] * rstd;
    }
    dot = warp_sum(dot) * invH;
    X.pin();
    D.pin();
    uint4* outr = reinterpret_cast<uint4*>(dx) + r * H8;
    OMNI_ROW_LOOP(k, c) {
      float f[8], d[8], g[8], a[8];
      unpack8(NPL > 0 ? X.v[NPL > 0 ? k : 0] : __ldg(xr + c), f);
      unpack8(NPL > 0 ? D.v[NPL > 0 ? k : 0] : __ldg(dr + c), d);
      unpack8(__ldg(reinterpret_cast<const uint4*>(w) + c), g);
      if (ar) unpack8(NPL > 0 ? A.v[NPL > 0 ? k : 0] : __ldg(ar + c), a);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float v = rstd * (d[i] * g[i] - f[i] * rstd * dot);
        if (ar) v += a[i];
        f[i] = v;
      }
      outr[c] = pack8(f);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// LayerNorm (fp32 statistics, eps inside sqrt), affine w/b:  y = bf16( (x-mean)*rstd * w + b )
// ------------------------------------------------------------------------------------------------
template <int NPL>
__global__ void __launch_bounds__(EW_THREADS, 1)
layernorm_fwd_kernel(const bf16* __restrict__ x, const bf16* __restrict__ w, const bf16* __restrict__ b,
                     bf16* __restrict__ y, float* __restrict__ mean_out, float* __restrict__ rstd_out,
                     long long rows, int H8, long long ldx, long long ldy, float eps) {
  const int lane = threadIdx.x & 31;
  const long long wg = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
  const long long nw = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  const float invH = 1.0f / static_cast<float>(H8 * 8);
  for (long long r = wg; r < rows; r += nw) {
    const uint4* xr = reinterpret_cast<const uint4*>(x + r * ldx);
    RowRegs<NPL> X;
    X.load(xr, lane);
    float s = 0.f;
    OMNI_ROW_LOOP(k, c) {
      float f[8];
      unpack8(NPL > 0 ? X.v[NPL > 0 ? k : 0] : __ldg(xr + c), f);
#pragma unroll
      for (int i = 0; i < 8; ++i) s += f[i];
    }
    const float mean = warp_sum(s) * invH;
    X.pin();
    float v = 0.f;
    OMNI_ROW_LOOP(k, c) {
      float f[8];
      unpack8(NPL > 0 ? X.v[NPL > 0 ? k : 0] : __ldg(xr + c), f);
#pragma unroll
      for (int i = 0; i < 8; ++i) { const float d = f[i] - mean; v += d * d; }
    }
    const float rstd = rsqrtf(warp_sum(v) * invH + eps);
    X.pin();
    if (lane == 0) {
      if (mean_out) mean_out[r] = mean;
      if (rstd_out) rstd_out[r] = rstd;
    }
    uint4* yr = reinterpret_cast<uint4*>(y + r * ldy);
    OMNI_ROW_LOOP(k, c) {
      float f[8], g[8], bb[8];
      unpack8(NPL > 0 ? X.v[NPL > 0 ? k : 0] : __ldg(xr + c), f);
      unpack8(__ldg(reinterpret_cast<const uint4*>(w) + c), g);
      unpack8(__ldg(reinterpret_cast<const uint4*>(b) + c), bb);
#pragma unroll
      for (int i = 0; i < 8; ++i) f[i] = (f[i] - mean) * rstd * g[i] + bb[i];
      yr[c] = pack8(f);
    }
  }
}

// dx = rstd * (g - mean(g) - xhat * mean(g*xhat)), g = dy*w   (frozen affine: no dw/db)
template <int NPL>
__global__ void __launch_bounds__(EW_THREADS, 1)
layernorm_bwd_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ x, const bf16* __restrict__ w,
                     const float* __restrict__ mean_in, const float* __restrict__ rstd_in, bf16* __restrict__ dx,
                     const bf16* __restrict__ dx_add, long long rows, int H8) {
  const int lane = threadIdx.x & 31;
  const long long wg = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
  const long long nw = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  const float invH = 1.0f / static_cast<float>(H8 * 8);
  for (long long r = wg; r < rows; r += nw) {
    const uint4* xr = reinterpret_cast<const uint4*>(x) + r * H8;
    const uint4* dr = reinterpret_cast<const uint4*>(dy) + r * H8;
    const uint4* ar = dx_add ? reinterpret_cast<const uint4*>(dx_add) + r * H8 : nullptr;
    RowRegs<NPL> X, D, A;
    X.load(xr, lane);
    D.load(dr, lane);
    if (ar) A.load(ar, lane);
    const float mean = mean_in[r], rstd = rstd_in[r];
    float s1 = 0.f, s2 = 0.f;
    OMNI_ROW_LOOP(k, c) {
      float f[8], d[8], g[8];
      unpack8(NPL > 0 ? X.v[NPL > 0 ? k : 0] : __ldg(xr + c), f);
      unpack8(NPL > 0 ? D.v[NPL > 0 ? k : 0] : __ldg(dr + c), d);
      unpack8(__ldg(reinterpret_cast<const uint4*>(w) + c), g);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float gi = d[i] * g[i];
        s1 += gi;
        s2 += gi * (f[i] - mean) * rstd;
      }
    }
    s1 = warp_sum(s1) * invH;
    s2 = warp_sum(s2) * invH;
    X.pin();
    D.pin();
    uint4* outr = reinterpret_cast<uint4*>(dx) + r * H8;
    OMNI_ROW_LOOP(k, c) {
      float f[8], d[8], g[8], a[8];
      unpack8(NPL > 0 ? X.v[NPL > 0 ? k : 0] : __ldg(xr + c), f);
      unpack8(NPL > 0 ? D.v[NPL > 0 ? k : 0] : __ldg(dr + c), d);
      unpack8(__ldg(reinterpret_cast<const uint4*>(w) + c), g);
      if (ar) unpack8(NPL > 0 ? A.v[NPL > 0 ? k : 0] : __ldg(ar + c), a);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float v = rstd * (d[i] * g[i] - s1 - (f[i] - mean) * rstd * s2);
        if (ar) v += a[i];
        f[i] = v;
      }
      outr[c] = pack8(f);
    }
  }
}
#undef OMNI_ROW_LOOP

// picks the register-resident instantiation for the row widths of the named architectures
// (H = 1024 and 2048; wider rows -- Llama-3.1-8B's 4096 -- would need more than the 128 registers two blocks per SM leave)
#define OMNI_NPL_DISPATCH(KERNEL, H8, ...)                                        \
  do {                                                                            \
    if ((H8) == 128) KERNEL<4><<<__VA_ARGS__;                                     \
    else if ((H8) == 256) KERNEL<8><<<__VA_ARGS__;                                \
    else KERNEL<0><<<__VA_ARGS__;                                                 \
  } while (0)

// ------------------------------------------------------------------------------------------------
// RoPE on the q and k parts of a packed [rows, ld] qkv buffer, in place.
//   out = bf16( bf16(x*cos) + bf16(rotate_half(x)*sin) ),  cos/sin bf16 tables [max_pos, head_dim]
// inverse = 1 applies the transposed rotation (backward).
// one thread = 8 consecutive dims of the first half of one head (and their partners in the second half)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(EW_THREADS)
rope_kernel(bf16* __restrict__ qkv, const bf16* __restrict__ cos_t, const bf16* __restrict__ sin_t,
            const int* __restrict__ pos, long long rows, long long ld, int n_heads_total, int head_dim, int inverse,
            long long total) {
  const int half8 = head_dim / 16;  // 16-byte chunks in half a head
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(idx % half8);
    const long long t = idx / half8;
    const int h = static_cast<int>(t % n_heads_total);
    const long long r = t / n_heads_total;
    const int p = pos[r];
    bf16* base = qkv + r * ld + static_cast<long long>(h) * head_dim;
    uint4* p1 = reinterpret_cast<uint4*>(base) + c;
    uint4* p2 = reinterpret_cast<uint4*>(base + head_dim / 2) + c;
    float x1[8], x2[8], cs[8], sn[8];
    unpack8(*p1, x1);
    unpack8(*p2, x2);
    unpack8(__ldg(reinterpret_cast<const uint4*>(cos_t + static_cast<long long>(p) * head_dim) + c), cs);
    unpack8(__ldg(reinterpret_cast<const uint4*>(sin_t + static_cast<long long>(p) * head_dim) + c), sn);
    float o1[8], o2[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (!inverse) {
        o1[i] = rbf(x1[i] * cs[i]) + rbf(-x2[i] * sn[i]);
        o2[i] = rbf(x2[i] * cs[i]) + rbf(x1[i] * sn[i]);
      } else {
        o1[i] = x1[i] * cs[i] + x2[i] * sn[i];
        o2[i] = x2[i] * cs[i] - x1[i] * sn[i];
      }
    }
    *p1 = pack8(o1);
    *p2 = pack8(o2);
  }
}

// ------------------------------------------------------------------------------------------------
// SwiGLU: gu [rows, 2I] = [gate | up]  ->  act [rows, I] = bf16( bf16(silu(g)) * u )
// ------------------------------------------------------------------------------------------------
// One block walks rows (grid-stride); a thread owns up to four 16-byte column chunks of the row, 256 chunks apart, and has
// all of their loads in flight before the first use (no per-element 64-bit division, 8 / 12 independent loads per thread).
__global__ void __launch_bounds__(EW_THREADS)
swiglu_fwd_kernel(const bf16* __restrict__ gu, bf16* __restrict__ act, long long rows, int I8) {
  for (long long r = blockIdx.x; r < rows; r += gridDim.x) {
    const uint4* g4 = reinterpret_cast<const uint4*>(gu) + r * (2LL * I8);
    uint4* o4 = reinterpret_cast<uint4*>(act) + r * static_cast<long long>(I8);
    for (int c0 = threadIdx.x; c0 < I8; c0 += 4 * EW_THREADS) {
      uint4 gv[4], uv[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int c = c0 + k * EW_THREADS;
        if (c < I8) {
          gv[k] = ld_nc_u4(g4 + c);
          uv[k] = ld_nc_u4(g4 + I8 + c);
        }
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int c = c0 + k * EW_THREADS;
        if (c < I8) {
          float g[8], u[8];
          unpack8(gv[k], g);
          unpack8(uv[k], u);
#pragma unroll
          for (int i = 0; i < 8; ++i) g[i] = rbf(silu(g[i])) * u[i];
          st_na_u4(o4 + c, pack8(g));
        }
      }
    }
  }
}

// d_gu [rows, 2I] from d_act [rows, I] and the saved gu.  gu / d_gu columns are [gate blk | up blk] blocks of blk8 16-byte
// chunks (blk8 = I8: the plain [gate | up] halves; blk8 = 8: the 64-column interleave of OMNI_ACT_SWIGLU64).
__global__ void __launch_bounds__(EW_THREADS)
swiglu_bwd_kernel(const bf16* __restrict__ dact, const bf16* __restrict__ gu, bf16* __restrict__ dgu, long long rows,
                  int I8, int blk8, long long total) {
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(idx % I8);
    const long long r = idx / I8;
    const int cg = (c / blk8) * 2 * blk8 + (c % blk8);
    const uint4* g4 = reinterpret_cast<const uint4*>(gu) + r * (2LL * I8);
    uint4* o4 = reinterpret_cast<uint4*>(dgu) + r * (2LL * I8);
    float g[8], u[8], d[8], dg[8], du[8];
    unpack8(ld_nc_u4(g4 + cg), g);
    unpack8(ld_nc_u4(g4 + cg + blk8), u);
    unpack8(ld_nc_u4(reinterpret_cast<const uint4*>(dact) + idx), d);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float sg = 1.0f / (1.0f + __expf(-g[i]));
      const float s = g[i] * sg;
      du[i] = d[i] * s;
      dg[i] = d[i] * u[i] * (sg * (1.0f + g[i] * (1.0f - sg)));
    }
    st_na_u4(o4 + cg, pack8(dg));
    st_na_u4(o4 + cg + blk8, pack8(du));
  }
}

// ------------------------------------------------------------------------------------------------
// GELU (erf):  y = bf16(gelu(fp32(x)))   and its backward
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(EW_THREADS)
gelu_fwd_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, long long total8) {
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total8;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    float f[8];
    unpack8(ld_nc_u4(reinterpret_cast<const uint4*>(x) + idx), f);
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] = gelu_erf(f[i]);
    st_na_u4(reinterpret_cast<uint4*>(y) + idx, pack8(f));
  }
}
__global__ void __launch_bounds__(EW_THREADS)
gelu_bwd_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ x, bf16* __restrict__ dx, long long total8) {
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total8;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    float f[8], d[8];
    unpack8(ld_nc_u4(reinterpret_cast<const uint4*>(x) + idx), f);
    unpack8(ld_nc_u4(reinterpret_cast<const uint4*>(dy) + idx), d);
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] = d[i] * gelu_grad_fast(f[i]);
    st_na_u4(reinterpret_cast<uint4*>(dx) + idx, pack8(f));
  }
}

// ------------------------------------------------------------------------------------------------
// Row gather: out[i, :] = table[idx[i], :]   (embedding lookup of the decode step / label-row selection)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(EW_THREADS)
gather_rows_kernel(const bf16* __restrict__ table, const int64_t* __restrict__ idx, bf16* __restrict__ out,
                   long long n, int H8, long long ld_table, long long table_rows, int* status) {
  const int lane = threadIdx.x & 31;
  const long long wg = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
  const long long nw = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  for (long long r = wg; r < n; r += nw) {
    const long long src = idx[r];
    uint4* o = reinterpret_cast<uint4*>(out) + r * H8;
    if (src < 0 || src >= table_rows) {
      if (status && lane == 0) *status = 1;
      for (int c = lane; c < H8; c += 32) o[c] = make_uint4(0u, 0u, 0u, 0u);
      continue;
    }
    const uint4* s = reinterpret_cast<const uint4*>(table + src * ld_table);
    for (int c = lane; c < H8; c += 32) st_na_u4(o + c, ld_nc_u4(s + c));
  }
}

// out[idx[i], :] += src[i, :]  with unique idx (scatter-add of row gradients back to a packed buffer); zero elsewhere is
// the caller's job.
__global__ void __launch_bounds__(EW_THREADS)
scatter_rows_kernel(const bf16* __restrict__ src, const int64_t* __restrict__ idx, bf16* __restrict__ out, long long n,
                    int H8, long long ld_out) {
  const int lane = threadIdx.x & 31;
  const long long wg = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
  const long long nw = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  for (long long r = wg; r < n; r += nw) {
    const long long dst = idx[r];
    const uint4* s = reinterpret_cast<const uint4*>(src) + r * H8;
    uint4* o = reinterpret_cast<uint4*>(out + dst * ld_out);
    for (int c = lane; c < H8; c += 32) st_na_u4(o + c, ld_nc_u4(s + c));
  }
}

// ------------------------------------------------------------------------------------------------
// Batched 2-D transpose of bf16 matrices: in [Z, R, C] -> out [Z, C, R], 64 x 64 tiles through padded shared memory with
// 4-byte accesses on both sides.  Used for the transposed copies of the TRAINABLE matrices (LoRA down / up, projector
// weights) that the dgrad GEMMs need every step (the frozen weights keep a cached transposed copy instead).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
transpose_bf16_kernel(const bf16* __restrict__ in, bf16* __restrict__ out, int R, int C) {
  __shared__ uint32_t tile[64][33];               // 64 rows x 64 bf16 (32 words) + 1 word of padding
  const long long z = blockIdx.z;
  const bf16* src = in + z * static_cast<long long>(R) * C;
  bf16* dst = out + z * static_cast<long long>(R) * C;
  const int r0 = blockIdx.y * 64, c0 = blockIdx.x * 64;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;          // 32 x 8 threads
  const bool c_even = (C & 1) == 0, r_even = (R & 1) == 0;
  for (int i = ty; i < 64; i += 8) {
    const int r = r0 + i, c = c0 + 2 * tx;
    uint32_t w = 0;
    if (r < R) {
      const bf16* pp = src + static_cast<long long>(r) * C + c;
      if (c + 1 < C && c_even) {
        w = *reinterpret_cast<const uint32_t*>(pp);
      } else {
        const uint32_t lo = c < C ? *reinterpret_cast<const uint16_t*>(pp) : 0u;
        const uint32_t hi = c + 1 < C ? *reinterpret_cast<const uint16_t*>(pp + 1) : 0u;
        w = lo | (hi << 16);
      }
    }
    tile[i][tx] = w;
  }
  __syncthreads();
  // out row = input column c0 + j, out columns = input rows r0 + 2*tx, r0 + 2*tx + 1
  for (int j = ty; j < 64; j += 8) {
    const int oc = c0 + j, orow = r0 + 2 * tx;
    if (oc >= C) continue;
    const uint32_t a = tile[2 * tx][j >> 1], b = tile[2 * tx + 1][j >> 1];
    const uint32_t lo = (j & 1) ? (a >> 16) : (a & 0xffffu);
    const uint32_t hi = (j & 1) ? (b >> 16) : (b & 0xffffu);
    bf16* pp = dst + static_cast<long long>(oc) * R + orow;
    if (orow + 1 < R && r_even) {
      *reinterpret_cast<uint32_t*>(pp) = lo | (hi << 16);
    } else {
      if (orow < R) *reinterpret_cast<uint16_t*>(pp) = static_cast<uint16_t>(lo);
      if (orow + 1 < R) *reinterpret_cast<uint16_t*>(pp + 1) = static_cast<uint16_t>(hi);
    }
  }
}

}  // namespace omni

using namespace omni;

extern "C" int omni_transpose_bf16(const void* in, void* out, int32_t Z, int32_t R, int32_t C, void* stream) {
  OMNI_CHECK_ARG(in && out && Z > 0 && R > 0 && C > 0 && Z <= 65535);
  OMNI_CHECK_ARG((reinterpret_cast<uintptr_t>(in) & 3) == 0 && (reinterpret_cast<uintptr_t>(out) & 3) == 0);
  dim3 grid((C + 63) / 64, (R + 63) / 64, Z);
  transpose_bf16_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const bf16*)in, (bf16*)out, R, C);
  OMNI_LAUNCH_CHECK();
  return OMNI_OK;
}

extern "C" int omni_rmsnorm_fwd(const void* x, const void* w, void* y, float* rstd, int64_t rows, int32_t H, int64_t ldx,
                                int64_t ldy, float eps, void* stream) {
  OMNI_CHECK_ARG(x && w && y && rows >= 0 && H > 0 && (H % 8) == 0 && (ldx % 8) == 0 && (ldy % 8) == 0);
  if (rows == 0) return OMNI_OK;
  // launched with the programmatic-dependent-launch attribute: inside the decode step its launch overlaps the tail of
  // the GEMM that produced x (the kernel waits for it before its first load)
  const dim3 grid(rows_grid(rows)), block(EW_THREADS);
  cudaStream_t st = (cudaStream_t)stream;
  const int h8 = H / 8;
  const long long rows_ll = rows, ldx_ll = ldx, ldy_ll = ldy;
  cudaError_t e;
  if (h8 == 128)
    e = omni_launch_pdl(rmsnorm_fwd_kernel<4>, grid, block, 0, st, (const bf16*)x, (const bf16*)w, (bf16*)y, rstd, rows_ll, h8,
                        ldx_ll, ldy_ll, eps);
  else if (h8 == 256)
    e = omni_launch_pdl(rmsnorm_fwd_kernel<8>, grid, block, 0, st, (const bf16*)x, (const bf16*)w, (bf16*)y, rstd, rows_ll, h8,
                        ldx_ll, ldy_ll, eps);
  else
    e = omni_launch_pdl(rmsnorm_fwd_kernel<0>, grid, block, 0, st, (const bf16*)x, (const bf16*)w, (bf16*)y, rstd, rows_ll, h8,
                        ldx_ll, ldy_ll, eps);
  if (e != cudaSuccess) return OMNI_ERR_CUDA;
  return OMNI_OK;
}

extern "C" int omni_rmsnorm_bwd(const void* dy, const void* x, const void* w, const float* rstd, void* dx,
                                const void* dx_add, int64_t rows, int32_t H, void* stream) {
  OMNI_CHECK_ARG(dy && x && w && rstd && dx && rows >= 0 && H > 0 && (H % 8) == 0);
  if (rows == 0) return OMNI_OK;
  OMNI_NPL_DISPATCH(rmsnorm_bwd_kernel, H / 8, rows_grid(rows), EW_THREADS, 0, (cudaStream_t)stream>>>(
      (const bf16*)dy, (const bf16*)x, (const bf16*)w, rstd, (bf16*)dx, (const bf16*)dx_add, rows, H / 8));
  OMNI_LAUNCH_CHECK();
  return OMNI_OK;
}

extern "C" int omni_layernorm_fwd(const void* x, const void* w, const void* b, void* y, float* mean, float* rstd,
                                  int64_t rows, int32_t H, int64_t ldx, int64_t ldy, float eps, void* stream) {
  OMNI_CHECK_ARG(x && w && b && y && rows >= 0 && H > 0 && (H % 8) == 0 && (ldx % 8) == 0 && (ldy % 8) == 0);
  if (rows == 0) return OMNI_OK;
  OMNI_NPL_DISPATCH(layernorm_fwd_kernel, H / 8, rows_grid(rows), EW_THREADS, 0, (cudaStream_t)stream>>>(
      (const bf16*)x, (const bf16*)w, (const bf16*)b, (bf16*)y, mean, rstd, rows, H / 8, ldx, ldy, eps));
  OMNI_LAUNCH_CHECK();
  return OMNI_OK;
}

extern "C" int omni_layernorm_bwd(const void* dy, const void* x, const void* w, const float* mean, const float* rstd,
                                  void* dx, const void* dx_add, int64_t rows, int32_t H, void* stream) {
  OMNI_CHECK_ARG(dy && x && w && mean && rstd && dx && rows >= 0 && H > 0 && (H % 8) == 0);
  if (rows == 0) return OMNI_OK;
  OMNI_NPL_DISPATCH(layernorm_bwd_kernel, H / 8, rows_grid(rows), EW_THREADS, 0, (cudaStream_t)stream>>>(
      (const bf16*)dy, (const bf16*)x, (const bf16*)w, mean, rstd, (bf16*)dx, (const bf16*)dx_add, rows, H / 8));
  OMNI_LAUNCH_CHECK();
  return OMNI_OK;
}

extern "C" int omni_rope(void* qkv, const void* cos_t, const void* sin_t, const int32_t* pos, int64_t rows, int64_t ld,
                         int32_t n_heads_total, int32_t head_dim, int32_t inverse, void* stream) {
  OMNI_CHECK_ARG(qkv && cos_t && sin_t && pos && rows >= 0 && n_heads_total > 0);
  OMNI_CHECK_ARG(head_dim > 0 && (head_dim % 16) == 0 && (ld % 8) == 0);
  if (rows == 0) return OMNI_OK;
  const long long total = rows * n_heads_total * (head_dim / 16);
  long long blocks = ceil_div_ll(total, EW_THREADS);
  if (blocks > kNumSMs * 16LL) blocks = kNumSMs * 16LL;
  rope_kernel<<<(int)blocks, EW_THREADS, 0, (cudaStream_t)stream>>>((bf16*)qkv, (const bf16*)cos_t, (const bf16*)sin_t,
                                                                     pos, rows, ld, n_heads_total, head_dim, inverse,
                                                                     total);
  OMNI_LAUNCH_CHECK();
  return OMNI_OK;
}

extern "C" int omni_swiglu_fwd(const void* gu, void* act, int64_t rows, int32_t I, void* stream) {
  OMNI_CHECK_ARG(gu && act && rows >= 0 && I > 0 && (I % 8) == 0);
  if (rows == 0) return OMNI_OK;
  const long long blocks = rows < kNumSMs * 8LL ? rows : kNumSMs * 8LL;
  swiglu_fwd_kernel<<<(int)blocks, EW_THREADS, 0, (cudaStream_t)stream>>>((const bf16*)gu, (bf16*)act, rows, I / 8);
  OMNI_LAUNCH_CHECK();
  return OMNI_OK;
}

extern "C" int omni_swiglu_bwd_blocked(const void* dact, const void* gu, void* dgu, int64_t rows, int32_t I, int32_t blk,
                                       void* stream) {
  OMNI_CHECK_ARG(dact && gu && dgu && rows >= 0 && I > 0 && (I % 8) == 0 && blk > 0 && (blk % 8) == 0 && (I % blk) == 0);
  if (rows == 0) return OMNI_OK;
  const long long total = rows * (I / 8);
  long long blocks = ceil_div_ll(total, EW_THREADS);
  if (blocks > kNumSMs * 16LL) blocks = kNumSMs * 16LL;
  swiglu_bwd_kernel<<<(int)blocks, EW_THREADS, 0, (cudaStream_t)stream>>>((const bf16*)dact, (const bf16*)gu,
                                                                           (bf16*)dgu, rows, I / 8, blk / 8, total);
  OMNI_LAUNCH_CHECK();
  return OMNI_OK;
}

extern "C" int omni_swiglu_bwd(const void* dact, const void* gu, void* dgu, int64_t rows, int32_t I, void* stream) {
  return omni_swiglu_bwd_blocked(dact, gu, dgu, rows, I, I, stream);
}

extern "C" int omni_gelu_fwd(const void* x, void* y, int64_t n, void* stream) {
  OMNI_CHECK_ARG(x && y && n >= 0 && (n % 8) == 0);
  if (n == 0) return OMNI_OK;
  long long blocks = ceil_div_ll(n / 8, EW_THREADS);
  if (blocks > kNumSMs * 16LL) blocks = kNumSMs * 16LL;
  gelu_fwd_kernel<<<(int)blocks, EW_THREADS, 0, (cudaStream_t)stream>>>((const bf16*)x, (bf16*)y, n / 8);
  OMNI_LAUNCH_CHECK();
  return OMNI_OK;
}

extern "C" int omni_gelu_bwd(const void* dy, const void* x, void* dx, int64_t n, void* stream) {
  OMNI_CHECK_ARG(dy && x && dx && n >= 0 && (n % 8) == 0);
  if (n == 0) return OMNI_OK;
  long long blocks = ceil_div_ll(n / 8, EW_THREADS);
  if (blocks > kNumSMs * 16LL) blocks = kNumSMs * 16LL;
  gelu_bwd_kernel<<<(int)blocks, EW_THREADS, 0, (cudaStream_t)stream>>>((const bf16*)dy, (const bf16*)x, (bf16*)dx,
                                                                         n / 8);
  OMNI_LAUNCH_CHECK();
  return OMNI_OK;
}

extern "C" int omni_gather_rows(const void* table, const int64_t* idx, void* out, int64_t n, int32_t H, int64_t ld_table,
                                int64_t table_rows, int32_t* status, void* stream) {
  OMNI_CHECK_ARG(table && idx && out && n >= 0 && H > 0 && (H % 8) == 0 && (ld_table % 8) == 0);
  if (n == 0) return OMNI_OK;
  gather_rows_kernel<<<rows_grid(n), EW_THREADS, 0, (cudaStream_t)stream>>>((const bf16*)table, idx, (bf16*)out, n,
                                                                             H / 8, ld_table, table_rows, status);
  OMNI_LAUNCH_CHECK();
  return OMNI_OK;
}

extern "C" int omni_scatter_rows(const void* src, const int64_t* idx, void* out, int64_t n, int32_t H, int64_t ld_out,
                                 void* stream) {
  OMNI_CHECK_ARG(src && idx && out && n >= 0 && H > 0 && (H % 8) == 0 && (ld_out % 8) == 0);
  if (n == 0) return OMNI_OK;
  scatter_rows_kernel<<<rows_grid(n), EW_THREADS, 0, (cudaStream_t)stream>>>((const bf16*)src, idx, (bf16*)out, n,
                                                                              H / 8, ld_out);
  OMNI_LAUNCH_CHECK();
  return OMNI_OK;
}

// ------------------------------------------------------------------------------------------------
// ResNet front-end glue (av_hubert/avhubert/resnet.py:35-74,131-169), channels-last activations [rows, C]:
//   prelu_res_kernel:      x <- PReLU(x (+ residual)) in place, per-channel slope          (relu1 / `out += residual; relu2`)
//   prelu_maxpool_kernel:  y[n, ho, wo, :] = max_{3x3, stride 2, pad 1} PReLU(x[n, h, w, :])  (frontend3D PReLU + MaxPool3d(1,3,3))
// ------------------------------------------------------------------------------------------------
namespace omni {

__global__ void __launch_bounds__(EW_THREADS)
prelu_res_kernel(bf16* __restrict__ x, const bf16* __restrict__ res, const bf16* __restrict__ slope,
                 const bf16* __restrict__ bias, const bf16* __restrict__ res_bias, int C8, long long total8) {
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total8;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(idx % C8);
    float f[8], s[8];
    unpack8(reinterpret_cast<const uint4*>(x)[idx], f);
    unpack8(__ldg(reinterpret_cast<const uint4*>(slope) + c), s);
    if (bias) {                                                // folded-BatchNorm shift of the convolution that made x
      float b[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(bias) + c), b);
#pragma unroll
      for (int i = 0; i < 8; ++i) f[i] = rbf(f[i] + b[i]);
    }
    if (res) {
      float r[8];
      unpack8(ld_nc_u4(reinterpret_cast<const uint4*>(res) + idx), r);
      if (res_bias) {                                          // ... and of the 1x1 downsample convolution
        float b[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(res_bias) + c), b);
#pragma unroll
        for (int i = 0; i < 8; ++i) r[i] = rbf(r[i] + b[i]);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) f[i] = rbf(f[i] + r[i]);     // `out += residual` rounds to bf16 before the PReLU
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] = f[i] > 0.f ? f[i] : f[i] * s[i];
    reinterpret_cast<uint4*>(x)[idx] = pack8(f);
  }
}

__global__ void __launch_bounds__(EW_THREADS)
prelu_maxpool_kernel(const bf16* __restrict__ x, const bf16* __restrict__ slope, bf16* __restrict__ y, int H, int W,
                     int Ho, int Wo, int C8, long long total8) {
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total8;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(idx % C8);
    long long t = idx / C8;
    const int wo = static_cast<int>(t % Wo); t /= Wo;
    const int ho = static_cast<int>(t % Ho);
    const long long n = t / Ho;
    float s[8], m[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(slope) + c), s);
#pragma unroll
    for (int i = 0; i < 8; ++i) m[i] = -INFINITY;
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
      const int h = 2 * ho - 1 + dy;
      if (h < 0 || h >= H) continue;
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        const int w = 2 * wo - 1 + dx;
        if (w < 0 || w >= W) continue;
        float f[8];
        unpack8(ld_nc_u4(reinterpret_cast<const uint4*>(x) + ((n * H + h) * W + w) * C8 + c), f);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float v = rbf(f[i] > 0.f ? f[i] : f[i] * s[i]);
          m[i] = fmaxf(m[i], v);
        }
      }
    }
    st_na_u4(reinterpret_cast<uint4*>(y) + idx, pack8(m));
  }
}

}  // namespace omni

extern "C" int omni_prelu_res(void* x, const void* residual, const void* slope, const void* bias, const void* res_bias,
                              int64_t rows, int32_t C, void* stream) {
  OMNI_CHECK_ARG(x && slope && rows >= 0 && C > 0 && (C % 8) == 0);
  if (rows == 0) return OMNI_OK;
  const long long total8 = rows * (C / 8);
  long long blocks = ceil_div_ll(total8, EW_THREADS);
  if (blocks > kNumSMs * 16LL) blocks = kNumSMs * 16LL;
  prelu_res_kernel<<<(int)blocks, EW_THREADS, 0, (cudaStream_t)stream>>>((bf16*)x, (const bf16*)residual,
                                                                          (const bf16*)slope, (const bf16*)bias,
                                                                          (const bf16*)res_bias, C / 8, total8);
  OMNI_LAUNCH_CHECK();
  return OMNI_OK;
}

extern "C" int omni_prelu_maxpool3x3s2(const void* x, const void* slope, void* y, int64_t N, int32_t H, int32_t W,
                                       int32_t C, void* stream) {
  OMNI_CHECK_ARG(x && slope && y && N >= 0 && H > 0 && W > 0 && C > 0 && (C % 8) == 0);
  if (N == 0) return OMNI_OK;
  const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  const long long total8 = N * Ho * Wo * (C / 8);
  long long blocks = ceil_div_ll(total8, EW_THREADS);
  if (blocks > kNumSMs * 16LL) blocks = kNumSMs * 16LL;
  prelu_maxpool_kernel<<<(int)blocks, EW_THREADS, 0, (cudaStream_t)stream>>>((const bf16*)x, (const bf16*)slope, (bf16*)y,
                                                                              H, W, Ho, Wo, C / 8, total8);
  OMNI_LAUNCH_CHECK();
  return OMNI_OK;
}

// ------------------------------------------------------------------------------------------------
// im2col of AV-HuBERT's video front-end convolution (resnet.py:137: Conv3d(1, 64, (5,7,7), stride (1,2,2), pad (2,3,3))).
// video [B, T, H, W] bf16 (C_in = 1)  ->  A [B*T*Ho*Wo, 256] bf16, column k = (kt*7 + ky)*7 + kx (245 taps, 11 zero pads),
// so that the convolution becomes one tcgen05 GEMM against the [64, 256] (BatchNorm-folded) filter matrix.
// One thread = 8 consecutive taps of one output position (one 16-byte store).
// ------------------------------------------------------------------------------------------------
namespace omni {

__global__ void __launch_bounds__(EW_THREADS)
im2col_front3d_kernel(const bf16* __restrict__ video, bf16* __restrict__ out, int T, int H, int W, int Ho, int Wo,
                      long long total_chunks) {
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total_chunks;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int kc = static_cast<int>(idx & 31);          // 32 chunks of 8 taps per output position
    long long m = idx >> 5;
    const int xo = static_cast<int>(m % Wo); m /= Wo;
    const int yo = static_cast<int>(m % Ho); m /= Ho;
    const int t = static_cast<int>(m % T);
    const long long b = m / T;
    const bf16* clip = video + b * static_cast<long long>(T) * H * W;
    __align__(16) bf16 v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int k = kc * 8 + i;
      bf16 val = __float2bfloat16_rn(0.f);
      if (k < 245) {
        const int kt = k / 49;
        const int r = k - kt * 49;
        const int ky = r / 7;
        const int kx = r - ky * 7;
        const int tt = t + kt - 2, yy = 2 * yo + ky - 3, xx = 2 * xo + kx - 3;
        if (tt >= 0 && tt < T && yy >= 0 && yy < H && xx >= 0 && xx < W)
          val = clip[(static_cast<long long>(tt) * H + yy) * W + xx];
      }
      v[i] = val;
    }
    st_na_u4(reinterpret_cast<uint4*>(out) + idx, *reinterpret_cast<const uint4*>(v));
  }
}

}  // namespace omni

extern "C" int omni_im2col_front3d(const void* video, void* out, int32_t B, int32_t T, int32_t H, int32_t W,
                                   void* stream) {
  OMNI_CHECK_ARG(video && out && B > 0 && T > 0 && H > 0 && W > 0);
  const int Ho = (H + 6 - 7) / 2 + 1, Wo = (W + 6 - 7) / 2 + 1;
  const long long total = static_cast<long long>(B) * T * Ho * Wo * 32;
  long long blocks = ceil_div_ll(total, EW_THREADS);
  if (blocks > kNumSMs * 32LL) blocks = kNumSMs * 32LL;
  im2col_front3d_kernel<<<(int)blocks, EW_THREADS, 0, (cudaStream_t)stream>>>((const bf16*)video, (bf16*)out, T, H, W, Ho,
                                                                               Wo, total);
  OMNI_LAUNCH_CHECK();
  return OMNI_OK;
}

// ------------------------------------------------------------------------------------------------
// Time-major variant of the front-end convolution (resnet.py:137), 4x less HBM traffic than the 245-tap im2col above:
// only the 7x7 spatial taps are materialised, A2[b][yo][xo][tt][64] (49 taps + 15 zero columns, tt = t + 2 in a
// zero-padded time axis of T + 4), so that the five temporal taps of an output position are five CONSECUTIVE rows:
// the GEMM reads A2 through an overlapping-row view [rows, 320] with row stride 64 (TMA global stride 128 B) against the
// [C, 5*64] filter matrix.  Output row (b, yo, xo, t) for t < T is the convolution at time t; the 4 trailing rows of
// every (b, yo, xo) line are scratch.  omni_prelu_maxpool_front then applies PReLU + MaxPool(1,3,3)/(1,2,2) reading that
// layout and writes the channels-last [B*T, Hp, Wp, C] activation of the ResNet trunk.
// ------------------------------------------------------------------------------------------------
namespace omni {

__global__ void __launch_bounds__(EW_THREADS)
im2col_front2d_kernel(const bf16* __restrict__ video, bf16* __restrict__ out, int T, int H, int W, int Ho, int Wo,
                      long long total) {
  const int Tp = T + 4;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    long long m = idx;
    const int xo = static_cast<int>(m % Wo); m /= Wo;       // lanes run along x: neighbouring lanes read neighbouring pixels
    const int yo = static_cast<int>(m % Ho); m /= Ho;
    const int tt = static_cast<int>(m % Tp);
    const long long b = m / Tp;
    const int t = tt - 2;
    const bool t_ok = t >= 0 && t < T;
    const bf16* frame = video + (b * T + (t_ok ? t : 0)) * static_cast<long long>(H) * W;
    uint32_t w[32];
#pragma unroll
    for (int k2 = 0; k2 < 32; ++k2) {
      unsigned short lo = 0, hi = 0;
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int k = 2 * k2 + e;
        if (k < 49) {
          const int ky = k / 7, kx = k % 7;
          const int yy = 2 * yo + ky - 3, xx = 2 * xo + kx - 3;
          unsigned short v = 0;
          if (t_ok && yy >= 0 && yy < H && xx >= 0 && xx < W)
            v = __ldg(reinterpret_cast<const unsigned short*>(frame) + yy * W + xx);
          if (e == 0) lo = v; else hi = v;
        }
      }
      w[k2] = static_cast<uint32_t>(lo) | (static_cast<uint32_t>(hi) << 16);
    }
    const long long row = ((b * Ho + yo) * Wo + xo) * Tp + tt;
    uint4* dst = reinterpret_cast<uint4*>(out) + row * 8;
#pragma unroll
    for (int i = 0; i < 8; ++i) dst[i] = make_uint4(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
  }
}

// x: conv output [B][H][W][Tp][C] (rows of the time-major GEMM), y: [B*T][Ho][Wo][C] channels-last
__global__ void __launch_bounds__(EW_THREADS)
prelu_maxpool_front_kernel(const bf16* __restrict__ x, const bf16* __restrict__ slope, bf16* __restrict__ y, int T, int Tp,
                           int H, int W, int Ho, int Wo, int C8, long long total8, int ring) {
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total8;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(idx % C8);
    long long r = idx / C8;
    const int t = static_cast<int>(r % T); r /= T;         // consecutive threads: channels, then time (contiguous reads)
    const int wo = static_cast<int>(r % Wo); r /= Wo;
    const int ho = static_cast<int>(r % Ho);
    const long long b = r / Ho;
    // max-pool(PReLU(x)) = max(PReLU(max x), PReLU(min x)) for ANY slope (PReLU is monotone for slope >= 0 and V-shaped for
    // slope < 0: its maximum over a set sits at one of the two extremes; the bf16 rounding is monotone, so it commutes with
    // the maximum): the nine taps only cost packed bf16 max / min instructions (2 per 32-bit word), PReLU runs twice per
    // output instead of nine times -- the kernel was issue-bound at ~36 fp32 instructions per output element.
    uint32_t mx[4] = {0xff80ff80u, 0xff80ff80u, 0xff80ff80u, 0xff80ff80u};      // -inf pairs
    uint32_t mn[4] = {0x7f807f80u, 0x7f807f80u, 0x7f807f80u, 0x7f807f80u};      // +inf pairs
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
      const int h = 2 * ho - 1 + dy;
      if (h < 0 || h >= H) continue;
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        const int w = 2 * wo - 1 + dx;
        if (w < 0 || w >= W) continue;
        const uint4 u = ld_nc_u4(reinterpret_cast<const uint4*>(x) + (((b * H + h) * W + w) * Tp + t) * C8 + c);
        const uint32_t uw[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          asm("max.bf16x2 %0, %0, %1;" : "+r"(mx[i]) : "r"(uw[i]));
          asm("min.bf16x2 %0, %0, %1;" : "+r"(mn[i]) : "r"(uw[i]));
        }
      }
    }
    float s[8], hi[8], lo[8], m[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(slope) + c), s);
    unpack8(make_uint4(mx[0], mx[1], mx[2], mx[3]), hi);
    unpack8(make_uint4(mn[0], mn[1], mn[2], mn[3]), lo);
#pragma unroll
    for (int i = 0; i < 8; ++i)
      m[i] = fmaxf(rbf(hi[i] > 0.f ? hi[i] : hi[i] * s[i]), rbf(lo[i] > 0.f ? lo[i] : lo[i] * s[i]));
    // ring = 1: the frame is stored with a one-pixel border [Ho + 2, Wo + 2] that the caller keeps at zero (the layout
    // the 3x3 convolutions of the trunk read as an overlapping-row GEMM operand, csrc/resnet_trunk.cu)
    st_na_u4(reinterpret_cast<uint4*>(y) +
                 ((((b * T + t) * (Ho + 2 * ring) + ho + ring) * (Wo + 2 * ring) + wo + ring) * C8 + c),
             pack8(m));
  }
}

}  // namespace omni

extern "C" int omni_im2col_front2d(const void* video, void* out, int32_t B, int32_t T, int32_t H, int32_t W,
                                   void* stream) {
  OMNI_CHECK_ARG(video && out && B > 0 && T > 0 && H > 0 && W > 0);
  const int Ho = (H + 6 - 7) / 2 + 1, Wo = (W + 6 - 7) / 2 + 1;
  const long long total = static_cast<long long>(B) * (T + 4) * Ho * Wo;
  long long blocks = ceil_div_ll(total, EW_THREADS);
  if (blocks > kNumSMs * 32LL) blocks = kNumSMs * 32LL;
  im2col_front2d_kernel<<<(int)blocks, EW_THREADS, 0, (cudaStream_t)stream>>>((const bf16*)video, (bf16*)out, T, H, W, Ho,
                                                                               Wo, total);
  OMNI_LAUNCH_CHECK();
  return OMNI_OK;
}

extern "C" int omni_prelu_maxpool_front(const void* x, const void* slope, void* y, int32_t B, int32_t T, int32_t H,
                                        int32_t W, int32_t C, void* stream) {
  OMNI_CHECK_ARG(x && slope && y && B > 0 && T > 0 && H > 0 && W > 0 && C > 0 && (C % 8) == 0);
  const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  const long long total8 = static_cast<long long>(B) * T * Ho * Wo * (C / 8);
  long long blocks = ceil_div_ll(total8, EW_THREADS);
  if (blocks > kNumSMs * 16LL) blocks = kNumSMs * 16LL;
  prelu_maxpool_front_kernel<<<(int)blocks, EW_THREADS, 0, (cudaStream_t)stream>>>(
      (const bf16*)x, (const bf16*)slope, (bf16*)y, T, T + 4, H, W, Ho, Wo, C / 8, total8, 0);
  OMNI_LAUNCH_CHECK();
  return OMNI_OK;
}

extern "C" int omni_prelu_maxpool_front_ring(const void* x, const void* slope, void* y, int32_t B, int32_t T, int32_t H,
                                             int32_t W, int32_t C, void* stream) {
  OMNI_CHECK_ARG(x && slope && y && B > 0 && T > 0 && H > 0 && W > 0 && C > 0 && (C % 8) == 0);
  const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  const long long total8 = static_cast<long long>(B) * T * Ho * Wo * (C / 8);
  long long blocks = ceil_div_ll(total8, EW_THREADS);
  if (blocks > kNumSMs * 16LL) blocks = kNumSMs * 16LL;
  prelu_maxpool_front_kernel<<<(int)blocks, EW_THREADS, 0, (cudaStream_t)stream>>>(
      (const bf16*)x, (const bf16*)slope, (bf16*)y, T, T + 4, H, W, Ho, Wo, C / 8, total8, 1);
  OMNI_LAUNCH_CHECK();
  return OMNI_OK;
}
