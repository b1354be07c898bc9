// Beam search on the device (the reference's evaluation default: eval_OmniAVSR.py:216-226 -> num_beams = 15, driven through
// modeling_OmniAVSR.py:313-322 -> HF transformers 4.43.1 `GenerationMixin._beam_search` + `BeamSearchScorer.process`).
//
// Two launches per step, both inside the CUDA graph of the step (no host synchronisation, no library kernel):
//   omni_beam_topk_rows   one CTA per beam row: fp32 log-softmax statistics of the bf16 logits row (online max / sum) and the
//                         row's 2K best tokens.  The global top-2K over the K*V candidates of an utterance is contained in the
//                         union of the per-row top-2K lists, so the K*V matrix is never ranked as a whole.
//   omni_beam_select      one CTA per utterance: merges the K lists (score = log-prob + running beam score), walks the 2K best
//                         candidates exactly like BeamSearchScorer.process (EOS candidates among the first K become finished
//                         hypotheses, kept as the K best by score / length; the first K non-EOS candidates continue), decides
//                         `done`, and writes everything the next forward needs: beam scores, token history, the embedding
//                         rows of the chosen tokens and the KV-cache INDIRECTION table.
// KV cache: instead of gathering the whole cache by beam index after every step (HF `_reorder_cache`: ~245 MB of copies per
// step at K = 15, 500 cached positions), every physical cache row keeps what its forward passes wrote; ind[row][t] names the
// physical row that holds generated position t of the hypothesis now living in `row`, and the prompt is stored once per
// utterance (row b*K).  The single-token attention kernel resolves the row per key position (decode_attention.cu, BEAM).
#include "common.cuh"
#include "../../include/omni_avsr.h"

namespace omni {

constexpr int BT_THREADS = 1024;
constexpr int BT_CAP = 2048;          // candidate buffer of the fast path

__device__ __forceinline__ uint32_t bf16_key(uint32_t bits) {          // monotone: larger bf16 value -> larger 16-bit key
  return (bits & 0x8000u) ? (~bits & 0xffffu) : (bits | 0x8000u);
}
__device__ __forceinline__ float key_value(uint32_t key) {
  const uint32_t bits = (key & 0x8000u) ? (key & 0x7fffu) : (~key & 0xffffu);
  return __uint_as_float(bits << 16);
}
__device__ __forceinline__ uint32_t f32_key(float f) {                  // monotone 32-bit key of an fp32 value
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
// (key, index) -> one 64-bit word ordered like "key descending, then index ascending" under plain > comparison
__device__ __forceinline__ unsigned long long comp_of(uint32_t key, uint32_t idx) {
  return (static_cast<unsigned long long>(key) << 32) | (0xffffffffu - idx);
}

__device__ __forceinline__ unsigned long long block_max_u64(unsigned long long v, unsigned long long* s_red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long w = __shfl_xor_sync(0xffffffffu, v, o);
    v = w > v ? w : v;
  }
  __syncthreads();                                  // s_red may still be read from the previous round
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
  __syncthreads();
  v = s_red[threadIdx.x & 31];                       // BT_THREADS / 32 == 32 partial maxima
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long w = __shfl_xor_sync(0xffffffffu, v, o);
    v = w > v ? w : v;
  }
  return v;
}

__global__ void __launch_bounds__(BT_THREADS)
beam_topk_rows_kernel(const bf16* __restrict__ logits, long long ld, int V, const float* __restrict__ beam_scores, int n_cand,
                      float* __restrict__ cand_score, int* __restrict__ cand_tok) {
  __shared__ float s_m[32], s_s[32];
  __shared__ unsigned long long s_cand[BT_CAP];
  __shared__ unsigned long long s_red[32];
  __shared__ int s_count;
  __shared__ uint32_t s_L;
  __shared__ float s_stats[2];

  const int row = blockIdx.x, tid = threadIdx.x;
  const bf16* x = logits + static_cast<long long>(row) * ld;
  const uint16_t* xb = reinterpret_cast<const uint16_t*>(x);
  const int nvec = V >> 3;

  // ---- pass 1: online max / sum-exp per thread, per-thread maximal key ----
  float m = -INFINITY, s = 0.f;
  uint32_t tk = 0;
  constexpr int UN = 4;                 // 16-byte loads in flight per thread (the row passes are latency-bound otherwise)
  for (int i0 = tid; i0 < nvec; i0 += UN * BT_THREADS) {
    uint4 uu[UN];
#pragma unroll
    for (int k = 0; k < UN; ++k)
      if (i0 + k * BT_THREADS < nvec) uu[k] = ld_nc_u4(reinterpret_cast<const uint4*>(x) + i0 + k * BT_THREADS);
#pragma unroll
    for (int k = 0; k < UN; ++k) {
      if (i0 + k * BT_THREADS >= nvec) break;
      const uint32_t w[4] = {uu[k].x, uu[k].y, uu[k].z, uu[k].w};
      float f[8];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 t = bf2_to_f2(w[j]);
        f[2 * j] = t.x; f[2 * j + 1] = t.y;
        const uint32_t k0 = bf16_key(w[j] & 0xffffu), k1 = bf16_key(w[j] >> 16);
        tk = max(tk, max(k0, k1));
      }
      float cm = f[0];
#pragma unroll
      for (int j = 1; j < 8; ++j) cm = fmaxf(cm, f[j]);
      if (cm > m) { s *= __expf(m - cm); m = cm; }
#pragma unroll
      for (int j = 0; j < 8; ++j) s += __expf(f[j] - m);
    }
  }
  for (int i = (nvec << 3) + tid; i < V; i += BT_THREADS) {            // tail (V not a multiple of 8)
    const uint32_t bits = xb[i];
    const float f = __uint_as_float(bits << 16);
    tk = max(tk, bf16_key(bits));
    if (f > m) { s *= __expf(m - f); m = f; }
    s += __expf(f - m);
  }
  // block statistics
  {
    float wm = warp_max(m);
    float ws = warp_sum(m == -INFINITY ? 0.f : s * __expf(m - wm));
    if ((tid & 31) == 0) { s_m[tid >> 5] = wm; s_s[tid >> 5] = ws; }
    if (tid == 0) s_count = 0;
    __syncthreads();
    if (tid < 32) {
      const float pm = s_m[tid], ps = s_s[tid];
      const float bm = warp_max(pm);
      const float bs = warp_sum(pm == -INFINITY ? 0.f : ps * __expf(pm - bm));
      if (tid == 0) { s_stats[0] = bm; s_stats[1] = logf(bs); }
    }
  }
  // the n_cand-th largest per-thread maximum is a lower bound of the row's n_cand-th largest element: bitwise radix select
  // over the 1024 16-bit keys (one __syncthreads_count per bit; a rank-by-counting loop over all pairs cost 26 us per row)
  {
    const uint32_t mine = tk;
    uint32_t prefix = 0;
    int remaining = min(n_cand, BT_THREADS);
#pragma unroll 1
    for (int bit = 15; bit >= 0; --bit) {
      const uint32_t cand = prefix | (1u << bit);
      const int c = __syncthreads_count((mine >> bit) == (cand >> bit));      // keys with this prefix and the bit set
      if (c >= remaining) prefix = cand;
      else remaining -= c;
    }
    if (tid == 0) s_L = prefix;
  }
  __syncthreads();
  const uint32_t L = s_L;
  const float mx = s_stats[0], logsum = s_stats[1];
  const float bscore = beam_scores[row];

  // ---- pass 2: collect every element >= L ----
  for (int i0 = tid; i0 < nvec; i0 += UN * BT_THREADS) {
    uint4 uu[UN];
#pragma unroll
    for (int k = 0; k < UN; ++k)
      if (i0 + k * BT_THREADS < nvec) uu[k] = ld_nc_u4(reinterpret_cast<const uint4*>(x) + i0 + k * BT_THREADS);
#pragma unroll
    for (int k = 0; k < UN; ++k) {
      const int i = i0 + k * BT_THREADS;
      if (i >= nvec) break;
      const uint32_t w[4] = {uu[k].x, uu[k].y, uu[k].z, uu[k].w};
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const uint32_t kk = bf16_key((w[j >> 1] >> ((j & 1) * 16)) & 0xffffu);
        if (kk >= L) {
          const int p = atomicAdd(&s_count, 1);
          if (p < BT_CAP) s_cand[p] = comp_of(kk, static_cast<uint32_t>(i * 8 + j));
        }
      }
    }
  }
  for (int i = (nvec << 3) + tid; i < V; i += BT_THREADS) {
    const uint32_t k = bf16_key(xb[i]);
    if (k >= L) {
      const int p = atomicAdd(&s_count, 1);
      if (p < BT_CAP) s_cand[p] = comp_of(k, static_cast<uint32_t>(i));
    }
  }
  __syncthreads();
  const int C = s_count;
  float* os = cand_score + static_cast<long long>(row) * n_cand;
  int* ot = cand_tok + static_cast<long long>(row) * n_cand;
  if (C <= BT_CAP) {
    // exact ranking of the collected candidates (key descending, index ascending): rank = number of better candidates
    for (int i = tid; i < C; i += BT_THREADS) {
      const unsigned long long mine = s_cand[i];
      int r = 0;
      for (int j = 0; j < C; ++j) r += s_cand[j] > mine;
      if (r < n_cand) {
        os[r] = ((key_value(static_cast<uint32_t>(mine >> 32)) - mx) - logsum) + bscore;
        ot[r] = static_cast<int>(0xffffffffu - static_cast<uint32_t>(mine));
      }
    }
    for (int r = C + tid; r < n_cand; r += BT_THREADS) { os[r] = -INFINITY; ot[r] = -1; }     // V < n_cand
  } else {
    // degenerate rows (massive ties at the threshold): n_cand rounds of "best element after the previous pick"
    unsigned long long prev = ~0ull;
    for (int r = 0; r < n_cand; ++r) {
      unsigned long long best = 0ull;
      for (int i = tid; i < V; i += BT_THREADS) {
        const unsigned long long c = comp_of(bf16_key(xb[i]), static_cast<uint32_t>(i));
        if (c < prev && c > best) best = c;
      }
      best = block_max_u64(best, s_red);
      if (tid == 0) {
        if (best == 0ull) { os[r] = -INFINITY; ot[r] = -1; }
        else {
          os[r] = ((key_value(static_cast<uint32_t>(best >> 32)) - mx) - logsum) + bscore;
          ot[r] = static_cast<int>(0xffffffffu - static_cast<uint32_t>(best));
        }
      }
      prev = best == 0ull ? 0ull : best;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
constexpr int BS_THREADS = 256;
constexpr int BS_MAX_CAND = 2048;     // K * 2K candidates per utterance (K <= 32)
constexpr int BS_MAX_K = 32;

struct BeamSelectK {
  const float* cand_score; const int* cand_tok;
  float* beam_scores;
  const long long* step_idx; const long long* eos; const long long* pad;
  int* seqs; int* ind;
  int* hyp_seq; int* hyp_len; double* hyp_score; int* hyp_order; int* hyp_count; double* hyp_worst;
  int* done; int* n_done; int* status;
  const bf16* embed; bf16* x_next;
  long long ld_embed, ld_x;
  int B, K, n_cand, V, max_new, ind_ld, H8;
};

__device__ void hyp_add(const BeamSelectK& p, int b, const int* seq, int len, float sum_logprobs, int cur_len) {
  // _Hyps.add of the host implementation == BeamHypotheses.add (length_penalty 1.0): score = sum_logprobs / cur_len
  const int K = p.K;
  const double score = static_cast<double>(sum_logprobs) / static_cast<double>(cur_len);
  int cnt = p.hyp_count[b];
  double worst = p.hyp_worst[b];
  int* order = p.hyp_order + b * (K + 1);
  double* hs = p.hyp_score + b * (K + 1);
  if (cnt < K || score > worst) {
    const int slot = order[cnt];
    int* dst = p.hyp_seq + (static_cast<long long>(b) * (K + 1) + slot) * p.max_new;
    for (int t = 0; t < len; ++t) dst[t] = seq[t];
    p.hyp_len[b * (K + 1) + slot] = len;
    hs[slot] = score;
    ++cnt;
    if (cnt > K) {
      int imin = 0;
      for (int i = 1; i < cnt; ++i)
        if (hs[order[i]] < hs[order[imin]]) imin = i;                   // ties: the earliest insertion goes first
      const int freed = order[imin];
      for (int i = imin; i < cnt - 1; ++i) order[i] = order[i + 1];
      order[cnt - 1] = freed;
      --cnt;
      worst = hs[order[0]];
      for (int i = 1; i < cnt; ++i) worst = fmin(worst, hs[order[i]]);
    } else {
      worst = fmin(score, worst);
    }
    p.hyp_count[b] = cnt;
    p.hyp_worst[b] = worst;
  }
}

__global__ void __launch_bounds__(BS_THREADS)
beam_select_kernel(const BeamSelectK p) {
  __shared__ unsigned long long s_comp[BS_MAX_CAND];
  __shared__ float s_top_score[2 * BS_MAX_K];
  __shared__ int s_top_flat[2 * BS_MAX_K];
  __shared__ float s_new_score[BS_MAX_K];
  __shared__ int s_new_tok[BS_MAX_K], s_new_src[BS_MAX_K];

  const int b = blockIdx.x, tid = threadIdx.x, K = p.K, nc = p.n_cand;
  const int step = static_cast<int>(*p.step_idx);
  if (step >= p.max_new) return;
  const int cur_len = step + 1;
  const int par = step & 1;
  const long long plane_seq = static_cast<long long>(p.B) * K * p.max_new;
  const long long plane_ind = static_cast<long long>(p.B) * K * p.ind_ld;
  const int* seqs_in = p.seqs + par * plane_seq;
  int* seqs_out = p.seqs + (par ^ 1) * plane_seq;
  const int* ind_in = p.ind + par * plane_ind;
  int* ind_out = p.ind + (par ^ 1) * plane_ind;
  const int pad = static_cast<int>(*p.pad), eos = static_cast<int>(*p.eos);
  const bool was_done = p.done[b] != 0;

  if (!was_done) {
    const int n = K * nc;
    for (int i = tid; i < n; i += BS_THREADS) {
      const int k = i / nc;
      const long long src = static_cast<long long>(b * K + k) * nc + (i - k * nc);
      const int tok = p.cand_tok[src];
      const float sc = p.cand_score[src];
      // flat index k * V + tok orders ties like a row-major top-k over the [K * V] candidates; absent candidates last
      s_comp[i] = tok < 0 ? 0ull : comp_of(f32_key(sc), static_cast<uint32_t>(k * p.V + tok));
    }
    __syncthreads();
    for (int i = tid; i < n; i += BS_THREADS) {
      const unsigned long long mine = s_comp[i];
      int r = 0;
      for (int j = 0; j < n; ++j) r += (s_comp[j] > mine) || (s_comp[j] == mine && j < i);
      if (r < nc) {
        const int k = i / nc;
        const long long src = static_cast<long long>(b * K + k) * nc + (i - k * nc);
        s_top_score[r] = p.cand_score[src];
        s_top_flat[r] = mine == 0ull ? -1 : static_cast<int>(0xffffffffu - static_cast<uint32_t>(mine));
      }
    }
    __syncthreads();
    if (tid == 0) {
      int taken = 0;
      for (int rank = 0; rank < nc && taken < K; ++rank) {
        const int flat = s_top_flat[rank];
        if (flat < 0) break;
        const float sc = s_top_score[rank];
        const int k = flat / p.V, tok = flat - k * p.V, row = b * K + k;
        if (tok == eos) {
          if (rank < K) hyp_add(p, b, seqs_in + static_cast<long long>(row) * p.max_new, step, sc, cur_len);
          continue;
        }
        s_new_score[taken] = sc; s_new_tok[taken] = tok; s_new_src[taken] = row;
        ++taken;
      }
      if (taken < K) {                 // HF raises "At most K tokens can be equal to eos_token_id"; reported through status
        atomicExch(p.status, 1);
        for (; taken < K; ++taken) { s_new_score[taken] = -1e9f; s_new_tok[taken] = pad; s_new_src[taken] = b * K + taken; }
      }
      const bool fin = p.hyp_count[b] >= K &&
                       p.hyp_worst[b] >= static_cast<double>(s_top_score[0]) / static_cast<double>(cur_len);
      if (fin) { p.done[b] = 1; atomicAdd(p.n_done, 1); }
    }
  } else if (tid < K) {                // finished utterance: padded, ignored by every later step
    s_new_score[tid] = 0.f; s_new_tok[tid] = pad; s_new_src[tid] = b * K + tid;
  }
  __syncthreads();

  // all copies as flat loops over (beam, element): independent loads in flight instead of one round trip per beam
  const int n_hist = K * (step + 1);
  for (int idx = tid; idx < n_hist; idx += BS_THREADS) {
    const int j = idx / (step + 1), t = idx - j * (step + 1);
    const int src = s_new_src[j], dst = b * K + j;
    seqs_out[static_cast<long long>(dst) * p.max_new + t] =
        t < step ? seqs_in[static_cast<long long>(src) * p.max_new + t] : s_new_tok[j];
    if (t < p.ind_ld)
      ind_out[static_cast<long long>(dst) * p.ind_ld + t] = t < step ? ind_in[static_cast<long long>(src) * p.ind_ld + t] : dst;
  }
  if (tid < K) p.beam_scores[b * K + tid] = s_new_score[tid];
  const int n_emb = K * p.H8;
#pragma unroll 4
  for (int idx = tid; idx < n_emb; idx += BS_THREADS) {
    const int j = idx / p.H8, h = idx - j * p.H8;
    const uint4* e = reinterpret_cast<const uint4*>(p.embed + static_cast<long long>(s_new_tok[j]) * p.ld_embed);
    reinterpret_cast<uint4*>(p.x_next + static_cast<long long>(b * K + j) * p.ld_x)[h] = __ldg(e + h);
  }
}

}  // namespace omni

extern "C" int omni_beam_topk_rows(const void* logits, int64_t ld, int32_t rows, int32_t V, const float* beam_scores,
                                   int32_t n_cand, float* cand_score, int32_t* cand_tok, void* stream) {
  using namespace omni;
  OMNI_CHECK_ARG(logits && beam_scores && cand_score && cand_tok && rows > 0 && V > 0 && n_cand > 0 && n_cand <= 1024);
  OMNI_CHECK_ARG((ld % 8) == 0 && ld >= V && (reinterpret_cast<uintptr_t>(logits) & 15) == 0);
  beam_topk_rows_kernel<<<rows, BT_THREADS, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const bf16*>(logits), ld, V, beam_scores, n_cand, cand_score, cand_tok);
  OMNI_LAUNCH_CHECK();
  return OMNI_OK;
}

extern "C" int omni_beam_select(const omni_beam_select_args* a, void* stream) {
  using namespace omni;
  OMNI_CHECK_ARG(a && a->cand_score && a->cand_tok && a->beam_scores && a->step_idx && a->eos && a->pad && a->seqs && a->ind &&
                 a->hyp_seq && a->hyp_len && a->hyp_score && a->hyp_order && a->hyp_count && a->hyp_worst && a->done &&
                 a->n_done && a->status && a->embed && a->x_next);
  OMNI_CHECK_ARG(a->B > 0 && a->K > 0 && a->K <= BS_MAX_K && a->n_cand == 2 * a->K && a->V > 0 && a->max_new > 0 &&
                 a->ind_ld > 0 && a->H > 0 && (a->H % 8) == 0 && (a->ld_embed % 8) == 0 && (a->ld_x % 8) == 0);
  OMNI_CHECK_ARG(static_cast<long long>(a->K) * a->V < 0x7fffffffLL);
  BeamSelectK p;
  p.cand_score = a->cand_score; p.cand_tok = a->cand_tok; p.beam_scores = a->beam_scores;
  p.step_idx = reinterpret_cast<const long long*>(a->step_idx);
  p.eos = reinterpret_cast<const long long*>(a->eos); p.pad = reinterpret_cast<const long long*>(a->pad);
  p.seqs = a->seqs; p.ind = a->ind;
  p.hyp_seq = a->hyp_seq; p.hyp_len = a->hyp_len; p.hyp_score = a->hyp_score; p.hyp_order = a->hyp_order;
  p.hyp_count = a->hyp_count; p.hyp_worst = a->hyp_worst; p.done = a->done; p.n_done = a->n_done; p.status = a->status;
  p.embed = reinterpret_cast<const bf16*>(a->embed); p.x_next = reinterpret_cast<bf16*>(a->x_next);
  p.ld_embed = a->ld_embed; p.ld_x = a->ld_x;
  p.B = a->B; p.K = a->K; p.n_cand = a->n_cand; p.V = a->V; p.max_new = a->max_new; p.ind_ld = a->ind_ld; p.H8 = a->H / 8;
  beam_select_kernel<<<a->B, BS_THREADS, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p);
  OMNI_LAUNCH_CHECK();
  return OMNI_OK;
}
