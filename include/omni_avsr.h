/*
 * omni_avsr.h -- C ABI of the B200 (sm_100a) kernels behind the Omni-AVSR hot path.
 *
 * The reference (umbertocappellazzo/Omni-AVSR) has no FFI layer: its hot path is a chain of
 * PyTorch library calls inside nn.Module.forward. Each entry point below replaces the op sequence
 * at the cited reference lines (paths relative to the reference root). Host code (Python) keeps the
 * reference's module/class API and calls these through ctypes.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless named h_*; tensors are row-major, bf16 unless noted;
 *   - `stream` is a cudaStream_t passed as void*; all work is stream-ordered, nothing synchronises;
 *   - functions never allocate or free caller memory, keep no mutable global state and are re-entrant;
 *   - return 0 on success, <0 on error (OMNI_ERR_*); no exceptions cross the ABI.
 */
#ifndef OMNI_AVSR_H_
#define OMNI_AVSR_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OMNI_ABI_VERSION 1

#define OMNI_ACT_NONE 0
#define OMNI_ACT_RELU 1
#define OMNI_ACT_GELU 2
/* SwiGLU pair epilogue (LlamaMLP / Qwen2MLP, transformers modeling_llama.py `down_proj(act_fn(gate_proj(x)) * up_proj(x))`):
 * the N columns are [gate 64 | up 64] blocks (weight rows interleaved in 64-row blocks by the caller); `out` receives the
 * bf16 gate|up tile as usual and `out2` [M, N/2] = bf16( bf16(silu(gate)) * up ) computed from the ROUNDED gate / up values,
 * i.e. bit-identical to omni_swiglu_fwd on `out`.  CTA-pair kernel only (N % 256 == 0, no bias / residual / fp32 output);
 * OMNI_ERR_UNSUPPORTED otherwise. */
#define OMNI_ACT_SWIGLU64 3
/* GELU with the pre-activation kept (AV-HuBERT fc1 while its LoRA adapters train: the backward needs x): `out` [M, N] =
 * bf16(A.B^T + bias), `out2` [M, N] (ld = ldo2) = bf16(gelu(out)) with the same erf formula as OMNI_ACT_GELU.  CTA-pair
 * kernel only; OMNI_ERR_UNSUPPORTED otherwise. */
#define OMNI_ACT_GELU_KEEP 4
/* ResNet BasicBlock epilogue of the ring-padded convolution GEMMs (resnet.py:35-74; csrc/resnet_trunk.cu): the GEMM row is
 * `ring_group` consecutive pixels x `ring_c` channels (N = ring_group * ring_c).  Per element, with the rounding points of
 * the unfused sequence conv -> omni_prelu_res_ring:  f = bf16(bf16(acc) + bias[ch]);  if residual: f = bf16(f +
 * bf16(residual + res_bias[ch]));  f = PReLU(f, slope[ch]);  pixels on the one-pixel ring of their [ring_h + 2, ring_w + 2]
 * frame are stored as 0.  bias / slope / res_bias are [ring_c].  CTA-pair kernel only; OMNI_ERR_UNSUPPORTED otherwise. */
#define OMNI_ACT_PRELU_RING 5
/* Fused backward epilogues of the dgrad GEMMs that produce d(activation) (CTA-pair kernel only; OMNI_ERR_UNSUPPORTED
 * otherwise; `residual` is NOT added, it carries the tensor saved by the forward):
 *   OMNI_ACT_SWIGLU_BWD64: A = dY, B = W_down^T [I, H] -> d(act) [M, I = N] never reaches memory; residual = gate|up [M, 2I]
 *     in the 64-column interleave of OMNI_ACT_SWIGLU64, out = d(gate|up) [M, 2I] (same layout, ldo >= 2N):
 *     the arithmetic of omni_swiglu_bwd_blocked on the bf16-rounded d(act)  (LlamaMLP backward, Llama_LoRA.py MLP).
 *   OMNI_ACT_GELU_BWD: residual = pre-activation [M, N], out = d(pre) = bf16(d(act)) * gelu'(pre)  (omni_gelu_bwd; the
 *     AV-HuBERT fc2 -> fc1 backward under LoRA fine-tuning, wav2vec2.py:1003-1006). */
#define OMNI_ACT_SWIGLU_BWD64 6
#define OMNI_ACT_GELU_BWD 7

#define OMNI_COMPRESS_AVG 0   /* nn.AvgPool1d(r)  : modeling_OmniAVSR.py:544-546 (audio), :469-471 (video) */
#define OMNI_COMPRESS_STACK 1 /* frame stacking   : modeling_OmniAVSR.py:562-568 (audio), :487-493 (video) */

/* ABI version + build info (no GPU needed). */
int omni_abi_version(void);
/* Returns the compute capability (major*10+minor) of the current device, or <0. */
int omni_device_cc(void);

/* ------------------------------------------------------------------------------------------------
 * tcgen05 bf16 GEMM:  out = epi( alpha * (A . B^T  [+ K-extension blocks A2 . B2^T]) )
 * Replaces: every nn.Linear on the path; with tile_group/ext_table it is the Omni-LoRA-adapted
 * q/v projection (Llama_LoRA.py:246-259, Qwen_LoRA.py:557-570), the adapter picked per 128-token
 * tile by task id.
 * ---------------------------------------------------------------------------------------------- */
typedef struct omni_gemm_args {
  const void* A;   /* [M, K]  ld = lda */
  const void* B;   /* [b_rows, K] ld = ldb ; row of B for an output tile = b_row_table ? table : n0 */
  const void* A2;  /* [M, a2_cols] ld = lda2 (K-extension A operand, e.g. s*x.A_lora^T) */
  const void* B2;  /* [b2_rows, b2_cols] ld = ldb2 (K-extension B operand, e.g. LoRA up weights) */
  void* out;       /* [M, N] ld = ldo ; bf16, or fp32 if out_fp32 */
  const void* bias;     /* [N] bf16 or NULL */
  const void* residual; /* [M, N] ld = ldr, bf16 or NULL (added after activation) */
  const int32_t* tile_group;  /* [ceil(M/128)] task id per M tile or NULL (=group 0) */
  const int32_t* b_row_table; /* [groups * n_tiles] B row per (group, n_tile) or NULL */
  const int32_t* ext_table;   /* [groups * n_tiles * n_ext][4] = {a2_col, b2_row, b2_col, 0}; b2_row<0: skip */
  int64_t lda, ldb, lda2, ldb2, ldo, ldr;
  int32_t M, N, K;         /* K = 0 with an ext_table: the reduction is the K-extension list alone (CTA-pair kernel) */
  int32_t b_rows;          /* rows of B visible to the tensor map (>= N; more when grouped) */
  int32_t a2_cols, b2_rows, b2_cols;
  int32_t n_ext;           /* extension slots per (group, n_tile) */
  int32_t block_n;         /* 0 = auto, else 64 / 128 / 256 (tables are indexed with this tile width) */
  int32_t act;             /* OMNI_ACT_* */
  int32_t out_fp32;
  float alpha;
  int32_t pair_aligned;    /* 1: tile_group is constant over every pair of consecutive 128-row tiles (segments start on
                              256-row boundaries), which lets the K-extended GEMM run on the CTA-pair kernel */
  void* out2;              /* OMNI_ACT_SWIGLU64: [M, N/2] bf16; OMNI_ACT_GELU_KEEP: [M, N] bf16; ld = ldo2; else NULL */
  int64_t ldo2;
  const void* slope;       /* OMNI_ACT_PRELU_RING: per-channel PReLU slope [ring_c] */
  const void* res_bias;    /* OMNI_ACT_PRELU_RING: folded-BatchNorm shift of the residual branch [ring_c] or NULL */
  int32_t ring_h, ring_w, ring_group, ring_c;
  void* workspace;         /* omni_gemm_skinny_bf16 only: split-K exchange buffer (256-byte aligned device memory, zero-filled
                              once by the caller; the kernel leaves its counters at zero), or NULL = no split-K */
  int64_t workspace_bytes; /* >= omni_gemm_skinny_workspace_bytes() */
  /* omni_gemm_skinny_bf16 with split-K only (N = the full row width, N % 8 == 0, act NONE): RMSNorm of the finished rows in
   * the SAME launch.  The CTA that completes the last 128-feature tile of a token slice re-reads those rows from L2 and
   * writes norm_out[t, :] = norm_weight * bf16(out[t, :] * rsqrt(mean(out[t, :]^2) + norm_eps)) -- the arithmetic (and the
   * bits) of omni_rmsnorm_fwd.  Replaces the LlamaRMSNorm that follows o_proj / down_proj in a decode step
   * (post_attention_layernorm, the next layer's input_layernorm, the final norm).  norm_out NULL = off. */
  const void* norm_weight; /* [N] bf16 */
  void* norm_out;          /* [M, N] bf16, ld = norm_ld */
  int64_t norm_ld;
  float norm_eps;
} omni_gemm_args;

int omni_gemm_bf16(const omni_gemm_args* args, void* stream);

/* Weight-streaming variant for the decode step (M <= 128 token rows; every nn.Linear of LlamaDecoderLayer_lora under HF
 * generate, Llama_LoRA.py:580-655, one token per sequence): same argument block and the same results up to fp32
 * summation order.  The weights are the M operand of the MMA (128 output features per CTA), K is split over a cluster of up
 * to 8 co-resident CTAs whose partial sums are exchanged through `workspace` (L2-resident), so that ~all SMs stream weight bytes.
 * Differences to omni_gemm_bf16: K % 64 == 0; bf16 output only; b_row_table is indexed per 64-feature block and
 * ext_table per 128-feature tile (block_n must be 128 when given); OMNI_ACT_SWIGLU64 needs N % 128 == 0 and `out` may be
 * NULL (only out2 = the activation is written); OMNI_ACT_GELU_KEEP is not available. */
int omni_gemm_skinny_bf16(const omni_gemm_args* args, void* stream);
/* Upper bound of the workspace any omni_gemm_skinny_bf16 call needs on the current device (fp32 partial tiles of at most
 * SM-count CTAs x 128 tokens x 128 features + the per-tile counters). */
int64_t omni_gemm_skinny_workspace_bytes(void);

/* ------------------------------------------------------------------------------------------------
 * Whisper log-mel front end on the device: audio [B, T] (fp32 or bf16, batch stride audio_bs) -> bf16 [B, 80, 3000].
 * Replaces the host WhisperFeatureExtractor call + D2H/H2D round trip of modeling_OmniAVSR.py:531-534.
 * mel_filters: fp32 [80, 201] slaney filter bank; workspace >= omni_logmel_workspace_bytes(B).
 * ---------------------------------------------------------------------------------------------- */
int64_t omni_logmel_workspace_bytes(int32_t B);
int omni_logmel(const void* audio, int32_t audio_is_bf16, int64_t audio_bs, int32_t B, int32_t T,
                const float* mel_filters, void* out, void* workspace, int64_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Weight-gradient GEMM (reduction over tokens, both operands token-major / "MN-major" for tcgen05):
 *   out[z][i, j] = alpha * sum_{k in [k0[z], k1[z])} A[k, a_col0 + i] * B[k, b_col0 + j]   (+ out if accumulate)
 * Replaces what autograd derives for the trainable tensors of the path: LoRA down/up of
 * Llama_LoRA.py:246-259 / multihead_attention.py:485-494 (one token range per task run) and the projector
 * linears of modeling_OmniAVSR.py:353,366.  omni_colsum_bf16 = bias gradient (sum over tokens).
 * ---------------------------------------------------------------------------------------------- */
#define OMNI_WGRAD_MAX_RANGES 8
typedef struct omni_wgrad_args {
  const void* A;  /* [K, a_cols] ld = lda, bf16 (e.g. dY) */
  const void* B;  /* [K, b_cols] ld = ldb, bf16 (e.g. X)  */
  void* out;      /* [n_ranges][Mo, No] ld = ldo, z stride out_zstride (elements); bf16 or fp32 */
  int64_t lda, ldb, ldo, out_zstride;
  int32_t K, a_cols, b_cols;
  int32_t Mo, No, a_col0, b_col0;
  int32_t n_ranges;
  int32_t k0[OMNI_WGRAD_MAX_RANGES], k1[OMNI_WGRAD_MAX_RANGES];
  int32_t out_fp32, accumulate;
  float alpha;
} omni_wgrad_args;
int omni_gemm_wgrad_bf16(const omni_wgrad_args* args, void* stream);
int omni_colsum_bf16(const void* x, void* out, int64_t rows, int32_t cols, int64_t ld, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Matryoshka compression of encoder features (truncate to n_tok, drop the remainder, pool/stack).
 *   x   [B, t_stride_rows, D] (only the first n_tok rows of each clip are read; batch stride x_bs elements)
 *   out [B, n_tok / rate, D]  (avg)   or   [B, n_tok / rate, rate*D]  (stack)
 * avg: out = bf16( (sum_{i<rate} fp32(x[b, j*rate+i, :])) / rate ), sequential order -- bit-exact with
 * transpose -> nn.AvgPool1d(rate) -> transpose  (modeling_OmniAVSR.py:537,544-546 / 469-471).
 * ---------------------------------------------------------------------------------------------- */
int omni_matryoshka_compress(const void* x, void* out, int32_t B, int32_t n_tok, int32_t D, int64_t x_bs,
                             int32_t rate, int32_t mode, void* stream);
/* backward of the avg mode: dx[b, j*rate+i, :] = bf16(fp32(dout[b,j,:]) / rate) for j < n_tok/rate, 0 for the
 * dropped remainder rows (rows >= n_tok are not written). dx has batch stride dx_bs. */
int omni_matryoshka_compress_bwd(const void* dout, void* dx, int32_t B, int32_t n_tok, int32_t D, int64_t dx_bs,
                                 int32_t rate, int32_t mode, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Prompt splice: builds the LLM input embeddings and labels of the ASR / VSR / AVSR sequences in one
 * launch, straight into their final rows (modeling_OmniAVSR.py:270-299 + 337-395 train, :406-458 infer).
 *   has_bos = 1 (Llama):  [X[0], A?, V?, P_t, X[1:]]       has_bos = 0 (Qwen): [A?, V?, P_t, X]
 *   A = [e(<audio>), audio_tok[b], e(</audio>)], V likewise, P_t = prompt of task t, X = embed(tokens[b]).
 *   labels_t = [lab[0], -100 x (P_t + media rows), lab[1:]]  (Qwen: [-100 x ..., lab]).
 * task_mask bit0 = ASR, bit1 = VSR, bit2 = AVSR. Decode prefill = the same layout with L = 1 (Llama) / 0 (Qwen).
 * out_t is [B, S_t, H]; S_t is implied by the lengths and returned through omni_splice_seq_len().
 * ---------------------------------------------------------------------------------------------- */
typedef struct omni_splice_args {
  const int64_t* tokens; /* [B, L] */
  const int64_t* labels; /* [B, L] or NULL (no labels written) */
  const void* embed;     /* [V, H] embedding table */
  const void* audio_tok; /* [B, n_a, H] projected audio tokens or NULL */
  const void* video_tok; /* [B, n_v, H] projected video tokens or NULL */
  const void* prompt[3]; /* [P_t, H] per task (audio, video, audiovisual) */
  void* out[3];          /* [B, S_t, H] per task or NULL */
  int64_t* out_labels[3];/* [B, S_t] per task or NULL */
  int32_t prompt_len[3];
  int32_t B, L, H, n_a, n_v;
  int32_t id_audio_sos, id_audio_eos, id_video_sos, id_video_eos;
  int32_t has_bos;
  int32_t task_mask;
  int64_t vocab;         /* rows of embed (bounds check for token ids) */
  int32_t* status;       /* optional device int: set to 1 when a token id is outside [0, vocab) (row zero-filled) */
} omni_splice_args;

int32_t omni_splice_seq_len(const omni_splice_args* args, int32_t task);
int omni_splice_prompt(const omni_splice_args* args, void* stream);
/* backward: accumulates d(audio_tok) / d(video_tok) = sum over the sequences that consumed them
 * (own task + AVSR) from dout[3]; fp32 accumulation, bf16 result. */
int omni_splice_prompt_bwd(const omni_splice_args* args, const void* const dout[3], void* d_audio_tok,
                           void* d_video_tok, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Fused Matryoshka path (north_star subsystem 1): compression -> projector MLP -> splice, ONE persistent launch.
 * Replaces, for both modalities at once, the reference's op chain
 *   transpose -> AvgPool1d(rate) -> transpose  or  the frame-stacking views   (modeling_OmniAVSR.py:544-546,562-568 audio,
 *                                                                              :469-471,487-493 video)
 *   Linear(K1 -> I) + ReLU + Linear(I -> H)                                   (:366 audio_proj, :353 video_proj)
 *   cat(<audio>, tokens, </audio>) / cat(<video>, tokens, </video>)            (:347-368)
 *   cat(bos, media, prompt, text) per task + the label tensors                 (:270-299, :373-387; infer :406-458)
 * Kernel structure (csrc/pool_project_splice.cu): CTA pairs (tcgen05 cta_group::2) walk a static work list
 * [GEMM-1 tiles | GEMM-2 tiles] of both modalities; four extra warps per CTA pool the encoder rows (TMA-staged r-row boxes,
 * fp32 sum in window order, /rate, bf16: bit-exact with AvgPool1d) into `pooled`, then copy the marker / prompt / text
 * embedding rows and write the labels.  GEMM-1 tiles wait on per-256-row-block "pooled" counters, GEMM-2 tiles on the
 * "hidden" counters of GEMM-1; `pooled` and `hidden` stay in L2 between producer and consumer.  GEMM-2's epilogue
 * (+bias, round to bf16) writes every projected token straight to its row in the own-task sequence and in the AVSR
 * sequence -- the projected tokens never exist as a separate tensor unless `tok` is given.
 * `splice` describes the destination exactly as for omni_splice_prompt; its audio_tok / video_tok pointers are ignored
 * (presence = audio.x / video.x != NULL) and n_a / n_v must equal n_tok / rate of the modality.
 * Limits: rate <= 256; D, I, H multiples of 8.  workspace: omni_pps_workspace_bytes() bytes of device memory (dependency
 * counters; zeroed by the call itself).  mode = OMNI_COMPRESS_AVG / OMNI_COMPRESS_STACK (K1 = D / rate * D).
 * ---------------------------------------------------------------------------------------------- */
typedef struct omni_pps_modality {
  const void* x;      /* [B, >= n_tok, D] encoder output, batch stride x_bs elements; NULL = modality absent */
  int64_t x_bs;
  int32_t n_tok, rate, D;
  const void* w1;     /* [I, K1] */
  const void* b1;     /* [I] */
  const void* w2;     /* [H, I] */
  const void* b2;     /* [H] */
  void* pooled;       /* [B * (n_tok / rate), K1] compressed features (side output: backward, parity checks) */
  void* hidden;       /* [B * (n_tok / rate), I]  relu(pooled . w1^T + b1) (side output) */
  void* tok;          /* optional [B * (n_tok / rate), H] dense copy of the projected tokens, or NULL */
} omni_pps_modality;

typedef struct omni_pps_args {
  omni_pps_modality audio, video;
  omni_splice_args splice;
  int32_t I;          /* projector intermediate width */
  int32_t mode;       /* OMNI_COMPRESS_* */
  void* workspace;
  int64_t workspace_bytes;
} omni_pps_args;

int64_t omni_pps_workspace_bytes(const omni_pps_args* args);
int omni_pool_project_splice(const omni_pps_args* args, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Flash-attention forward (tcgen05, head_dim 64 or 128) over one segment of the packed q|k|v rows:
 *   qkv [M, ld] bf16, every row = [q heads | k heads | v heads]; the segment is B clips x S tokens from row0.
 *   out [M, out_ld] bf16 (n_heads*head_dim columns written for the segment's rows); lse optional fp32 [n_heads, M].
 * Replaces F.scaled_dot_product_attention at Llama_LoRA.py:300 / Qwen_LoRA.py:606 (causal GQA), the HF Whisper encoder
 * self-attention and fairseq multihead_attention.py:619-654 (non-causal).  Returns OMNI_ERR_UNSUPPORTED for
 * any other head_dim (none of the named architectures).
 * ---------------------------------------------------------------------------------------------- */
int omni_attention_fwd(const void* qkv, int64_t M, int64_t ld, void* out, int64_t out_ld, float* lse, int32_t row0,
                       int32_t B, int32_t S, int32_t n_heads, int32_t n_kv_heads, int32_t head_dim, int32_t causal,
                       float scale, void* stream);

/* Flash-attention backward (tcgen05) of the same segment: dqkv [M, dqkv_ld] bf16 receives dQ | dK | dV in the column
 * layout of qkv (the segment's rows only; GQA group sums included).  out / lse are the forward results, dout the
 * gradient of out [M, dout_ld]; delta is caller-owned scratch of omni_attention_bwd_scratch_floats(B, S, n_heads) fp32
 * values, 16-byte aligned (the pre-pass writes per 64-query step the rows' lse * log2(e) and rowsum(dO o O) there, in
 * 512-byte blocks the dK/dV kernel fetches with one bulk copy per step).  Replaces what autograd derives for
 * the SDPA calls listed above (Llama_LoRA.py:300, Qwen_LoRA.py:606, multihead_attention.py:619-654). */
int64_t omni_attention_bwd_scratch_floats(int32_t B, int32_t S, int32_t n_heads);
int omni_attention_bwd(const void* qkv, int64_t M, int64_t ld, const void* out, int64_t out_ld, const void* dout,
                       int64_t dout_ld, const float* lse, float* delta, void* dqkv, int64_t dqkv_ld, int32_t row0,
                       int32_t B, int32_t S, int32_t n_heads, int32_t n_kv_heads, int32_t head_dim, int32_t causal,
                       float scale, void* stream);

/* Decode-step attention over the static KV cache (one new token per clip): appends the token's K / V (read from its
 * packed q|k|v row, RoPE already applied) to k_cache / v_cache [B, n_kv_heads, max_len, head_dim] at position *len_idx
 * (device memory: the step is replayed from a CUDA graph) and attends over positions 0..*len_idx.  out [B, out_ld].
 * Replaces the cache update + SDPA of Llama_LoRA.py:284-300 / Qwen_LoRA.py:590-606 under HF generate.  GQA group sizes
 * 1..8 (every architecture of the reference's table) and head_dim 64/128; anything else returns OMNI_ERR_UNSUPPORTED. */
int omni_decode_attention(const void* qkv, int64_t ld, void* k_cache, void* v_cache, const int64_t* len_idx, void* out,
                          int64_t out_ld, int32_t B, int32_t n_heads, int32_t n_kv_heads, int32_t head_dim,
                          int32_t max_len, float scale, void* stream);
/* Same with the rotary embedding fused in: the packed row holds q|k|v BEFORE RoPE; the q heads and the new key are rotated
 * at position *len_idx with the bf16 tables cos_t / sin_t [table_rows >= max_len, head_dim] (rounding points of omni_rope /
 * apply_rotary_pos_emb, Llama_LoRA.py:277) before the key is appended and attended.  Saves the separate omni_rope launch of
 * every layer in a decode step. */
int omni_decode_attention_rope(const void* qkv, int64_t ld, void* k_cache, void* v_cache, const int64_t* len_idx, void* out,
                               int64_t out_ld, int32_t B, int32_t n_heads, int32_t n_kv_heads, int32_t head_dim,
                               int32_t max_len, float scale, const void* cos_t, const void* sin_t, int32_t table_rows,
                               void* stream);
/* Beam-search variant (HF 4.43.1 `_beam_search` under eval_OmniAVSR.py:216-226, num_beams = K): the B rows are B/K
 * utterances x K beams.  Instead of HF's `_reorder_cache` gather of the whole cache after every step, each physical cache row
 * keeps what its own forwards appended and the kernel resolves the row per key position p:
 *     p <  *prefill_len : row (b / K) * K               (the prompt is stored once per utterance)
 *     p >= *prefill_len : beam_ind[par][b][p - *prefill_len],  par = (*len_idx - *prefill_len + 1) & 1
 * beam_ind is the double-buffered table [2][B][ind_ld] (int32) maintained by omni_beam_select; the new token (position
 * *len_idx) is appended to row b itself.  cos_t / sin_t may be NULL (RoPE already applied). */
int omni_decode_attention_beam(const void* qkv, int64_t ld, void* k_cache, void* v_cache, const int64_t* len_idx, void* out,
                               int64_t out_ld, int32_t B, int32_t n_heads, int32_t n_kv_heads, int32_t head_dim,
                               int32_t max_len, float scale, const void* cos_t, const void* sin_t, int32_t table_rows,
                               const int32_t* beam_ind, int32_t ind_ld, const int64_t* prefill_len, int32_t K,
                               void* stream);

/* ------------------------------------------------------------------------------------------------
 * Beam-search ranking + scorer bookkeeping of one decode step, on the device (no host synchronisation per step).
 * Replaces, inside HF transformers 4.43.1 `GenerationMixin._beam_search` (reached from modeling_OmniAVSR.py:313-322 with the
 * evaluation default num_beams = 15 of eval_OmniAVSR.py:216-226): log_softmax + beam-score add + topk(2K) over [K*V], and
 * `BeamSearchScorer.process` (length_penalty 1.0, early_stopping False).
 *   omni_beam_topk_rows: per logits row (bf16 [rows, ld], V columns): fp32 log-softmax statistics and the row's n_cand best
 *     tokens, cand_score[row][j] = (x - max - log(sum exp)) + beam_scores[row] in descending order, ties by lowest token id;
 *     cand_tok = -1 / score = -inf past V candidates.
 *   omni_beam_select: per utterance, merge the K candidate lists (ties by lowest k*V + tok), walk the n_cand = 2K best like
 *     the scorer, and write the next step's state (see the struct).  All counters live in device memory so that the step can
 *     be replayed from a CUDA graph; `step_idx` is advanced by omni_decode_advance after the step's forward.
 * ---------------------------------------------------------------------------------------------- */
typedef struct omni_beam_select_args {
  const float* cand_score;   /* [B*K, n_cand] from omni_beam_topk_rows */
  const int32_t* cand_tok;   /* [B*K, n_cand] */
  float* beam_scores;        /* [B*K] running sum of log-probs (in: read by omni_beam_topk_rows; out: the chosen beams') */
  const int64_t* step_idx;   /* [1] tokens generated so far (cur_len - 1) */
  const int64_t* eos;        /* [1] */
  const int64_t* pad;        /* [1] */
  int32_t* seqs;             /* [2][B*K, max_new] token history, buffer (step & 1) is read, the other written */
  int32_t* ind;              /* [2][B*K, ind_ld] KV-cache row of every generated position (see omni_decode_attention_beam) */
  int32_t* hyp_seq;          /* [B, K+1, max_new] finished hypotheses (slot storage) */
  int32_t* hyp_len;          /* [B, K+1] */
  double* hyp_score;         /* [B, K+1] sum_logprobs / length */
  int32_t* hyp_order;        /* [B, K+1] slots in insertion order (first hyp_count entries are live); init 0..K */
  int32_t* hyp_count;        /* [B] */
  double* hyp_worst;         /* [B] init 1e9 */
  int32_t* done;             /* [B] */
  int32_t* n_done;           /* [1] number of finished utterances */
  int32_t* status;           /* [1] set to 1 when fewer than K non-EOS candidates exist (HF raises ValueError there) */
  const void* embed;         /* [vocab, H] bf16 embedding table, ld_embed */
  void* x_next;              /* [B*K, H] bf16 input rows of the next forward, ld_x */
  int64_t ld_embed, ld_x;
  int32_t B, K, n_cand, V, max_new, ind_ld, H;
} omni_beam_select_args;
int omni_beam_topk_rows(const void* logits, int64_t ld, int32_t rows, int32_t V, const float* beam_scores, int32_t n_cand,
                        float* cand_score, int32_t* cand_tok, void* stream);
int omni_beam_select(const omni_beam_select_args* args, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Row kernels of the decoder / encoder blocks (all bf16 in/out, fp32 statistics).
 * Replace the transformers==4.43.1 ops reached from Llama_LoRA.py:624,643-644 (RMSNorm), :277 (RoPE),
 * LlamaMLP (SwiGLU), and the fairseq LayerNorm/GELU of the AV-HuBERT blocks
 * (av_hubert/fairseq/fairseq/models/wav2vec/wav2vec2.py:979-1006).
 * ---------------------------------------------------------------------------------------------- */
/* y = w * bf16(x * rsqrt(mean(x^2)+eps)); rstd [rows] fp32 is optional (needed by the backward). */
int omni_rmsnorm_fwd(const void* x, const void* w, void* y, float* rstd, int64_t rows, int32_t H, int64_t ldx,
                     int64_t ldy, float eps, void* stream);
/* dx = rstd*(dy*w - xhat*mean(dy*w*xhat)) (+ dx_add if given); x, dy, dx contiguous [rows, H]. */
int omni_rmsnorm_bwd(const void* dy, const void* x, const void* w, const float* rstd, void* dx, const void* dx_add,
                     int64_t rows, int32_t H, void* stream);
int omni_layernorm_fwd(const void* x, const void* w, const void* b, void* y, float* mean, float* rstd, int64_t rows,
                       int32_t H, int64_t ldx, int64_t ldy, float eps, void* stream);
int omni_layernorm_bwd(const void* dy, const void* x, const void* w, const float* mean, const float* rstd, void* dx,
                       const void* dx_add, int64_t rows, int32_t H, void* stream);
/* In-place rotary embedding of the first n_heads_total heads of every row of a packed [rows, ld] q|k|v buffer;
 * cos/sin are bf16 tables [max_pos, head_dim], pos [rows] int32. inverse=1 applies the transposed rotation. */
int omni_rope(void* qkv, const void* cos_t, const void* sin_t, const int32_t* pos, int64_t rows, int64_t ld,
              int32_t n_heads_total, int32_t head_dim, int32_t inverse, void* stream);
/* gu [rows, 2I] = [gate | up] -> act [rows, I] = bf16(bf16(silu(gate)) * up), and its backward. */
int omni_swiglu_fwd(const void* gu, void* act, int64_t rows, int32_t I, void* stream);
int omni_swiglu_bwd(const void* dact, const void* gu, void* dgu, int64_t rows, int32_t I, void* stream);
/* Same for the [gate blk | up blk] block-interleaved layout OMNI_ACT_SWIGLU64 produces (blk = 64; blk = I is the plain
 * [gate | up] layout of omni_swiglu_bwd). */
int omni_swiglu_bwd_blocked(const void* dact, const void* gu, void* dgu, int64_t rows, int32_t I, int32_t blk, void* stream);
int omni_gelu_fwd(const void* x, void* y, int64_t n, void* stream);
int omni_gelu_bwd(const void* dy, const void* x, void* dx, int64_t n, void* stream);
/* ResNet front-end glue of AV-HuBERT (av_hubert/avhubert/resnet.py:35-74,131-169), channels-last [rows, C]:
 * x <- PReLU((x + bias) (+ residual + res_bias)) in place (per-channel slope; bias / res_bias = optional folded-BatchNorm
 * shifts of the convolutions that produced x / residual), and PReLU + MaxPool(3x3, stride 2, pad 1) in one pass
 * (frontend3D's PReLU + MaxPool3d((1,3,3),(1,2,2),(0,1,1))): x [N, H, W, C] -> y [N, Ho, Wo, C]. */
int omni_prelu_res(void* x, const void* residual, const void* slope, const void* bias, const void* res_bias,
                   int64_t rows, int32_t C, void* stream);
int omni_prelu_maxpool3x3s2(const void* x, const void* slope, void* y, int64_t N, int32_t H, int32_t W, int32_t C,
                            void* stream);
/* im2col of the video front-end Conv3d(1,64,(5,7,7),stride (1,2,2),pad (2,3,3)) (resnet.py:137): video [B,T,H,W] bf16 ->
 * out [B*T*Ho*Wo, 256] bf16 (245 taps + 11 zero columns), so the convolution runs as one omni_gemm_bf16 call. */
int omni_im2col_front3d(const void* video, void* out, int32_t B, int32_t T, int32_t H, int32_t W, void* stream);

/* Time-major variant (4x less traffic): out2 [B][Ho][Wo][T+4][64] holds the 49 spatial taps of every (position, padded
 * time) row; the five temporal taps are the five consecutive rows, consumed by omni_gemm_bf16 through an overlapping-row
 * view (lda = 64, K = 320).  omni_prelu_maxpool_front = PReLU + MaxPool3d((1,3,3),(1,2,2),(0,1,1)) (resnet.py:139-140)
 * reading the GEMM output [B][H][W][T+4][C] and writing channels-last [B*T][Hp][Wp][C]. */
int omni_im2col_front2d(const void* video, void* out2, int32_t B, int32_t T, int32_t H, int32_t W, void* stream);
int omni_prelu_maxpool_front(const void* x, const void* slope, void* y, int32_t B, int32_t T, int32_t H, int32_t W,
                             int32_t C, void* stream);
/* ResNet-18 trunk without library convolutions (resnet.py:35-74,77-129,156-164): activations are channels-last frames with a
 * one-pixel ZERO RING, [N, H+2, W+2, C].  A 3x3 stride-1 convolution is ONE omni_gemm_bf16 call on an overlapping-row view
 * of that buffer (row stride C; main K = the dy=-1 taps, K-extension blocks = the dy=0 / +1 taps; see csrc/resnet_trunk.cu).
 *   omni_prelu_maxpool_front_ring: omni_prelu_maxpool_front writing the ring-padded layout (interior only; the caller keeps
 *     the ring at zero);
 *   omni_prelu_res_ring: omni_prelu_res on the interior pixels, zeros on the ring (the GEMM leaves garbage there);
 *   omni_gather_s2_ring: GEMM operand rows of the strided convolutions for every position of the ring-padded OUTPUT grid,
 *     taps = 9 (conv3x3 pad 1: out [N (Ho+2)(Wo+2), 9 C], tap-major) or 1 (1x1 downsample: [.., C]); stride 2 in the trunk,
 *     stride 1 for channel counts the overlapping-row GEMM cannot take (3 C not a multiple of 64);
 *   omni_avgpool_ring: AdaptiveAvgPool2d(1) over the interior -> [N, C]. */
int omni_prelu_maxpool_front_ring(const void* x, const void* slope, void* y, int32_t B, int32_t T, int32_t H, int32_t W,
                                  int32_t C, void* stream);
int omni_prelu_res_ring(void* x, const void* residual, const void* slope, const void* bias, const void* res_bias, int64_t N,
                        int32_t H, int32_t W, int32_t C, void* stream);
int omni_gather_s2_ring(const void* x, void* out, int64_t N, int32_t H, int32_t W, int32_t C, int32_t taps, int32_t stride,
                        void* stream);
int omni_avgpool_ring(const void* x, void* out, int64_t N, int32_t H, int32_t W, int32_t C, void* stream);
/* Table-driven convolution of the small late-stage grids (layers 2-4 of the trunk, resnet.py:108-110): the GEMM row is ONE
 * FRAME (plain channels-last [N, P_alloc, C], no ring), an N tile is the channels of one output pixel (or of 256 / C_out
 * consecutive output pixels), and the whole reduction is the K-extension list of that tile -- omni_gemm_bf16 with K = 0
 * (A / B may be NULL) and one {column of the input pixel, 0, column of the filter-pattern block} entry per 64 input
 * channels of every contributing input pixel; taps that fall on the zero padding are simply absent.  OMNI_ACT_PRELU_RING
 * with ring_h = ring_w = 0 is the same BasicBlock epilogue without ring zeroing.  omni_avgpool_frames: mean over the first
 * P pixels of every frame -> [N, C] (AdaptiveAvgPool2d(1), resnet.py:163). */
int omni_avgpool_frames(const void* x, void* out, int64_t N, int32_t P, int32_t P_alloc, int32_t C, void* stream);
/* out[i,:] = table[idx[i],:] (embed_tokens of the decode step, label-row selection); status as in the splice. */
int omni_gather_rows(const void* table, const int64_t* idx, void* out, int64_t n, int32_t H, int64_t ld_table,
                     int64_t table_rows, int32_t* status, void* stream);
/* Batched transpose: in [Z, R, C] bf16 -> out [Z, C, R] (per-step transposed copies of the trainable LoRA / projector
 * matrices for the dgrad GEMMs; autograd's `.t().contiguous()`). */
int omni_transpose_bf16(const void* in, void* out, int32_t Z, int32_t R, int32_t C, void* stream);
/* out[idx[i],:] = src[i,:] for unique idx. */
int omni_scatter_rows(const void* src, const int64_t* idx, void* out, int64_t n, int32_t H, int64_t ld_out,
                      void* stream);

/* ------------------------------------------------------------------------------------------------
 * Loss / decode head / optimizer.
 * omni_ce_*: `logits.float()` + CrossEntropyLoss (Llama_LoRA.py:373-386) on the label rows only:
 *   loss[r] = logsumexp(logits[r]) - logits[r, target[r]]  (0 for ignore_index); the backward overwrites the
 *   logits with bf16((softmax - onehot) * scale[r]).
 * omni_argmax: greedy token choice of HF generate (modeling_OmniAVSR.py:313-322 with num_beams=1).
 * omni_sumsq + omni_adamw: gradient_clip_val (train_OmniAVSR.py:53) + AdamW (lightning_OmniAVSR.py:153) fused
 *   over the flat trainable buffer: grad scaled by grad_scale*min(1, max_norm/(grad_scale*sqrt(*sumsq)+1e-6)).
 * ---------------------------------------------------------------------------------------------- */
int omni_ce_fwd(const void* logits, const int64_t* targets, float* loss, float* lse, int64_t rows, int32_t V,
                int64_t ld, int64_t ignore_index, void* stream);
int omni_ce_bwd(void* logits, const int64_t* targets, const float* lse, const float* scale, int64_t rows, int32_t V,
                int64_t ld, int64_t ignore_index, void* stream);
int omni_argmax(const void* logits, int64_t* out, int64_t rows, int32_t V, int64_t ld, void* stream);
/* Token bookkeeping of one greedy decode step in ONE launch (HF `_sample` semantics of transformers 4.43.1, driven by
 * modeling_OmniAVSR.py:313-322): tok = argmax(logits[b, :V]) (first maximal index), pad for finished sequences;
 * out[step, b] = tok; unfinished[b] &= tok != eos; alive[step] = any unfinished (caller zeroes alive once per decode);
 * x_next[b, :] = embed[tok, :] (the next forward's input row).  step / eos / pad are device scalars (CUDA-graph replay).
 * omni_decode_advance: step_idx += 1, len_idx += 1, pos[0..n_pos) += 1 after the step's forward. */
int omni_decode_pick(const void* logits, int32_t V, int64_t ld, int64_t* unfinished, const int64_t* eos, const int64_t* pad,
                     const int64_t* step_idx, int64_t* out, int32_t B, int64_t* alive, const void* embed, int64_t ld_embed,
                     void* x_next, int64_t ld_x, int32_t H, void* stream);
int omni_decode_advance(int64_t* step_idx, int64_t* len_idx, int32_t* pos, int32_t n_pos, void* stream);
int omni_sumsq(const void* g, int64_t n, float* acc, void* stream);
int omni_adamw(void* p, const void* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps,
               float weight_decay, int32_t step, float grad_scale, float max_norm, const float* sumsq, void* stream);

/* ------------------------------------------------------------------------------------------------
 * On-device input pipeline (datamodule/transforms.py of the reference), one utterance per call.
 * omni_video_transform: VideoTransform (:83-104): uint8 frames [T, C (1 or 3), H, W] -> x/255 -> crop 88x88 at
 *   (crop_i, crop_j) (RandomCrop / CenterCrop offsets chosen by the caller) -> torchvision Grayscale -> AdaptiveTimeMask
 *   (:36-56: frames t in [spans[2k], spans[2k+1]) zeroed) -> Normalize(0.421, 0.165); out [T, 1, 88, 88] fp32 (bit-exact
 *   with the reference's fp32 result) or bf16 (out_bf16 = 1: the cast Lightning's bf16-true applies next).
 * omni_audio_transform: AudioTransform (:107-131): wave [T] fp32 -> AdaptiveTimeMask (samples in the spans zeroed) ->
 *   AddNoise (:59-80, torchaudio.functional.add_noise at snr_db; noise = NULL skips it) -> layer_norm over the whole
 *   utterance (eps 1e-8).  `spans` are HOST arrays of [start, end) pairs (at most OMNI_MAX_MASK_SPANS, passed by value to
 *   the kernels); workspace >= omni_audio_transform_workspace_bytes() device bytes.
 * ---------------------------------------------------------------------------------------------- */
#define OMNI_MAX_MASK_SPANS 48
int omni_video_transform(const void* frames, int32_t T, int32_t C, int32_t H, int32_t W, int32_t crop_i, int32_t crop_j,
                         const int32_t* spans, int32_t n_spans, void* out, int32_t out_bf16, void* stream);
int64_t omni_audio_transform_workspace_bytes(void);
int omni_audio_transform(const float* wave, const float* noise, int64_t T, float snr_db, const int32_t* spans,
                         int32_t n_spans, float* out, void* workspace, int64_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* OMNI_AVSR_H_ */
