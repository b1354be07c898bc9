// Splice layout shared by compress_splice.cu (stand-alone splice kernel) and pool_project_splice.cu (fused path):
// destination row (task, clip, position) -> source row + label, as the reference builds the sequences with torch.cat
// (Omni_AVSR/modeling_OmniAVSR.py:270-299, :337-395 train, :406-458 infer).
#pragma once
#include "common.cuh"
#include "../../include/omni_avsr.h"

namespace omni {

__device__ __forceinline__ void acc8(float (&a)[8], const uint4& u) {
  float2 f;
  f = bf2_to_f2(u.x); a[0] += f.x; a[1] += f.y;
  f = bf2_to_f2(u.y); a[2] += f.x; a[3] += f.y;
  f = bf2_to_f2(u.z); a[4] += f.x; a[5] += f.y;
  f = bf2_to_f2(u.w); a[6] += f.x; a[7] += f.y;
}

struct SpliceK {
  const int64_t* tokens;
  const int64_t* labels;
  const bf16* embed;
  const bf16* audio_tok;
  const bf16* video_tok;
  const bf16* prompt[3];
  bf16* out[3];
  int64_t* out_labels[3];
  int prompt_len[3];
  int S[3];          // sequence length per task (0 = disabled)
  int has_a[3], has_v[3];
  long long row_end[3];  // cumulative B*S
  int B, L, H8, n_a, n_v;
  int id_as, id_ae, id_vs, id_ve;
  int has_bos;
  long long vocab;
  int* status;
};

struct RowSrc {
  const bf16* ptr;   // nullptr => zero row
  long long label;
  int media;         // 1: a projected audio / video token (the fused path writes those rows from the projector epilogue)
};

// Resolves destination row (task t, clip b, position pos) to its source row and label.
__device__ __forceinline__ RowSrc splice_resolve(const SpliceK& k, int t, int b, int pos) {
  RowSrc r;
  r.label = -100;
  r.media = 0;
  long long tok = -1;
  const long long H = static_cast<long long>(k.H8) * 8;
  int p = pos;
  if (k.has_bos) {
    if (p == 0) {
      tok = k.tokens[static_cast<long long>(b) * k.L];
      if (k.labels) r.label = k.labels[static_cast<long long>(b) * k.L];
      goto from_embed;
    }
    p -= 1;
  }
  if (k.has_a[t]) {
    const int seg = k.n_a + 2;
    if (p < seg) {
      if (p == 0) { tok = k.id_as; goto from_embed; }
      if (p == seg - 1) { tok = k.id_ae; goto from_embed; }
      r.ptr = k.audio_tok + (static_cast<long long>(b) * k.n_a + (p - 1)) * H;
      r.media = 1;
      return r;
    }
    p -= seg;
  }
  if (k.has_v[t]) {
    const int seg = k.n_v + 2;
    if (p < seg) {
      if (p == 0) { tok = k.id_vs; goto from_embed; }
      if (p == seg - 1) { tok = k.id_ve; goto from_embed; }
      r.ptr = k.video_tok + (static_cast<long long>(b) * k.n_v + (p - 1)) * H;
      r.media = 1;
      return r;
    }
    p -= seg;
  }
  if (p < k.prompt_len[t]) {
    r.ptr = k.prompt[t] + static_cast<long long>(p) * H;
    return r;
  }
  p -= k.prompt_len[t];
  {
    const long long ti = static_cast<long long>(b) * k.L + k.has_bos + p;
    tok = k.tokens[ti];
    if (k.labels) r.label = k.labels[ti];
  }
from_embed:
  if (tok < 0 || tok >= k.vocab) {
    if (k.status) *k.status = 1;
    r.ptr = nullptr;
  } else {
    r.ptr = k.embed + tok * H;
  }
  return r;
}

// has_audio / has_video < 0: presence = (a->audio_tok / a->video_tok != NULL), the stand-alone splice convention
static inline int fill_splice(const omni_splice_args* a, SpliceK* k, int has_audio = -1, int has_video = -1) {
  if (a->B <= 0 || a->L < 0 || a->H <= 0 || (a->H % 8) != 0) return OMNI_ERR_BAD_ARG;
  if (!a->embed) return OMNI_ERR_BAD_ARG;
  if (a->L > 0 && !a->tokens) return OMNI_ERR_BAD_ARG;
  if (a->has_bos && a->L < 1) return OMNI_ERR_BAD_ARG;
  k->tokens = a->tokens; k->labels = a->labels;
  k->embed = reinterpret_cast<const bf16*>(a->embed);
  k->audio_tok = reinterpret_cast<const bf16*>(a->audio_tok);
  k->video_tok = reinterpret_cast<const bf16*>(a->video_tok);
  k->B = a->B; k->L = a->L; k->H8 = a->H / 8; k->n_a = a->n_a; k->n_v = a->n_v;
  k->id_as = a->id_audio_sos; k->id_ae = a->id_audio_eos; k->id_vs = a->id_video_sos; k->id_ve = a->id_video_eos;
  k->has_bos = a->has_bos ? 1 : 0; k->vocab = a->vocab; k->status = a->status;
  long long acc = 0;
  for (int t = 0; t < 3; ++t) {
    k->prompt[t] = reinterpret_cast<const bf16*>(a->prompt[t]);
    k->out[t] = reinterpret_cast<bf16*>(a->out[t]);
    k->out_labels[t] = a->out_labels[t];
    k->prompt_len[t] = a->prompt_len[t];
    const bool pa = has_audio < 0 ? (a->audio_tok != nullptr) : (has_audio != 0);
    const bool pv = has_video < 0 ? (a->video_tok != nullptr) : (has_video != 0);
    k->has_a[t] = ((t == 0 || t == 2) && pa) ? 1 : 0;
    k->has_v[t] = ((t == 1 || t == 2) && pv) ? 1 : 0;
    const bool on = (a->task_mask >> t) & 1;
    if (on && a->prompt_len[t] > 0 && !a->prompt[t]) return OMNI_ERR_BAD_ARG;
    k->S[t] = on ? ((a->has_bos ? 1 : 0) + k->has_a[t] * (a->n_a + 2) + k->has_v[t] * (a->n_v + 2) + a->prompt_len[t] +
                    (a->L - (a->has_bos ? 1 : 0)))
                 : 0;
    acc += static_cast<long long>(a->B) * k->S[t];
    k->row_end[t] = acc;
  }
  return OMNI_OK;
}


}  // namespace omni
