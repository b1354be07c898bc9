"""ORACLE (test infrastructure only): CPU restatement of the beam search the reference runs at evaluation time.

The reference calls HF `generate(inputs_embeds=..., max_new_tokens=32, num_beams=15, eos/bos/pad ids, modality=...)`
(Omni_AVSR/modeling_OmniAVSR.py:313-322, eval_OmniAVSR.py:216-226).  The algorithm lives in the un-vendored dependency
transformers==4.43.1 (requirements.txt:7): `GenerationMixin._beam_search` (generation/utils.py) driving a
`BeamSearchScorer` (generation/beam_search.py) with the library defaults length_penalty = 1.0, early_stopping = False,
num_beam_groups = 1, num_return_sequences = 1, no logits processors.  This file restates that published algorithm:

  * inputs expanded to B*K rows; beam_scores = [0, -1e9, ..., -1e9] per utterance;
  * per step: fp32 log_softmax of the last-position logits + beam score, top-2K over the K*V candidates (sorted);
  * BeamSearchScorer.process: walk the candidates in rank order; an EOS candidate ranked < K closes a hypothesis with score
    sum_logprobs / generated_len (generated_len counts the EOS position), a non-EOS candidate fills the next beam slot
    until K are taken; an utterance is done when it holds K hypotheses and the worst of them is >= best running score /
    cur_len; finished utterances are fed pad tokens;
  * stop when every utterance is done or after max_new_tokens; finalize adds the open beams as hypotheses, returns the
    best one per utterance, EOS appended if it fits, right-padded with pad_token_id.

Pinned by tests/test_oracle_beam.py against the installed transformers' own `generate(num_beams=K)` on small random
models.  Only `inputs_embeds` is passed, so decoder_prompt_len = 0 and the returned ids hold the new tokens only.
(With a real Llama-3.2 checkpoint the hub's generation_config.json sets do_sample=True; there is no checkpoint here and the
random-init model carries the default GenerationConfig, i.e. deterministic beam search -- noted in DESIGN.md.)
"""
from __future__ import annotations

from typing import Callable, List, Optional, Tuple

import torch


class BeamHypotheses:
    """generation/beam_search.py `BeamHypotheses` (length_penalty 1.0, early_stopping False)."""

    def __init__(self, num_beams: int, length_penalty: float = 1.0, early_stopping=False):
        self.num_beams, self.length_penalty, self.early_stopping = num_beams, length_penalty, early_stopping
        self.beams: List[Tuple[float, torch.Tensor]] = []
        self.worst_score = 1e9

    def __len__(self):
        return len(self.beams)

    def add(self, hyp: torch.Tensor, sum_logprobs: float, generated_len: int):
        score = sum_logprobs / (generated_len ** self.length_penalty)
        if len(self) < self.num_beams or score > self.worst_score:
            self.beams.append((score, hyp))
            if len(self) > self.num_beams:
                ranked = sorted((s, i) for i, (s, _) in enumerate(self.beams))
                del self.beams[ranked[0][1]]
                self.worst_score = ranked[1][0]
            else:
                self.worst_score = min(score, self.worst_score)

    def is_done(self, best_sum_logprobs: float, cur_len: int) -> bool:
        if len(self) < self.num_beams:
            return False
        if self.early_stopping is True:
            return True
        highest_attainable = best_sum_logprobs / cur_len ** self.length_penalty
        return self.worst_score >= highest_attainable


class BeamScorer:
    """generation/beam_search.py `BeamSearchScorer.process` / `.finalize` for one beam group, prompt length 0."""

    def __init__(self, batch_size: int, num_beams: int, length_penalty: float = 1.0, early_stopping=False):
        self.B, self.K = batch_size, num_beams
        self.hyps = [BeamHypotheses(num_beams, length_penalty, early_stopping) for _ in range(batch_size)]
        self.done = [False] * batch_size

    @property
    def is_done(self) -> bool:
        return all(self.done)

    def process(self, input_ids: torch.Tensor, next_scores, next_tokens, next_indices, pad_id: int, eos_id: int):
        """input_ids [B*K, cur] (cpu int64); next_* nested lists [B][2K].  Returns (scores, tokens, beam_idx) lists [B*K]."""
        K = self.K
        cur_len = input_ids.shape[-1] + 1
        out_s, out_t, out_i = [], [], []
        for b in range(self.B):
            if self.done[b]:
                out_s += [0.0] * K
                out_t += [pad_id] * K
                out_i += [0] * K
                continue
            taken = 0
            for rank, (tok, sc, idx) in enumerate(zip(next_tokens[b], next_scores[b], next_indices[b])):
                row = b * K + idx
                if tok == eos_id:
                    if rank >= K:
                        continue
                    self.hyps[b].add(input_ids[row].clone(), sc, cur_len)
                else:
                    out_s.append(sc)
                    out_t.append(tok)
                    out_i.append(row)
                    taken += 1
                if taken == K:
                    break
            if taken < K:
                raise ValueError(f"At most {K} tokens in {next_tokens[b]} can be equal to `eos_token_id: {eos_id}`.")
            self.done[b] = self.done[b] or self.hyps[b].is_done(max(next_scores[b]), cur_len)
        return out_s, out_t, out_i

    def finalize(self, input_ids: torch.Tensor, final_scores, pad_id: int, eos_id: int, max_length: int) -> torch.Tensor:
        K = self.K
        for b in range(self.B):
            if self.done[b]:
                continue
            for k in range(K):
                row = b * K + k
                self.hyps[b].add(input_ids[row], final_scores[row], input_ids.shape[-1])
        best = []
        for b in range(self.B):
            best.append(sorted(self.hyps[b].beams, key=lambda x: x[0])[-1][1])
        lengths = [int(h.shape[0]) for h in best]
        sent_max = min(max(lengths) + 1, max_length)
        fill = pad_id if min(lengths) != max(lengths) else 0
        out = torch.full((self.B, sent_max), fill, dtype=torch.int64)
        for b, h in enumerate(best):
            out[b, : lengths[b]] = h
            if lengths[b] < sent_max:
                out[b, lengths[b]] = eos_id
        return out


def topk_candidates(scores: torch.Tensor, B: int, K: int):
    """scores [B*K, V] fp32 (log-probs + beam score) -> nested lists (scores, tokens, beam indices) of the top 2K."""
    V = scores.shape[-1]
    top_s, top_i = torch.topk(scores.view(B, K * V), 2 * K, dim=1, largest=True, sorted=True)
    return top_s.tolist(), (top_i % V).tolist(), torch.div(top_i, V, rounding_mode="floor").tolist()


@torch.no_grad()
def beam_search(step_logits: Callable, reorder: Callable, B: int, K: int, max_new_tokens: int, eos_id: int, pad_id: int,
                length_penalty: float = 1.0, early_stopping=False, return_scores: bool = False):
    """Driver of generation/utils.py `_beam_search`.

    step_logits(tokens_or_None) -> last-position logits [B*K, V] (any float dtype; None = first step on the expanded
    prompt); reorder(beam_idx LongTensor [B*K]) reorders the model's cache rows before the next step."""
    scorer = BeamScorer(B, K, length_penalty, early_stopping)
    beam_scores = torch.zeros(B, K, dtype=torch.float32)
    beam_scores[:, 1:] = -1e9
    beam_scores = beam_scores.view(-1)
    input_ids = torch.zeros((B * K, 0), dtype=torch.int64)
    tokens = None
    cur_len = 0
    while True:
        logits = step_logits(tokens)
        logp = torch.log_softmax(logits.float().cpu(), dim=-1) + beam_scores[:, None]
        s, t, i = topk_candidates(logp, B, K)
        out_s, out_t, out_i = scorer.process(input_ids, s, t, i, pad_id, eos_id)
        beam_scores = torch.tensor(out_s, dtype=torch.float32)
        tokens = torch.tensor(out_t, dtype=torch.int64)
        beam_idx = torch.tensor(out_i, dtype=torch.int64)
        input_ids = torch.cat([input_ids[beam_idx], tokens[:, None]], dim=-1)
        reorder(beam_idx)
        cur_len += 1
        if scorer.is_done or cur_len >= max_new_tokens:
            break
    out = scorer.finalize(input_ids, beam_scores.tolist(), pad_id, eos_id, max_new_tokens)
    if return_scores:
        best = [sorted(h.beams, key=lambda x: x[0])[-1][0] for h in scorer.hyps]
        return out, best
    return out
