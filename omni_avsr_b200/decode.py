"""Greedy decode with a static KV cache and a CUDA-graph-captured decode step (the `eval_OmniAVSR.py` path:
modeling_OmniAVSR.py:308-323 -> HF GenerationMixin.generate with inputs_embeds, reproduced per SURVEY A.5):

  * only `inputs_embeds` is given, so the returned ids contain ONLY the new tokens;
  * step 0 consumes the embeddings, later steps embed the previously chosen id (Llama_LoRA.py:429-432);
  * `modality` selects the adapter at every step (Llama_LoRA.py:441);
  * greedy = argmax(fp32(logits[:, -1])); a row that emitted EOS is padded with pad_token_id afterwards;
  * generation stops when every row is finished or after max_new_tokens.

Execution: the prefill runs eagerly; the per-token step (embed gather -> N decoder layers -> lm_head -> argmax ->
finished-row bookkeeping -> KV/mask/position advance) is captured ONCE per (batch, max_len) in a CUDA graph and
replayed, so a step costs one graph launch instead of ~250 kernel launches from Python.  Everything the step needs to
advance (cache write index, key mask, RoPE positions, output slot) lives in device tensors updated inside the graph; the
"all rows finished" cut is applied once at the end (same ids as HF's per-step check, one host sync per decode).
"""
from __future__ import annotations

import os

import torch

from . import ops
from .Llama_LoRA import TILE, KVCache, PackedRows, pack_segments


class _StepRows:
    """PackedRows-like layout of the single-token step (B rows, static device tensors).  No padding to the 128-row tile:
    every kernel of the step takes any row count, and the weight-streaming GEMM picks its token-tile width (64 / 128)
    from it."""

    def __init__(self, B, device, max_pos):
        self.M = B
        self.segments = [(0, B, 1, 0)]
        self.valid_rows = B
        self.max_pos = max_pos
        self.runs = [(0, 0, self.M)]
        self.pair_aligned = True
        self.tile_group = torch.zeros((self.M + TILE - 1) // TILE, dtype=torch.int32, device=device)
        self.pos = torch.zeros(self.M, dtype=torch.int32, device=device)


class GraphedGreedyStep:
    """One decode step captured in a CUDA graph (static buffers; see module docstring)."""

    def __init__(self, llm, B, max_len, max_new, device):
        a = llm.config
        self.llm, self.B, self.max_len, self.max_new = llm, B, max_len, max_new
        self.cache = KVCache(a, B, max_len, device)
        self.rows = _StepRows(B, device, max_len)
        self.tok = torch.zeros(B, dtype=torch.int64, device=device)
        self.unfinished = torch.ones(B, dtype=torch.int64, device=device)
        self.out = torch.zeros((max_new, B), dtype=torch.int64, device=device)
        self.alive = torch.zeros(max_new, dtype=torch.int64, device=device)
        self.step_idx = torch.zeros(1, dtype=torch.int64, device=device)
        self.eos = torch.zeros(1, dtype=torch.int64, device=device)
        self.pad = torch.zeros(1, dtype=torch.int64, device=device)
        self.xpad = torch.zeros((self.rows.M, a.hidden_size), dtype=torch.bfloat16, device=device)
        self.h_last = torch.zeros((B, a.hidden_size), dtype=torch.bfloat16, device=device)
        self.graph = None
        llm.model.rope(max_len)          # make sure the RoPE tables cover max_len before capture

    def _body(self):
        """pick token from h_last -> bookkeeping -> (embed -> layers) -> new h_last, advance device state."""
        llm, B = self.llm, self.B
        logits = llm.logits_rows(self.h_last)
        # argmax + pad-after-EOS + output slot + unfinished / alive flags + embedding row of the chosen token: one launch
        ops.decode_pick(logits, llm.config.vocab_size, self.unfinished, self.eos, self.pad, self.step_idx, self.out,
                        self.alive, llm.model.embed_tokens.weight.data, self.xpad)
        # forward of the chosen token
        hid = llm.model.forward_packed(self.xpad, self.rows, self.cache)
        self.h_last.copy_(hid[:B])
        ops.decode_advance(self.step_idx, self.cache.len_idx, self.rows.pos)

    def start(self, task, prefill_len, h_last, eos, pad):
        self.rows.tile_group.fill_(task)
        self.rows.pos.fill_(prefill_len)
        self.cache.len = prefill_len
        self.cache.sync_device_state()
        self.cache.graph_mode = True
        self.tok.zero_()
        self.unfinished.fill_(1)
        self.step_idx.zero_()
        self.alive.zero_()
        self.eos.fill_(eos)
        self.pad.fill_(pad)
        self.h_last.copy_(h_last)

    def run(self, steps):
        if self.graph is None:
            # warm-up on a side stream (allocator / lazy init), restoring the device state afterwards, then capture
            snap = [t.clone() for t in self._state()]
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                self._body()
            torch.cuda.current_stream().wait_stream(s)
            for t, v in zip(self._state(), snap):
                t.copy_(v)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._body()
            self.graph = g
            for t, v in zip(self._state(), snap):
                t.copy_(v)
        for _ in range(steps):
            self.graph.replay()

    def _state(self):
        return [self.unfinished, self.out, self.alive, self.step_idx, self.cache.len_idx, self.rows.pos,
                self.h_last, self.cache.k, self.cache.v]

    def finish(self):
        self.cache.graph_mode = False


def _get_step(llm, B, max_len, max_new, device):
    cache = getattr(llm, "_graphed_steps", None)
    if cache is None:
        cache = llm._graphed_steps = {}
    key = (B, max_len, max_new)
    if key not in cache:
        if len(cache) >= 4:
            cache.clear()
        cache[key] = GraphedGreedyStep(llm, B, max_len, max_new, device)
    return cache[key]


@torch.no_grad()
def greedy_generate(llm, inputs_embeds, max_new_tokens, eos_token_id, pad_token_id, modality=None, trim=True,
                    use_graph=True):
    ops.require_cuda(inputs_embeds)
    B, S0, H = inputs_embeds.shape
    dev = inputs_embeds.device
    task = llm._task_of(modality)
    if pad_token_id is None:
        pad_token_id = eos_token_id
    # bucket the cache length so that different prefill lengths share one captured graph
    max_len = (S0 + max_new_tokens + 127) // 128 * 128
    step = _get_step(llm, B, max_len, max_new_tokens, dev)
    cache = step.cache
    cache.graph_mode = False
    cache.len = 0
    rows = PackedRows.get([(task, B, S0)], dev)
    hid = llm.model.forward_packed(pack_segments([inputs_embeds.to(torch.bfloat16)], rows), rows, cache)   # prefill
    cache.advance(S0)
    last = (torch.arange(B, device=dev, dtype=torch.int64) * S0 + (S0 - 1)).contiguous()
    h_last = ops.gather_rows(hid, last)
    step.start(task, S0, h_last, eos_token_id, pad_token_id)
    if use_graph and not os.environ.get("OMNI_DECODE_NO_GRAPH"):     # (eager steps: profiling aid)
        step.run(max_new_tokens)
    else:
        for _ in range(max_new_tokens):
            step._body()
    step.finish()
    out = step.out.t().contiguous()
    if trim:
        alive = step.alive.tolist()                # the only host sync of the whole decode
        n = next((i + 1 for i, v in enumerate(alive) if v == 0), len(alive))
        out = out[:, :n]
    return out.clone()


# ---------------------------------------------------------------------------------------------------------------
# Beam search (the evaluation default of the reference: eval_OmniAVSR.py:216-226 -> num_beams = 15, max 32 new tokens;
# HF transformers==4.43.1 `_beam_search` + `BeamSearchScorer` semantics with length_penalty 1.0, early_stopping False).
# Device side: the B*K beam rows run through the same packed single-token step as greedy decode (split-K GEMMs, the
# single-token attention kernel, lm_head GEMM); the candidate ranking (fp32 log-softmax + running beam score, top-2K
# over K*V) stays on the device; only the 2K candidates per utterance cross to the host, where the hypothesis
# bookkeeping of the scorer runs (it is inherently sequential and tiny).
# ---------------------------------------------------------------------------------------------------------------
class _Hyps:
    """Finished hypotheses of one utterance (at most K, ranked by sum_logprobs / length)."""

    def __init__(self, K):
        self.K, self.items, self.worst = K, [], 1e9

    def add(self, ids, sum_logprobs, length):
        score = sum_logprobs / length
        if len(self.items) < self.K or score > self.worst:
            self.items.append((score, ids))
            if len(self.items) > self.K:
                order = sorted((s, i) for i, (s, _) in enumerate(self.items))
                del self.items[order[0][1]]
                self.worst = order[1][0]
            else:
                self.worst = min(score, self.worst)

    def done(self, best_running, cur_len):
        return len(self.items) >= self.K and self.worst >= best_running / cur_len


@torch.no_grad()
def beam_generate(llm, inputs_embeds, max_new_tokens, num_beams, eos_token_id, pad_token_id, modality=None):
    ops.require_cuda(inputs_embeds)
    B, S0, H = inputs_embeds.shape
    K = int(num_beams)
    BK = B * K
    dev = inputs_embeds.device
    a = llm.config
    task = llm._task_of(modality)
    if pad_token_id is None:
        pad_token_id = eos_token_id
    max_len = (S0 + max_new_tokens + 127) // 128 * 128
    cache = KVCache(a, BK, max_len, dev)
    # prefill on the expanded prompt (HF expands inputs_embeds to B*K rows before the first forward)
    rows = PackedRows.get([(task, BK, S0)], dev)
    x = inputs_embeds.to(torch.bfloat16).repeat_interleave(K, dim=0)
    hid = llm.model.forward_packed(pack_segments([x], rows), rows, cache)
    cache.advance(S0)
    last = (torch.arange(BK, device=dev, dtype=torch.int64) * S0 + (S0 - 1)).contiguous()
    h_last = ops.gather_rows(hid, last)
    # single-token step state (same layout as the graphed greedy step, run eagerly: the beam permutation changes per step)
    srows = _StepRows(BK, dev, max_len)
    srows.tile_group.fill_(task)
    xpad = torch.zeros((srows.M, a.hidden_size), dtype=torch.bfloat16, device=dev)
    llm.model.rope(max_len)
    cache.graph_mode = True

    beam_scores = torch.zeros((B, K), dtype=torch.float32, device=dev)
    beam_scores[:, 1:] = -1e9
    beam_scores = beam_scores.view(-1)
    hyps = [_Hyps(K) for _ in range(B)]
    finished = [False] * B
    seqs = [[] for _ in range(BK)]          # token history per beam row (host)
    V = a.vocab_size
    cur_len = 0
    while True:
        logits = llm.logits_rows(h_last)                                        # [BK, V] bf16 (lm_head GEMM)
        logp = torch.log_softmax(logits.float(), dim=-1) + beam_scores[:, None]
        top_s, top_i = torch.topk(logp.view(B, K * V), 2 * K, dim=1, largest=True, sorted=True)
        top_s, top_i = top_s.tolist(), top_i.tolist()                           # the step's only host sync
        cur_len += 1
        new_scores, new_tokens, new_rows = [], [], []
        for b in range(B):
            if finished[b]:
                new_scores += [0.0] * K
                new_tokens += [pad_token_id] * K
                new_rows += [0] * K
                continue
            taken = 0
            for rank in range(2 * K):
                tok, row, sc = top_i[b][rank] % V, b * K + top_i[b][rank] // V, top_s[b][rank]
                if tok == eos_token_id:
                    if rank < K:
                        hyps[b].add(list(seqs[row]), sc, cur_len)
                    continue
                new_scores.append(sc)
                new_tokens.append(tok)
                new_rows.append(row)
                taken += 1
                if taken == K:
                    break
            if taken < K:
                raise ValueError(f"At most {K} tokens can be equal to `eos_token_id: {eos_token_id}`.")
            finished[b] = hyps[b].done(max(top_s[b]), cur_len)
        seqs = [seqs[r] + [t] for r, t in zip(new_rows, new_tokens)]
        beam_scores = torch.tensor(new_scores, dtype=torch.float32, device=dev)
        if all(finished) or cur_len >= max_new_tokens:
            break
        # next step: reorder the cache rows by beam, embed the chosen tokens, one packed single-token forward
        idx = torch.tensor(new_rows, dtype=torch.int64, device=dev)
        n = cache.len
        cache.k[:, :, :, :n] = cache.k[:, :, :, :n].index_select(1, idx)
        cache.v[:, :, :, :n] = cache.v[:, :, :, :n].index_select(1, idx)
        tok = torch.tensor(new_tokens, dtype=torch.int64, device=dev)
        xpad[:BK].copy_(ops.gather_rows(llm.model.embed_tokens.weight.data, tok))
        cache.sync_device_state()
        srows.pos.fill_(cache.len)
        hid = llm.model.forward_packed(xpad, srows, cache)
        cache.advance(1)
        h_last = hid[:BK].contiguous()
    cache.graph_mode = False
    # finalize: open beams become hypotheses, best one per utterance, EOS appended if it fits, right-padded
    final = beam_scores.tolist()
    for b in range(B):
        if not finished[b]:
            for k in range(K):
                hyps[b].add(list(seqs[b * K + k]), final[b * K + k], cur_len)
    best = [sorted(h.items, key=lambda t: t[0])[-1][1] for h in hyps]
    lengths = [len(h) for h in best]
    sent_max = min(max(lengths) + 1, max_new_tokens)
    out = torch.full((B, sent_max), pad_token_id, dtype=torch.int64)
    for b, h in enumerate(best):
        out[b, : lengths[b]] = torch.tensor(h, dtype=torch.int64)
        if lengths[b] < sent_max:
            out[b, lengths[b]] = eos_token_id
    return out.to(dev)
