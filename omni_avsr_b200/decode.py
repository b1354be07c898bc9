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
# Everything of a step runs on the device inside ONE CUDA graph: lm_head GEMM -> omni_beam_topk_rows (fp32 log-softmax +
# running beam score, the 2K best tokens of every beam row) -> omni_beam_select (merge per utterance, the scorer's walk over
# the 2K best candidates, finished-hypothesis heap, `done`, token history, embedding rows of the chosen tokens, KV-cache
# indirection) -> the packed single-token forward of the B*K rows -> counter advance.  Differences from HF's execution
# that do not change the result: the prompt is prefilled once per utterance (HF expands it to B*K identical rows), the
# KV cache is never gathered by beam index (the attention kernel follows an indirection table instead), and the host
# looks at the number of finished utterances every 8 steps only.
# ---------------------------------------------------------------------------------------------------------------
class _Hyps:
    """Finished hypotheses of one utterance (at most K, ranked by sum_logprobs / length); host mirror of the device heap
    (csrc/beam_search.cu: hyp_add), used to close the open beams after the last step."""

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


BEAM_MAX_NEW = 128          # DM_MAX_NEW of csrc/decode_attention.cu (indirection row held in shared memory)
BEAM_MAX_K = 32             # BS_MAX_K of csrc/beam_search.cu


class BeamState:
    """Device-side state of one beam search (the fields of omni_beam_select_args, include/omni_avsr.h)."""

    def __init__(self, B, K, V, max_new, device):
        BK = B * K
        self.B, self.K, self.V, self.max_new = B, K, V, max_new

        def z(shape, dtype):
            return torch.zeros(shape, dtype=dtype, device=device)
        self.cand_score = z((BK, 2 * K), torch.float32)
        self.cand_tok = z((BK, 2 * K), torch.int32)
        self.beam_scores = z(BK, torch.float32)
        self.step_idx = z(1, torch.int64)
        self.eos, self.pad, self.prefill_len = z(1, torch.int64), z(1, torch.int64), z(1, torch.int64)
        self.seqs = z((2, BK, max_new), torch.int32)
        self.ind = z((2, BK, max_new), torch.int32)
        self.hyp_seq = z((B, K + 1, max_new), torch.int32)
        self.hyp_len = z((B, K + 1), torch.int32)
        self.hyp_score = z((B, K + 1), torch.float64)
        self.hyp_order = z((B, K + 1), torch.int32)
        self.hyp_count = z(B, torch.int32)
        self.hyp_worst = z(B, torch.float64)
        self.done, self.n_done, self.status = z(B, torch.int32), z(1, torch.int32), z(1, torch.int32)
        self._order0 = torch.arange(K + 1, dtype=torch.int32, device=device).repeat(B, 1).contiguous()
        first = torch.full((B, K), -1e9, dtype=torch.float32, device=device)
        first[:, 0] = 0.0                    # HF: only the first beam of an utterance is live before the first step
        self._scores0 = first.view(-1).contiguous()

    def reset(self, eos, pad, prefill_len):
        self.beam_scores.copy_(self._scores0)
        self.hyp_order.copy_(self._order0)
        self.hyp_worst.fill_(1e9)
        for t in (self.step_idx, self.seqs, self.ind, self.hyp_count, self.done, self.n_done, self.status):
            t.zero_()
        self.eos.fill_(eos)
        self.pad.fill_(pad)
        self.prefill_len.fill_(prefill_len)

    def tensors(self):
        return [self.beam_scores, self.step_idx, self.seqs, self.ind, self.hyp_seq, self.hyp_len, self.hyp_score,
                self.hyp_order, self.hyp_count, self.hyp_worst, self.done, self.n_done, self.status]


class GraphedBeamStep:
    """One beam-search step (ranking + scorer + forward of the B*K rows) captured in a CUDA graph."""

    def __init__(self, llm, B, K, max_len, max_new, device):
        a = llm.config
        self.llm, self.B, self.K, self.max_len, self.max_new = llm, B, K, max_len, max_new
        BK = B * K
        self.cache = KVCache(a, BK, max_len, device)
        self.cache.row_stride = K                      # the prefill of utterance u lands in cache row u * K
        self.state = BeamState(B, K, a.vocab_size, max_new, device)
        self.cache.beam = (self.state.ind, self.state.prefill_len, K)
        self.rows = _StepRows(BK, device, max_len)
        self.xpad = torch.zeros((BK, a.hidden_size), dtype=torch.bfloat16, device=device)
        self.h_last = torch.zeros((BK, a.hidden_size), dtype=torch.bfloat16, device=device)
        self.graph = None
        llm.model.rope(max_len)

    def _body(self):
        llm, st = self.llm, self.state
        logits = llm.logits_rows(self.h_last)                                    # [B*K, V] bf16 (lm_head GEMM)
        ops.beam_topk_rows(logits, llm.config.vocab_size, st.beam_scores, st.cand_score, st.cand_tok)
        ops.beam_select(st, llm.model.embed_tokens.weight.data, self.xpad)
        hid = llm.model.forward_packed(self.xpad, self.rows, self.cache)
        self.h_last.copy_(hid[: self.B * self.K])
        ops.decode_advance(st.step_idx, self.cache.len_idx, self.rows.pos)

    def start(self, task, prefill_len, h_last, eos, pad):
        self.rows.tile_group.fill_(task)
        self.rows.pos.fill_(prefill_len)
        self.cache.len = prefill_len
        self.cache.sync_device_state()
        self.cache.graph_mode = True
        self.state.reset(eos, pad, prefill_len)
        self.h_last.copy_(h_last)

    def _mutable(self):
        return self.state.tensors() + [self.cache.len_idx, self.rows.pos, self.h_last, self.xpad]

    def run(self, steps, use_graph=True):
        """Runs up to `steps` ranking steps; returns how many were executed (== steps unless every utterance finished)."""
        if use_graph and self.graph is None:
            snap = [t.clone() for t in self._mutable()]
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                self._body()                           # allocator / lazy-init warm-up outside the capture
            torch.cuda.current_stream().wait_stream(s)
            for t, v in zip(self._mutable(), snap):
                t.copy_(v)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._body()
            self.graph = g
            for t, v in zip(self._mutable(), snap):
                t.copy_(v)
        n = 0
        while n < steps:
            if use_graph:
                self.graph.replay()
            else:
                self._body()
            n += 1
            if n % 8 == 0 and n < steps and int(self.state.n_done.item()) == self.B:
                break                                  # HF leaves the loop as soon as every utterance is done
        return n

    def finish(self):
        self.cache.graph_mode = False


def _get_beam_step(llm, B, K, max_len, max_new, device):
    cache = getattr(llm, "_graphed_beam_steps", None)
    if cache is None:
        cache = llm._graphed_beam_steps = {}
    key = (B, K, max_len, max_new)
    if key not in cache:
        if len(cache) >= 2:
            cache.clear()
        cache[key] = GraphedBeamStep(llm, B, K, max_len, max_new, device)
    return cache[key]


@torch.no_grad()
def beam_generate(llm, inputs_embeds, max_new_tokens, num_beams, eos_token_id, pad_token_id, modality=None, use_graph=True):
    ops.require_cuda(inputs_embeds)
    B, S0, H = inputs_embeds.shape
    K = int(num_beams)
    if K > BEAM_MAX_K or max_new_tokens > BEAM_MAX_NEW:
        raise NotImplementedError(f"beam search kernels: num_beams <= {BEAM_MAX_K}, max_new_tokens <= {BEAM_MAX_NEW}")
    dev = inputs_embeds.device
    task = llm._task_of(modality)
    if eos_token_id is None:
        raise ValueError("beam search needs an eos_token_id (HF BeamSearchScorer closes hypotheses on it)")
    if pad_token_id is None:
        pad_token_id = eos_token_id
    max_len = (S0 + max_new_tokens + 127) // 128 * 128
    step = _get_beam_step(llm, B, K, max_len, max_new_tokens, dev)
    cache = step.cache
    cache.graph_mode = False
    cache.len = 0
    # prefill once per utterance (HF expands inputs_embeds to B*K identical rows before the first forward)
    rows = PackedRows.get([(task, B, S0)], dev)
    hid = llm.model.forward_packed(pack_segments([inputs_embeds.to(torch.bfloat16)], rows), rows, cache)
    cache.advance(S0)
    last = (torch.arange(B, device=dev, dtype=torch.int64) * S0 + (S0 - 1)).contiguous()
    h_last = ops.gather_rows(hid, last).repeat_interleave(K, dim=0)
    step.start(task, S0, h_last, eos_token_id, pad_token_id)
    cur_len = step.run(max_new_tokens, use_graph=use_graph and not os.environ.get("OMNI_DECODE_NO_GRAPH"))
    step.finish()

    # finalize on the host (one read-back per decode): open beams become hypotheses, best one per utterance, EOS appended if
    # it fits, right-padded
    st = step.state
    if int(st.status.item()) != 0:
        raise ValueError(f"At most {K} tokens can be equal to `eos_token_id: {eos_token_id}`.")
    hyp_seq, hyp_len, hyp_score = st.hyp_seq.tolist(), st.hyp_len.tolist(), st.hyp_score.tolist()
    hyp_order, hyp_count, hyp_worst = st.hyp_order.tolist(), st.hyp_count.tolist(), st.hyp_worst.tolist()
    done = st.done.tolist()
    seqs = st.seqs[cur_len & 1].tolist()               # step s writes buffer (s + 1) & 1
    final = st.beam_scores.tolist()
    best = []
    for b in range(B):
        h = _Hyps(K)
        h.worst = hyp_worst[b]
        for i in range(hyp_count[b]):
            slot = hyp_order[b][i]
            h.items.append((hyp_score[b][slot], hyp_seq[b][slot][: hyp_len[b][slot]]))
        if not done[b]:
            for k in range(K):
                h.add(seqs[b * K + k][:cur_len], final[b * K + k], cur_len)
        best.append(sorted(h.items, key=lambda t: t[0])[-1][1])
    lengths = [len(h) for h in best]
    sent_max = min(max(lengths) + 1, max_new_tokens)
    out = torch.full((B, sent_max), pad_token_id, dtype=torch.int64)
    for b, h in enumerate(best):
        out[b, : lengths[b]] = torch.tensor(h, dtype=torch.int64)
        if lengths[b] < sent_max:
            out[b, lengths[b]] = eos_token_id
    return out.to(dev)
