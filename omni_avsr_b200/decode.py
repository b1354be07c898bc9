"""Greedy decode with a static KV cache (the `eval_OmniAVSR.py` path: modeling_OmniAVSR.py:308-323 ->
HF GenerationMixin.generate with inputs_embeds, reproduced per SURVEY A.5):

  * only `inputs_embeds` is given, so the returned ids contain ONLY the new tokens;
  * step 0 consumes the embeddings, later steps embed the previously chosen id (Llama_LoRA.py:429-432);
  * `modality` selects the adapter at every step (Llama_LoRA.py:441);
  * greedy = argmax(fp32(logits[:, -1])); a row that emitted EOS is padded with pad_token_id afterwards;
  * generation stops when every row is finished or after max_new_tokens.

All max_new_tokens steps are enqueued without a host sync; the "all rows finished" cut is applied once at the end,
which yields the same ids as HF's per-step check.
"""
from __future__ import annotations

import torch

from . import ops
from .Llama_LoRA import KVCache, PackedRows, pack_segments


@torch.no_grad()
def greedy_generate(llm, inputs_embeds, max_new_tokens, eos_token_id, pad_token_id, modality=None, trim=True):
    ops.require_cuda(inputs_embeds)
    B, S0, H = inputs_embeds.shape
    a = llm.config
    dev = inputs_embeds.device
    task = llm._task_of(modality)
    cache = KVCache(a, B, S0 + max_new_tokens, dev)
    rows = PackedRows.get([(task, B, S0)], dev)
    hid = llm.model.forward_packed(pack_segments([inputs_embeds.to(torch.bfloat16)], rows), rows, cache)
    cache.advance(S0)
    last = (torch.arange(B, device=dev, dtype=torch.int64) * S0 + (S0 - 1)).contiguous()
    h_last = ops.gather_rows(hid, last)
    if pad_token_id is None:
        pad_token_id = eos_token_id
    unfinished = torch.ones(B, dtype=torch.int64, device=dev)
    toks, alive = [], []
    for step in range(max_new_tokens):
        logits = llm.logits_rows(h_last)
        nxt = ops.argmax_rows(logits)
        nxt = nxt * unfinished + pad_token_id * (1 - unfinished)
        toks.append(nxt)
        unfinished = unfinished * (nxt != eos_token_id).long()
        alive.append(unfinished.max())
        if step == max_new_tokens - 1:
            break
        x = ops.gather_rows(llm.model.embed_tokens.weight.data, nxt.contiguous())
        rows1 = PackedRows.get([(task, B, 1)], dev, pos_offset=cache.len)
        hid = llm.model.forward_packed(pack_segments([x.view(B, 1, H)], rows1), rows1, cache)
        cache.advance(1)
        h_last = hid[:B]
    out = torch.stack(toks, dim=1)
    if trim:
        alive = torch.stack(alive).tolist()        # the only host sync of the whole decode
        n = next((i + 1 for i, v in enumerate(alive) if v == 0), len(alive))
        out = out[:, :n]
    return out
