"""ORACLE (test infrastructure, not product code) -- CPU restatement of the Llama-AVSR / Llama-MTSK model
(Omni_AVSR/modeling_LlamaAVSR.py), SURVEY.md §8(f) rank 2.  Only tests/ may import this.

Follows the reference op for op:
  prepare_inputs ............ modeling_LlamaAVSR.py:272-468 (non-Matryoshka :418-468, Matryoshka single modality
                              :350-411, Matryoshka audiovisual :297-348)
  forward (train) ........... :238-248  (one LLM call per Matryoshka sequence, mean of the losses)
  encode_audio/encode_video . :470-610  (every rate of the list in train mode, one rate at inference)
Parity status: PINNED -- tests/golden/make_reference_golden.py executes the unmodified reference class and
tests/test_reference_golden.py holds these functions to its sequences / labels (bit-exact) and losses.
"""
from __future__ import annotations

import torch

from . import matryoshka as om

IGNORE_INDEX = -100


def compress_all(enc, rates, mode):
    """Train-mode encode_*: the list of compressed features, one per rate (:476-487 / :503-511)."""
    return [om.compress(enc, r, mode) for r in rates]


def prepare_inputs(embed, tokens, labels, audio_tok, video_tok, prompt_ids, marker_ids, is_qwen, modality, is_matryoshka,
                   is_trainval):
    """audio_tok / video_tok: PROJECTED media tokens -- in Matryoshka train mode a list per rate, else one tensor (or
    None when the modality does not use them).  Returns (embeddings, labels): lists in Matryoshka train mode."""
    id_as, id_ae, id_vs, id_ve = marker_ids
    B = tokens.shape[0]
    text_ = embed(tokens)                                                                  # :276
    prompt = embed(prompt_ids.expand(B, -1))                                               # :279
    if is_trainval:
        if is_qwen:
            text = torch.cat([prompt, text_], dim=1)                                       # :283
        else:
            text = torch.cat([torch.cat([text_[:, 0, :].unsqueeze(1), prompt], dim=1), text_[:, 1:, :]], dim=1)  # :285-287
    else:
        text = prompt if is_qwen else torch.cat([text_[:, 0, :].unsqueeze(1), prompt], dim=1)   # :289-292
    ignore = prompt.shape[1]                                                               # :294

    def block(feats, sos, eos):
        s = embed(torch.tensor([sos]).expand(B, -1))
        e = embed(torch.tensor([eos]).expand(B, -1))
        return torch.cat((s, feats, e), dim=1)

    def lab(n):
        pre = torch.tensor([IGNORE_INDEX] * n).expand(B, -1)
        if is_qwen:
            return torch.cat([pre, labels], dim=1)                                         # :458-459
        return torch.cat((labels[:, 0].unsqueeze(1), pre, labels[:, 1:]), dim=1)           # :337, :400, :461-463

    def insert(seq, media):
        if is_qwen:
            return torch.cat([media, seq], dim=1)                                          # :427, :444
        return torch.cat((seq[:, 0, :].unsqueeze(1), media, seq[:, 1:, :]), dim=1)

    if is_matryoshka and is_trainval:
        assert not is_qwen, "the reference's Matryoshka branches index the BOS embedding"
        seqs, labs = [], []
        if modality == "audiovisual":                                                      # :297-348
            for v in video_tok:
                vin = block(v, id_vs, id_ve)
                with_v = insert(text, vin)
                for a in audio_tok:
                    ain = block(a, id_as, id_ae)
                    seqs.append(insert(with_v, ain))
                    labs.append(lab(ignore + ain.shape[1] + vin.shape[1]) if labels is not None else None)
        else:                                                                              # :350-411
            feats = video_tok if modality == "video" else audio_tok
            sos, eos = (id_vs, id_ve) if modality == "video" else (id_as, id_ae)
            for f in feats:
                m = block(f, sos, eos)
                seqs.append(insert(text, m))
                labs.append(lab(ignore + m.shape[1]) if labels is not None else None)
        return seqs, (labs if labels is not None else None)
    seq = text
    if video_tok is not None:                                                              # :419-432 / :326-329
        vin = block(video_tok, id_vs, id_ve)
        seq = insert(seq, vin)
        ignore += vin.shape[1]
    if audio_tok is not None:                                                              # :436-449 / :331-334
        ain = block(audio_tok, id_as, id_ae)
        seq = insert(seq, ain)
        ignore += ain.shape[1]
    return seq, (lab(ignore) if labels is not None else None)


def train_loss(llm, embeddings, labels, is_matryoshka):
    """forward, train branch (:240-248): mean over the Matryoshka sequences of the LLM's own mean CE."""
    if not is_matryoshka:
        return llm(inputs_embeds=embeddings, labels=labels).loss
    total = 0.0
    for e, l in zip(embeddings, labels):
        total = total + llm(inputs_embeds=e, labels=l).loss
    return total / len(embeddings)
