"""ORACLE (test infrastructure, not product code) -- CPU restatement of the Matryoshka compression,
projector and prompt-splice arithmetic of the reference, op for op.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import this.

Parity status: PINNED against outputs of the reference itself: tests/golden/reference_golden.pt holds what the
unmodified AVSR_LLMs.encode_audio / encode_video / prepare_inputs / forward of /root/reference produced on seeded
inputs (generator: tests/golden/make_reference_golden.py) and tests/test_reference_golden.py requires this file to
reproduce every byte of the three LLM input sequences, every label and the compressed features (avg-pooling and
stack, Llama and Qwen layouts, train and infer branches).  The older oracle-generated fixture
(tests/golden/omni_golden.pt) and the structural anchors of SURVEY.md §8c are kept as regression checks.

Reference lines (relative to /root/reference):
  token-count rule / truncation ... Omni_AVSR/modeling_OmniAVSR.py:537
  avg-pool compression ............. :544-546 (audio), :469-471 (video)
  stack compression ................ :562-568 (audio), :487-493 (video)
  projector modules ................ :65-111 (audio), :152-196 (video)
  splice (train) ................... :337-395 + :270-299
  splice (infer) ................... :406-458
"""
from __future__ import annotations

import torch
from torch import nn

IGNORE_INDEX = -100


def num_audio_tokens(max_len) -> int:
    """modeling_OmniAVSR.py:537 -- `max(int(max_len/16000*50), 25)`; max_len is a 0-d int64 tensor in the
    reference (max(inputs["lengths"])), so the arithmetic is float32 tensor arithmetic."""
    if not torch.is_tensor(max_len):
        max_len = torch.tensor(max_len, dtype=torch.int64)
    return max(int(max_len / 16000 * 50), 25)


def compress(enc: torch.Tensor, rate: int, mode: str) -> torch.Tensor:
    """enc [B, n_tok, D] (already truncated) -> compressed features."""
    if mode == "avg-pooling":
        x = enc.transpose(1, 2).contiguous()          # :544
        x = nn.AvgPool1d(rate)(x)                      # :545
        return x.transpose(1, 2).contiguous()          # :546
    if mode == "stack":
        temp = enc                                     # :562
        chunks = [temp[:, x:x + rate, :].reshape(temp.shape[0], 1, -1) for x in range(0, temp.shape[1], rate)]  # :563
        rest = temp.shape[1] % rate                    # :564
        if rest == 0:
            if not chunks:
                return temp.new_zeros((temp.shape[0], 0, temp.shape[2] * rate))
            return torch.stack(chunks, dim=1).squeeze(2)   # :566
        if len(chunks) <= 1:
            return temp.new_zeros((temp.shape[0], 0, temp.shape[2] * rate))
        return torch.stack(chunks[:-1], dim=1).squeeze(2)  # :568
    raise ValueError(mode)


def make_projector(in_dim: int, intermediate: int, hidden: int, layernorm: bool) -> nn.Sequential:
    """Linear+ReLU+Linear(+LayerNorm). In Matryoshka multi-projector avg-pool mode the reference passes the
    LayerNorm as the `bias` argument of nn.Linear (:104,:188), i.e. there is NO LayerNorm and the bias is on."""
    layers = [nn.Linear(in_dim, intermediate), nn.ReLU(), nn.Linear(intermediate, hidden)]
    if layernorm:
        layers.append(nn.LayerNorm(hidden))
    return nn.Sequential(*layers)


def media_block(embed: nn.Embedding, feats: torch.Tensor, sos_id: int, eos_id: int) -> torch.Tensor:
    """[e(<m>), feats, e(</m>)]  (:347-355 video, :360-368 audio)."""
    B = feats.shape[0]
    starts = embed(torch.tensor([sos_id], device=feats.device).expand(B, -1))
    ends = embed(torch.tensor([eos_id], device=feats.device).expand(B, -1))
    return torch.cat([starts, feats, ends], dim=1)


def build_train_sequences(embed: nn.Embedding, tokens, labels, audio_tok, video_tok, prompts, marker_ids, is_qwen):
    """Returns ({task: inputs_embeds}, {task: labels}) exactly as prepare_inputs(:337-395) + forward(:270-299).

    audio_tok / video_tok are the *projected* features [B, n, H]; prompts = dict task -> [1, P, H] buffer.
    """
    id_as, id_ae, id_vs, id_ve = marker_ids
    text = embed(tokens)                                                   # :337
    B = tokens.shape[0]
    video_inputs = media_block(embed, video_tok, id_vs, id_ve)             # :347-355
    audio_inputs = media_block(embed, audio_tok, id_as, id_ae)             # :360-368
    ign = {
        "audio": prompts["audio"].shape[1] + audio_inputs.shape[1],
        "video": prompts["video"].shape[1] + video_inputs.shape[1],
        "audiovisual": prompts["audiovisual"].shape[1] + audio_inputs.shape[1] + video_inputs.shape[1],
    }
    lab = {}
    for k, n in ign.items():
        pre = torch.tensor([IGNORE_INDEX] * n, device=text.device).expand(B, -1)      # :373-375
        if is_qwen:
            lab[k] = torch.cat([pre, labels], dim=1)                                   # :378-380
        else:
            lab[k] = torch.cat([labels[:, 0].unsqueeze(1), pre, labels[:, 1:]], dim=1)  # :382-387
    media = {"audio": [audio_inputs], "video": [video_inputs], "audiovisual": [audio_inputs, video_inputs]}
    seqs = {}
    for k in ("audio", "video", "audiovisual"):
        p = prompts[k].expand(B, -1, -1)
        if is_qwen:
            seqs[k] = torch.cat([*media[k], p, text], dim=1)                            # :270-275
        else:
            seqs[k] = torch.cat([text[:, 0, :].unsqueeze(1), *media[k], p, text[:, 1:, :]], dim=1)  # :278-283
    return seqs, lab


def build_infer_sequence(embed: nn.Embedding, tokens, audio_tok, video_tok, prompt, marker_ids, is_qwen):
    """prepare_inputs infer branch (:406-458): [bos, prompt] -> insert video after bos -> insert audio after bos."""
    id_as, id_ae, id_vs, id_ve = marker_ids
    B = tokens.shape[0]
    text_ = embed(tokens)                                                  # :406
    prompt = prompt.expand(B, -1, -1)
    if is_qwen:
        te = prompt                                                        # :417
    else:
        te = torch.cat([text_[:, 0, :].unsqueeze(1), prompt], dim=1)       # :419
    if video_tok is not None:
        v = media_block(embed, video_tok, id_vs, id_ve)
        te = torch.cat([v, te], dim=1) if is_qwen else torch.cat([te[:, 0, :].unsqueeze(1), v, te[:, 1:, :]], dim=1)
    if audio_tok is not None:
        a = media_block(embed, audio_tok, id_as, id_ae)
        te = torch.cat([a, te], dim=1) if is_qwen else torch.cat([te[:, 0, :].unsqueeze(1), a, te[:, 1:, :]], dim=1)
    return te
