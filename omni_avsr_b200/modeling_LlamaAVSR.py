"""Drop-in mirror of the reference's Omni_AVSR/modeling_LlamaAVSR.py (class AVSR_LLMs: Llama-AVSR, and Llama-MTSK when
`is_matryoshka`) on the same sm_100a kernels as the Omni-AVSR path.

Reference (file:line in /root/reference/Omni_AVSR/modeling_LlamaAVSR.py): __init__ :28-210, _unfreeze_PETF :212-236,
forward :238-270, prepare_inputs :272-468, encode_video :470-533, encode_audio :535-610.

What differs from Omni-AVSR: ONE modality and ONE prompt per model, a shared (not task-specific) LoRA, and in Matryoshka
mode EVERY rate (audio or video), or every (video rate, audio rate) pair, is trained in the same step: the reference calls
the LLM once per sequence (:241-246) and averages the losses.  Here the sequences of a step are packed into one row
buffer (256-row aligned segments) and go through the LLM in ONE pass; the splice / label kernel is launched once per
sequence and the encoder output is compressed once per rate (the encoders run once, as in the reference).

Reference quirks kept: the Matryoshka layouts assume a BOS token (:297-392 index `text_embeddings[:, 0]`), so Matryoshka +
Qwen is rejected; `test_ratio_matry` is `[video_rate, audio_rate]` for the audiovisual modality (:313-318, :489, :563) and a
scalar otherwise; stack-mode Matryoshka needs `remove_layernorm_from_projector=True` (the LayerNorm variant multiplies a
list by an int, :72); the prompt is embedded at call time from the tokenizer (:278-279), it is not a buffer.
"""
from __future__ import annotations

from typing import Optional

import torch
from torch import nn

from . import ops
from .Llama_LoRA import PackedRows, TASKS, pack_segments
from .modeling_OmniAVSR import AVSR_LLMs as _OmniAVSR
from .modeling_OmniAVSR import Projector, SpliceFn, compress

IGNORE_INDEX = -100


class AVSR_LLMs(_OmniAVSR):
    def __init__(self, modality, pretrain_avhubert_enc_video, use_lora_avhubert, llm_model, hidden_size,
                 intermediate_size, tokenizer, prompt, pad_id, downsample_ratio_audio, downsample_ratio_video,
                 audio_encoder_name, compression_mode, unfrozen_modules, max_dec_tokens, num_beams, PETF_LLM_name=None,
                 peft_config_llm=None, remove_layernorm_from_projector=False, is_matryoshka=False,
                 device="cuda", llm_overrides: Optional[dict] = None, audio_arch=None, video_arch=None):
        if modality not in TASKS:
            raise ValueError(modality)
        if is_matryoshka and "Qwen" in llm_model:
            raise NotImplementedError("the reference's Matryoshka layouts index the BOS embedding (:297-392): Llama only")
        if is_matryoshka and compression_mode == "stack" and not remove_layernorm_from_projector:
            raise TypeError("can't multiply sequence by non-int: the reference's stack + LayerNorm Matryoshka projector "
                            "(:72) cannot be constructed; pass remove_layernorm_from_projector=True")
        if getattr(peft_config_llm, "IS_TASK_SPECIFIC", False):
            raise ValueError("Llama-AVSR / Llama-MTSK use a shared LoRA (lightning_LlamaAVSR.py:104-113)")
        super().__init__(modality, pretrain_avhubert_enc_video, use_lora_avhubert, llm_model, hidden_size,
                         intermediate_size, tokenizer, prompt, prompt, prompt, pad_id, downsample_ratio_audio,
                         downsample_ratio_video, audio_encoder_name, compression_mode, unfrozen_modules, max_dec_tokens,
                         num_beams, PETF_LLM_name=PETF_LLM_name, peft_config_llm=peft_config_llm,
                         remove_layernorm_from_projector=remove_layernorm_from_projector, matry_weights=None,
                         is_task_specific=False, is_matryoshka=is_matryoshka, is_single_matry_projector=False,
                         device=device, llm_overrides=llm_overrides, audio_arch=audio_arch, video_arch=video_arch)
        self.prompt = prompt
        # the reference keeps no prompt buffers in this model (state-dict parity): embed the prompt per call
        for name in ("prompt_audio", "prompt_video", "prompt_audiovisual"):
            delattr(self, name)
        start = 0 if "Qwen" in llm_model else 1
        self._prompt_ids = self.tokenizer(self.prompt, return_tensors="pt").input_ids[:, start:-1].to(device)  # :278

    # projector construction rules of THIS model (:61-104 audio, :144-188 video): LayerNorm iff not remove_layernorm
    def _make_projectors(self, name, dim, rates, inter, hidden, is_audio):
        stack = self.compression_mode == "stack"
        ln = not self.remove_layernorm_from_projector
        if self.is_matryoshka:
            mm = {el: i for i, el in enumerate(rates)}
            proj = nn.ModuleList([Projector(dim * (r if stack else 1), inter, hidden, ln, self.flat, f"{name}.{i}")
                                  for i, r in enumerate(rates)])
            return proj, mm
        r = rates[0]
        return Projector(dim * r if stack else dim, inter, hidden, ln, self.flat, name), None

    # ------------------------------------------------------------------------------------------------
    def _unfreeze_PETF(self, unfrozen_modules):
        """modeling_LlamaAVSR.py:212-236: unlike Omni-AVSR, the AV-HuBERT adapters are unfrozen only when the model's
        modality is "video"; in the audiovisual Llama-AVSR / MTSK recipes they stay frozen (down = 0: an identity)."""
        super()._unfreeze_PETF(unfrozen_modules)
        if hasattr(self, "video_encoder") and self.modality != "video":
            for p in self.video_encoder.lora_parameters():
                p.requires_grad_(False)

    def _prompt_embeddings(self):
        return self.llm.model.embed_tokens(self._prompt_ids)[0].detach()          # [P, H]  (:279)

    def _splice_one(self, inputs, audio_tok, video_tok, with_labels):
        """One sequence [bos, <audio> a </audio>, <video> v </video>, prompt, text[1:]] (+ labels) as a dense block."""
        t = TASKS.index(self.modality)
        p = self._prompt_embeddings()
        tokens = inputs["tokens"]
        labels = inputs.get("labels") if with_labels else None
        layout = ops.SpliceLayout(tokens=tokens.contiguous(), labels=None if labels is None else labels.contiguous(),
                                  embed=self.llm.model.embed_tokens.weight.data, audio_tok=audio_tok, video_tok=video_tok,
                                  prompts=[p, p, p], marker_ids=self._marker_ids, has_bos=self._has_bos, task_mask=1 << t)
        B, S = tokens.shape[0], layout.seq_len[t]
        rows = PackedRows.get([(t, B, S)], tokens.device)
        out = SpliceFn.apply(audio_tok, video_tok, layout, rows, labels is not None)
        seq = out[0][: B * S].view(B, S, self.hidden_size)
        return seq, (out[1 + t] if labels is not None else None)

    def forward(self, inputs, is_trainval=True, test_ratio_matry=None):
        embeddings, labels = self.prepare_inputs(inputs, is_trainval, test_ratio_matry=test_ratio_matry)
        if is_trainval:
            seqs = embeddings if self.is_matryoshka else [embeddings]
            labs = labels if self.is_matryoshka else [labels]
            B = seqs[0].shape[0]
            rows = PackedRows.get([(0, B, s.shape[1]) for s in seqs], seqs[0].device)
            xp = pack_segments(seqs, rows)
            hid = self.llm.model.forward_packed(xp, rows)
            segs = [(b, s, off) for (_, b, s, off) in rows.segments]
            losses = self.llm.loss_from_hidden(hid, segs, labs, [1.0] * len(seqs))
            total = losses[0]
            for l in losses[1:]:
                total = total + l                                                    # :242-246
            return total / len(seqs)
        vocab = self.tokenizer.vocab
        trim = not getattr(self, "decode_no_trim", False)
        if "Qwen" in self.llm_model:                                                 # :264-268
            return self.llm.generate(inputs_embeds=embeddings, max_new_tokens=self.max_dec_tokens, num_beams=self.num_beams,
                                     eos_token_id=vocab["<|endoftext|>"], pad_token_id=vocab["<|endoftext|>"], trim=trim)
        return self.llm.generate(inputs_embeds=embeddings, max_new_tokens=self.max_dec_tokens, num_beams=self.num_beams,
                                 eos_token_id=vocab["<|end_of_text|>"], bos_token_id=vocab["<|begin_of_text|>"],
                                 pad_token_id=vocab["<pad>"], trim=trim)             # :253-258

    def prepare_inputs(self, inputs, is_trainval, test_ratio_matry=None):
        use_a = self.modality in ("audio", "audiovisual")
        use_v = self.modality in ("video", "audiovisual")
        af = self.encode_audio(inputs["audio"], max(inputs["lengths"]), is_trainval, test_ratio_matry=test_ratio_matry) \
            if use_a else None                                                        # :274
        vf = self.encode_video(inputs["video"], is_trainval, test_ratio_matry=test_ratio_matry) if use_v else None
        if self.is_matryoshka and is_trainval:
            # every rate (pair) of the step: video rates outer, audio rates inner (:312-324); one rate list otherwise
            a_tok = [self.audio_proj[i](f).contiguous() for i, f in enumerate(af)] if use_a else [None]
            v_tok = [self.video_proj[i](f).contiguous() for i, f in enumerate(vf)] if use_v else [None]
            seqs, labs = [], []
            for v in v_tok:
                for a in a_tok:
                    s, l = self._splice_one(inputs, a, v, inputs.get("labels") is not None)
                    seqs.append(s)
                    labs.append(l)
            return seqs, (labs if inputs.get("labels") is not None else None)
        a_tok = v_tok = None
        if self.is_matryoshka:                                                        # inference at one rate (pair)
            ra = test_ratio_matry[1] if self.modality == "audiovisual" else test_ratio_matry
            rv = test_ratio_matry[0] if self.modality == "audiovisual" else test_ratio_matry
            if use_a:
                a_tok = self.audio_proj[self.matry_map_audio[ra]](af).contiguous()    # KeyError for an unknown rate
            if use_v:
                v_tok = self.video_proj[self.matry_map_video[rv]](vf).contiguous()
        else:
            if use_a:
                a_tok = self.audio_proj(af).contiguous()
            if use_v:
                v_tok = self.video_proj(vf).contiguous()
        if is_trainval:
            return self._splice_one(inputs, a_tok, v_tok, inputs.get("labels") is not None)
        tokens = inputs["tokens"]
        one = {"tokens": (tokens[:, :1] if self._has_bos else tokens[:, :0]).contiguous()}   # only e(BOS) is used (:290)
        seq, _ = self._splice_one(one, a_tok, v_tok, False)
        return seq.contiguous(), None

    # ------------------------------------------------------------------------------------------------
    def _rates(self, rates, is_trainval, test_ratio, index):
        if is_trainval:
            return list(rates)
        return [test_ratio[index] if self.modality == "audiovisual" else test_ratio]

    def encode_video(self, videos, is_trainval, test_ratio_matry=None):
        src = torch.reshape(videos, (-1, videos.shape[2], videos.shape[1], videos.shape[3], videos.shape[-1]))  # :471
        video_enc, _, _ = self.video_encoder.extract_finetune(source={"video": src, "audio": None})
        n_tok = video_enc.shape[1]
        if self.is_matryoshka:
            rates = self._rates(self.downsample_ratio_video, is_trainval, test_ratio_matry, 0)
            outs = []
            for r in rates:
                if not is_trainval and self.compression_mode == "avg-pooling" and r not in self.matry_map_video:
                    raise KeyError(r)                                                # :518-521
                outs.append(video_enc if (r == 1 and self.compression_mode == "stack" and not is_trainval)
                            else compress(video_enc, n_tok, r, self.compression_mode))
            return outs if is_trainval else outs[0]
        if self.downsample_ratio_video != 1:
            return compress(video_enc, n_tok, self.downsample_ratio_video, self.compression_mode)
        return video_enc

    def encode_audio(self, audio, max_len, is_trainval, test_ratio_matry=None):
        feats = self.audio_frontend(audio.squeeze(-1))                               # :536-538 on the GPU
        audio_enc = self.audio_encoder(feats).last_hidden_state                      # :539
        ml = max_len if torch.is_tensor(max_len) else torch.tensor(max_len)
        n_tok = min(max(int(ml.detach().cpu().to(torch.int64) / 16000 * 50), 25), audio_enc.shape[1])   # :540
        if self.is_matryoshka:
            rates = self._rates(self.downsample_ratio_audio, is_trainval, test_ratio_matry, 1)
            outs = []
            for r in rates:
                if not is_trainval and self.compression_mode == "avg-pooling" and r not in self.matry_map_audio:
                    raise KeyError(r)
                outs.append(audio_enc[:, :n_tok].contiguous() if (r == 1 and self.compression_mode == "stack" and not is_trainval)
                            else compress(audio_enc, n_tok, r, self.compression_mode))
            return outs if is_trainval else outs[0]
        if self.downsample_ratio_audio != 1:
            return compress(audio_enc, n_tok, self.downsample_ratio_audio, self.compression_mode)
        return audio_enc[:, :n_tok].contiguous()
