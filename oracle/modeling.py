"""ORACLE (test infrastructure, not product code) -- CPU restatement of AVSR_LLMs
(Omni_AVSR/modeling_OmniAVSR.py:27-606) and of the step functions of Omni_AVSR/lightning_OmniAVSR.py:159-209,
composed from oracle/{matryoshka,llm_lora,encoders}.py.  Same op sequence as the reference, including the host
log-mel and the three separate LLM passes; encoders in eval mode, rates passed explicitly (the deterministic entry
the reference offers through validation_step, lightning_OmniAVSR.py:180).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
Parity status: the parts are PINNED against outputs of the reference's own sources (tests/test_reference_golden.py:
splice / labels / compression bit-exact, the three task losses incl. matry_weights, greedy ids of the inference
branch) and against transformers / resnet.py (tests/test_oracle_*.py).  The full composition with the real Whisper /
AV-HuBERT encoders cannot be run from the reference here (no fairseq, no checkpoints): that seam is unpinned.
"""
from __future__ import annotations

import torch
from torch import nn

from . import encoders as oe
from . import llm_lora as ol
from . import matryoshka as om

TASKS = ("audio", "video", "audiovisual")


class AVSR_LLMs(nn.Module):
    def __init__(self, llm_cfg: ol.LLMConfig, lora_cfg, whisper_cfg: oe.WhisperCfg, avh_cfg: oe.AVHubertCfg,
                 intermediate_size, rates_audio, rates_video, compression_mode, prompts_ids, marker_ids, is_qwen,
                 matry_weights=None, is_task_specific=True, resnet_widths=(64, 128, 256, 512), modality="audiovisual",
                 max_dec_tokens=32, eos_id=None, pad_id=None, single_projector=False, projector_layernorm=False):
        super().__init__()
        self.single_projector = single_projector
        self.compression_mode, self.is_qwen = compression_mode, is_qwen
        self.rates_audio, self.rates_video = list(rates_audio), list(rates_video)
        self.matry_weights, self.is_task_specific = matry_weights, is_task_specific
        self.marker_ids, self.modality = marker_ids, modality
        self.max_dec_tokens, self.eos_id, self.pad_id = max_dec_tokens, eos_id, pad_id
        self.audio_encoder = oe.WhisperEncoder(whisper_cfg)
        self.video_encoder = oe.AVHubertVideo(avh_cfg, resnet_widths)
        H = llm_cfg.hidden_size
        stack = compression_mode == "stack"
        if single_projector:
            # avg-pooling Matryoshka with ONE projector shared by every rate (:94-97, :101-102 audio; :178-186 video);
            # nn.LayerNorm(hidden) at the end unless remove_layernorm_from_projector
            assert not stack
            self.audio_proj = om.make_projector(whisper_cfg.d_model, intermediate_size, H, projector_layernorm)
            self.video_proj = om.make_projector(avh_cfg.embed_dim, intermediate_size, H, projector_layernorm)
        else:
            self.audio_proj = nn.ModuleList([om.make_projector(whisper_cfg.d_model * (r if stack else 1), intermediate_size,
                                                               H, False) for r in self.rates_audio])
            self.video_proj = nn.ModuleList([om.make_projector(avh_cfg.embed_dim * (r if stack else 1), intermediate_size,
                                                               H, False) for r in self.rates_video])
        self.llm = ol.ForCausalLM_lora(llm_cfg, lora_cfg)
        self.prompts_ids = prompts_ids          # dict task -> LongTensor [1, P]

    def prompts(self):
        e = self.llm.model.embed_tokens
        return {k: e(v.to(e.weight.device)) for k, v in self.prompts_ids.items()}

    def encode_audio(self, audio, max_len, rate):
        audios = audio.to(torch.float32).cpu()                                      # :531-532 (.cpu().numpy())
        feats = oe.log_mel(audios.squeeze(-1))                                      # :533 (host feature extractor)
        dev = next(self.audio_encoder.parameters()).device                          # :534 (.cuda() in the reference)
        enc = self.audio_encoder(feats.to(device=dev, dtype=audio.dtype))
        enc = enc[:, 0: om.num_audio_tokens(max_len), :]                            # :537
        return om.compress(enc, rate, self.compression_mode)

    def encode_video(self, videos, rate):
        src = torch.reshape(videos, (-1, videos.shape[2], videos.shape[1], videos.shape[3], videos.shape[-1]))  # :463
        enc = self.video_encoder(src)
        return om.compress(enc, rate, self.compression_mode)

    def media_tokens(self, inputs, rate_a, rate_v, need_a=True, need_v=True):
        a = v = None
        if need_a:
            pa = self.audio_proj if self.single_projector else self.audio_proj[self.rates_audio.index(rate_a)]   # :366
            a = pa(self.encode_audio(inputs["audio"], max(inputs["lengths"]), rate_a))
        if need_v:
            pv = self.video_proj if self.single_projector else self.video_proj[self.rates_video.index(rate_v)]   # :353
            v = pv(self.encode_video(inputs["video"], rate_v))
        return a, v

    def forward(self, inputs, rate_a, rate_v):
        """Train/val branch (:263-306) with explicit rates -> (audio_loss, video_loss, audiovisual_loss)."""
        a, v = self.media_tokens(inputs, rate_a, rate_v)
        seqs, labs = om.build_train_sequences(self.llm.model.embed_tokens, inputs["tokens"], inputs["labels"], a, v,
                                              self.prompts(), self.marker_ids, self.is_qwen)
        losses = []
        for i, t in enumerate(TASKS):
            out = self.llm(inputs_embeds=seqs[t], labels=labs[t], modality=t if self.is_task_specific else None)
            losses.append(out.loss * self.matry_weights[i] if self.matry_weights else out.loss)
        return tuple(losses)

    @torch.no_grad()
    def decode(self, inputs, task, rate_a=None, rate_v=None, return_margins=False, num_beams=1):
        """Inference branch (:308-323): greedy, or HF beam search with num_beams > 1 (eval default 15)."""
        a, v = self.media_tokens(inputs, rate_a, rate_v, task in ("audio", "audiovisual"), task in ("video", "audiovisual"))
        emb = om.build_infer_sequence(self.llm.model.embed_tokens, inputs["tokens"], a, v, self.prompts()[task],
                                      self.marker_ids, self.is_qwen)
        return self.llm.generate(emb, self.max_dec_tokens, self.eos_id, self.pad_id,
                                 modality=task if self.is_task_specific else None, return_margins=return_margins,
                                 num_beams=num_beams)


def training_step(model: AVSR_LLMs, batch, rate_a, rate_v, world_size=1, total_batch=None):
    """lightning_OmniAVSR.py:159-176."""
    la, lv, lav = model(batch, rate_a, rate_v)
    loss = (la + lv + lav) / 3
    B = batch["tokens"].shape[0]
    total = total_batch if total_batch is not None else B * world_size
    return loss * (world_size / total), (la, lv, lav)
