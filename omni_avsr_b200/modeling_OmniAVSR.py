"""Drop-in mirror of the reference's Omni_AVSR/modeling_OmniAVSR.py (class AVSR_LLMs) on the sm_100a kernels.

Same constructor / forward / prepare_inputs / encode_audio / encode_video signatures and return values as the
reference (file:line in /root/reference/Omni_AVSR/modeling_OmniAVSR.py):
  __init__ :28-231, _unfreeze_PETF :234-260, forward :263-323, prepare_inputs :326-458,
  encode_video :461-526, encode_audio :528-606.

Hot-path differences (same results, different execution):
  * log-mel on the GPU (no .cpu().numpy() round trip, :531-534);
  * compression straight from the [B, T, D] encoder output, truncation folded into the kernel (:537-588);
  * projector MLP on the tcgen05 GEMM (bias+ReLU / bias epilogues);
  * ONE splice launch writes the three task sequences and their labels into a packed, 128-row-aligned buffer;
  * ONE LLM pass over the packed rows, adapter chosen per tile by task id, lm_head + CE on label rows only.
"""
from __future__ import annotations

import random
from typing import Optional

import torch
from torch import nn

from . import autograd_ops as ag
from . import ops
from .encoders import AVHUBERT_ARCHS, AVHubertVideoEncoder, LogMel, WhisperEncoder
from .Llama_LoRA import (FlatParams, LlamaForCausalLM_lora, PackedRows, TASKS, arch_from_name)
from .Qwen_LoRA import Qwen2ForCausalLM_lora

IGNORE_INDEX = -100


class _TrainLinear(nn.Module):
    def __init__(self, in_f, out_f, flat: FlatParams, name: str, bias=True):
        super().__init__()
        self.weight = flat.alloc((out_f, in_f), name + ".weight")
        self.bias = flat.alloc((out_f,), name + ".bias") if bias else None
        with torch.no_grad():   # nn.Linear default init
            bound = 1.0 / (in_f ** 0.5)
            self.weight.uniform_(-bound, bound)
            if self.bias is not None:
                self.bias.uniform_(-bound, bound)


class _TrainLayerNorm(nn.Module):
    def __init__(self, d, flat: FlatParams, name: str):
        super().__init__()
        self.weight = flat.alloc((d,), name + ".weight")
        self.bias = flat.alloc((d,), name + ".bias")
        self.d = d
        with torch.no_grad():
            self.weight.fill_(1.0)
            self.bias.zero_()


class Projector(nn.Sequential):
    """nn.Sequential(Linear, ReLU, Linear[, LayerNorm]) with the reference's child names 0, 1, 2[, 3]
    (state-dict keys audio_proj.{k}.{0,2}.{weight,bias}); forward runs the tcgen05 GEMMs with fused bias(+ReLU)."""

    def __init__(self, in_dim, intermediate, hidden, layernorm, flat, name):
        mods = [_TrainLinear(in_dim, intermediate, flat, name + ".0"), nn.ReLU(),
                _TrainLinear(intermediate, hidden, flat, name + ".2")]
        if layernorm:
            mods.append(_TrainLayerNorm(hidden, flat, name + ".3"))
        super().__init__(*mods)

    def forward(self, x):
        shape = x.shape
        x2 = x.reshape(-1, shape[-1])
        h = ag.TrainableLinearFn.apply(x2, self[0].weight, self[0].bias, "relu")
        y = ag.TrainableLinearFn.apply(h, self[2].weight, self[2].bias, None)
        if len(self) == 4:   # nn.LayerNorm(hidden) of the non-Matryoshka / single-projector recipes (:85,:97,:111)
            y = ag.TrainableLayerNormFn.apply(y, self[3].weight, self[3].bias, 1e-5)
        return y.view(*shape[:-1], y.shape[-1])

    @property
    def fusable(self) -> bool:
        """Linear + ReLU + Linear only: the whole projector runs inside omni_pool_project_splice."""
        return len(self) == 3

    @staticmethod
    def param_count(in_dim, intermediate, hidden, layernorm):
        return in_dim * intermediate + intermediate + intermediate * hidden + hidden + (2 * hidden if layernorm else 0) + 64


class CompressFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, enc, n_tok, rate, mode):
        ctx.meta = (n_tok, enc.shape[1], rate, mode)
        return ops.matryoshka_compress(enc, n_tok, rate, mode)

    @staticmethod
    def backward(ctx, d):
        n_tok, t_full, rate, mode = ctx.meta
        return ops.matryoshka_compress_bwd(d.contiguous(), n_tok, t_full, rate, mode), None, None, None


def compress(enc, n_tok, rate, mode):
    enc = enc.contiguous()
    if enc.requires_grad:
        return CompressFn.apply(enc, n_tok, rate, mode)
    return ops.matryoshka_compress(enc, n_tok, rate, mode)


class SpliceFn(torch.autograd.Function):
    """One launch: media tokens + marker/prompt/text embedding rows -> packed LLM input rows + labels."""

    @staticmethod
    def forward(ctx, audio_tok, video_tok, layout, rows, want_labels):
        H = layout.H
        xp = torch.empty((rows.M, H), device=layout.keep[2].device, dtype=torch.bfloat16)
        outs, labs, lab_t = [None] * 3, [None] * 3, []
        seg_of = {task: (B, S, off) for (task, B, S, off) in rows.segments}
        _pad_rows_zero(xp, rows)
        for t in range(3):
            if t in seg_of:
                B, S, off = seg_of[t]
                outs[t] = xp[off: off + B * S]
                if want_labels:
                    labs[t] = torch.empty((B, S), device=xp.device, dtype=torch.int64)
        ops.splice_prompt(layout, outs, labs)
        ctx.layout, ctx.seg_of = layout, seg_of
        ctx.mark_non_differentiable(*[l for l in labs if l is not None])
        return (xp, *[l if l is not None else torch.empty(0, device=xp.device, dtype=torch.int64) for l in labs])

    @staticmethod
    def backward(ctx, dxp, *_):
        dxp = dxp.contiguous()
        douts = [None] * 3
        for t, (B, S, off) in ctx.seg_of.items():
            douts[t] = dxp[off: off + B * S]
        da, dv = ops.splice_prompt_bwd(ctx.layout, douts, ctx.needs_input_grad[0], ctx.needs_input_grad[1])
        return da, dv, None, None, None


def _pad_rows_zero(xp, rows):
    cur = 0
    for (task, B, S, off) in rows.segments:       # zero only the pad rows between the 128-aligned segments
        if off > cur:
            xp[cur:off].zero_()
        cur = off + B * S
    if cur < rows.M:
        xp[cur:].zero_()


class PoolProjectSpliceFn(torch.autograd.Function):
    """Fused compression -> projector -> splice (ops.pool_project_splice, ONE launch) with the backward of the unfused
    chain: splice_bwd -> Linear-2 (dgrad / wgrad / bias) -> ReLU mask -> Linear-1 (wgrad / bias, dgrad only where the
    encoder trains) -> compression backward.  Tensor inputs: audio_enc, video_enc, then (w1, b1, w2, b2) per modality."""

    @staticmethod
    def forward(ctx, audio_enc, video_enc, w1a, b1a, w2a, b2a, w1v, b1v, w2v, b2v, meta):
        layout, rows, n_tok_a, ra, n_tok_v, rv, mode, want_labels = meta
        H = layout.H
        xp = torch.empty((rows.M, H), device=layout.keep[2].device, dtype=torch.bfloat16)
        _pad_rows_zero(xp, rows)
        outs, labs = [None] * 3, [None] * 3
        seg_of = {task: (B, S, off) for (task, B, S, off) in rows.segments}
        for t in range(3):
            if t in seg_of:
                B, S, off = seg_of[t]
                outs[t] = xp[off: off + B * S]
                if want_labels:
                    labs[t] = torch.empty((B, S), device=xp.device, dtype=torch.int64)
        a_in = ops.PoolProjectInput(audio_enc.detach(), n_tok_a, ra, w1a.detach(), b1a.detach(), w2a.detach(), b2a.detach()) \
            if audio_enc is not None else None
        v_in = ops.PoolProjectInput(video_enc.detach(), n_tok_v, rv, w1v.detach(), b1v.detach(), w2v.detach(), b2v.detach()) \
            if video_enc is not None else None
        res = ops.pool_project_splice(layout, outs, labs, a_in, v_in, mode)
        saved = []
        for r, w1, w2 in ((res["audio"], w1a, w2a), (res["video"], w1v, w2v)):
            saved += [r[0], r[1], w1, w2] if r is not None else [None, None, None, None]
        ctx.save_for_backward(*saved)
        ctx.layout, ctx.seg_of, ctx.meta = layout, seg_of, meta
        ctx.enc_shapes = (None if audio_enc is None else audio_enc.shape[1], None if video_enc is None else video_enc.shape[1])
        ctx.mark_non_differentiable(*[l for l in labs if l is not None])
        return (xp, *[l if l is not None else torch.empty(0, device=xp.device, dtype=torch.int64) for l in labs])

    @staticmethod
    def backward(ctx, dxp, *_):
        layout, rows, n_tok_a, ra, n_tok_v, rv, mode, _wl = ctx.meta
        dxp = dxp.contiguous()
        douts = [None] * 3
        for t, (B, S, off) in ctx.seg_of.items():
            douts[t] = dxp[off: off + B * S]
        pa, ha, w1a, w2a, pv, hv, w1v, w2v = ctx.saved_tensors
        da, dv = ops.splice_prompt_bwd(layout, douts, pa is not None, pv is not None)
        grads = [None] * 10
        for k, (dtok, pooled, hidden, w1, w2, n_tok, rate, t_full, need_enc) in enumerate((
                (da, pa, ha, w1a, w2a, n_tok_a, ra, ctx.enc_shapes[0], ctx.needs_input_grad[0]),
                (dv, pv, hv, w1v, w2v, n_tok_v, rv, ctx.enc_shapes[1], ctx.needs_input_grad[1]))):
            if pooled is None:
                continue
            dy = dtok.view(-1, dtok.shape[-1])
            dh = ops.gemm(dy, ops.transpose(w2.detach().contiguous()))
            dh = dh * (hidden > 0)                                              # ReLU mask
            g = 2 + 4 * k
            grads[g + 2] = ops.gemm_wgrad(dy, hidden, mo=w2.shape[0], no=w2.shape[1])[0]
            grads[g + 3] = ops.colsum(dy)
            grads[g + 0] = ops.gemm_wgrad(dh, pooled, mo=w1.shape[0], no=w1.shape[1])[0]
            grads[g + 1] = ops.colsum(dh)
            if need_enc:
                dpool = ops.gemm(dh, ops.transpose(w1.detach().contiguous()))
                B = layout.B
                grads[k] = ops.matryoshka_compress_bwd(dpool.view(B, -1, dpool.shape[-1]), n_tok, t_full, rate, mode)
        return (*grads, None)


class AVSR_LLMs(nn.Module):
    def __init__(self, modality, pretrain_avhubert_enc_video, use_lora_avhubert, llm_model, hidden_size,
                 intermediate_size, tokenizer, prompt_audio, prompt_video, prompt_audiovisual, pad_id,
                 downsample_ratio_audio, downsample_ratio_video, audio_encoder_name, compression_mode,
                 unfrozen_modules, max_dec_tokens, num_beams, PETF_LLM_name=None, peft_config_llm=None,
                 remove_layernorm_from_projector=False, matry_weights=None,
                 is_task_specific=None, is_matryoshka=False, is_single_matry_projector=False,
                 device="cuda", llm_overrides: Optional[dict] = None, audio_arch=None, video_arch=None):
        super().__init__()
        self.modality = modality
        self.pretrain_avhubert_enc_video = pretrain_avhubert_enc_video
        self.max_dec_tokens = max_dec_tokens
        self.num_beams = num_beams
        self.downsample_ratio_audio = downsample_ratio_audio
        self.downsample_ratio_video = downsample_ratio_video
        self.audio_encoder_name = audio_encoder_name
        self.llm_model = llm_model
        self.peft_config_llm = peft_config_llm
        self.PETF_LLM_name = PETF_LLM_name
        self.compression_mode = compression_mode
        self.hidden_size = hidden_size
        self.remove_layernorm_from_projector = remove_layernorm_from_projector
        self.matry_weights = matry_weights
        self.is_task_specific = is_task_specific
        self.is_matryoshka = is_matryoshka
        self.is_single_matry_projector = is_single_matry_projector
        self.device_ = torch.device(device)
        self.fused_projector = True      # compression + projector + splice in one launch whenever the projectors allow it
        if compression_mode not in ("stack", "avg-pooling"):
            raise ValueError(compression_mode)
        if PETF_LLM_name != "lora":
            raise NotImplementedError("the Omni-AVSR hot path is the LoRA-adapted LLM (PETF_LLM_name='lora')")

        has_a = modality in ("audio", "audiovisual")
        has_v = modality in ("video", "audiovisual")
        rates_a = list(downsample_ratio_audio) if is_matryoshka else [downsample_ratio_audio]
        rates_v = list(downsample_ratio_video) if is_matryoshka else [downsample_ratio_video]

        # ---- one flat store for everything trainable (LLM LoRA, AV-HuBERT LoRA, projectors) ----------------
        larch = arch_from_name(llm_model, **(llm_overrides or {}))
        if hidden_size != larch.hidden_size:
            raise ValueError(f"hidden_size {hidden_size} != {llm_model} hidden {larch.hidden_size}")
        v_arch = video_arch or AVHUBERT_ARCHS["large" if (pretrain_avhubert_enc_video and "large" in
                                                          str(pretrain_avhubert_enc_video)) else "base"]
        audio_dim = (audio_arch.d_model if audio_arch else None)
        cap = LlamaForCausalLM_lora.lora_param_count(larch, peft_config_llm) + 4096
        if has_v:
            cap += AVHubertVideoEncoder.lora_param_count(v_arch)
        stack = compression_mode == "stack"

        # encoders first (the projector input widths come from them)
        if has_a:
            self.audio_encoder = (WhisperEncoder(audio_arch, device) if audio_arch is not None
                                  else WhisperEncoder.from_pretrained(audio_encoder_name, device))
            self.audio_frontend = LogMel(device)
            audio_dim = self.audio_encoder.config.hidden_size
        video_dim = v_arch.encoder_embed_dim
        for present, dim, rates in ((has_a, audio_dim, rates_a), (has_v, video_dim, rates_v)):
            if present:
                for r in rates:   # generous: one projector per rate, stack width
                    cap += Projector.param_count(dim * (r if stack else 1), intermediate_size, hidden_size, True)
        self.flat = FlatParams(device, cap)

        if has_v:
            self.video_encoder = AVHubertVideoEncoder(v_arch, device, self.flat, use_lora=bool(use_lora_avhubert))

        # ---- projectors (:65-111 audio, :152-196 video) ----------------------------------------------------
        if has_a:
            self.audio_proj, self.matry_map_audio = self._make_projectors("audio_proj", audio_dim, rates_a, intermediate_size,
                                                                         hidden_size, is_audio=True)
        if has_v:
            self.video_proj, self.matry_map_video = self._make_projectors("video_proj", video_dim, rates_v, intermediate_size,
                                                                         hidden_size, is_audio=False)

        # ---- LLM (:198-216) ---------------------------------------------------------------------------------
        cls = Qwen2ForCausalLM_lora if "Qwen" in llm_model else LlamaForCausalLM_lora
        self.llm = cls(larch, peft_config_llm, device=device, flat=self.flat)
        self.tokenizer = tokenizer
        if "llama" in llm_model:
            self.llm.config.pad_token_id = pad_id
        self.llm.resize_token_embeddings(len(self.tokenizer))
        for p in self.llm.parameters():
            p.requires_grad_(False)

        start = 0 if "Qwen" in self.llm_model else 1                       # :218
        emb = self.llm.model.embed_tokens
        self._prompt_ids = {}
        for name, prompt in (("prompt_audio", prompt_audio), ("prompt_video", prompt_video),
                             ("prompt_audiovisual", prompt_audiovisual)):
            ids = self.tokenizer(prompt, return_tensors="pt").input_ids[:, start:-1].to(device)
            self._prompt_ids[name] = ids
            self.register_buffer(name, emb(ids).detach().clone())        # [1, P, H] buffers (:219-221)
        self.prompt_audio_len = self.prompt_audio.shape[1]
        self.prompt_video_len = self.prompt_video.shape[1]
        self.prompt_audiovisual_len = self.prompt_audiovisual.shape[1]
        v = self.tokenizer.vocab
        self._marker_ids = (v["<audio>"], v["</audio>"], v["<video>"], v["</video>"])
        self._has_bos = "Qwen" not in self.llm_model
        self._unfreeze_PETF(unfrozen_modules)

    # ------------------------------------------------------------------------------------------------
    def _make_projectors(self, name, dim, rates, inter, hidden, is_audio):
        stack = self.compression_mode == "stack"
        mm = {el: i for i, el in enumerate(rates)} if self.is_matryoshka else None
        rm = self.remove_layernorm_from_projector
        if self.is_matryoshka:
            if stack:
                # audio :74-77 (condition inverted in the reference: LN only when remove_layernorm is True);
                # video :159-162 (LayerNorm passed as Linear's bias argument => never a LayerNorm)
                ln = rm if is_audio else False
                proj = nn.ModuleList([Projector(dim * r, inter, hidden, ln, self.flat, f"{name}.{i}")
                                      for i, r in enumerate(rates)])
            elif self.is_single_matry_projector:
                proj = Projector(dim, inter, hidden, not rm, self.flat, name)                      # :96-97,:101-102
            else:
                proj = nn.ModuleList([Projector(dim, inter, hidden, False, self.flat, f"{name}.{i}")  # :99,:104 (no LN)
                                      for i, _ in enumerate(rates)])
        else:
            r = rates[0]
            ln = (not rm)
            proj = Projector(dim * r if stack else dim, inter, hidden, ln, self.flat, name)        # :81-85,:107-111
        return proj, mm

    def _unfreeze_PETF(self, unfrozen_modules):
        """:234-260.  Projectors are always trainable (never frozen in the reference)."""
        if unfrozen_modules is None or None in unfrozen_modules:
            unfrozen_modules = []
        llm_on = "peft_llm" in unfrozen_modules
        for layer in self.llm.model.layers:
            layer.self_attn.lora_down.requires_grad_(llm_on)
            layer.self_attn.lora_up.requires_grad_(llm_on)
        if hasattr(self, "video_encoder"):
            avh_on = "lora_avhubert" in unfrozen_modules
            for p in self.video_encoder.lora_parameters():
                p.requires_grad_(avh_on)

    # Parts of the reference's `video_encoder` (the full fairseq AVHubertModel: `remove_pretraining_modules` is never called,
    # modeling_OmniAVSR.py:123-126) that the video-only `extract_finetune` path never touches.  A real `model_avg_N.pth`
    # carries them; this mirror has no such modules, so they are dropped before the (strict) load.
    _UNUSED_AVHUBERT = ("video_encoder.mask_emb", "video_encoder.label_embs_concat", "video_encoder.final_proj.",
                        "video_encoder.feature_extractor_audio.", "video_encoder.target_glu.",
                        "video_encoder.encoder.layers.", "video_encoder.feature_extractor_video.resnet.")

    def load_state_dict(self, state_dict, strict=True, assign=False):
        """Accepts the reference's bare AVSR_LLMs state dict (lightning_OmniAVSR.py:148-150); the packed / transposed
        copies used by the kernels are rebuilt lazily afterwards."""
        own = set(self.state_dict().keys())
        sd = {}
        for k, v in state_dict.items():
            if k not in own and k.startswith(self._UNUSED_AVHUBERT[:5]):
                continue                                   # pre-training heads / audio front-end of AV-HuBERT
            if k not in own and k.endswith("num_batches_tracked"):
                continue                                   # BatchNorm counters (the mirror runs eval-mode BN)
            sd[k] = v
        out = super().load_state_dict(sd, strict=strict)
        self.refresh_prompts()
        for mod in self.modules():
            if hasattr(mod, "_wt"):
                mod._wt = None
            if hasattr(mod, "_head_t"):
                mod._head_t = None
        return out

    def refresh_prompts(self):
        """Re-embed the three task prompts from the CURRENT embedding table.  The reference computes the prompt buffers
        after `from_pretrained` (:218-221); here weights arrive later (checkpoints.load_llm / load_state_dict), so the
        buffers are recomputed whenever the table changes -- unless the state dict itself provided them."""
        if not isinstance(self._prompt_ids, dict):
            return                                          # Llama-AVSR embeds its prompt at call time (no buffers)
        emb = self.llm.model.embed_tokens
        with torch.no_grad():
            for name, ids in self._prompt_ids.items():
                getattr(self, name).copy_(emb(ids).detach())

    def trainable_parameter_count(self) -> int:
        return sum(p.numel() for p in self.parameters() if p.requires_grad)

    # ------------------------------------------------------------------------------------------------
    def _project(self, feats, which, rate):
        proj = self.audio_proj if which == "audio" else self.video_proj
        mm = self.matry_map_audio if which == "audio" else self.matry_map_video
        if isinstance(proj, nn.ModuleList):
            proj = proj[mm[rate]]                      # KeyError for an unknown rate, as in the reference (:353,:366)
        return proj(feats)

    def _layout(self, inputs, audio_tok, video_tok, task_mask, with_labels):
        tokens = inputs["tokens"]
        labels = inputs.get("labels") if with_labels else None
        return ops.SpliceLayout(tokens=tokens.contiguous(), labels=None if labels is None else labels.contiguous(),
                                embed=self.llm.model.embed_tokens.weight.data, audio_tok=audio_tok, video_tok=video_tok,
                                prompts=[self.prompt_audio[0], self.prompt_video[0], self.prompt_audiovisual[0]],
                                marker_ids=self._marker_ids, has_bos=self._has_bos, task_mask=task_mask)

    def forward(self, inputs, is_trainval=True, modality=None, test_ratio_matry_audio=None, test_ratio_matry_video=None):
        if is_trainval:
            out = self.prepare_inputs(inputs, is_trainval, test_ratio_matry_audio=test_ratio_matry_audio,
                                      test_ratio_matry_video=test_ratio_matry_video)
            xp, rows, labels = out["packed"], out["rows"], out["labels"]
            hook = getattr(self, "_llm_input_hook", None)
            if hook is not None and xp.requires_grad:
                xp.register_hook(hook)             # fires when every LLM layer has finished its backward (dp.GradReducer)
            hid = self.llm.model.forward_packed(xp, rows)
            w = self.matry_weights if self.matry_weights else (1.0, 1.0, 1.0)
            segs = [(B, S, off) for (_, B, S, off) in rows.segments]
            losses = self.llm.loss_from_hidden(hid, segs, labels, [float(x) for x in w])
            return losses[0], losses[1], losses[2]                          # :302-306
        embeddings = self.prepare_inputs(inputs, is_trainval, test_ratio_matry_audio=test_ratio_matry_audio,
                                         test_ratio_matry_video=test_ratio_matry_video)
        vocab = self.tokenizer.vocab
        if "Qwen" in self.llm_model:                                        # :318-322
            return self.llm.generate(inputs_embeds=embeddings, max_new_tokens=self.max_dec_tokens, num_beams=self.num_beams,
                                     eos_token_id=vocab["<|endoftext|>"], pad_token_id=vocab["<|endoftext|>"],
                                     modality=modality, trim=not getattr(self, "decode_no_trim", False))
        return self.llm.generate(inputs_embeds=embeddings, max_new_tokens=self.max_dec_tokens, num_beams=self.num_beams,
                                 eos_token_id=vocab["<|end_of_text|>"], bos_token_id=vocab["<|begin_of_text|>"],
                                 pad_token_id=vocab["<pad>"], modality=modality,                # :313-317
                                 trim=not getattr(self, "decode_no_trim", False))

    # ------------------------------------------------------------------------------------------------
    def _projector_for(self, which, rate):
        proj = self.audio_proj if which == "audio" else self.video_proj
        mm = self.matry_map_audio if which == "audio" else self.matry_map_video
        if isinstance(proj, nn.ModuleList):
            proj = proj[mm[rate]]                      # KeyError for an unknown rate, as in the reference (:353,:366)
        return proj

    def _media(self, inputs, which, is_trainval, test_ratio):
        """(encoder output [B, T, D], n_tok, rate, projector) of one modality -- the inputs of the fused
        compression -> projector -> splice launch (rate 1 = no compression)."""
        if which == "audio":
            enc, n_tok = self._audio_encoder_output(inputs["audio"], max(inputs["lengths"]))
            rates = self.downsample_ratio_audio
            mm = self.matry_map_audio
        else:
            enc = self._video_encoder_output(inputs["video"])
            n_tok = enc.shape[1]
            rates = self.downsample_ratio_video
            mm = self.matry_map_video
        if self.is_matryoshka:
            rate = self._pick_rate(rates, test_ratio, is_trainval)
            if rate not in mm:
                raise KeyError(rate)
        else:
            rate = rates
        return enc.contiguous(), n_tok, rate, self._projector_for(which, rate)

    def _fused_inputs(self, inputs, is_trainval, use_a, use_v, ra_test, rv_test):
        a = self._media(inputs, "audio", is_trainval, ra_test) if use_a else None
        v = self._media(inputs, "video", is_trainval, rv_test) if use_v else None
        return a, v

    def prepare_inputs(self, inputs, is_trainval, test_ratio_matry_audio=None, test_ratio_matry_video=None):
        use_a = True if is_trainval else self.modality in ("audio", "audiovisual")
        use_v = True if is_trainval else self.modality in ("video", "audiovisual")
        fused = self.fused_projector and all(
            isinstance(p, Projector) and p.fusable
            for p in ([*(self.audio_proj if isinstance(self.audio_proj, nn.ModuleList) else [self.audio_proj])] if use_a else []) +
                     ([*(self.video_proj if isinstance(self.video_proj, nn.ModuleList) else [self.video_proj])] if use_v else []))
        if fused:
            return self._prepare_inputs_fused(inputs, is_trainval, use_a, use_v, test_ratio_matry_audio, test_ratio_matry_video)
        return self._prepare_inputs_unfused(inputs, is_trainval, test_ratio_matry_audio, test_ratio_matry_video)

    def _prepare_inputs_fused(self, inputs, is_trainval, use_a, use_v, ra_test, rv_test):
        """ONE launch (omni_pool_project_splice): compression + projector MLP + marker / prompt / text splice + labels."""
        a, v = self._fused_inputs(inputs, is_trainval, use_a, use_v, ra_test, rv_test)
        tokens = inputs["tokens"]
        B = tokens.shape[0]
        mode = self.compression_mode
        na = a[1] // a[2] if a is not None else None
        nv = v[1] // v[2] if v is not None else None
        for x in (a, v):
            if x is not None and mode == "avg-pooling" and x[1] // x[2] == 0:
                raise RuntimeError(f"Given input size: ({x[0].shape[2]}x1x{x[1]}). Calculated output size: "
                                   f"({x[0].shape[2]}x1x0). Output size is too small")       # nn.AvgPool1d (:545)
        prompts = [self.prompt_audio[0], self.prompt_video[0], self.prompt_audiovisual[0]]
        embed = self.llm.model.embed_tokens.weight.data

        def weights(x):
            if x is None:
                return (None,) * 4
            p = x[3]
            return p[0].weight, p[0].bias, p[2].weight, p[2].bias

        if is_trainval:
            labels = inputs.get("labels")
            layout = ops.SpliceLayout(tokens=tokens.contiguous(), labels=None if labels is None else labels.contiguous(),
                                      embed=embed, audio_tok=None, video_tok=None, prompts=prompts,
                                      marker_ids=self._marker_ids, has_bos=self._has_bos, task_mask=7, n_audio=na, n_video=nv)
            rows = PackedRows.get([(t, B, layout.seq_len[t]) for t in range(3)], embed.device)
            meta = (layout, rows, a[1], a[2], v[1], v[2], mode, True)
            xp, la, lv, lav = PoolProjectSpliceFn.apply(a[0], v[0], *weights(a), *weights(v), meta)
            return {"packed": xp, "rows": rows, "labels": [la, lv, lav], "labels_audio": la, "labels_video": lv,
                    "labels_audiovisual": lav, "selected_rates": (a[2], v[2])}
        t = TASKS.index(self.modality)
        tok1 = tokens[:, :1] if self._has_bos else tokens[:, :0]          # only e(BOS) is used (:419)
        layout = ops.SpliceLayout(tokens=tok1.contiguous(), labels=None, embed=embed, audio_tok=None, video_tok=None,
                                  prompts=prompts, marker_ids=self._marker_ids, has_bos=self._has_bos, task_mask=1 << t,
                                  n_audio=na, n_video=nv)
        out = [None] * 3
        out[t] = torch.empty((B, layout.seq_len[t], self.hidden_size), device=tokens.device, dtype=torch.bfloat16)
        with torch.no_grad():
            mk = lambda x: None if x is None else ops.PoolProjectInput(x[0], x[1], x[2], *[w.data for w in weights(x)])
            ops.pool_project_splice(layout, out, [None] * 3, mk(a), mk(v), mode)
        return out[t]

    def _prepare_inputs_unfused(self, inputs, is_trainval, test_ratio_matry_audio=None, test_ratio_matry_video=None):
        """Separate launches (compress -> GEMM -> GEMM [-> LayerNorm] -> splice): projectors with a LayerNorm (the
        non-Matryoshka and single-projector recipes, :85,:97,:111) and the parity tests of the individual kernels."""
        if is_trainval:
            if self.is_matryoshka:
                audio_features, ra = self.encode_audio(inputs["audio"], max(inputs["lengths"]), is_trainval=is_trainval,
                                                       test_ratio_matry_audio=test_ratio_matry_audio)
                video_features, rv = self.encode_video(inputs["video"], is_trainval=is_trainval,
                                                       test_ratio_matry_video=test_ratio_matry_video)
            else:
                audio_features, ra = self.encode_audio(inputs["audio"], max(inputs["lengths"])), self.downsample_ratio_audio
                video_features, rv = self.encode_video(inputs["video"]), self.downsample_ratio_video
            audio_tok = self._project(audio_features, "audio", ra).contiguous()
            video_tok = self._project(video_features, "video", rv).contiguous()
            layout = self._layout(inputs, audio_tok, video_tok, 7, True)
            B = inputs["tokens"].shape[0]
            rows = PackedRows.get([(t, B, layout.seq_len[t]) for t in range(3)], audio_tok.device)
            xp, la, lv, lav = SpliceFn.apply(audio_tok, video_tok, layout, rows, True)
            return {"packed": xp, "rows": rows, "labels": [la, lv, lav], "labels_audio": la, "labels_video": lv,
                    "labels_audiovisual": lav, "selected_rates": (ra, rv)}
        # ---- inference (:397-458): [bos, A?, V?, prompt_t] ------------------------------------------------------
        use_a = self.modality in ("audio", "audiovisual")
        use_v = self.modality in ("video", "audiovisual")
        audio_tok = video_tok = None
        if use_a:
            if self.is_matryoshka:
                af = self.encode_audio(inputs["audio"], max(inputs["lengths"]), is_trainval=is_trainval,
                                       test_ratio_matry_audio=test_ratio_matry_audio)
                ra = test_ratio_matry_audio
            else:
                af, ra = self.encode_audio(inputs["audio"], max(inputs["lengths"])), self.downsample_ratio_audio
            audio_tok = self._project(af, "audio", ra).contiguous()
        if use_v:
            if self.is_matryoshka:
                vf = self.encode_video(inputs["video"], is_trainval=is_trainval, test_ratio_matry_video=test_ratio_matry_video)
                rv = test_ratio_matry_video
            else:
                vf, rv = self.encode_video(inputs["video"]), self.downsample_ratio_video
            video_tok = self._project(vf, "video", rv).contiguous()
        t = TASKS.index(self.modality)
        tokens = inputs["tokens"]
        tok1 = tokens[:, :1] if self._has_bos else tokens[:, :0]          # only e(BOS) is used (:419)
        layout = self._layout({"tokens": tok1.contiguous()}, audio_tok, video_tok, 1 << t, False)
        B = tokens.shape[0]
        out = [None] * 3
        out[t] = torch.empty((B, layout.seq_len[t], self.hidden_size), device=tokens.device, dtype=torch.bfloat16)
        ops.splice_prompt(layout, out, [None] * 3)
        return out[t]

    # ------------------------------------------------------------------------------------------------
    def _pick_rate(self, rates, test_ratio, is_trainval):
        if is_trainval and not test_ratio:
            return random.choice(rates)                                     # :474,:549 (python RNG)
        return test_ratio

    def _video_encoder_output(self, videos):
        src = torch.reshape(videos, (-1, videos.shape[2], videos.shape[1], videos.shape[3], videos.shape[-1]))  # :463
        video_enc, _, _ = self.video_encoder.extract_finetune(source={"video": src, "audio": None})
        return video_enc

    def _audio_encoder_output(self, audio, max_len):
        feats = self.audio_frontend(audio.squeeze(-1))                                  # :531-533 on the GPU
        audio_enc = self.audio_encoder(feats).last_hidden_state                         # :534
        ml = max_len if torch.is_tensor(max_len) else torch.tensor(max_len)
        n_tok = max(int(ml.detach().cpu().to(torch.int64) / 16000 * 50), 25)            # :537 (float32 tensor arithmetic)
        return audio_enc, min(n_tok, audio_enc.shape[1])

    def encode_video(self, videos, is_trainval=None, test_ratio_matry_video=None):
        video_enc = self._video_encoder_output(videos)
        n_tok = video_enc.shape[1]
        if self.is_matryoshka:
            rate = self._pick_rate(self.downsample_ratio_video, test_ratio_matry_video, is_trainval)
            if rate not in self.matry_map_video:
                raise KeyError(rate)
            out = compress(video_enc, n_tok, rate, self.compression_mode)
            return (out, rate) if is_trainval else out
        if self.downsample_ratio_video != 1:
            return compress(video_enc, n_tok, self.downsample_ratio_video, self.compression_mode)
        return video_enc

    def encode_audio(self, audio, max_len, is_trainval=None, test_ratio_matry_audio=None):
        audio_enc, n_tok = self._audio_encoder_output(audio, max_len)
        if self.is_matryoshka:
            rate = self._pick_rate(self.downsample_ratio_audio, test_ratio_matry_audio, is_trainval)
            if rate not in self.matry_map_audio:
                raise KeyError(rate)
            out = compress(audio_enc, n_tok, rate, self.compression_mode)
            return (out, rate) if is_trainval else out
        if self.downsample_ratio_audio != 1:
            return compress(audio_enc, n_tok, self.downsample_ratio_audio, self.compression_mode)
        return audio_enc[:, :n_tok].contiguous()
