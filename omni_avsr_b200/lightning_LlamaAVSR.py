"""Mirror of the reference's Omni_AVSR/lightning_LlamaAVSR.py (ModelModule_LLM for Llama-AVSR / Llama-MTSK) without
pytorch-lightning.  Reference (file:line in /root/reference/Omni_AVSR/lightning_LlamaAVSR.py): prompt choice :83-90,
shared LoRA config :94-105, AVSR_LLMs construction :107-127, training_step :146-157, validation_step :159-164, test_step
:166-177.  Optimizer / DDP semantics are those of lightning_OmniAVSR.ModelModule_LLM (one all-reduce of the flat
trainable-gradient buffer + fused clip + AdamW)."""
from __future__ import annotations

from typing import Optional

import torch

from . import dp
from .lightning_OmniAVSR import (DEFAULT_ARGS, DEFAULT_PAD_TOKEN, ModelModule_LLM as _OmniModule, SyntheticTokenizer,
                                 compute_word_level_distance, llm_size, make_args as _make_args)
from .Llama_LoRA import LoRA_config
from .modeling_LlamaAVSR import AVSR_LLMs
from .Qwen_LoRA import QwenLoRA_config


def make_args(**kw):
    d = dict(modality="audio", is_task_specific=False, use_shared_lora_task_specific=False, is_matryoshka=False,
             downsample_ratio_audio=4, downsample_ratio_video=2, downsample_ratio_test_matry=None, matry_weights=None)
    d.update(kw)
    return _make_args(**d)


class ModelModule_LLM(_OmniModule):
    def __init__(self, args, tokenizer=None, device="cuda", model_kwargs: Optional[dict] = None):
        torch.nn.Module.__init__(self)
        self.args = args
        if args.use_lora_avhubert:
            assert "lora_avhubert" in args.unfrozen_modules, "LoRA modules for the AV-HuBERT encoder must be unfrozen!!"
        self.tokenizer = tokenizer if tokenizer is not None else SyntheticTokenizer(args.llm_model)
        pad_id = self.tokenizer.convert_tokens_to_ids(DEFAULT_PAD_TOKEN) if "llama" in args.llm_model else None
        prompt = {"audio": args.prompt_audio, "video": args.prompt_video}.get(args.modality, args.prompt_audiovisual)  # :83-90
        n = args.llm_model
        if "Qwen" in n:                                                                                       # :95-100
            lora_config_llm = QwenLoRA_config(args.rank, args.alpha, n == "Qwen/Qwen2.5-0.5B", n == "Qwen/Qwen2.5-1.5B",
                                              n == "Qwen/Qwen2.5-3B", n == "Qwen/Qwen2.5-7B")
        else:                                                                                                 # :102-105
            is_l3 = n in ("meta-llama/Meta-Llama-3-8B", "meta-llama/Meta-Llama-3.1-8B", "meta-llama/Llama-3.2-1B")
            lora_config_llm = LoRA_config(args.rank, args.alpha, is_l3, n == "meta-llama/Llama-3.2-3B")
        mk = dict(model_kwargs or {})
        hidden = mk.pop("hidden_size_override", None) or llm_size[n]
        self.model = AVSR_LLMs(
            modality=args.modality, pretrain_avhubert_enc_video=args.pretrain_avhubert_enc_video_path,
            use_lora_avhubert=args.use_lora_avhubert, llm_model=n, hidden_size=hidden,
            intermediate_size=args.intermediate_size, tokenizer=self.tokenizer, prompt=prompt, pad_id=pad_id,
            downsample_ratio_audio=args.downsample_ratio_audio, downsample_ratio_video=args.downsample_ratio_video,
            audio_encoder_name=args.audio_encoder_name, compression_mode=args.compression_mode,
            unfrozen_modules=args.unfrozen_modules, max_dec_tokens=args.max_dec_tokens, num_beams=args.num_beams,
            PETF_LLM_name=args.add_PETF_LLM, peft_config_llm=lora_config_llm,
            remove_layernorm_from_projector=args.no_layernorm_projector, is_matryoshka=args.is_matryoshka,
            device=device, **mk)
        self.model._unfreeze_PETF(args.unfrozen_modules)
        if getattr(args, "pretrained_model_path", None):
            self.model.load_state_dict(torch.load(args.pretrained_model_path, map_location=device))
        self.total_length = 0
        self.total_edit_distance = 0
        self.global_step = 0
        self._opt = None

    def training_step(self, batch, batch_idx=0, rates=None):
        train_loss = self.model(batch, is_trainval=True)                                      # :147
        self.last_losses = (train_loss.detach(),)
        return train_loss * dp.loss_scale(batch["tokens"].shape[0], device=train_loss.device)  # :152-154

    def validation_step(self, batch, batch_idx=0):
        with torch.no_grad():
            return self.model(batch, is_trainval=True)                                        # :160

    def on_test_epoch_start(self):
        self.total_length = 0
        self.total_edit_distance = 0

    def test_step(self, batch, batch_idx=0):
        if self.args.is_matryoshka:                                                            # :167-170
            generated_ids = self.model(batch, is_trainval=False, test_ratio_matry=self.args.downsample_ratio_test_matry)
        else:
            generated_ids = self.model(batch, is_trainval=False)
        if "gold_text" in batch:
            text = self.tokenizer.batch_decode(generated_ids, skip_special_tokens=True)[0]
            self.total_edit_distance += compute_word_level_distance(batch["gold_text"], text)
            self.total_length += len(batch["gold_text"].split())
        return generated_ids
