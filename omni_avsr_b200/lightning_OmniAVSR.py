"""Mirror of the reference's Omni_AVSR/lightning_OmniAVSR.py (ModelModule_LLM) without pytorch-lightning.

Reference (file:line in /root/reference/Omni_AVSR/lightning_OmniAVSR.py): llm_size :28-37, tokenizer surgery
:53-91, LoRA config :99-113, AVSR_LLMs construction :115-140, configure_optimizers :152-157, training_step
:159-176, validation_step :178-192, test_step :194-209, WER :40-42,:206-219.
Trainer semantics folded in (train_OmniAVSR.py:40-56): precision bf16-true, gradient_clip_val 10, DDP all-reduce of
the trainable gradients -- here ONE NCCL all-reduce of the flat gradient buffer followed by ONE fused
clip + AdamW kernel.
"""
from __future__ import annotations

import math
import zlib
from types import SimpleNamespace
from typing import Optional

import torch
import torch.distributed as dist

from . import dp, ops
from .Llama_LoRA import LoRA_config
from .modeling_OmniAVSR import AVSR_LLMs
from .Qwen_LoRA import QwenLoRA_config

DEFAULT_PAD_TOKEN = "<pad>"
AUDIO_SOS, AUDIO_EOS, VIDEO_SOS, VIDEO_EOS = "<audio>", "</audio>", "<video>", "</video>"

llm_size = {"meta-llama/Meta-Llama-3.1-8B": 4096, "meta-llama/Llama-3.2-1B": 2048, "meta-llama/Llama-3.2-3B": 3072,
            "Qwen/Qwen2.5-0.5B": 896, "Qwen/Qwen2.5-1.5B": 1536, "Qwen/Qwen2.5-3B": 2048, "Qwen/Qwen2.5-7B": 3584,
            "Qwen/Qwen2.5-14B": 5120, "Qwen/Qwen2.5-32B": 5120}


def compute_word_level_distance(seq1: str, seq2: str) -> int:
    """Word-level Levenshtein distance on lower-cased split text (:40-42; torchaudio.functional.edit_distance)."""
    a, b = seq1.lower().split(), seq2.lower().split()
    prev = list(range(len(b) + 1))
    for i, wa in enumerate(a, 1):
        cur = [i]
        for j, wb in enumerate(b, 1):
            cur.append(min(prev[j] + 1, cur[j - 1] + 1, prev[j - 1] + (wa != wb)))
        prev = cur
    return prev[-1]


class _Encoded:
    def __init__(self, ids):
        self.input_ids = ids


class SyntheticTokenizer:
    """Stand-in for AutoTokenizer when no tokenizer files are available (this environment has no hub access).

    Reproduces what the model path needs from the reference's tokenizer (:53-91): the special-token ids in the order
    `add_special_tokens` assigns them, BOS/EOS templating (Llama: BOS ... EOS, Qwen: ... EOS), right padding, and a
    deterministic word -> id map.  Constants follow SURVEY.md §8(d)."""

    def __init__(self, llm_model: str):
        self.is_qwen = "Qwen" in llm_model
        self.padding_side = "right"
        if self.is_qwen:
            self.base_vocab = 151643
            self.eos_token, self.eos_token_id = "<|endoftext|>", 151643
            self.bos_token, self.bos_token_id = None, None
            self.vocab = {"<|endoftext|>": 151643, AUDIO_SOS: 151665, AUDIO_EOS: 151666, VIDEO_SOS: 151667,
                          VIDEO_EOS: 151668}
            self.pad_token_id = 151643
            self._len = 151669
        else:
            self.base_vocab = 128000
            self.bos_token, self.bos_token_id = "<|begin_of_text|>", 128000
            self.eos_token, self.eos_token_id = "<|end_of_text|>", 128001
            self.vocab = {"<|begin_of_text|>": 128000, "<|end_of_text|>": 128001, DEFAULT_PAD_TOKEN: 128256,
                          AUDIO_SOS: 128257, AUDIO_EOS: 128258, VIDEO_SOS: 128259, VIDEO_EOS: 128260}
            self.pad_token_id = 128256
            self._len = 128261

    def __len__(self):
        return self._len

    def convert_tokens_to_ids(self, tok):
        return self.vocab[tok]

    def _words(self, text: str):
        out = []
        for w in text.replace(".", " .").split():
            pieces = [w] if len(w) <= 8 else [w[: len(w) // 2], w[len(w) // 2:]]   # "Transcribe" -> 2 pieces, as BPE does
            for p in pieces:
                out.append(zlib.crc32(p.encode()) % self.base_vocab)
        return out

    def encode(self, text: str):
        ids = self._words(text) + [self.eos_token_id]
        return ids if self.is_qwen else [self.bos_token_id] + ids

    def __call__(self, text, return_tensors=None, padding=None):
        if isinstance(text, str):
            return _Encoded(torch.tensor([self.encode(text)], dtype=torch.int64))
        rows = [self.encode(t) for t in text]
        L = max(len(r) for r in rows)
        return _Encoded(torch.tensor([r + [self.pad_token_id] * (L - len(r)) for r in rows], dtype=torch.int64))

    def batch_decode(self, ids, skip_special_tokens=True, clean_up_tokenization_spaces=False):
        special = set(self.vocab.values())
        out = []
        for row in ids.tolist():
            out.append(" ".join(f"w{t}" for t in row if not (skip_special_tokens and (t in special or t >= self.base_vocab))))
        return out


class WarmupCosineScheduler:
    """utils/cosine.py:6-25: linear warm-up for warmup_epochs then cosine to 0 over the remaining steps (per step)."""

    def __init__(self, base_lr, warmup_epochs, num_epochs, iter_per_epoch):
        self.base_lr = base_lr
        self.warmup_iter = warmup_epochs * iter_per_epoch
        self.total_iter = num_epochs * iter_per_epoch
        self.step_num = 0

    def lr(self):
        s = self.step_num
        if s < self.warmup_iter:
            return self.base_lr * s / max(self.warmup_iter, 1)
        decay = self.total_iter - self.warmup_iter
        return 0.5 * self.base_lr * (1 + math.cos(math.pi * (s - self.warmup_iter) / max(decay, 1)))

    def step(self):
        self.step_num += 1


DEFAULT_ARGS = dict(
    modality="audiovisual", pretrain_avhubert_enc_video_path="large_vox_iter5.pt", use_lora_avhubert=True,
    llm_model="meta-llama/Llama-3.2-1B", intermediate_size=2048, prompt_audio="Transcribe speech to text.",
    prompt_video="Transcribe video to text.", prompt_audiovisual="Transcribe speech and video to text.",
    downsample_ratio_audio=[4, 16], downsample_ratio_video=[2, 5], audio_encoder_name="openai/whisper-medium.en",
    compression_mode="avg-pooling", unfrozen_modules=["peft_llm", "lora_avhubert"], max_dec_tokens=32, num_beams=1,
    add_PETF_LLM="lora", rank=32, alpha=4, no_layernorm_projector=False, matry_weights=[1.0, 1.5, 1.0],
    is_task_specific=True, use_shared_lora_task_specific=True, is_matryoshka=True, is_single_matry_projector=False,
    lr=1e-3, weight_decay=0.1, warmup_epochs=1, max_epochs=8, pretrained_model_path=None,
    downsample_ratio_test_matry_audio=None, downsample_ratio_test_matry_video=None, gradient_clip_val=10.0,
    steps_per_epoch=1000,
)


def make_args(**kw) -> SimpleNamespace:
    d = dict(DEFAULT_ARGS)
    d.update(kw)
    return SimpleNamespace(**d)


class ModelModule_LLM(torch.nn.Module):
    def __init__(self, args, tokenizer=None, device="cuda", model_kwargs: Optional[dict] = None):
        super().__init__()
        self.args = args
        if args.use_lora_avhubert:
            assert "lora_avhubert" in args.unfrozen_modules, "LoRA modules for the AV-HuBERT encoder must be unfrozen!!"
        self.tokenizer = tokenizer if tokenizer is not None else SyntheticTokenizer(args.llm_model)
        pad_id = self.tokenizer.convert_tokens_to_ids(DEFAULT_PAD_TOKEN) if "llama" in args.llm_model else None
        if "Qwen" in args.llm_model:                                                          # :101-108
            n = args.llm_model
            lora_config_llm = QwenLoRA_config(args.rank, args.alpha, n == "Qwen/Qwen2.5-0.5B", n == "Qwen/Qwen2.5-1.5B",
                                              n == "Qwen/Qwen2.5-3B", n == "Qwen/Qwen2.5-7B", n == "Qwen/Qwen2.5-14B",
                                              n == "Qwen/Qwen2.5-32B", args.is_task_specific,
                                              args.use_shared_lora_task_specific)
        else:                                                                                  # :110-113
            is_l3 = args.llm_model in ("meta-llama/Meta-Llama-3-8B", "meta-llama/Meta-Llama-3.1-8B", "meta-llama/Llama-3.2-1B")
            lora_config_llm = LoRA_config(args.rank, args.alpha, is_l3, args.llm_model == "meta-llama/Llama-3.2-3B",
                                          args.is_task_specific, args.use_shared_lora_task_specific)
        hidden = (model_kwargs or {}).get("hidden_size_override") or llm_size[args.llm_model]
        mk = dict(model_kwargs or {})
        mk.pop("hidden_size_override", None)
        self.model = AVSR_LLMs(
            modality=args.modality, pretrain_avhubert_enc_video=args.pretrain_avhubert_enc_video_path,
            use_lora_avhubert=args.use_lora_avhubert, llm_model=args.llm_model, hidden_size=hidden,
            intermediate_size=args.intermediate_size, tokenizer=self.tokenizer, prompt_audio=args.prompt_audio,
            prompt_video=args.prompt_video, prompt_audiovisual=args.prompt_audiovisual, pad_id=pad_id,
            downsample_ratio_audio=args.downsample_ratio_audio, downsample_ratio_video=args.downsample_ratio_video,
            audio_encoder_name=args.audio_encoder_name, compression_mode=args.compression_mode,
            unfrozen_modules=args.unfrozen_modules, max_dec_tokens=args.max_dec_tokens, num_beams=args.num_beams,
            PETF_LLM_name=args.add_PETF_LLM, peft_config_llm=lora_config_llm,
            remove_layernorm_from_projector=args.no_layernorm_projector, matry_weights=args.matry_weights,
            is_task_specific=args.is_task_specific, is_matryoshka=args.is_matryoshka,
            is_single_matry_projector=args.is_single_matry_projector, device=device, **mk)
        self.model._unfreeze_PETF(args.unfrozen_modules)
        if getattr(args, "pretrained_model_path", None):
            self.model.load_state_dict(torch.load(args.pretrained_model_path, map_location=device))   # :148-150
        self.total_length = 0
        self.total_edit_distance = 0
        self.global_step = 0
        self._opt = None

    # ---- optimizer: NCCL all-reduce of the flat gradient buffer, then ONE global-norm clip + AdamW kernel over the flat trainable buffer ---------------
    def configure_optimizers(self):
        flat = self.model.flat
        n = flat.used
        dev = flat.data.device
        self._opt = dict(m=torch.zeros(n, device=dev, dtype=torch.float32), v=torch.zeros(n, device=dev, dtype=torch.float32),
                         sumsq=torch.zeros(1, device=dev, dtype=torch.float32), step=0)
        self.scheduler = WarmupCosineScheduler(self.args.lr, self.args.warmup_epochs, self.args.max_epochs,
                                               getattr(self.args, "steps_per_epoch", 1000))
        return self._opt

    def zero_grad_flat(self):
        self.model.flat.grad[: self.model.flat.used].zero_()

    def zero_grad(self, set_to_none: bool = False):
        """nn.Module.zero_grad(set_to_none=True) would detach the `.grad` views from the flat gradient buffer (the optimizer
        kernel would then read zeros forever); the gradients are zeroed in place instead."""
        self.zero_grad_flat()

    def optimizer_step(self, lr: Optional[float] = None):
        if self._opt is None:
            self.configure_optimizers()
        o, flat = self._opt, self.model.flat
        p, g = flat.flat()
        red = getattr(self, "_reducer", None)
        if red is not None:                        # started from the autograd hook on the LLM input (see training_step)
            grad_scale = red.finish()
            self._reducer = None
        else:
            grad_scale = dp.allreduce_flat_grad(g)     # the single NCCL all-reduce of the step (sum; 1/W applied below)
        o["step"] += 1
        o["sumsq"].zero_()
        ops.sumsq_(g, o["sumsq"])
        if lr is None:
            self.scheduler.step()
            lr = self.scheduler.lr()
        # frozen tensors (requires_grad=False: e.g. the AV-HuBERT adapters of an audiovisual Llama-AVSR run) are left alone,
        # weight decay included, as torch.optim.AdamW does for parameters without a gradient
        for (a, b) in flat.trainable_spans():
            ops.adamw_(p[a:b], g[a:b], o["m"][a:b], o["v"][a:b], lr=lr, beta1=0.9, beta2=0.98, eps=1e-8,
                       weight_decay=self.args.weight_decay, step=o["step"], grad_scale=grad_scale,
                       max_norm=float(getattr(self.args, "gradient_clip_val", 10.0)), sumsq=o["sumsq"])
        self.global_step += 1

    # ---- steps -------------------------------------------------------------------------------------------------
    def training_step(self, batch, batch_idx=0, rates=None):
        ra, rv = rates if rates is not None else (None, None)
        # :171-173: loss *= W / sum(batch sizes).  The gather of the batch sizes is started here and waited for after the
        # forward (one rank: the host constant 1 / B)
        scale = dp.LossScale(batch["tokens"].shape[0], device=batch["tokens"].device)
        if scale.w > 1 and not getattr(self, "no_sync", False):
            # (no_sync = True: a micro-batch of a gradient-accumulation step -- Lightning's accumulate_grad_batches under DDP --
            # whose gradients only accumulate in the flat buffer; the last micro-batch runs with no_sync = False)
            # gradient all-reduce in two pieces: the LLM adapters' range goes out as soon as the gradient of the LLM input
            # exists, under the backward of the projectors / AV-HuBERT encoder (see dp.GradReducer)
            flat = self.model.flat
            self._reducer = dp.GradReducer(flat.grad[: flat.used], self._llm_grad_offset())
            self.model._llm_input_hook = self._reducer.hook
        audio_loss, video_loss, audiovisual_loss = self.model(batch, is_trainval=True, test_ratio_matry_audio=ra,
                                                              test_ratio_matry_video=rv)
        self.model._llm_input_hook = None
        train_loss = (audio_loss + video_loss + audiovisual_loss) / 3                       # :162
        self.last_losses = (audio_loss.detach(), video_loss.detach(), audiovisual_loss.detach())
        return train_loss * scale.value()

    def _llm_grad_offset(self) -> int:
        """First element of the LLM adapters in the flat buffer (they are allocated last: [AV-HuBERT LoRA | projectors | LLM])."""
        off = getattr(self, "_llm_off", None)
        if off is None:
            offs = [o for (name, o, _) in self.model.flat.names if name.startswith("layers.")]
            off = self._llm_off = min(offs) if offs else self.model.flat.used
        return off

    def train_step(self, batch, rates=None, lr=None):
        """zero_grad -> training_step -> backward -> all-reduce + clip + AdamW.  Returns the (detached) loss."""
        self.zero_grad_flat()
        loss = self.training_step(batch, 0, rates)
        loss.backward()
        self.optimizer_step(lr)
        return loss.detach()

    def validation_step(self, batch, batch_idx=0):
        with torch.no_grad():
            if self.args.is_matryoshka:
                la, lv, lav = self.model(batch, is_trainval=True,
                                         test_ratio_matry_audio=self.args.downsample_ratio_test_matry_audio,
                                         test_ratio_matry_video=self.args.downsample_ratio_test_matry_video)
            else:
                la, lv, lav = self.model(batch, is_trainval=True)
        return (la + lv + lav) / 3

    def on_test_epoch_start(self):
        self.total_length = 0
        self.total_edit_distance = 0
        self.model.modality = self.args.modality                                              # :216

    def test_step(self, batch, batch_idx=0):
        a = self.args
        kw = {}
        if a.is_matryoshka:
            kw = dict(test_ratio_matry_audio=a.downsample_ratio_test_matry_audio,
                      test_ratio_matry_video=a.downsample_ratio_test_matry_video)
        if a.is_task_specific:
            kw["modality"] = self.model.modality
        generated_ids = self.model(batch, is_trainval=False, **kw)
        if "gold_text" in batch:
            text = self.tokenizer.batch_decode(generated_ids, skip_special_tokens=True)[0]
            self.total_edit_distance += compute_word_level_distance(batch["gold_text"], text)
            self.total_length += len(batch["gold_text"].split())
        return generated_ids

    def on_test_epoch_end(self):
        return self.total_edit_distance / max(self.total_length, 1)
