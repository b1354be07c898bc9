"""Drop-in mirror of the reference's Omni_AVSR/Qwen_LoRA.py (Qwen2.5 0.5B-32B with Omni-LoRA on q/v).

Reference symbols (file:line in /root/reference/Omni_AVSR/Qwen_LoRA.py): QwenLoRA_config :92-103,
Qwen2SdpaAttention_lora :452-620 (per-size GQA factor table :464-475, LoRA math :557-570),
Qwen2ForCausalLM_lora :105-251.  Everything below the class names is shared with Llama_LoRA.py: the only Qwen
specifics are the q/k/v bias (added in the GEMM epilogue), rms_norm_eps 1e-6, rope theta 1e6 without scaling and
the absence of a BOS token (handled by the splice kernel's has_bos flag).
"""
from __future__ import annotations

from dataclasses import dataclass

from .Llama_LoRA import (LlamaDecoderLayer_lora, LlamaForCausalLM_lora, LlamaModel_lora, LlamaSdpaAttention_lora,
                         LLMArch)


@dataclass
class QwenLoRA_config:
    RANK: int
    ALPHA: int = 1
    IS_QWEN25_0_5B: bool = False
    IS_QWEN25_1_5B: bool = False
    IS_QWEN25_3B: bool = False
    IS_QWEN25_7B: bool = False
    IS_QWEN25_14B: bool = False
    IS_QWEN25_32B: bool = False
    IS_TASK_SPECIFIC: bool = False
    SHARED_LORA: bool = False


def _gqa_factor(lc: QwenLoRA_config) -> int:
    if lc.IS_QWEN25_0_5B:
        return 7
    if lc.IS_QWEN25_1_5B:
        return 6
    if lc.IS_QWEN25_3B:
        return 8
    if lc.IS_QWEN25_7B:
        return 7
    if lc.IS_QWEN25_14B or lc.IS_QWEN25_32B:
        return 5
    raise AssertionError("Only Qwen2.5 0.5B, 1.5B, 3B, 7B, 14B, 32B models are supported")


class Qwen2SdpaAttention_lora(LlamaSdpaAttention_lora):
    def __init__(self, config: LLMArch, lora_config: QwenLoRA_config, layer_idx=None, flat=None, device="cuda"):
        super().__init__(config, lora_config, layer_idx, flat, device,
                         kv_out_dim=config.hidden_size // _gqa_factor(lora_config))


class Qwen2DecoderLayer_lora(LlamaDecoderLayer_lora):
    attention_cls = Qwen2SdpaAttention_lora


class Qwen2Model_lora(LlamaModel_lora):
    layer_cls = Qwen2DecoderLayer_lora


class Qwen2ForCausalLM_lora(LlamaForCausalLM_lora):
    model_cls = Qwen2Model_lora
