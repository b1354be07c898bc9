"""ORACLE (test infrastructure, not product code) -- CPU restatement of the Omni-LoRA LLMs.

Follows, op for op (same bf16 rounding points):
  * Omni_AVSR/Llama_LoRA.py:103-110 (LoRA_config), :113-316 (LlamaSdpaAttention_lora), :318-398
    (LlamaForCausalLM_lora.forward: lm_head on all positions -> fp32 -> shifted CE), :446-578 (LlamaModel_lora),
    :580-655 (LlamaDecoderLayer_lora);
  * Omni_AVSR/Qwen_LoRA.py:92-103 (QwenLoRA_config), :452-620 (Qwen2SdpaAttention_lora), :105-204.
The arithmetic of the base classes lives in the un-vendored dependency transformers==4.43.1
(requirements.txt:7): LlamaRMSNorm / LlamaMLP / LlamaRotaryEmbedding(+llama3 scaling) / apply_rotary_pos_emb /
repeat_kv and the Qwen2 twins; their published algorithm is restated here and cross-checked in
tests/test_oracle_llm.py against the installed transformers (5.5.0) LlamaForCausalLM / Qwen2ForCausalLM built
from config with the adapters switched off (lora_down == 0 => adapted model == base model, Llama_LoRA.py:166-175).

Parity status: PINNED against outputs of the reference itself.  The reference ships no tests / golden vectors, so
tests/golden/make_reference_golden.py executes the unmodified Llama_LoRA.py / Qwen_LoRA.py from /root/reference in the
build container (name-only transformers 4.43.1 -> 5.5 shims, tests/golden/_ref_compat.py) and freezes logits, losses
and greedy tokens for the S / T / ST adapter modes; tests/test_reference_golden.py holds this file to them (observed:
bit-identical logits on CPU).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Optional

import torch
import torch.nn.functional as F
from torch import nn

TASKS = ("audio", "video", "audiovisual")


@dataclass
class LoRA_config:  # Llama_LoRA.py:103-110
    RANK: int
    ALPHA: int = 1
    IS_LLAMA3: bool = False
    IS_LLAMA3_2_3B: bool = False
    IS_TASK_SPECIFIC: bool = False
    SHARED_LORA: bool = False


@dataclass
class QwenLoRA_config:  # Qwen_LoRA.py:92-103
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


@dataclass
class LLMConfig:
    """Shape spec of the named architectures."""
    family: str                 # "llama" | "qwen2"
    hidden_size: int
    intermediate_size: int
    num_hidden_layers: int
    num_attention_heads: int
    num_key_value_heads: int
    vocab_size: int
    rms_norm_eps: float
    rope_theta: float
    head_dim: Optional[int] = None
    rope_scaling: Optional[dict] = None     # llama3: {factor, low_freq_factor, high_freq_factor, original_max_position_embeddings}
    attention_bias: bool = False            # Qwen2: bias on q/k/v
    tie_word_embeddings: bool = True
    max_position_embeddings: int = 131072
    pad_token_id: Optional[int] = None
    # reference quirk: Lightning bf16-true casts Llama's inv_freq buffer to bf16; Qwen2 (4.43.1) caches
    # cos/sin computed from the fp32 inv_freq at construction time.
    inv_freq_dtype: str = "bf16"

    def __post_init__(self):
        if self.head_dim is None:
            self.head_dim = self.hidden_size // self.num_attention_heads


def llama_3_2_1b(vocab=128261):
    return LLMConfig("llama", 2048, 8192, 16, 32, 8, vocab, 1e-5, 500000.0, 64,
                     dict(factor=32.0, low_freq_factor=1.0, high_freq_factor=4.0, original_max_position_embeddings=8192),
                     False, True, inv_freq_dtype="bf16")


def llama_3_1_8b(vocab=128261):
    return LLMConfig("llama", 4096, 14336, 32, 32, 8, vocab, 1e-5, 500000.0, 128,
                     dict(factor=8.0, low_freq_factor=1.0, high_freq_factor=4.0, original_max_position_embeddings=8192),
                     False, False, inv_freq_dtype="bf16")


def qwen25_3b(vocab=151669):
    return LLMConfig("qwen2", 2048, 11008, 36, 16, 2, vocab, 1e-6, 1000000.0, 128, None, True, True,
                     max_position_embeddings=32768, inv_freq_dtype="fp32")


def compute_inv_freq(cfg: LLMConfig) -> torch.Tensor:
    """transformers modeling_rope_utils: default and `llama3` (_compute_llama3_parameters)."""
    dim = cfg.head_dim
    inv_freq = 1.0 / (cfg.rope_theta ** (torch.arange(0, dim, 2, dtype=torch.int64).to(torch.float) / dim))
    if cfg.rope_scaling is None:
        return inv_freq
    rs = cfg.rope_scaling
    factor, lo, hi, old = rs["factor"], rs["low_freq_factor"], rs["high_freq_factor"], rs["original_max_position_embeddings"]
    low_freq_wavelen = old / lo
    high_freq_wavelen = old / hi
    wavelen = 2 * math.pi / inv_freq
    inv_freq_llama = torch.where(wavelen > low_freq_wavelen, inv_freq / factor, inv_freq)
    smooth = (old / wavelen - lo) / (hi - lo)
    smoothed = (1 - smooth) * inv_freq_llama / factor + smooth * inv_freq_llama
    is_medium = ~(wavelen < high_freq_wavelen) * ~(wavelen > low_freq_wavelen)
    return torch.where(is_medium, smoothed, inv_freq_llama)


def rope_cos_sin(cfg: LLMConfig, position_ids: torch.Tensor, dtype: torch.dtype):
    """LlamaRotaryEmbedding.forward: fp32 angles -> cat(freqs, freqs) -> cos/sin -> cast to the activations' dtype."""
    inv_freq = compute_inv_freq(cfg)
    if cfg.inv_freq_dtype == "bf16" and dtype == torch.bfloat16:
        inv_freq = inv_freq.to(torch.bfloat16)
    inv = inv_freq[None, :, None].float().expand(position_ids.shape[0], -1, 1)
    pos = position_ids[:, None, :].float()
    freqs = (inv @ pos).transpose(1, 2)
    emb = torch.cat((freqs, freqs), dim=-1)
    return emb.cos().to(dtype), emb.sin().to(dtype)


def rotate_half(x):
    x1 = x[..., : x.shape[-1] // 2]
    x2 = x[..., x.shape[-1] // 2:]
    return torch.cat((-x2, x1), dim=-1)


def apply_rotary_pos_emb(q, k, cos, sin):
    cos = cos.unsqueeze(1)
    sin = sin.unsqueeze(1)
    return (q * cos) + (rotate_half(q) * sin), (k * cos) + (rotate_half(k) * sin)


def repeat_kv(x, n_rep):
    b, h, s, d = x.shape
    if n_rep == 1:
        return x
    return x[:, :, None, :, :].expand(b, h, n_rep, s, d).reshape(b, h * n_rep, s, d)


class RMSNorm(nn.Module):
    def __init__(self, hidden, eps):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(hidden))
        self.variance_epsilon = eps

    def forward(self, x):
        dt = x.dtype
        x = x.to(torch.float32)
        var = x.pow(2).mean(-1, keepdim=True)
        x = x * torch.rsqrt(var + self.variance_epsilon)
        return self.weight * x.to(dt)


class MLP(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.gate_proj = nn.Linear(cfg.hidden_size, cfg.intermediate_size, bias=False)
        self.up_proj = nn.Linear(cfg.hidden_size, cfg.intermediate_size, bias=False)
        self.down_proj = nn.Linear(cfg.intermediate_size, cfg.hidden_size, bias=False)

    def forward(self, x):
        return self.down_proj(F.silu(self.gate_proj(x)) * self.up_proj(x))


def _kv_out_dim(cfg: LLMConfig, lc) -> int:
    """Output width of lora_up_V as the reference computes it (Llama_LoRA.py:143-163, Qwen_LoRA.py:464-475)."""
    h = cfg.hidden_size
    if isinstance(lc, QwenLoRA_config):
        if lc.IS_QWEN25_0_5B: g = 7
        elif lc.IS_QWEN25_1_5B: g = 6
        elif lc.IS_QWEN25_3B: g = 8
        elif lc.IS_QWEN25_7B: g = 7
        elif lc.IS_QWEN25_14B or lc.IS_QWEN25_32B: g = 5
        else:
            raise AssertionError("Only Qwen2.5 0.5B, 1.5B, 3B, 7B, 14B, 32B models are supported")
        return h // g
    if lc.IS_LLAMA3:
        return h // 4
    if lc.IS_LLAMA3_2_3B:
        return h // 3
    return h


class Attention_lora(nn.Module):
    """LlamaSdpaAttention_lora (Llama_LoRA.py:113-316) / Qwen2SdpaAttention_lora (Qwen_LoRA.py:452-620)."""

    def __init__(self, cfg: LLMConfig, lc, layer_idx):
        super().__init__()
        self.cfg, self.lora_config, self.layer_idx = cfg, lc, layer_idx
        h, hd = cfg.hidden_size, cfg.head_dim
        self.num_heads, self.num_key_value_heads, self.head_dim = cfg.num_attention_heads, cfg.num_key_value_heads, hd
        self.num_key_value_groups = self.num_heads // self.num_key_value_heads
        self.q_proj = nn.Linear(h, self.num_heads * hd, bias=cfg.attention_bias)
        self.k_proj = nn.Linear(h, self.num_key_value_heads * hd, bias=cfg.attention_bias)
        self.v_proj = nn.Linear(h, self.num_key_value_heads * hd, bias=cfg.attention_bias)
        self.o_proj = nn.Linear(self.num_heads * hd, h, bias=False)
        self.rank = lc.RANK
        self.scaling = lc.ALPHA / self.rank                       # :120
        r = round(h / self.rank)                                  # :125
        vo = _kv_out_dim(cfg, lc)
        if lc.IS_TASK_SPECIFIC:                                   # :124-175
            self.lora_down_Q = nn.ModuleDict({t: nn.Linear(h, r, bias=False) for t in TASKS})
            self.lora_down_V = nn.ModuleDict({t: nn.Linear(h, r, bias=False) for t in TASKS})
            self.lora_up_Q = nn.ModuleDict({t: nn.Linear(r, h, bias=False) for t in TASKS})
            self.lora_up_V = nn.ModuleDict({t: nn.Linear(r, vo, bias=False) for t in TASKS})
            if lc.SHARED_LORA:
                self.lora_down_Q_shared = nn.Linear(h, r, bias=False)
                self.lora_down_V_shared = nn.Linear(h, r, bias=False)
                self.lora_up_Q_shared = nn.Linear(r, h, bias=False)
                self.lora_up_V_shared = nn.Linear(r, vo, bias=False)
            for t in TASKS:
                nn.init.zeros_(self.lora_down_Q[t].weight)
                nn.init.kaiming_uniform_(self.lora_up_Q[t].weight, a=math.sqrt(5))
                nn.init.zeros_(self.lora_down_V[t].weight)
                nn.init.kaiming_uniform_(self.lora_up_V[t].weight, a=math.sqrt(5))
            if lc.SHARED_LORA:
                nn.init.zeros_(self.lora_down_Q_shared.weight)
                nn.init.kaiming_uniform_(self.lora_up_Q_shared.weight, a=math.sqrt(5))
                nn.init.zeros_(self.lora_down_V_shared.weight)
                nn.init.kaiming_uniform_(self.lora_up_V_shared.weight, a=math.sqrt(5))
        else:                                                     # :176-192
            self.lora_down_Q = nn.Linear(h, r, bias=False)
            self.lora_down_V = nn.Linear(h, r, bias=False)
            self.lora_up_Q = nn.Linear(r, h, bias=False)
            self.lora_up_V = nn.Linear(r, vo, bias=False)
            nn.init.zeros_(self.lora_down_Q.weight)
            nn.init.kaiming_uniform_(self.lora_up_Q.weight, a=math.sqrt(5))
            nn.init.zeros_(self.lora_down_V.weight)
            nn.init.kaiming_uniform_(self.lora_up_V.weight, a=math.sqrt(5))

    def forward(self, hidden_states, cos, sin, past_kv=None, modality=None):
        lc = self.lora_config
        bsz, q_len, _ = hidden_states.size()
        query_states = self.q_proj(hidden_states)                 # :246-248
        key_states = self.k_proj(hidden_states)
        value_states = self.v_proj(hidden_states)
        if lc.IS_TASK_SPECIFIC:                                   # :250-251 (KeyError if modality is missing)
            Q_lora = self.lora_up_Q[modality](self.lora_down_Q[modality](hidden_states))
            V_lora = self.lora_up_V[modality](self.lora_down_V[modality](hidden_states))
        else:
            Q_lora = self.lora_up_Q(self.lora_down_Q(hidden_states))
            V_lora = self.lora_up_V(self.lora_down_V(hidden_states))
        if lc.SHARED_LORA:                                        # :254-259
            Q_sh = self.lora_up_Q_shared(self.lora_down_Q_shared(hidden_states))
            V_sh = self.lora_up_V_shared(self.lora_down_V_shared(hidden_states))
            query_states = query_states + (Q_lora + Q_sh) * self.scaling
            value_states = value_states + (V_lora + V_sh) * self.scaling
        else:
            query_states = query_states + Q_lora * self.scaling
            value_states = value_states + V_lora * self.scaling
        query_states = query_states.view(bsz, q_len, self.num_heads, self.head_dim).transpose(1, 2)
        key_states = key_states.view(bsz, q_len, self.num_key_value_heads, self.head_dim).transpose(1, 2)
        value_states = value_states.view(bsz, q_len, self.num_key_value_heads, self.head_dim).transpose(1, 2)
        query_states, key_states = apply_rotary_pos_emb(query_states, key_states, cos, sin)   # :277
        if past_kv is not None:                                   # :279-282 DynamicCache.update == cat-append
            if past_kv[self.layer_idx] is not None:
                pk, pv = past_kv[self.layer_idx]
                key_states = torch.cat([pk, key_states], dim=2)
                value_states = torch.cat([pv, value_states], dim=2)
            past_kv[self.layer_idx] = (key_states, value_states)
        k = repeat_kv(key_states, self.num_key_value_groups)      # :284-285
        v = repeat_kv(value_states, self.num_key_value_groups)
        is_causal = q_len > 1                                     # :298 (no mask is ever passed on this path)
        attn = F.scaled_dot_product_attention(query_states, k, v, attn_mask=None, dropout_p=0.0, is_causal=is_causal)
        attn = attn.transpose(1, 2).contiguous().view(bsz, q_len, -1)
        return self.o_proj(attn)                                  # :314


class DecoderLayer_lora(nn.Module):  # Llama_LoRA.py:580-655
    def __init__(self, cfg, lc, layer_idx):
        super().__init__()
        self.self_attn = Attention_lora(cfg, lc, layer_idx)
        self.mlp = MLP(cfg)
        self.input_layernorm = RMSNorm(cfg.hidden_size, cfg.rms_norm_eps)
        self.post_attention_layernorm = RMSNorm(cfg.hidden_size, cfg.rms_norm_eps)

    def forward(self, x, cos, sin, past_kv=None, modality=None):
        residual = x
        x = self.input_layernorm(x)
        x = self.self_attn(x, cos, sin, past_kv, modality)
        x = residual + x
        residual = x
        x = self.post_attention_layernorm(x)
        x = self.mlp(x)
        return residual + x


class Model_lora(nn.Module):  # Llama_LoRA.py:446-578
    def __init__(self, cfg, lc):
        super().__init__()
        self.cfg = cfg
        self.embed_tokens = nn.Embedding(cfg.vocab_size, cfg.hidden_size, cfg.pad_token_id)
        self.layers = nn.ModuleList([DecoderLayer_lora(cfg, lc, i) for i in range(cfg.num_hidden_layers)])
        self.norm = RMSNorm(cfg.hidden_size, cfg.rms_norm_eps)

    def forward(self, input_ids=None, inputs_embeds=None, past_kv=None, modality=None, position_ids=None):
        if (input_ids is None) == (inputs_embeds is None):
            raise ValueError("You cannot specify both input_ids and inputs_embeds at the same time, and must specify either one")
        if inputs_embeds is None:
            inputs_embeds = self.embed_tokens(input_ids)          # :489-490
        past = 0
        if past_kv is not None and past_kv[0] is not None:
            past = past_kv[0][0].shape[2]
        if position_ids is None:                                  # :503-509
            position_ids = torch.arange(past, past + inputs_embeds.shape[1]).unsqueeze(0)
        cos, sin = rope_cos_sin(self.cfg, position_ids, inputs_embeds.dtype)   # :517
        cos, sin = cos.to(inputs_embeds.device), sin.to(inputs_embeds.device)  # (tables built on the host; GPU-eager runs)
        h = inputs_embeds
        for layer in self.layers:
            h = layer(h, cos, sin, past_kv, modality)
        return self.norm(h)


@dataclass
class CausalLMOutput:
    loss: Optional[torch.Tensor]
    logits: torch.Tensor


class ForCausalLM_lora(nn.Module):  # Llama_LoRA.py:318-444 / Qwen_LoRA.py:105-251
    def __init__(self, cfg: LLMConfig, lora_config):
        super().__init__()
        self.config, self.lora_config = cfg, lora_config
        self.model = Model_lora(cfg, lora_config)
        self.lm_head = nn.Linear(cfg.hidden_size, cfg.vocab_size, bias=False)
        if cfg.tie_word_embeddings:
            self.lm_head.weight = self.model.embed_tokens.weight

    def forward(self, input_ids=None, inputs_embeds=None, labels=None, past_kv=None, modality=None, position_ids=None):
        h = self.model(input_ids, inputs_embeds, past_kv, modality, position_ids)
        logits = self.lm_head(h)                                  # :372
        logits = logits.float()                                   # :373
        loss = None
        if labels is not None:                                    # :376-386
            shift_logits = logits[..., :-1, :].contiguous()
            shift_labels = labels[..., 1:].contiguous()
            loss = nn.CrossEntropyLoss()(shift_logits.view(-1, self.config.vocab_size), shift_labels.view(-1))
        return CausalLMOutput(loss, logits)

    @torch.no_grad()
    def generate(self, inputs_embeds, max_new_tokens, eos_token_id, pad_token_id, modality=None, num_beams=1,
                 bos_token_id=None, return_margins=False):
        """Greedy branch of HF GenerationMixin.generate as used at modeling_OmniAVSR.py:313-322:
        inputs_embeds only => returns only the new tokens; step 0 consumes the embeddings, later steps embed the
        previous token (Llama_LoRA.py:429-432); finished rows are padded with pad_token_id; stops when all rows are
        finished or after max_new_tokens."""
        if num_beams != 1:
            return self.beam_generate(inputs_embeds, max_new_tokens, num_beams, eos_token_id, pad_token_id, modality)
        B = inputs_embeds.shape[0]
        past = [None] * self.config.num_hidden_layers
        unfinished = torch.ones(B, dtype=torch.long)
        out, margins = [], []
        cur = dict(inputs_embeds=inputs_embeds)
        for _ in range(max_new_tokens):
            logits = self.forward(past_kv=past, modality=modality, **cur).logits[:, -1, :]
            nxt = torch.argmax(logits.float(), dim=-1)
            top2 = logits.float().topk(2, dim=-1).values
            margins.append((top2[:, 0] - top2[:, 1]) / logits.float().abs().max(dim=-1).values)
            nxt = nxt * unfinished + pad_token_id * (1 - unfinished)
            out.append(nxt)
            unfinished = unfinished * (nxt != eos_token_id).long()
            if unfinished.max() == 0:
                break
            cur = dict(input_ids=nxt[:, None])
        if return_margins:
            return torch.stack(out, dim=1), torch.stack(margins, dim=1)
        return torch.stack(out, dim=1)


def _beam_generate(self, inputs_embeds, max_new_tokens, num_beams, eos_token_id, pad_token_id, modality=None,
                   return_scores=False):
    """Beam branch of HF generate (oracle/beam_search.py drives it): prompt expanded to B*K rows, cache reordered by
    beam index every step."""
    from .beam_search import beam_search
    B, K = inputs_embeds.shape[0], num_beams
    past = [None] * self.config.num_hidden_layers
    expanded = inputs_embeds.repeat_interleave(K, dim=0)

    def step_logits(tokens):
        cur = dict(inputs_embeds=expanded) if tokens is None else dict(input_ids=tokens[:, None])
        return self.forward(past_kv=past, modality=modality, **cur).logits[:, -1, :]

    def reorder(beam_idx):
        for i, kv in enumerate(past):
            past[i] = (kv[0][beam_idx], kv[1][beam_idx])

    return beam_search(step_logits, reorder, B, K, max_new_tokens, eos_token_id, pad_token_id, return_scores=return_scores)


ForCausalLM_lora.beam_generate = _beam_generate


def make_lora_config(cfg: LLMConfig, name: str, rank: int, alpha: int, task_specific: bool, shared: bool):
    """lightning_OmniAVSR.py:99-113."""
    if "Qwen" in name:
        return QwenLoRA_config(rank, alpha, name == "Qwen/Qwen2.5-0.5B", name == "Qwen/Qwen2.5-1.5B",
                               name == "Qwen/Qwen2.5-3B", name == "Qwen/Qwen2.5-7B", name == "Qwen/Qwen2.5-14B",
                               name == "Qwen/Qwen2.5-32B", task_specific, shared)
    is_l3 = name in ("meta-llama/Meta-Llama-3-8B", "meta-llama/Meta-Llama-3.1-8B", "meta-llama/Llama-3.2-1B")
    return LoRA_config(rank, alpha, is_l3, name == "meta-llama/Llama-3.2-3B", task_specific, shared)
