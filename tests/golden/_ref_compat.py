"""Import shims that let the UNMODIFIED reference sources under /root/reference be imported and executed in this
container, for golden-vector generation only (tests/golden/make_reference_golden.py).  Nothing here restates reference
arithmetic: the shims only restore *names* that the reference's pinned dependencies had and this image's do not.

* transformers 4.43.1 -> 5.5.0: `LlamaSdpaAttention` / `Qwen2SdpaAttention` (= the attention module with the 4.43.1
  attribute names `num_heads`, `num_key_value_heads`, `hidden_size`), `LlamaModel._update_causal_mask` (returns None
  for the mask-free SDPA path the reference's training forward takes, as 4.43.1's
  `AttentionMaskConverter._ignore_causal_mask_sdpa` does for attention_mask=None), `config.use_return_dict`,
  `DynamicCache.from_legacy_cache/get_usable_length`; for Qwen2 the 4.43.1 `Qwen2RotaryEmbedding` (cached cos/sin tables)
  and 5-argument `apply_rotary_pos_emb`, restated from that release because the reference's attention calls them with
  signatures 5.5.0 no longer has (this is dependency code, not reference code).
* fairseq / omegaconf / hydra: empty stub modules so `modeling_OmniAVSR.py`'s top-level imports succeed; the stubs that
  `multihead_attention.py` needs (`FairseqDropout`, `quant_noise`, `with_incremental_state`, `utils.softmax`) are the
  identity / thin torch calls they are in eval mode.

Used on this container only; the GPU box never imports it (no /root/reference there)."""
import importlib.util
import os
import sys
import types

import torch
import torch.nn.functional as F
from torch import nn

REF = "/root/reference"


def available():
    return os.path.isdir(os.path.join(REF, "Omni_AVSR"))


def _stub(name, **attrs):
    m = sys.modules.get(name)
    if m is None:
        m = types.ModuleType(name)
        m.__path__ = []          # behave like a package so sub-imports resolve through sys.modules
        sys.modules[name] = m
    for k, v in attrs.items():
        setattr(m, k, v)
    return m


def install_transformers_shims():
    import transformers.models.llama.modeling_llama as ml
    import transformers.models.qwen2.modeling_qwen2 as mq

    def make(base):
        class _SdpaAttention(base):
            def __init__(self, config, layer_idx=None):
                super().__init__(config, layer_idx)
                self.hidden_size = config.hidden_size
                self.num_heads = config.num_attention_heads
                self.num_key_value_heads = config.num_key_value_heads
                self.num_key_value_groups = self.num_heads // self.num_key_value_heads
        return _SdpaAttention

    if not hasattr(ml, "LlamaSdpaAttention"):
        ml.LlamaSdpaAttention = make(ml.LlamaAttention)
    if not hasattr(mq, "Qwen2SdpaAttention"):
        base = make(mq.Qwen2Attention)

        class _Qwen2SdpaAttention(base):
            def __init__(self, config, layer_idx=None):
                super().__init__(config, layer_idx)
                rope_theta = getattr(config, "rope_theta", None) or config.rope_parameters["rope_theta"]
                self.rotary_emb = Qwen2RotaryEmbedding443(self.head_dim, config.max_position_embeddings, rope_theta)
        mq.Qwen2SdpaAttention = _Qwen2SdpaAttention

    def _update_causal_mask(self, attention_mask, input_tensor, cache_position, past_key_values, output_attentions):
        assert attention_mask is None, "shim covers the mask-free (training / unpadded) path only"
        return None

    from transformers.cache_utils import DynamicCache
    if not hasattr(DynamicCache, "get_usable_length"):      # 4.43.1 DynamicCache: == get_seq_length(layer_idx)
        DynamicCache.get_usable_length = lambda self, new_seq_length, layer_idx=0: self.get_seq_length(layer_idx)
    if not hasattr(DynamicCache, "from_legacy_cache"):      # 4.43.1: from_legacy_cache(None) == empty cache
        def from_legacy_cache(cls, past_key_values=None):
            assert past_key_values is None
            return cls()
        DynamicCache.from_legacy_cache = classmethod(from_legacy_cache)
        DynamicCache.to_legacy_cache = lambda self: self

    for cls in (ml.LlamaModel, mq.Qwen2Model):
        if not hasattr(cls, "_update_causal_mask"):
            cls._update_causal_mask = _update_causal_mask
        if not hasattr(cls, "_attn_implementation"):
            cls._attn_implementation = "sdpa"


class Qwen2RotaryEmbedding443(nn.Module):
    """transformers==4.43.1 `Qwen2RotaryEmbedding` (un-vendored dependency of Qwen_LoRA.py:579-581), published algorithm:
    cos/sin tables for `max_position_embeddings` positions built at construction from the fp32 inv_freq in the default
    dtype and kept as buffers (so a later `.bfloat16()` rounds the TABLES); forward slices `[:seq_len]`."""

    def __init__(self, dim, max_position_embeddings=2048, base=10000):
        super().__init__()
        self.dim, self.base = dim, base
        inv_freq = 1.0 / (base ** (torch.arange(0, dim, 2, dtype=torch.int64).float() / dim))
        self.register_buffer("inv_freq", inv_freq, persistent=False)
        self._set(max_position_embeddings, torch.get_default_dtype())

    def _set(self, seq_len, dtype):
        self.max_seq_len_cached = seq_len
        t = torch.arange(seq_len, dtype=torch.int64).type_as(self.inv_freq)
        emb = torch.outer(t, self.inv_freq)
        emb = torch.cat((emb, emb), dim=-1)
        self.register_buffer("cos_cached", emb.cos().to(dtype), persistent=False)
        self.register_buffer("sin_cached", emb.sin().to(dtype), persistent=False)

    def forward(self, x, seq_len=None):
        assert seq_len <= self.max_seq_len_cached
        return self.cos_cached[:seq_len].to(dtype=x.dtype), self.sin_cached[:seq_len].to(dtype=x.dtype)


def apply_rotary_pos_emb_443(q, k, cos, sin, position_ids, unsqueeze_dim=1):
    """transformers==4.43.1 qwen2 `apply_rotary_pos_emb` (5-argument form the reference calls, Qwen_LoRA.py:583)."""
    def rotate_half(x):
        x1, x2 = x[..., : x.shape[-1] // 2], x[..., x.shape[-1] // 2:]
        return torch.cat((-x2, x1), dim=-1)
    cos = cos[position_ids].unsqueeze(unsqueeze_dim)
    sin = sin[position_ids].unsqueeze(unsqueeze_dim)
    return (q * cos) + (rotate_half(q) * sin), (k * cos) + (rotate_half(k) * sin)


def install_fairseq_stubs():
    class FairseqDropout(nn.Module):
        def __init__(self, p, module_name=None):
            super().__init__()
            self.p = p

        def forward(self, x, inplace=False):
            return F.dropout(x, p=self.p, training=self.training, inplace=inplace) if self.p > 0 and self.training else x

    def quant_noise(module, p, block_size):
        assert p <= 0
        return module

    def with_incremental_state(cls):
        return cls

    utils = _stub("fairseq.utils", softmax=lambda x, dim, onnx_trace=False: F.softmax(x, dim=dim, dtype=torch.float32))
    _stub("fairseq", utils=utils, checkpoint_utils=types.SimpleNamespace(load_model_ensemble_and_task=None))
    _stub("fairseq.incremental_decoding_utils", with_incremental_state=with_incremental_state)
    _stub("fairseq.modules")
    _stub("fairseq.modules.fairseq_dropout", FairseqDropout=FairseqDropout)
    _stub("fairseq.modules.quant_noise", quant_noise=quant_noise)
    _stub("av_hubert")
    _stub("av_hubert.avhubert")
    _stub("av_hubert.avhubert.hubert_asr", AVHubertSeq2Seq=None, AVHubertSeq2SeqConfig=None)
    _stub("av_hubert.avhubert.hubert_lora", AVHubertModel_lora=None)


def load_by_path(name, relpath, package=None):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, relpath))
    mod = importlib.util.module_from_spec(spec)
    if package:
        mod.__package__ = package
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def import_reference():
    """Returns (Llama_LoRA, Qwen_LoRA, modeling_OmniAVSR, multihead_attention) — the reference's own modules."""
    install_transformers_shims()
    install_fairseq_stubs()
    pkg = _stub("Omni_AVSR")
    pkg.__path__ = [os.path.join(REF, "Omni_AVSR")]
    ll = load_by_path("Omni_AVSR.Llama_LoRA", "Omni_AVSR/Llama_LoRA.py", "Omni_AVSR")
    ql = load_by_path("Omni_AVSR.Qwen_LoRA", "Omni_AVSR/Qwen_LoRA.py", "Omni_AVSR")
    ql.apply_rotary_pos_emb = apply_rotary_pos_emb_443
    mo = load_by_path("Omni_AVSR.modeling_OmniAVSR", "Omni_AVSR/modeling_OmniAVSR.py", "Omni_AVSR")
    mha = load_by_path("ref_multihead_attention", "av_hubert/fairseq/fairseq/modules/multihead_attention.py")
    return ll, ql, mo, mha
