"""Pins the LLM oracle: (1) against the installed transformers LlamaForCausalLM / Qwen2ForCausalLM with the
adapters off (zero-init lora_down => adapted == base, reference Llama_LoRA.py:166-175), (2) the LoRA math against
a literal re-evaluation of reference lines Llama_LoRA.py:246-259, (3) greedy decode against HF generate."""
import copy

import pytest
import torch

from oracle import llm_lora as ol

transformers = pytest.importorskip("transformers")


def _tiny(family):
    if family == "llama":
        cfg = ol.LLMConfig("llama", 64, 128, 2, 4, 1, 97, 1e-5, 500000.0, 16,
                           dict(factor=32.0, low_freq_factor=1.0, high_freq_factor=4.0,
                                original_max_position_embeddings=8192), False, True, inv_freq_dtype="fp32")
        hf_cfg = transformers.LlamaConfig(
            hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=1,
            vocab_size=97, rms_norm_eps=1e-5, head_dim=16, tie_word_embeddings=True, max_position_embeddings=131072,
            rope_parameters=dict(rope_type="llama3", rope_theta=500000.0, factor=32.0, low_freq_factor=1.0,
                                 high_freq_factor=4.0, original_max_position_embeddings=8192),
            attention_bias=False, mlp_bias=False)
        hf = transformers.LlamaForCausalLM(hf_cfg)
        lc = ol.LoRA_config(RANK=8, ALPHA=4, IS_LLAMA3=True, IS_TASK_SPECIFIC=True, SHARED_LORA=True)
    else:
        cfg = ol.LLMConfig("qwen2", 128, 256, 2, 8, 1, 97, 1e-6, 1000000.0, 16, None, True, True,
                           max_position_embeddings=32768, inv_freq_dtype="fp32")
        hf_cfg = transformers.Qwen2Config(
            hidden_size=128, intermediate_size=256, num_hidden_layers=2, num_attention_heads=8, num_key_value_heads=1,
            vocab_size=97, rms_norm_eps=1e-6, tie_word_embeddings=True, max_position_embeddings=32768,
            rope_parameters=dict(rope_type="default", rope_theta=1000000.0))
        hf = transformers.Qwen2ForCausalLM(hf_cfg)
        lc = ol.QwenLoRA_config(RANK=8, ALPHA=4, IS_QWEN25_3B=True, IS_TASK_SPECIFIC=False, SHARED_LORA=False)
    hf.eval()
    torch.manual_seed(0)
    model = ol.ForCausalLM_lora(cfg, lc)
    missing, unexpected = model.load_state_dict(hf.state_dict(), strict=False)
    assert not unexpected
    assert all("lora_" in k for k in missing), missing
    return cfg, hf, model


@pytest.mark.parametrize("family", ["llama", "qwen2"])
def test_base_arithmetic_matches_transformers(family):
    cfg, hf, model = _tiny(family)
    torch.manual_seed(1)
    x = torch.randn(2, 11, cfg.hidden_size)
    labels = torch.randint(0, cfg.vocab_size, (2, 11))
    labels[:, :4] = -100
    with torch.no_grad():
        want = hf(inputs_embeds=x, labels=labels)
        got = model(inputs_embeds=x, labels=labels, modality="audio")
    assert torch.allclose(got.logits, want.logits.float(), atol=2e-5, rtol=1e-5)
    assert abs(got.loss.item() - want.loss.item()) < 1e-5


def test_llama3_inv_freq_matches_transformers():
    cfg = ol.llama_3_2_1b()
    from transformers.modeling_rope_utils import ROPE_INIT_FUNCTIONS
    hf_cfg = transformers.LlamaConfig(hidden_size=2048, num_attention_heads=32, head_dim=64,
                                      rope_parameters=dict(rope_type="llama3", rope_theta=500000.0, factor=32.0,
                                                           low_freq_factor=1.0, high_freq_factor=4.0,
                                                           original_max_position_embeddings=8192))
    want, _ = ROPE_INIT_FUNCTIONS["llama3"](hf_cfg, "cpu")
    assert torch.equal(ol.compute_inv_freq(cfg), want)


def test_lora_math_literal():
    """Evaluate Llama_LoRA.py:246-259 literally on the oracle's own weights and compare q/v pre-RoPE."""
    cfg, hf, model = _tiny("llama")
    att = model.model.layers[0].self_attn
    torch.manual_seed(2)
    for p in att.parameters():
        if p.dim() == 2:
            p.data = torch.randn_like(p) * 0.05
    model = model.bfloat16()
    att = model.model.layers[0].self_attn
    x = torch.randn(2, 7, cfg.hidden_size).bfloat16()
    s = att.scaling
    assert s == 4 / 8
    for t in ol.TASKS:
        q = att.q_proj(x) + (att.lora_up_Q[t](att.lora_down_Q[t](x)) + att.lora_up_Q_shared(att.lora_down_Q_shared(x))) * s
        v = att.v_proj(x) + (att.lora_up_V[t](att.lora_down_V[t](x)) + att.lora_up_V_shared(att.lora_down_V_shared(x))) * s
        captured = {}

        def grab(q_, k_, v_, **kw):
            captured["v"] = v_
            return torch.zeros_like(q_)
        orig = ol.F.scaled_dot_product_attention
        ol.F.scaled_dot_product_attention = grab
        try:
            pos = torch.arange(7).unsqueeze(0)
            cos, sin = ol.rope_cos_sin(cfg, pos, torch.bfloat16)
            att(x, cos, sin, None, t)
        finally:
            ol.F.scaled_dot_product_attention = orig
        v_heads = v.view(2, 7, 1, 16).transpose(1, 2)
        assert torch.equal(captured["v"][:, :1], v_heads)   # repeat_kv duplicates the kv head


def test_task_specific_requires_modality():
    cfg, hf, model = _tiny("llama")
    with pytest.raises(KeyError):
        model(inputs_embeds=torch.randn(1, 3, cfg.hidden_size))


@pytest.mark.parametrize("family", ["llama", "qwen2"])
def test_greedy_matches_transformers_generate(family):
    cfg, hf, model = _tiny(family)
    torch.manual_seed(3)
    x = torch.randn(3, 9, cfg.hidden_size)
    eos, pad = 5, 7
    with torch.no_grad():
        want = hf.generate(inputs_embeds=x, max_new_tokens=12, num_beams=1, do_sample=False, eos_token_id=eos,
                           pad_token_id=pad)
    got = model.generate(x, 12, eos, pad, modality="audio")
    assert torch.equal(got, want[:, : got.shape[1]])
    assert got.shape[1] == want.shape[1]
