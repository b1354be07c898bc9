"""CPU: pins oracle/beam_search.py (restatement of transformers==4.43.1 `_beam_search` + `BeamSearchScorer`, the algorithm
behind the reference's `generate(num_beams=15)` call at modeling_OmniAVSR.py:313-322) against the installed transformers'
own `generate(num_beams=K)` on small random Llama models driven through `inputs_embeds` only, with an EOS token that is
likely enough to close hypotheses at different lengths."""
import pytest
import torch

from oracle.beam_search import beam_search


def _tiny(seed, vocab=24):
    from transformers import LlamaConfig, LlamaForCausalLM
    torch.manual_seed(seed)
    cfg = LlamaConfig(vocab_size=vocab, hidden_size=32, intermediate_size=64, num_hidden_layers=2, num_attention_heads=4,
                      num_key_value_heads=2, max_position_embeddings=128, tie_word_embeddings=True)
    m = LlamaForCausalLM(cfg).eval()
    with torch.no_grad():
        for p in m.parameters():
            p.mul_(4.0)            # sharper distributions: beams separate, EOS shows up in the top-2K
    return m


@pytest.mark.parametrize("seed,B,K,max_new", [(0, 1, 4, 12), (1, 2, 3, 10), (2, 3, 5, 16), (3, 1, 15, 32), (4, 2, 2, 6),
                                              (5, 4, 4, 9), (6, 1, 6, 20)])
def test_oracle_beam_search_matches_transformers_generate(seed, B, K, max_new):
    m = _tiny(seed)
    eos, pad = 3, 2
    g = torch.Generator().manual_seed(100 + seed)
    emb = torch.randn(B, 5, 32, generator=g)
    with torch.no_grad():
        want = m.generate(inputs_embeds=emb, max_new_tokens=max_new, num_beams=K, do_sample=False, eos_token_id=eos,
                          pad_token_id=pad, bos_token_id=1, early_stopping=False, length_penalty=1.0)
    state = {"emb": emb.repeat_interleave(K, dim=0)}

    def step_logits(tokens):
        if tokens is not None:
            state["emb"] = torch.cat([state["emb"], m.get_input_embeddings()(tokens)[:, None]], dim=1)
        with torch.no_grad():
            return m(inputs_embeds=state["emb"]).logits[:, -1, :]

    def reorder(beam_idx):
        state["emb"] = state["emb"][beam_idx]

    got = beam_search(step_logits, reorder, B, K, max_new, eos, pad)
    assert got.shape == want.shape, (got, want)
    assert torch.equal(got, want), (got, want)


def test_oracle_llm_beam_generate_runs_and_greedy_is_k1_consistent():
    """ForCausalLM_lora.generate(num_beams=K) goes through the beam driver; with K = 1 beam search returns the greedy
    continuation up to the first EOS (HF semantics: same tokens, EOS kept, nothing after it)."""
    from oracle import llm_lora as ol
    torch.manual_seed(0)
    cfg = ol.LLMConfig("llama", 64, 128, 2, 4, 1, 40, 1e-5, 500000.0, 16, None, False, True)
    lc = ol.LoRA_config(4, 2, True, False, True, True)
    model = ol.ForCausalLM_lora(cfg, lc)
    emb = torch.randn(2, 6, 64)
    greedy = model.generate(emb, 8, eos_token_id=5, pad_token_id=0, modality="audio")
    beam1 = model.generate(emb, 8, eos_token_id=5, pad_token_id=0, modality="audio", num_beams=1)
    assert torch.equal(greedy, beam1)
    out = model.generate(emb, 8, eos_token_id=5, pad_token_id=0, modality="audiovisual", num_beams=3)
    assert out.shape[0] == 2 and 1 <= out.shape[1] <= 8
