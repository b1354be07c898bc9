"""Generates tests/golden/omni_golden.pt from the CPU oracle (seeded).  The reference ships no golden vectors for this
path and cannot be imported in this image, so these fixtures freeze the oracle restatement itself:
  python tests/golden/make_golden.py
Re-generation must be bit-identical (tests/test_oracle_golden.py checks that on CPU; the GPU tests check the CUDA
path against the same file)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import llm_lora as ol  # noqa: E402
from oracle import matryoshka as om  # noqa: E402


def tiny_llm(seed=11):
    torch.manual_seed(seed)
    cfg = ol.LLMConfig("llama", 256, 512, 2, 4, 1, 300, 1e-5, 500000.0, 64,
                       dict(factor=32.0, low_freq_factor=1.0, high_freq_factor=4.0, original_max_position_embeddings=8192),
                       False, True, inv_freq_dtype="bf16")
    lc = ol.LoRA_config(4, 2, True, False, True, True)
    m = ol.ForCausalLM_lora(cfg, lc)
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for n, p in m.named_parameters():
            if p.dim() == 2:
                p.copy_(torch.randn(p.shape, generator=g) * (0.05 if "lora" in n else 0.02))
    return m.bfloat16().eval(), cfg


def build():
    out = {}
    g = torch.Generator().manual_seed(1)
    x = torch.randn(2, 40, 64, generator=g).bfloat16()
    out["compress_x"] = x
    for rate in (4, 5):
        for mode in ("avg-pooling", "stack"):
            out[f"compress_{mode}_{rate}"] = om.compress(x[:, :33], rate, mode)
    V, H, B, L = 300, 64, 2, 7
    embed = torch.nn.Embedding(V, H)
    embed.weight.data = torch.randn(V, H, generator=g).bfloat16()
    tokens = torch.randint(0, V - 8, (B, L), generator=g)
    labels = tokens.clone()
    labels[0, -1] = -100
    a = torch.randn(B, 5, H, generator=g).bfloat16()
    v = torch.randn(B, 3, H, generator=g).bfloat16()
    prompts = {k: torch.randn(1, p, H, generator=g).bfloat16() for k, p in (("audio", 6), ("video", 6), ("audiovisual", 8))}
    marker = (V - 4, V - 3, V - 2, V - 1)
    out.update(splice_embed=embed.weight.data, splice_tokens=tokens, splice_labels=labels, splice_a=a, splice_v=v,
               splice_prompts=prompts, splice_marker=marker)
    with torch.no_grad():
        for is_qwen in (False, True):
            seqs, labs = om.build_train_sequences(embed, tokens, labels, a, v, prompts, marker, is_qwen)
            out[f"splice_seqs_qwen{int(is_qwen)}"] = seqs
            out[f"splice_labs_qwen{int(is_qwen)}"] = labs
    m, cfg = tiny_llm()
    xe = (torch.randn(2, 9, cfg.hidden_size, generator=g) * 0.5).bfloat16()
    lab = torch.randint(0, cfg.vocab_size, (2, 9), generator=g)
    lab[:, :4] = -100
    out["llm_x"], out["llm_labels"] = xe, lab
    with torch.no_grad():
        for t in ol.TASKS:
            o = m(inputs_embeds=xe, labels=lab, modality=t)
            out[f"llm_logits_{t}"] = o.logits
            out[f"llm_loss_{t}"] = o.loss
    out["token_rule"] = {n: om.num_audio_tokens(torch.tensor(n)) for n in (256000, 255999, 7999, 16000, 160000)}
    return out


if __name__ == "__main__":
    torch.save(build(), os.path.join(os.path.dirname(os.path.abspath(__file__)), "omni_golden.pt"))
    print("written")
