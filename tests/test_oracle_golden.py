"""CPU: the oracle reproduces the committed golden fixtures bit for bit, plus the structural anchors the reference
does determine: avg-pool == mean over r of the truncated rows, stack == reshape, label prefix lengths,
the token-count rule max(int(L/16000*50), 25)."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import make_golden  # noqa: E402

GOLD = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "omni_golden.pt"), weights_only=False)


def _same(a, b):
    if isinstance(a, dict):
        return all(_same(a[k], b[k]) for k in a)
    if torch.is_tensor(a):
        if a.dtype == torch.bfloat16:
            return torch.equal(a.view(torch.int16), b.view(torch.int16))
        return torch.equal(a, b)
    return a == b


def test_regeneration_is_bit_identical():
    new = make_golden.build()
    assert set(new) == set(GOLD)
    for k in GOLD:
        if k.startswith("llm_"):
            continue            # the LLM part depends on CPU matmul blocking: checked with a tolerance below
        assert _same(GOLD[k], new[k]), k
    for t in ("audio", "video", "audiovisual"):
        assert torch.allclose(new[f"llm_logits_{t}"], GOLD[f"llm_logits_{t}"], atol=2e-2, rtol=2e-2)


def test_structural_anchors():
    x = GOLD["compress_x"]
    for rate in (4, 5):
        n = 33 // rate
        mean = x[:, : n * rate].float().view(2, n, rate, 64).sum(2) / rate
        assert torch.equal(GOLD[f"compress_avg-pooling_{rate}"].view(torch.int16), mean.bfloat16().view(torch.int16))
        assert torch.equal(GOLD[f"compress_stack_{rate}"].view(torch.int16),
                           x[:, : n * rate].reshape(2, n, rate * 64).view(torch.int16))
    labs = GOLD["splice_labs_qwen0"]
    assert (labs["audio"][:, 1: 1 + 6 + 7] == -100).all() and labs["audio"].shape[1] == 1 + 7 + 6 + 6
    assert labs["audiovisual"].shape[1] == 1 + 7 + 5 + 8 + 6
    assert GOLD["token_rule"] == {256000: 800, 255999: 799, 7999: 25, 16000: 50, 160000: 500}
