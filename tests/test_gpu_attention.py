"""GPU parity of the tcgen05 flash-attention forward against torch SDPA (math reference in fp32) on the packed q|k|v
layout: causal GQA (LLM), non-causal (Whisper 1500 keys, AV-HuBERT 400 keys), ragged lengths.
Tolerance: max|a-b| <= 2e-2 * max|b| (bf16 probabilities and outputs)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _ref(qkv, B, S, row0, nh, nkv, hd, causal):
    blk = qkv[row0: row0 + B * S].float()
    q = blk[:, : nh * hd].view(B, S, nh, hd).transpose(1, 2)
    k = blk[:, nh * hd: (nh + nkv) * hd].view(B, S, nkv, hd).transpose(1, 2)
    v = blk[:, (nh + nkv) * hd:].view(B, S, nkv, hd).transpose(1, 2)
    rep = nh // nkv
    k = k.repeat_interleave(rep, dim=1)
    v = v.repeat_interleave(rep, dim=1)
    s = (q @ k.transpose(-1, -2)) * hd ** -0.5
    if causal:
        s = s.masked_fill(torch.ones(S, S, dtype=torch.bool, device=s.device).triu(1), float("-inf"))
    p = torch.softmax(s, dim=-1)
    o = p @ v
    lse = torch.logsumexp(s, dim=-1)                       # [B, nh, S]
    return o.transpose(1, 2).reshape(B * S, nh * hd), lse


@pytest.mark.parametrize("B,S,nh,nkv,causal,hd", [(2, 128, 4, 4, False, 64), (2, 256, 8, 2, True, 64),
                                                  (3, 460, 32, 8, True, 64), (2, 1500, 16, 16, False, 64),
                                                  (2, 400, 16, 16, False, 64), (1, 57, 4, 1, True, 64),
                                                  (2, 190, 32, 8, True, 64), (2, 412, 16, 2, True, 128),
                                                  (2, 300, 8, 8, False, 128), (1, 129, 32, 8, True, 128)])
def test_attention_forward(B, S, nh, nkv, causal, hd):
    from omni_avsr_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(S + nh)
    row0 = 128
    M = row0 + B * S + 70
    qkv = torch.randn(M, (nh + 2 * nkv) * hd, device="cuda", generator=g).bfloat16()
    out = torch.zeros(M, nh * hd, device="cuda", dtype=torch.bfloat16)
    lse = torch.zeros(nh, M, device="cuda", dtype=torch.float32)
    ops.attention_fwd(qkv, out, [(0, B, S, row0)], nh, nkv, hd, causal, lse=lse)
    want, want_lse = _ref(qkv, B, S, row0, nh, nkv, hd, causal)
    got = out[row0: row0 + B * S].float()
    err = (got - want).abs().max().item()
    assert err <= 2e-2 * want.abs().max().item(), err
    assert out[:row0].abs().max().item() == 0 and out[row0 + B * S:].abs().max().item() == 0
    got_lse = lse[:, row0: row0 + B * S].view(nh, B, S).transpose(0, 1)
    assert (got_lse - want_lse).abs().max().item() <= 2e-2
