"""GPU parity: tcgen05 flash-attention forward / backward (csrc/attention.cu, csrc/attention_bwd.cu) vs an fp32 torch
restatement of the same op on the same packed q|k|v rows.  Tolerance: max|a-b| <= 2e-2 * max|b| (bf16 P / dS operands,
fp32 accumulation)."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _split(qkv, B, S, row0, nh, nkv, hd):
    blk = qkv[row0: row0 + B * S].float()
    q = blk[:, : nh * hd].view(B, S, nh, hd).transpose(1, 2)
    k = blk[:, nh * hd: (nh + nkv) * hd].view(B, S, nkv, hd).transpose(1, 2)
    v = blk[:, (nh + nkv) * hd:].view(B, S, nkv, hd).transpose(1, 2)
    return q, k, v


def _ref(qkv, B, S, row0, nh, nkv, hd, causal):
    q, k, v = _split(qkv, B, S, row0, nh, nkv, hd)
    g = nh // nkv
    k = k.repeat_interleave(g, dim=1)
    v = v.repeat_interleave(g, dim=1)
    s = (q @ k.transpose(-1, -2)) / math.sqrt(hd)
    if causal:
        s = s.masked_fill(torch.ones(S, S, device=s.device, dtype=torch.bool).triu(1), float("-inf"))
    lse = torch.logsumexp(s, dim=-1)                    # [B, nh, S]
    o = torch.softmax(s, dim=-1) @ v
    return o.transpose(1, 2).reshape(B * S, nh * hd), lse


CASES = [(2, 128, 4, 4, False, 64), (2, 256, 8, 2, True, 64), (3, 460, 32, 8, True, 64), (2, 1500, 16, 16, False, 64),
         (2, 400, 16, 16, False, 64), (1, 57, 4, 1, True, 64), (2, 190, 32, 8, True, 64), (2, 412, 16, 2, True, 128),
         (2, 300, 8, 8, False, 128), (1, 129, 32, 8, True, 128)]


def _make(B, S, nh, nkv, hd):
    g = torch.Generator(device="cuda").manual_seed(S + nh)
    row0 = 128
    M = row0 + B * S + 70
    qkv = (torch.randn(M, (nh + 2 * nkv) * hd, device="cuda", generator=g) * 1.5).to(torch.bfloat16)
    return g, row0, M, qkv


@pytest.mark.parametrize("B,S,nh,nkv,causal,hd", CASES)
def test_attention_forward(B, S, nh, nkv, causal, hd):
    from omni_avsr_b200 import ops
    g, row0, M, qkv = _make(B, S, nh, nkv, hd)
    out = torch.full((M, nh * hd), 7.0, device="cuda", dtype=torch.bfloat16)
    lse = torch.zeros(nh, M, device="cuda")
    ops.attention_fwd(qkv, out, [(0, B, S, row0)], nh, nkv, hd, causal, lse=lse)
    torch.cuda.synchronize()
    want, want_lse = _ref(qkv, B, S, row0, nh, nkv, hd, causal)
    got = out[row0: row0 + B * S].float()
    err = (got - want).abs().max().item()
    assert err <= 2e-2 * max(1.0, want.abs().max().item()), err
    # rows outside the segment are untouched
    assert (out[:row0] == 7.0).all() and (out[row0 + B * S:] == 7.0).all()
    got_lse = lse[:, row0: row0 + B * S].view(nh, B, S).permute(1, 0, 2)
    assert (got_lse - want_lse).abs().max().item() <= 2e-2


@pytest.mark.parametrize("B,S,nh,nkv,causal,hd", CASES)
def test_attention_backward(B, S, nh, nkv, causal, hd):
    """dQ | dK | dV of the packed rows against autograd through the fp32 restatement."""
    from omni_avsr_b200 import ops
    g, row0, M, qkv = _make(B, S, nh, nkv, hd)
    out = torch.zeros((M, nh * hd), device="cuda", dtype=torch.bfloat16)
    lse = torch.zeros(nh, M, device="cuda")
    ops.attention_fwd(qkv, out, [(0, B, S, row0)], nh, nkv, hd, causal, lse=lse)
    dout = torch.randn(M, nh * hd, device="cuda", generator=g).to(torch.bfloat16)
    dqkv = torch.full_like(qkv, 3.0)
    ops.attention_bwd(qkv, out, dout, lse, dqkv, [(0, B, S, row0)], nh, nkv, hd, causal)
    torch.cuda.synchronize()
    x = qkv.float().requires_grad_(True)
    want_o, _ = _ref(x, B, S, row0, nh, nkv, hd, causal)
    want_o.backward(dout[row0: row0 + B * S].float())
    want = x.grad[row0: row0 + B * S]
    got = dqkv[row0: row0 + B * S].float()
    assert torch.isfinite(got).all()
    for name, lo, hi in (("dq", 0, nh * hd), ("dk", nh * hd, (nh + nkv) * hd), ("dv", (nh + nkv) * hd, (nh + 2 * nkv) * hd)):
        a, b = got[:, lo:hi], want[:, lo:hi]
        err = (a - b).abs().max().item()
        assert err <= 2e-2 * max(1.0, b.abs().max().item()), (name, err, b.abs().max().item())
    assert (dqkv[:row0] == 3.0).all() and (dqkv[row0 + B * S:] == 3.0).all()


def test_packed_sdpa_autograd_matches_reference():
    """PackedSdpaFn (the call the model makes) over two segments: gradient of the packed rows end to end."""
    from omni_avsr_b200.Llama_LoRA import PackedSdpaFn
    nh, nkv, hd = 8, 2, 64
    segs = [(0, 2, 100, 0), (1, 2, 140, 256)]
    M = 640
    g = torch.Generator(device="cuda").manual_seed(5)
    qkv = torch.randn(M, (nh + 2 * nkv) * hd, device="cuda", generator=g).to(torch.bfloat16).requires_grad_(True)
    out = PackedSdpaFn.apply(qkv, segs, nh, nkv, hd, True)
    w = torch.randn(M, nh * hd, device="cuda", generator=g).to(torch.bfloat16)
    (out.float() * w.float()).sum().backward()
    x = qkv.detach().float().requires_grad_(True)
    tot = 0
    for (_, B, S, off) in segs:
        o, _ = _ref(x, B, S, off, nh, nkv, hd, True)
        tot = tot + (o * w[off: off + B * S].float()).sum()
    tot.backward()
    err = (qkv.grad.float() - x.grad).abs().max().item()
    assert err <= 2e-2 * max(1.0, x.grad.abs().max().item()), err
    assert (qkv.grad[200:256] == 0).all() and (qkv.grad[536:] == 0).all()


@pytest.mark.parametrize("B,nh,nkv,hd,max_len,pos", [(3, 32, 8, 64, 384, 300), (2, 16, 2, 128, 256, 255), (5, 4, 4, 64, 128, 0),
                                                     (64, 32, 8, 64, 512, 37), (2, 8, 4, 128, 640, 513)])
def test_decode_attention_step(B, nh, nkv, hd, max_len, pos):
    """Single-token attention over the static KV cache (csrc/decode_attention.cu): cache append + softmax(q.K^T).V for all
    heads of a GQA group, against fp32 torch on the same cache contents."""
    from omni_avsr_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(pos + nh)
    kc = torch.randn(B, nkv, max_len, hd, device="cuda", generator=g).bfloat16()
    vc = torch.randn(B, nkv, max_len, hd, device="cuda", generator=g).bfloat16()
    kc0, vc0 = kc.clone(), vc.clone()
    M = 128
    qkv = (torch.randn(M, (nh + 2 * nkv) * hd, device="cuda", generator=g) * 1.2).bfloat16()
    out = torch.full((M, nh * hd), 5.0, device="cuda", dtype=torch.bfloat16)
    len_idx = torch.tensor([pos], device="cuda", dtype=torch.int64)
    ops.decode_attention(qkv[:B], kc, vc, len_idx, out[:B], B, nh, nkv, hd)
    torch.cuda.synchronize()
    q = qkv[:B, : nh * hd].float().view(B, nh, hd)
    k_new = qkv[:B, nh * hd: (nh + nkv) * hd].view(B, nkv, hd)
    v_new = qkv[:B, (nh + nkv) * hd:].view(B, nkv, hd)
    # the cache now holds the new token at `pos`, everything else untouched
    assert torch.equal(kc[:, :, pos], k_new) and torch.equal(vc[:, :, pos], v_new)
    keep = torch.ones(max_len, dtype=torch.bool, device="cuda")
    keep[pos] = False
    assert torch.equal(kc[:, :, keep], kc0[:, :, keep]) and torch.equal(vc[:, :, keep], vc0[:, :, keep])
    G = nh // nkv
    K = kc[:, :, : pos + 1].float().repeat_interleave(G, dim=1)          # [B, nh, n, hd]
    V = vc[:, :, : pos + 1].float().repeat_interleave(G, dim=1)
    s = torch.einsum("bhd,bhnd->bhn", q, K) / math.sqrt(hd)
    want = torch.einsum("bhn,bhnd->bhd", torch.softmax(s, dim=-1), V).reshape(B, nh * hd)
    err = (out[:B].float() - want).abs().max().item()
    assert err <= 1e-2 * max(1.0, want.abs().max().item()), err
    assert (out[B:] == 5.0).all()


@pytest.mark.parametrize("B,nh,nkv,hd,max_len,pos", [(3, 32, 8, 64, 384, 300), (2, 16, 2, 128, 256, 255), (4, 6, 2, 64, 128, 0),
                                                     (64, 32, 8, 64, 512, 445), (2, 14, 2, 64, 256, 77), (1, 12, 2, 128, 256, 9),
                                                     (2, 40, 8, 128, 256, 100)])
def test_decode_attention_fused_rope_equals_rope_then_attention(B, nh, nkv, hd, max_len, pos):
    """omni_decode_attention_rope (rotary embedding of the q heads and of the new key inside the single-token kernel) must
    give, bit for bit, the output AND the cache contents of the two-launch sequence omni_rope -> omni_decode_attention.
    Covers the GQA group sizes 3 (Llama-3.2-3B), 5 (Qwen2.5-14B/32B), 6 (1.5B), 7 (0.5B/7B) next to 4 and 8."""
    from omni_avsr_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(pos * 7 + nh)
    kc = torch.randn(B, nkv, max_len, hd, device="cuda", generator=g).bfloat16()
    vc = torch.randn(B, nkv, max_len, hd, device="cuda", generator=g).bfloat16()
    qkv = (torch.randn(B, (nh + 2 * nkv) * hd, device="cuda", generator=g) * 1.2).bfloat16()
    ang = torch.rand(max_len, hd // 2, device="cuda", generator=g) * 6.28
    emb = torch.cat([ang, ang], dim=-1)
    cos_t, sin_t = emb.cos().bfloat16().contiguous(), emb.sin().bfloat16().contiguous()
    len_idx = torch.tensor([pos], device="cuda", dtype=torch.int64)
    # two launches
    kc1, vc1, q1 = kc.clone(), vc.clone(), qkv.clone()
    ops.rope_(q1, cos_t, sin_t, torch.full((B,), pos, device="cuda", dtype=torch.int32), nh + nkv, hd)
    out1 = torch.empty(B, nh * hd, device="cuda", dtype=torch.bfloat16)
    ops.decode_attention(q1, kc1, vc1, len_idx, out1, B, nh, nkv, hd)
    # fused
    kc2, vc2 = kc.clone(), vc.clone()
    out2 = torch.empty(B, nh * hd, device="cuda", dtype=torch.bfloat16)
    ops.decode_attention(qkv, kc2, vc2, len_idx, out2, B, nh, nkv, hd, rope=(cos_t, sin_t))
    assert torch.equal(kc1.view(torch.int16), kc2.view(torch.int16)) and torch.equal(vc1.view(torch.int16), vc2.view(torch.int16))
    assert torch.equal(out1.view(torch.int16), out2.view(torch.int16))
    # and against fp32 torch
    q = q1[:, : nh * hd].float().view(B, nh, hd)
    G = nh // nkv
    K = kc1[:, :, : pos + 1].float().repeat_interleave(G, dim=1)
    V = vc1[:, :, : pos + 1].float().repeat_interleave(G, dim=1)
    s = torch.einsum("bhd,bhnd->bhn", q, K) / math.sqrt(hd)
    want = torch.einsum("bhn,bhnd->bhd", torch.softmax(s, dim=-1), V).reshape(B, nh * hd)
    assert (out2.float() - want).abs().max().item() <= 1e-2 * max(1.0, want.abs().max().item())
