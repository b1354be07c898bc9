"""Device-side beam search pieces (csrc/beam_search.cu, the BEAM variant of csrc/decode_attention.cu) against torch restatements
of the HF 4.43.1 operations they replace (`log_softmax` + `topk` of `GenerationMixin._beam_search`, `_reorder_cache`).
The end-to-end beam search is compared with the oracle in tests/test_gpu_llm.py / tests/test_gpu_model.py."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("rows,V,K,case", [(6, 128261, 15, "random"), (3, 151669, 4, "random"), (4, 1003, 5, "random"),
                                           (2, 20, 15, "random"), (3, 4096, 15, "ties"), (3, 16384, 15, "ties"), (2, 128261, 15, "equal"),
                                           (2, 50000, 32, "random")])
def test_beam_topk_rows(rows, V, K, case):
    from omni_avsr_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(V + K)
    Vp = (V + 7) // 8 * 8
    buf = torch.zeros(rows, Vp, device="cuda", dtype=torch.bfloat16)
    if case == "random":
        buf[:, :V] = (torch.randn(rows, V, device="cuda", generator=g) * 3).bfloat16()
    elif case == "ties":       # few distinct values: the threshold element is tied thousands of times (slow path)
        buf[:, :V] = torch.randint(0, 3, (rows, V), device="cuda", generator=g).bfloat16()
    else:
        buf[:, :V] = 1.5
    buf[:, V:] = 1e4           # padding columns must never be candidates
    logits = buf[:, :V]
    scores = torch.randn(rows, device="cuda", generator=g)
    n_cand = 2 * K
    cs = torch.empty(rows, n_cand, device="cuda", dtype=torch.float32)
    ct = torch.empty(rows, n_cand, device="cuda", dtype=torch.int32)
    ops.beam_topk_rows(logits, V, scores, cs, ct)
    torch.cuda.synchronize()
    logp = torch.log_softmax(logits.float(), dim=-1) + scores[:, None]
    n = min(n_cand, V)
    want_s, _ = torch.topk(logp, n, dim=1)
    assert torch.allclose(cs[:, :n], want_s, atol=2e-5, rtol=1e-6), (cs[:, :n] - want_s).abs().max()
    tok = ct[:, :n].long()
    assert (tok >= 0).all() and (tok < V).all()
    assert torch.allclose(logp.gather(1, tok), cs[:, :n], atol=2e-5, rtol=1e-6)        # the tokens carry those scores
    for r in range(rows):
        assert len(set(tok[r].tolist())) == n                                           # no duplicates
        # ties: equal bf16 logits are listed by ascending token id, and the lowest ids of a tied group are the ones taken
        raw = logits[r].float()
        vals = raw[tok[r]]
        assert (vals[1:] <= vals[:-1]).all()
        same = vals[1:] == vals[:-1]
        assert (tok[r][1:][same] > tok[r][:-1][same]).all()
        last = vals[-1]
        taken_of_last = tok[r][vals == last]
        all_of_last = torch.nonzero(raw == last).flatten()
        assert torch.equal(taken_of_last, all_of_last[: len(taken_of_last)])
    if V < n_cand:
        assert (ct[:, V:] == -1).all() and torch.isinf(cs[:, V:]).all()


@pytest.mark.parametrize("U,K,nh,nkv,hd,max_len,s0,t_new", [(2, 4, 32, 8, 64, 256, 100, 7), (1, 15, 32, 8, 64, 512, 413, 31),
                                                            (2, 3, 16, 2, 128, 256, 130, 0), (3, 5, 12, 4, 64, 128, 17, 12),
                                                            (1, 6, 40, 8, 128, 256, 60, 20)])
def test_decode_attention_beam_indirection(U, K, nh, nkv, hd, max_len, s0, t_new):
    """BEAM variant of the single-token attention: prompt keys from row (b // K) * K, generated position t from row
    ind[par][b][t], the new token appended to row b itself -- against fp32 torch on explicitly gathered K / V."""
    from omni_avsr_b200 import ops
    B = U * K
    g = torch.Generator(device="cuda").manual_seed(s0 + t_new)
    kc = torch.randn(B, nkv, max_len, hd, device="cuda", generator=g).bfloat16()
    vc = torch.randn(B, nkv, max_len, hd, device="cuda", generator=g).bfloat16()
    kc0, vc0 = kc.clone(), vc.clone()
    qkv = (torch.randn(B, (nh + 2 * nkv) * hd, device="cuda", generator=g) * 1.2).bfloat16()
    pos = s0 + t_new
    par = (t_new + 1) & 1
    ind = torch.full((2, B, 40), -7, device="cuda", dtype=torch.int32)          # the other buffer must not be read
    for b in range(B):
        u = b // K
        ind[par, b, :t_new] = torch.randint(u * K, (u + 1) * K, (t_new,), device="cuda", generator=g).int()
        ind[par, b, t_new] = b
    out = torch.empty(B, nh * hd, device="cuda", dtype=torch.bfloat16)
    len_idx = torch.tensor([pos], device="cuda", dtype=torch.int64)
    pl = torch.tensor([s0], device="cuda", dtype=torch.int64)
    ops.decode_attention(qkv, kc, vc, len_idx, out, B, nh, nkv, hd, beam=(ind, pl, K))
    torch.cuda.synchronize()
    k_new = qkv[:, nh * hd: (nh + nkv) * hd].view(B, nkv, hd)
    v_new = qkv[:, (nh + nkv) * hd:].view(B, nkv, hd)
    assert torch.equal(kc[:, :, pos], k_new) and torch.equal(vc[:, :, pos], v_new)
    keep = torch.ones(max_len, dtype=torch.bool, device="cuda")
    keep[pos] = False
    assert torch.equal(kc[:, :, keep], kc0[:, :, keep]) and torch.equal(vc[:, :, keep], vc0[:, :, keep])
    G = nh // nkv
    q = qkv[:, : nh * hd].float().view(B, nh, hd)
    want = torch.empty(B, nh * hd, device="cuda")
    for b in range(B):
        rows_of = [(b // K) * K] * s0 + ind[par, b, :t_new].tolist() + [b]
        idx = torch.tensor(rows_of, device="cuda")
        pp = torch.arange(pos + 1, device="cuda")
        Kb = kc[idx, :, pp].float().permute(1, 0, 2).repeat_interleave(G, dim=0)      # [nh, n, hd]
        Vb = vc[idx, :, pp].float().permute(1, 0, 2).repeat_interleave(G, dim=0)
        s = torch.einsum("hd,hnd->hn", q[b], Kb) / math.sqrt(hd)
        want[b] = torch.einsum("hn,hnd->hd", torch.softmax(s, dim=-1), Vb).reshape(-1)
    err = (out.float() - want).abs().max().item()
    assert err <= 1e-2 * max(1.0, want.abs().max().item()), err
