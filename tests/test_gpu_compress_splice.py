"""GPU parity (bit-exact) of the Matryoshka compression and prompt-splice kernels against the oracle."""
import pytest
import torch

from oracle import matryoshka as om

pytestmark = pytest.mark.gpu


def _ops():
    from omni_avsr_b200 import ops
    return ops


@pytest.mark.parametrize("B,T,n_tok,D,rate", [(2, 1500, 800, 1024, 4), (2, 1500, 800, 1024, 16), (3, 400, 400, 1024, 2),
                                              (3, 400, 400, 1024, 5), (1, 1500, 799, 768, 4), (2, 64, 25, 64, 16),
                                              (2, 64, 25, 64, 32), (1, 40, 33, 8, 3)])
@pytest.mark.parametrize("mode", ["avg-pooling", "stack"])
def test_compress_bit_exact(B, T, n_tok, D, rate, mode):
    ops = _ops()
    g = torch.Generator().manual_seed(B * 1000 + n_tok + rate)
    x = torch.randn(B, T, D, generator=g).bfloat16()
    if mode == "avg-pooling" and n_tok < rate:
        # nn.AvgPool1d refuses an empty output (the reference would raise here); the drop-in raises too
        with pytest.raises(RuntimeError):
            om.compress(x[:, :n_tok], rate, mode)
        with pytest.raises(RuntimeError):
            ops.matryoshka_compress(x.cuda(), n_tok, rate, mode)
        return
    want = om.compress(x[:, :n_tok], rate, mode)
    got = ops.matryoshka_compress(x.cuda(), n_tok, rate, mode).cpu()
    assert got.shape == want.shape
    assert torch.equal(got.view(torch.int16), want.view(torch.int16))


@pytest.mark.parametrize("mode", ["avg-pooling", "stack"])
def test_compress_backward(mode):
    ops = _ops()
    B, T, n_tok, D, rate = 2, 50, 43, 64, 5
    g = torch.Generator().manual_seed(3)
    x = torch.randn(B, T, D, generator=g).bfloat16().requires_grad_(True)
    y = om.compress(x[:, :n_tok], rate, mode)
    dy = torch.randn(y.shape, generator=g).bfloat16()
    y.backward(dy)
    got = ops.matryoshka_compress_bwd(dy.cuda(), n_tok, T, rate, mode).cpu()
    assert torch.equal(got.view(torch.int16), x.grad.view(torch.int16))


def _setup(B, L, H, n_a, n_v, V, seed, P=(6, 6, 8)):
    g = torch.Generator().manual_seed(seed)
    embed = torch.nn.Embedding(V, H)
    embed.weight.data = torch.randn(V, H, generator=g).bfloat16()
    tokens = torch.randint(0, V - 8, (B, L), generator=g)
    labels = tokens.clone()
    if L > 2:
        labels[:, -1] = -100   # a padded position
    a = torch.randn(B, n_a, H, generator=g).bfloat16() if n_a else None
    v = torch.randn(B, n_v, H, generator=g).bfloat16() if n_v else None
    prompts = {k: torch.randn(1, p, H, generator=g).bfloat16() for k, p in zip(("audio", "video", "audiovisual"), P)}
    marker = (V - 4, V - 3, V - 2, V - 1)
    return embed, tokens, labels, a, v, prompts, marker


@pytest.mark.parametrize("is_qwen", [False, True])
@pytest.mark.parametrize("B,L,H,n_a,n_v", [(2, 12, 256, 50, 80), (3, 48, 2048, 200, 200), (1, 5, 64, 1, 1)])
def test_splice_train_bit_exact(is_qwen, B, L, H, n_a, n_v):
    ops = _ops()
    V = 1000
    embed, tokens, labels, a, v, prompts, marker = _setup(B, L, H, n_a, n_v, V, seed=B + L)
    with torch.no_grad():
        seqs, labs = om.build_train_sequences(embed, tokens, labels, a, v, prompts, marker, is_qwen)
    lay = ops.SpliceLayout(tokens=tokens.cuda(), labels=labels.cuda(), embed=embed.weight.data.cuda(),
                           audio_tok=a.cuda(), video_tok=v.cuda(),
                           prompts=[prompts[k][0].cuda() for k in ("audio", "video", "audiovisual")],
                           marker_ids=marker, has_bos=not is_qwen)
    outs = [torch.empty(B, s, H, device="cuda", dtype=torch.bfloat16) for s in lay.seq_len]
    outl = [torch.empty(B, s, device="cuda", dtype=torch.int64) for s in lay.seq_len]
    status = torch.zeros(1, device="cuda", dtype=torch.int32)
    ops.splice_prompt(lay, outs, outl, status)
    assert status.item() == 0
    for t, k in enumerate(("audio", "video", "audiovisual")):
        assert outs[t].shape == seqs[k].shape
        assert torch.equal(outs[t].cpu().view(torch.int16), seqs[k].view(torch.int16))
        assert torch.equal(outl[t].cpu(), labs[k])


@pytest.mark.parametrize("is_qwen", [False, True])
@pytest.mark.parametrize("task", [0, 1, 2])
def test_splice_infer_bit_exact(is_qwen, task):
    ops = _ops()
    B, H, V = 2, 128, 500
    embed, tokens, labels, a, v, prompts, marker = _setup(B, 1, H, 13, 7, V, seed=task)
    key = ("audio", "video", "audiovisual")[task]
    if is_qwen:
        tokens = torch.empty(B, 0, dtype=torch.int64)
    aa = a if task in (0, 2) else None
    vv = v if task in (1, 2) else None
    with torch.no_grad():
        want = om.build_infer_sequence(embed, tokens, aa, vv, prompts[key], marker, is_qwen)
    lay = ops.SpliceLayout(tokens=tokens.cuda(), labels=None, embed=embed.weight.data.cuda(),
                           audio_tok=None if aa is None else aa.cuda(), video_tok=None if vv is None else vv.cuda(),
                           prompts=[prompts[k][0].cuda() for k in ("audio", "video", "audiovisual")],
                           marker_ids=marker, has_bos=not is_qwen, task_mask=1 << task)
    outs = [None, None, None]
    outs[task] = torch.empty(B, lay.seq_len[task], H, device="cuda", dtype=torch.bfloat16)
    ops.splice_prompt(lay, outs, [None, None, None])
    assert torch.equal(outs[task].cpu().view(torch.int16), want.view(torch.int16))


def test_splice_bad_token_sets_status():
    ops = _ops()
    embed, tokens, labels, a, v, prompts, marker = _setup(2, 6, 64, 3, 3, 100, seed=1)
    tokens[1, 3] = 100   # out of range
    lay = ops.SpliceLayout(tokens=tokens.cuda(), labels=labels.cuda(), embed=embed.weight.data.cuda(),
                           audio_tok=a.cuda(), video_tok=v.cuda(),
                           prompts=[prompts[k][0].cuda() for k in ("audio", "video", "audiovisual")],
                           marker_ids=marker, has_bos=True)
    outs = [torch.empty(2, s, 64, device="cuda", dtype=torch.bfloat16) for s in lay.seq_len]
    status = torch.zeros(1, device="cuda", dtype=torch.int32)
    ops.splice_prompt(lay, outs, [None] * 3, status)
    assert status.item() == 1


def test_splice_backward_matches_autograd():
    ops = _ops()
    B, L, H, n_a, n_v = 2, 7, 64, 5, 4
    embed, tokens, labels, a, v, prompts, marker = _setup(B, L, H, n_a, n_v, 200, seed=2)
    a.requires_grad_(True)
    v.requires_grad_(True)
    seqs, _ = om.build_train_sequences(embed, tokens, labels, a, v, prompts, marker, False)
    g = torch.Generator().manual_seed(4)
    douts = [torch.randn(seqs[k].shape, generator=g).bfloat16() for k in ("audio", "video", "audiovisual")]
    torch.autograd.backward([seqs[k] for k in ("audio", "video", "audiovisual")], douts)
    lay = ops.SpliceLayout(tokens=tokens.cuda(), labels=labels.cuda(), embed=embed.weight.data.cuda(),
                           audio_tok=a.detach().cuda(), video_tok=v.detach().cuda(),
                           prompts=[prompts[k][0].cuda() for k in ("audio", "video", "audiovisual")],
                           marker_ids=marker, has_bos=True)
    da, dv = ops.splice_prompt_bwd(lay, [d.cuda() for d in douts], True, True)
    assert torch.equal(da.cpu().view(torch.int16), a.grad.view(torch.int16))
    assert torch.equal(dv.cpu().view(torch.int16), v.grad.view(torch.int16))
