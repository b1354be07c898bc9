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


# ---------------------------------------------------------------------------------------------------------------
# fused compression -> projector MLP -> splice (omni_pool_project_splice, one persistent launch)
# ---------------------------------------------------------------------------------------------------------------
def _proj(g, K1, I, H):
    w1 = (torch.randn(I, K1, generator=g) / K1 ** 0.5).bfloat16()
    b1 = (torch.randn(I, generator=g) * 0.1).bfloat16()
    w2 = (torch.randn(H, I, generator=g) / I ** 0.5).bfloat16()
    b2 = (torch.randn(H, generator=g) * 0.1).bfloat16()
    return [t.cuda() for t in (w1, b1, w2, b2)]


def _fused_vs_unfused(B, L, H, I, Ta, n_tok_a, Da, ra, Tv, n_tok_v, Dv, rv, mode, is_qwen, task_mask, seed):
    """The fused launch must reproduce, bit for bit, what the separate kernels of the same library produce
    (compress -> GEMM+ReLU -> GEMM -> splice), including labels, and the compressed features must be bit-exact with
    the oracle's AvgPool1d / stacking."""
    ops = _ops()
    V = 1000
    g = torch.Generator().manual_seed(seed)
    use_a, use_v = bool(task_mask & 5), bool(task_mask & 6)
    train = task_mask == 7
    na, nv = n_tok_a // ra, n_tok_v // rv
    embed, tokens, labels, _, _, prompts, marker = _setup(B, L, H, 0, 0, V, seed)
    if not train:
        tokens = torch.empty(B, 0, dtype=torch.int64) if is_qwen else tokens[:, :1].contiguous()
        labels = None
    xa = torch.randn(B, Ta, Da, generator=g).bfloat16().cuda() if use_a else None
    xv = torch.randn(B, Tv, Dv, generator=g).bfloat16().cuda() if use_v else None
    stack = mode == "stack"
    pa = _proj(g, Da * (ra if stack else 1), I, H) if use_a else None
    pv = _proj(g, Dv * (rv if stack else 1), I, H) if use_v else None
    pr = [prompts[k][0].cuda() for k in ("audio", "video", "audiovisual")]
    tok_d = tokens.cuda()
    lab_d = None if labels is None else labels.cuda()
    emb_d = embed.weight.data.cuda()
    # ---- unfused chain
    ca = ops.matryoshka_compress(xa, n_tok_a, ra, mode) if use_a else None
    cv = ops.matryoshka_compress(xv, n_tok_v, rv, mode) if use_v else None
    if use_a:
        want = om.compress(xa.cpu()[:, :n_tok_a], ra, mode)
        assert torch.equal(ca.cpu().view(torch.int16), want.view(torch.int16))

    def project(c, p):
        h = ops.gemm(c.view(-1, c.shape[-1]), p[0], bias=p[1], act="relu")
        return h, ops.gemm(h, p[2], bias=p[3])
    ha, ta = project(ca, pa) if use_a else (None, None)
    hv, tv = project(cv, pv) if use_v else (None, None)
    lay_u = ops.SpliceLayout(tokens=tok_d, labels=lab_d, embed=emb_d, audio_tok=None if ta is None else ta.view(B, na, H),
                             video_tok=None if tv is None else tv.view(B, nv, H), prompts=pr, marker_ids=marker,
                             has_bos=not is_qwen, task_mask=task_mask)
    outs_u = [torch.empty(B, s, H, device="cuda", dtype=torch.bfloat16) if s else None for s in lay_u.seq_len]
    outl_u = [torch.empty(B, s, device="cuda", dtype=torch.int64) if (s and train) else None for s in lay_u.seq_len]
    ops.splice_prompt(lay_u, outs_u, outl_u)
    # ---- fused launch
    lay_f = ops.SpliceLayout(tokens=tok_d, labels=lab_d, embed=emb_d, audio_tok=None, video_tok=None, prompts=pr,
                             marker_ids=marker, has_bos=not is_qwen, task_mask=task_mask,
                             n_audio=na if use_a else None, n_video=nv if use_v else None)
    assert lay_f.seq_len == lay_u.seq_len
    outs_f = [torch.full((B, s, H), 7.0, device="cuda", dtype=torch.bfloat16) if s else None for s in lay_f.seq_len]
    outl_f = [torch.full((B, s), 12345, device="cuda", dtype=torch.int64) if (s and train) else None for s in lay_f.seq_len]
    status = torch.zeros(1, device="cuda", dtype=torch.int32)
    res = ops.pool_project_splice(lay_f, outs_f, outl_f,
                                  ops.PoolProjectInput(xa, n_tok_a, ra, *pa) if use_a else None,
                                  ops.PoolProjectInput(xv, n_tok_v, rv, *pv) if use_v else None, mode, status=status,
                                  want_tok=True)
    torch.cuda.synchronize()
    assert status.item() == 0
    for name, c, h, t in (("audio", ca, ha, ta), ("video", cv, hv, tv)):
        if c is None:
            assert res[name] is None
            continue
        pooled, hidden, tok = res[name]
        assert torch.equal(pooled.view(torch.int16), c.view(-1, c.shape[-1]).view(torch.int16)), name + " pooled"
        assert torch.equal(hidden.view(torch.int16), h.view(torch.int16)), name + " hidden"
        assert torch.equal(tok.view(torch.int16), t.view(torch.int16)), name + " tokens"
    for t in range(3):
        if outs_u[t] is None:
            continue
        assert torch.equal(outs_f[t].view(torch.int16), outs_u[t].view(torch.int16)), f"task {t} rows"
        if train:
            assert torch.equal(outl_f[t], outl_u[t]), f"task {t} labels"


@pytest.mark.parametrize("is_qwen", [False, True])
@pytest.mark.parametrize("mode", ["avg-pooling", "stack"])
def test_fused_pool_project_splice_small(is_qwen, mode):
    # the geometry of the reference-executed golden case: odd widths (I = 96: edge tiles, K tail), two clips
    _fused_vs_unfused(2, 12, 256, 96, 100, 62, 64, 4, 23, 23, 768, 2, mode, is_qwen, 7, seed=11)
    _fused_vs_unfused(3, 7, 128, 64, 70, 61, 40, 16, 30, 29, 72, 5, mode, is_qwen, 7, seed=12)


@pytest.mark.parametrize("ra,rv", [(4, 2), (16, 5)])
def test_fused_pool_project_splice_config2_geometry(ra, rv):
    # Whisper-medium / AV-HuBERT-Large / Llama-3.2-1B widths, 16 s clips (800 / 400 encoder tokens), 3 utterances
    _fused_vs_unfused(3, 48, 2048, 2048, 1500, 800, 1024, ra, 400, 400, 1024, rv, "avg-pooling", False, 7, seed=ra)


@pytest.mark.parametrize("is_qwen", [False, True])
@pytest.mark.parametrize("task", [0, 1, 2])
def test_fused_pool_project_splice_infer(is_qwen, task):
    _fused_vs_unfused(2, 1, 256, 128, 90, 77, 64, 4, 31, 31, 128, 2, "avg-pooling", is_qwen, 1 << task, seed=20 + task)


def test_fused_pool_project_splice_rate_one_and_ragged_tail():
    # rate 1 (no compression: the non-Matryoshka recipes with downsample_ratio 1) and n_tok not divisible by the rate
    _fused_vs_unfused(2, 9, 256, 128, 50, 25, 64, 1, 20, 19, 64, 3, "avg-pooling", False, 7, seed=31)


def test_fused_backward_matches_unfused_chain():
    """Gradients of the fused autograd function == gradients of the unfused modules (same kernels underneath)."""
    ops = _ops()
    from omni_avsr_b200.Llama_LoRA import PackedRows
    from omni_avsr_b200.modeling_OmniAVSR import PoolProjectSpliceFn, SpliceFn, compress
    from omni_avsr_b200 import autograd_ops as ag
    B, L, H, I, V = 2, 10, 256, 128, 500
    g = torch.Generator().manual_seed(5)
    embed, tokens, labels, _, _, prompts, marker = _setup(B, L, H, 0, 0, V, 5)
    xa = torch.randn(B, 100, 64, generator=g).bfloat16().cuda()
    xv = torch.randn(B, 40, 128, generator=g).bfloat16().cuda().requires_grad_(True)
    pa = [p.requires_grad_(True) for p in _proj(g, 64, I, H)]
    pv = [p.requires_grad_(True) for p in _proj(g, 128, I, H)]
    pr = [prompts[k][0].cuda() for k in ("audio", "video", "audiovisual")]
    common = dict(tokens=tokens.cuda(), labels=labels.cuda(), embed=embed.weight.data.cuda(), prompts=pr,
                  marker_ids=marker, has_bos=True)
    na, nv = 62 // 4, 40 // 2
    lay_f = ops.SpliceLayout(audio_tok=None, video_tok=None, n_audio=na, n_video=nv, **common)
    rows = PackedRows.get([(t, B, lay_f.seq_len[t]) for t in range(3)], "cuda")
    xp_f, *_ = PoolProjectSpliceFn.apply(xa, xv, *pa, *pv, (lay_f, rows, 62, 4, 40, 2, "avg-pooling", True))
    dxp = torch.randn(xp_f.shape, generator=g).bfloat16().cuda()
    xp_f.backward(dxp)
    got = [xv.grad.clone()] + [p.grad.clone() for p in pa + pv]
    xv.grad = None
    for p in pa + pv:
        p.grad = None
    ca = compress(xa, 62, 4, "avg-pooling")
    cv = compress(xv, 40, 2, "avg-pooling")
    ta = ag.TrainableLinearFn.apply(ag.TrainableLinearFn.apply(ca.view(-1, 64), pa[0], pa[1], "relu"), pa[2], pa[3], None)
    tv = ag.TrainableLinearFn.apply(ag.TrainableLinearFn.apply(cv.view(-1, 128), pv[0], pv[1], "relu"), pv[2], pv[3], None)
    lay_u = ops.SpliceLayout(audio_tok=ta.view(B, na, H).detach(), video_tok=tv.view(B, nv, H).detach(), **common)
    xp_u, *_ = SpliceFn.apply(ta.view(B, na, H), tv.view(B, nv, H), lay_u, rows, True)
    assert torch.equal(xp_u.view(torch.int16), xp_f.view(torch.int16))
    xp_u.backward(dxp)
    want = [xv.grad] + [p.grad for p in pa + pv]
    for a, b in zip(got, want):
        assert torch.equal(a.view(torch.int16), b.view(torch.int16))
