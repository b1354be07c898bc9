"""GPU parity of the row kernels (norms, RoPE, SwiGLU, GELU, CE, argmax, AdamW) against plain torch references that
restate the reference's op sequence (transformers LlamaRMSNorm / apply_rotary_pos_emb / LlamaMLP, fairseq LayerNorm +
gelu, CrossEntropyLoss on fp32 logits, torch.optim.AdamW + clip_grad_norm_)."""
import pytest
import torch
import torch.nn.functional as F

from oracle import llm_lora as ol

pytestmark = pytest.mark.gpu


def _ops():
    from omni_avsr_b200 import ops
    return ops


def _bits(t):
    return t.detach().cpu().view(torch.int16)


@pytest.mark.parametrize("rows,H", [(5, 64), (300, 2048), (129, 4096)])
def test_rmsnorm_fwd_bit_exact_and_bwd(rows, H):
    ops = _ops()
    g = torch.Generator().manual_seed(rows)
    x = torch.randn(rows, H, generator=g).bfloat16()
    w = (1 + 0.1 * torch.randn(H, generator=g)).bfloat16()
    norm = ol.RMSNorm(H, 1e-5)
    norm.weight.data = w.clone()
    xr = x.clone().requires_grad_(True)
    want = norm(xr)
    got, rstd = ops.rmsnorm_fwd(x.cuda(), w.cuda(), 1e-5, want_rstd=True)
    # rsqrt may differ by an ulp between CPU and GPU: allow <= 1 bf16 ulp on a handful of elements
    diff = (got.cpu().float() - want.float()).abs()
    assert diff.max().item() <= 2 ** -6 * want.float().abs().max().item()
    assert (diff > 0).float().mean().item() < 0.01
    dy = torch.randn(rows, H, generator=g).bfloat16()
    want.backward(dy)
    dx = ops.rmsnorm_bwd(dy.cuda(), x.cuda(), w.cuda(), rstd)
    err = (dx.cpu().float() - xr.grad.float()).abs().max().item()
    assert err <= 3e-2 * xr.grad.float().abs().max().item()


@pytest.mark.parametrize("rows,H", [(7, 1024), (400, 2048)])
def test_layernorm_fwd_bwd(rows, H):
    ops = _ops()
    g = torch.Generator().manual_seed(H)
    x = torch.randn(rows, H, generator=g).bfloat16()
    w = (1 + 0.1 * torch.randn(H, generator=g)).bfloat16()
    b = (0.1 * torch.randn(H, generator=g)).bfloat16()
    xr = x.clone().requires_grad_(True)
    want = F.layer_norm(xr, (H,), w, b, 1e-5)
    got, mean, rstd = ops.layernorm_fwd(x.cuda(), w.cuda(), b.cuda(), 1e-5, want_stats=True)
    assert (got.cpu().float() - want.float()).abs().max().item() <= 2 ** -6 * want.float().abs().max().item()
    dy = torch.randn(rows, H, generator=g).bfloat16()
    want.backward(dy)
    dx = ops.layernorm_bwd(dy.cuda(), x.cuda(), w.cuda(), mean, rstd)
    assert (dx.cpu().float() - xr.grad.float()).abs().max().item() <= 3e-2 * xr.grad.float().abs().max().item()


@pytest.mark.parametrize("hd,nh,nkv", [(64, 32, 8), (128, 16, 2)])
def test_rope_bit_exact(hd, nh, nkv):
    ops = _ops()
    from omni_avsr_b200.Llama_LoRA import LLMArch, rope_tables
    g = torch.Generator().manual_seed(hd)
    B, S = 2, 50
    rs = dict(factor=32.0, low_freq_factor=1.0, high_freq_factor=4.0, original_max_position_embeddings=8192)
    arch = LLMArch("llama", nh * hd, 4 * nh * hd, 1, nh, nkv, 100, 1e-5, 500000.0, hd, rs, inv_freq_dtype="bf16")
    cfg = ol.LLMConfig("llama", nh * hd, 4 * nh * hd, 1, nh, nkv, 100, 1e-5, 500000.0, hd, rs, inv_freq_dtype="bf16")
    qkv = torch.randn(B * S, (nh + 2 * nkv) * hd, generator=g).bfloat16()
    pos = torch.arange(S).repeat(B)
    cos, sin = ol.rope_cos_sin(cfg, torch.arange(S).unsqueeze(0), torch.bfloat16)
    q = qkv[:, : nh * hd].view(B, S, nh, hd).transpose(1, 2)
    k = qkv[:, nh * hd: (nh + nkv) * hd].view(B, S, nkv, hd).transpose(1, 2)
    qo, ko = ol.apply_rotary_pos_emb(q, k, cos, sin)
    cos_t, sin_t = rope_tables(arch, 64, "cuda")
    assert torch.equal(_bits(cos_t[:S]), _bits(cos[0]))
    buf = qkv.cuda().clone()
    ops.rope_(buf, cos_t, sin_t, pos.int().cuda(), nh + nkv, hd)
    out = buf.cpu()
    assert torch.equal(_bits(out[:, : nh * hd].view(B, S, nh, hd).transpose(1, 2)), _bits(qo))
    assert torch.equal(_bits(out[:, nh * hd: (nh + nkv) * hd].view(B, S, nkv, hd).transpose(1, 2)), _bits(ko))
    assert torch.equal(_bits(out[:, (nh + nkv) * hd:]), _bits(qkv[:, (nh + nkv) * hd:]))   # v untouched
    # inverse rotation == autograd of the forward
    qr = q.clone().float().requires_grad_(True)
    o, _ = ol.apply_rotary_pos_emb(qr, k.float(), cos.float(), sin.float())
    dy = torch.randn(o.shape, generator=g)
    o.backward(dy)
    dbuf = torch.zeros_like(qkv)
    dbuf[:, : nh * hd] = dy.transpose(1, 2).reshape(B * S, nh * hd).bfloat16()
    d = dbuf.cuda()
    ops.rope_(d, cos_t, sin_t, pos.int().cuda(), nh + nkv, hd, inverse=True)
    got = d.cpu()[:, : nh * hd].view(B, S, nh, hd).transpose(1, 2).float()
    assert (got - qr.grad).abs().max().item() <= 2e-2 * qr.grad.abs().max().item()


def test_swiglu_and_gelu():
    ops = _ops()
    g = torch.Generator().manual_seed(3)
    rows, I = 37, 512
    gu = torch.randn(rows, 2 * I, generator=g).bfloat16()
    gr = gu.clone().requires_grad_(True)
    want = F.silu(gr[:, :I]) * gr[:, I:]
    got = ops.swiglu_fwd(gu.cuda())
    assert (got.cpu().float() - want.float()).abs().max().item() <= 2 ** -7 * want.float().abs().max().item()
    d = torch.randn(rows, I, generator=g).bfloat16()
    want.backward(d)
    dgu = ops.swiglu_bwd(d.cuda(), gu.cuda())
    assert (dgu.cpu().float() - gr.grad.float()).abs().max().item() <= 3e-2 * gr.grad.float().abs().max().item()
    x = torch.randn(rows, I, generator=g).bfloat16()
    xr = x.clone().requires_grad_(True)
    wy = F.gelu(xr.float()).type_as(xr)          # fairseq gelu: fp32 then cast (modules/gelu.py)
    gy = ops.gelu_fwd(x.cuda())
    assert (gy.cpu().float() - wy.float()).abs().max().item() <= 2 ** -7 * wy.float().abs().max().item()
    wy.backward(d)
    gdx = ops.gelu_bwd(d.cuda(), x.cuda())
    assert (gdx.cpu().float() - xr.grad.float()).abs().max().item() <= 2e-2 * xr.grad.float().abs().max().item()


@pytest.mark.parametrize("R,V", [(9, 1005), (33, 128261)])
def test_cross_entropy_and_argmax(R, V):
    ops = _ops()
    g = torch.Generator().manual_seed(V)
    Vp = (V + 7) // 8 * 8
    buf = torch.zeros(R, Vp).bfloat16()
    buf[:, :V] = (torch.randn(R, V, generator=g) * 3).bfloat16()
    tgt = torch.randint(0, V, (R,), generator=g)
    tgt[1] = -100
    lr = buf[:, :V].float().requires_grad_(True)
    want = F.cross_entropy(lr, tgt, reduction="none", ignore_index=-100)
    d_logits = buf.cuda()[:, :V]
    loss, lse = ops.ce_fwd(d_logits, tgt.cuda())
    assert torch.allclose(loss.cpu(), want.detach(), atol=2e-4, rtol=1e-5)
    assert torch.equal(ops.argmax_rows(d_logits).cpu(), buf[:, :V].float().argmax(-1))
    scale = torch.rand(R, generator=g)
    (want * scale).sum().backward()
    ops.ce_bwd_(d_logits, tgt.cuda(), lse, scale.cuda())
    err = (d_logits.cpu().float() - lr.grad).abs().max().item()
    assert err <= 1e-2 * lr.grad.abs().max().item()
    assert d_logits[1].abs().max().item() == 0


def test_gather_scatter_rows():
    ops = _ops()
    g = torch.Generator().manual_seed(8)
    table = torch.randn(50, 128, generator=g).bfloat16()
    idx = torch.randperm(50, generator=g)[:17]
    got = ops.gather_rows(table.cuda(), idx.cuda())
    assert torch.equal(_bits(got), _bits(table[idx]))
    out = torch.zeros(50, 128, device="cuda", dtype=torch.bfloat16)
    ops.scatter_rows(got, idx.cuda(), out)
    want = torch.zeros(50, 128).bfloat16()
    want[idx] = table[idx]
    assert torch.equal(_bits(out), _bits(want))


def test_fused_clip_adamw_matches_torch():
    ops = _ops()
    g = torch.Generator().manual_seed(4)
    n = 10007
    p0 = torch.randn(n, generator=g).bfloat16()
    ref = torch.nn.Parameter(p0.float().clone())
    opt = torch.optim.AdamW([ref], lr=1e-3, weight_decay=0.1, betas=(0.9, 0.98))
    p = p0.cuda().clone()
    m = torch.zeros(n, device="cuda")
    v = torch.zeros(n, device="cuda")
    for step in range(1, 4):
        grad = (torch.randn(n, generator=g) * (30.0 if step == 2 else 0.01)).bfloat16()
        ref.grad = grad.float().clone()
        torch.nn.utils.clip_grad_norm_([ref], 10.0)
        opt.step()
        acc = torch.zeros(1, device="cuda")
        ops.sumsq_(grad.cuda(), acc)
        assert abs(acc.item() - grad.float().pow(2).sum().item()) <= 1e-3 * grad.float().pow(2).sum().item()
        ops.adamw_(p, grad.cuda(), m, v, lr=1e-3, beta1=0.9, beta2=0.98, eps=1e-8, weight_decay=0.1, step=step,
                   max_norm=10.0, sumsq=acc)
        # the product keeps bf16 params: compare against the fp32 reference rounded, 1 bf16 ulp slack per step
        assert (p.cpu().float() - ref.data).abs().max().item() <= step * 2 ** -7 * ref.data.abs().max().item()


def test_logmel_matches_oracle():
    ops = _ops()
    from oracle import encoders as oe
    from omni_avsr_b200.encoders import LogMel
    g = torch.Generator().manual_seed(12)
    audio = torch.randn(2, 16000 * 2 + 123, generator=g)
    audio[1, 20000:] = 0
    want = oe.log_mel(audio)
    fe = LogMel("cuda")
    got = fe(audio.cuda())
    assert got.shape == (2, 80, 3000)
    assert (got.float().cpu() - want).abs().max().item() <= 1e-2          # bf16 output of values in [-1, 2]
    got_bf = fe(audio.bfloat16().cuda().unsqueeze(-1).squeeze(-1))
    want_bf = oe.log_mel(audio.bfloat16().float())
    assert (got_bf.float().cpu() - want_bf).abs().max().item() <= 1e-2
    # full-length (30 s) input exercises the reflect padding at the far end
    long = torch.randn(1, 480000, generator=g)
    assert (fe(long.cuda()).float().cpu() - oe.log_mel(long)).abs().max().item() <= 1e-2


def test_prelu_kernels():
    ops = _ops()
    g = torch.Generator().manual_seed(13)
    x = torch.randn(6, 64, 10, 12, generator=g).bfloat16().contiguous(memory_format=torch.channels_last)
    r = torch.randn(6, 64, 10, 12, generator=g).bfloat16().contiguous(memory_format=torch.channels_last)
    slope = (torch.rand(64, generator=g) * 0.5).bfloat16()
    want = F.prelu(x + r, slope)
    got = ops.prelu_res_(x.cuda().clone(memory_format=torch.channels_last), slope.cuda(), r.cuda())
    assert torch.equal(_bits(got), _bits(want))
    want1 = F.prelu(x, slope)
    got1 = ops.prelu_res_(x.cuda().clone(memory_format=torch.channels_last), slope.cuda())
    assert torch.equal(_bits(got1), _bits(want1))
    # folded-BatchNorm shifts of both convolutions applied in the same pass (bit-exact with the separate bf16 adds)
    b1 = torch.randn(64, generator=g).bfloat16()
    b2 = torch.randn(64, generator=g).bfloat16()
    want2 = F.prelu((x + b1.view(1, -1, 1, 1)) + (r + b2.view(1, -1, 1, 1)), slope)
    got2 = ops.prelu_res_(x.cuda().clone(memory_format=torch.channels_last), slope.cuda(), r.cuda(), bias=b1.cuda(),
                          res_bias=b2.cuda())
    assert torch.equal(_bits(got2), _bits(want2))
    wantp = F.max_pool2d(F.prelu(x, slope), 3, 2, 1)
    gotp = ops.prelu_maxpool3x3s2(x.cuda(), slope.cuda())
    assert gotp.shape == wantp.shape
    assert torch.equal(_bits(gotp), _bits(wantp))


def test_conv3d_front_matches_torch():
    """im2col kernel + tcgen05 GEMM == Conv3d(1, C, (5,7,7), stride (1,2,2), pad (2,3,3)) (+ bias), incl. all borders."""
    ops = _ops()
    g = torch.Generator().manual_seed(21)
    B, T, H, W, C = 2, 7, 88, 88, 64
    video = torch.randn(B, T, H, W, generator=g).bfloat16()
    w3 = (torch.randn(C, 1, 5, 7, 7, generator=g) * 0.1).bfloat16()
    bias = torch.randn(C, generator=g).bfloat16()
    want = F.conv3d(video.float().unsqueeze(1), w3.float(), bias.float(), stride=(1, 2, 2), padding=(2, 3, 3))
    want = want.transpose(1, 2).reshape(B * T, C, 44, 44)
    wmat = torch.zeros(C, 256, dtype=torch.bfloat16)
    wmat[:, :245] = w3.reshape(C, 245)
    got = ops.conv3d_front(video.cuda(), wmat.cuda(), bias.cuda()).float().cpu()
    assert got.shape == want.shape
    err = (got - want).abs()
    assert err.max().item() <= 1e-2 * want.abs().max().item(), err.max().item()
    # borders and temporal edges individually
    for sl in (got[0] - want[0], got[T - 1] - want[T - 1], got[:, :, 0] - want[:, :, 0], got[:, :, :, 43] - want[:, :, :, 43]):
        assert sl.abs().max().item() <= 1e-2 * want.abs().max().item()


@pytest.mark.parametrize("B,T", [(2, 7), (1, 3), (3, 12)])
def test_front3d_prelu_maxpool_matches_torch(B, T):
    """Time-major im2col (49 taps) + overlapping-row tcgen05 GEMM (5 temporal K blocks) + PReLU/MaxPool kernel ==
    Conv3d(1, C, (5,7,7), (1,2,2), (2,3,3)) + bias -> PReLU -> MaxPool3d((1,3,3), (1,2,2), (0,1,1)) (resnet.py:137-140),
    including the temporal edges (zero padding) and all spatial borders."""
    ops = _ops()
    g = torch.Generator().manual_seed(23 + T)
    H, W, C = 88, 88, 64
    video = torch.randn(B, T, H, W, generator=g).bfloat16()
    w3 = (torch.randn(C, 1, 5, 7, 7, generator=g) * 0.1).bfloat16()
    bias = torch.randn(C, generator=g).bfloat16()
    # slopes of every kind: the pooling kernel takes max / min over the window first and applies PReLU to the two extremes,
    # which is exact for negative slopes (V-shaped PReLU) and slopes above 1 as well
    slope = torch.cat([torch.rand(C // 2, generator=g) * 0.5, torch.rand(C - C // 2, generator=g) * 2.0 - 0.75]).bfloat16()
    conv = F.conv3d(video.float().unsqueeze(1), w3.float(), bias.float(), stride=(1, 2, 2), padding=(2, 3, 3))
    conv = conv.bfloat16().float()                                     # the GEMM epilogue rounds to bf16
    act = F.prelu(conv, slope.float())
    want = F.max_pool3d(act, (1, 3, 3), (1, 2, 2), (0, 1, 1)).transpose(1, 2).reshape(B * T, C, 22, 22)
    wmat = torch.zeros(C, 5, 64, dtype=torch.bfloat16)
    wmat[:, :, :49] = w3.reshape(C, 5, 49)
    got = ops.front3d_prelu_maxpool(video.cuda(), wmat.view(C, 320).cuda(), bias.cuda(), slope.cuda())
    assert got.is_contiguous(memory_format=torch.channels_last)
    got = got.float().cpu()
    assert got.shape == want.shape
    tol = 1e-2 * want.abs().max().item()
    assert (got - want).abs().max().item() <= tol
    for sl in (got[0] - want[0], got[T - 1] - want[T - 1], got[B * T - 1] - want[B * T - 1],
               got[:, :, 0] - want[:, :, 0], got[:, :, :, 21] - want[:, :, :, 21]):
        assert sl.abs().max().item() <= tol


@pytest.mark.parametrize("shape", [(512, 2048), (3, 100, 64), (4, 2048, 64), (1, 67, 130), (2, 129, 63)])
def test_transpose_bit_exact(shape):
    from omni_avsr_b200 import ops
    g = torch.Generator().manual_seed(sum(shape))
    x = torch.randn(*shape, generator=g).bfloat16().cuda()
    got = ops.transpose(x)
    want = x.transpose(-1, -2).contiguous()
    assert got.shape == want.shape and torch.equal(got.view(torch.int16), want.view(torch.int16))


# ---------------------------------------------------------------------------------------------------------------
# ResNet-18 trunk convolutions on the tcgen05 GEMM (ring-padded channels-last frames, csrc/resnet_trunk.cu)
# ---------------------------------------------------------------------------------------------------------------
def _ring_from_nchw(x):
    from omni_avsr_b200 import ops
    N, C, H, W = x.shape
    r = ops.RingFrames(N, H, W, C, x.device, zero=True)
    r.rows.view(N, H + 2, W + 2, C)[:, 1:-1, 1:-1] = x.permute(0, 2, 3, 1)
    return r


def _nchw_from_ring(r):
    return r.rows.view(r.N, r.H + 2, r.W + 2, r.C)[:, 1:-1, 1:-1].permute(0, 3, 1, 2)


@pytest.mark.parametrize("N,H,W,Ci,Co", [(5, 22, 22, 64, 64), (3, 11, 11, 128, 128), (2, 6, 6, 256, 256), (3, 3, 3, 512, 512),
                                         (2, 7, 5, 64, 128), (4, 6, 6, 16, 32)])
def test_conv3x3_stride1_as_overlapping_row_gemm(N, H, W, Ci, Co):
    """3x3 / stride 1 / pad 1 convolution = one GEMM on the overlapping-row view of the ring-padded frames (main K = dy -1
    taps, K-extension blocks = dy 0 / +1 taps) vs torch conv2d in fp32; max|a-b| <= 1e-2 * max|b|.  (16 -> 32: the gather
    path of narrow test architectures.)"""
    from omni_avsr_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(N + H + Ci)
    x = torch.randn(N, Ci, H, W, device="cuda", generator=g).bfloat16()
    w = (torch.randn(Co, Ci, 3, 3, device="cuda", generator=g) / (3 * Ci ** 0.5)).bfloat16()
    want = torch.nn.functional.conv2d(x.float(), w.float(), padding=1)
    for g in (1, 2, 4):                      # g output pixels per GEMM row (wider N for the 64- / 128-channel layers)
        if g > 1 and ((g + 2) * Ci) % 64:
            continue
        wmat = ops.conv3x3_group_weights(w, g)
        if g == 1:
            assert torch.equal(wmat, w.permute(0, 2, 3, 1).reshape(Co, 9 * Ci))
        out = ops.conv3x3s1_ring(_ring_from_nchw(x), wmat, g)
        got = _nchw_from_ring(out).float()
        assert (got - want).abs().max().item() <= 1e-2 * want.abs().max().item(), g


@pytest.mark.parametrize("N,H,W,C,g", [(80, 22, 22, 64, 4), (120, 11, 11, 128, 2), (300, 6, 6, 256, 1), (900, 3, 3, 512, 1)])
@pytest.mark.parametrize("with_res", [False, True])
def test_conv_gemm_fused_prelu_ring_epilogue_is_bit_identical(N, H, W, C, g, with_res):
    """The BasicBlock tail in the epilogue of the CTA-pair convolution GEMM (OMNI_ACT_PRELU_RING: folded-BN shift, residual +
    its shift, PReLU, ring re-zeroing) == the unfused pair conv GEMM -> prelu_res_ring kernel, bit for bit."""
    from omni_avsr_b200 import ops
    gen = torch.Generator(device="cuda").manual_seed(C + H)
    x = torch.randn(N, C, H, W, device="cuda", generator=gen).bfloat16()
    w = (torch.randn(C, C, 3, 3, device="cuda", generator=gen) / (3 * C ** 0.5)).bfloat16()
    bias, rb = [(torch.randn(C, device="cuda", generator=gen) * 0.2).bfloat16() for _ in range(2)]
    slope = (torch.rand(C, device="cuda", generator=gen) * 0.5).bfloat16()
    res = _ring_from_nchw(torch.randn(N, C, H, W, device="cuda", generator=gen).bfloat16()) if with_res else None
    wmat = ops.conv3x3_group_weights(w, g)
    xin = _ring_from_nchw(x)
    launches = ops.LAUNCHES
    fused = ops.conv3x3s1_ring(xin, wmat, g, prelu=dict(slope=slope, bias=bias, residual=res, res_bias=rb if with_res else None))
    assert ops.LAUNCHES - launches == 1                      # one launch: the epilogue variant was taken
    fused_rows = fused.rows.clone()
    plain = ops.conv3x3s1_ring(xin, wmat, g)
    ops.prelu_res_ring_(plain, slope, res, bias=bias, res_bias=rb if with_res else None)
    assert torch.equal(fused_rows.view(torch.int16), plain.rows.view(torch.int16))


@pytest.mark.parametrize("N,H,W,Ci,Co", [(3, 22, 22, 64, 128), (2, 11, 11, 128, 256), (2, 6, 6, 256, 512), (3, 5, 8, 16, 32)])
def test_stride2_convs_prelu_ring_and_avgpool(N, H, W, Ci, Co):
    """The stride-2 3x3 convolution and the 1x1 stride-2 downsample (gather + GEMM on the ring-padded output grid), the
    PReLU / residual / folded-BN-shift kernel with ring re-zeroing and the final average pool, composed as one BasicBlock
    with a downsample branch (resnet.py:35-74), vs the same ops in fp32 torch."""
    from omni_avsr_b200 import ops
    F = torch.nn.functional
    g = torch.Generator(device="cuda").manual_seed(H * 31 + Ci)
    x = torch.randn(N, Ci, H, W, device="cuda", generator=g).bfloat16()
    w1 = (torch.randn(Co, Ci, 3, 3, device="cuda", generator=g) / (3 * Ci ** 0.5)).bfloat16()
    w2 = (torch.randn(Co, Co, 3, 3, device="cuda", generator=g) / (3 * Co ** 0.5)).bfloat16()
    wd = (torch.randn(Co, Ci, 1, 1, device="cuda", generator=g) / Ci ** 0.5).bfloat16()
    b1, b2, bd = [(torch.randn(Co, device="cuda", generator=g) * 0.1).bfloat16() for _ in range(3)]
    s1, s2 = [(torch.rand(Co, device="cuda", generator=g) * 0.5).bfloat16() for _ in range(2)]
    tap = lambda w: w.permute(0, 2, 3, 1).reshape(w.shape[0], -1).contiguous()
    y = _ring_from_nchw(x)
    o = ops.conv_s2_ring(y, tap(w1), 9)
    ops.prelu_res_ring_(o, s1, bias=b1)
    ring = o.rows.view(N, o.H + 2, o.W + 2, Co)
    assert (ring[:, 0] == 0).all() and (ring[:, -1] == 0).all() and (ring[:, :, 0] == 0).all() and (ring[:, :, -1] == 0).all()
    o2 = ops.conv3x3s1_ring(o, ops.conv3x3_group_weights(w2, 1))
    res = ops.conv_s2_ring(y, wd.reshape(Co, Ci).contiguous(), 1)
    out = ops.prelu_res_ring_(o2, s2, res, bias=b2, res_bias=bd)
    pooled = ops.avgpool_ring(out)
    xf = x.float()
    t = F.conv2d(xf, w1.float(), stride=2, padding=1) + b1.float().view(1, -1, 1, 1)
    t = F.prelu(t.bfloat16().float(), s1.float()).bfloat16().float()
    t = F.conv2d(t, w2.float(), padding=1).bfloat16().float() + b2.float().view(1, -1, 1, 1)
    r = F.conv2d(xf, wd.float(), stride=2).bfloat16().float() + bd.float().view(1, -1, 1, 1)
    want = F.prelu((t.bfloat16().float() + r.bfloat16().float()).bfloat16().float(), s2.float())
    got = _nchw_from_ring(out).float()
    assert got.shape == want.shape
    assert (got - want).abs().max().item() <= 2e-2 * want.abs().max().item()
    assert (pooled.float() - want.mean(dim=(2, 3))).abs().max().item() <= 2e-2 * want.mean(dim=(2, 3)).abs().max().item()


# ---------------------------------------------------------------------------------------------------------------
# Table-driven convolutions of the late trunk stages (ops.conv_frames: GEMM row = one frame, K-extension list per output pixel)
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("N,H,W,Ci,Co", [(300, 22, 22, 64, 128), (260, 11, 11, 128, 256), (700, 6, 6, 256, 512), (40, 11, 11, 128, 256),
                                         (130, 7, 5, 64, 128)])
def test_conv_frames_basic_blocks_and_avgpool(N, H, W, Ci, Co):
    """Two BasicBlocks (resnet.py:35-74) on the table-driven path: [3x3 stride-2 conv + PReLU, 1x1 stride-2 downsample, 3x3
    conv + residual + PReLU] reading ring-padded frames, then [3x3, 3x3 + identity residual] on plain frame rows, then the
    average pool -- vs the same ops in fp32 torch with the product's bf16 rounding points."""
    from omni_avsr_b200 import ops
    F = torch.nn.functional
    g = torch.Generator(device="cuda").manual_seed(H * 31 + Ci)
    x = torch.randn(N, Ci, H, W, device="cuda", generator=g).bfloat16()
    mk = lambda co, ci, k: (torch.randn(co, ci, k, k, device="cuda", generator=g) / (k * ci ** 0.5)).bfloat16()
    w1, w2, wd, w3, w4 = mk(Co, Ci, 3), mk(Co, Co, 3), mk(Co, Ci, 1), mk(Co, Co, 3), mk(Co, Co, 3)
    b1, b2, bd, b3, b4 = [(torch.randn(Co, device="cuda", generator=g) * 0.1).bfloat16() for _ in range(5)]
    s1, s2, s3, s4 = [(torch.rand(Co, device="cuda", generator=g) * 0.5).bfloat16() for _ in range(4)]
    y = _ring_from_nchw(x)
    c1 = ops.ConvFramesSpec(w1, H, W, 2, True)
    cd = ops.ConvFramesSpec(wd, H, W, 2, True)
    c2 = ops.ConvFramesSpec(w2, c1.Hout, c1.Wout, 1, False)
    c3 = ops.ConvFramesSpec(w3, c1.Hout, c1.Wout, 1, False)
    c4 = ops.ConvFramesSpec(w4, c1.Hout, c1.Wout, 1, False)
    launches = ops.LAUNCHES
    o = ops.conv_frames(y, c1, prelu=dict(slope=s1, bias=b1))
    res = ops.conv_frames(y, cd)
    blk1 = ops.conv_frames(o, c2, prelu=dict(slope=s2, bias=b2, residual=res, res_bias=bd))
    o = ops.conv_frames(blk1, c3, prelu=dict(slope=s3, bias=b3))
    blk2 = ops.conv_frames(o, c4, prelu=dict(slope=s4, bias=b4, residual=blk1))
    pooled = ops.avgpool_frames(blk2)
    assert ops.LAUNCHES - launches == 6                      # one launch per convolution + the pool

    rb = lambda t: t.bfloat16().float()
    bias = lambda b: b.float().view(1, -1, 1, 1)
    xf = x.float()
    t = F.prelu(rb(rb(F.conv2d(xf, w1.float(), stride=2, padding=1)) + bias(b1)), s1.float())
    r = rb(rb(F.conv2d(xf, wd.float(), stride=2)) + bias(bd))
    t = rb(rb(F.conv2d(rb(t), w2.float(), padding=1)) + bias(b2))
    want1 = rb(F.prelu(rb(t + r), s2.float()))
    t = rb(F.prelu(rb(rb(F.conv2d(want1, w3.float(), padding=1)) + bias(b3)), s3.float()))
    t = rb(rb(F.conv2d(t, w4.float(), padding=1)) + bias(b4))
    want2 = F.prelu(rb(t + want1), s4.float())

    def nchw(fr):
        return fr.buf.view(fr.N, fr.PA, fr.C)[:, : fr.H * fr.W].reshape(fr.N, fr.H, fr.W, fr.C).permute(0, 3, 1, 2).float()
    assert (nchw(blk1) - want1).abs().max().item() <= 2e-2 * want1.abs().max().item()
    assert (nchw(blk2) - want2).abs().max().item() <= 3e-2 * want2.abs().max().item()
    assert (pooled.float() - want2.mean(dim=(2, 3))).abs().max().item() <= 3e-2 * want2.mean(dim=(2, 3)).abs().max().item()
