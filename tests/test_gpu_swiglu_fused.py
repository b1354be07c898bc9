"""SwiGLU fused into the gate_up GEMM epilogue (OMNI_ACT_SWIGLU64) against the unfused gemm + swiglu_fwd pair and against
an fp32 torch reference.  The fused epilogue computes from the bf16-ROUNDED gate / up values, so the activation must be
bit-identical to the unfused kernels; the backward reorders the K dimension of the dgrad GEMM (interleaved columns), so
dx is compared with a tolerance (max|a-b| <= 1e-2*max|b|)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _bits(a, b):
    return torch.equal(a.view(torch.int16), b.view(torch.int16))


@pytest.mark.parametrize("M,H,I", [(9600, 256, 1024), (31232, 2048, 8192), (9473, 512, 1536)])
def test_fused_forward_is_bit_identical_and_backward_matches(M, H, I):
    from omni_avsr_b200 import autograd_ops as ag
    from omni_avsr_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(M)
    x = (torch.randn(M, H, device="cuda", generator=g) * 0.5).bfloat16()
    W = (torch.randn(2 * I, H, device="cuda", generator=g) * 0.05).bfloat16()
    assert ag.gate_up_swiglu_supported(M, 2 * I)
    W_il = ag.interleave_gate_up(W)
    # unfused
    gu = ops.gemm(x, W, block_n=256)
    act = ops.swiglu_fwd(gu)
    # fused
    xg = x.clone().requires_grad_(True)
    act_f = ag.GateUpSwigluFn.apply(xg, W_il, W_il.t().contiguous())
    assert _bits(act_f.detach(), act)
    # fp32 yardstick on a slice
    ref = torch.nn.functional.silu(x[:512].float() @ W[:I].float().t()) * (x[:512].float() @ W[I:].float().t())
    assert ((act_f[:512].float() - ref).abs().max() / ref.abs().max()).item() <= 2e-2
    # backward
    dact = (torch.randn(M, I, device="cuda", generator=g) * 0.1).bfloat16()
    act_f.backward(dact)
    dgu = ops.swiglu_bwd(dact, gu)
    dx = ops.gemm(dgu, W.t().contiguous(), block_n=256)
    rel = ((xg.grad.float() - dx.float()).abs().max() / dx.float().abs().max()).item()
    assert rel <= 1e-2, rel


def test_blocked_swiglu_bwd_equals_plain_on_permuted_columns():
    from omni_avsr_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(1)
    rows, I = 777, 512
    gu = torch.randn(rows, 2 * I, device="cuda", generator=g).bfloat16()
    dact = torch.randn(rows, I, device="cuda", generator=g).bfloat16()
    want = ops.swiglu_bwd(dact, gu)
    il = lambda t: torch.stack((t[:, :I].reshape(rows, I // 64, 64), t[:, I:].reshape(rows, I // 64, 64)), dim=2).reshape(rows, 2 * I).contiguous()
    got = ops.swiglu_bwd(dact, il(gu), blk=64)
    assert _bits(got, il(want))


def test_unsupported_shapes_are_refused():
    from omni_avsr_b200 import ops
    from omni_avsr_b200._lib import OmniKernelError
    x = torch.zeros(64, 256, device="cuda", dtype=torch.bfloat16)          # one row tile: not a CTA-pair launch
    W = torch.zeros(512, 256, device="cuda", dtype=torch.bfloat16)
    with pytest.raises(OmniKernelError):
        ops.gemm(x, W, act="swiglu64", block_n=256, out2=torch.empty(64, 256, device="cuda", dtype=torch.bfloat16))


def test_gelu_keep_epilogue_matches_separate_kernels():
    """OMNI_ACT_GELU_KEEP: pre-activation bit-identical to the plain bias GEMM, activation bit-identical to the fused-GELU
    GEMM (same erf formula), and within 1e-2 of torch's exact GELU; backward through FrozenLinearGeluFn matches the unfused
    chain bit for bit (same kernels, same order)."""
    from omni_avsr_b200 import autograd_ops as ag
    from omni_avsr_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(5)
    M, K, N = 12800, 1024, 4096
    x = (torch.randn(M, K, device="cuda", generator=g) * 0.5).bfloat16()
    W = (torch.randn(N, K, device="cuda", generator=g) * 0.05).bfloat16()
    b = (torch.randn(N, device="cuda", generator=g) * 0.1).bfloat16()
    assert ag.pair_kernel_shape(M, N)
    pre = ops.gemm(x, W, bias=b, block_n=256)
    act = ops.gemm(x, W, bias=b, act="gelu", block_n=256)
    xg = x.clone().requires_grad_(True)
    WT = W.t().contiguous()
    act_f = ag.FrozenLinearGeluFn.apply(xg, W, WT, b)
    assert _bits(act_f.detach(), act)
    ref = torch.nn.functional.gelu(pre[:256].float())
    assert ((act_f[:256].float() - ref).abs().max() / ref.abs().max()).item() <= 1e-2
    dact = (torch.randn(M, N, device="cuda", generator=g) * 0.1).bfloat16()
    act_f.backward(dact)
    dx = ops.gemm(ops.gelu_bwd(dact, pre), WT, block_n=256)
    assert _bits(xg.grad, dx)


@pytest.mark.parametrize("M,H,I", [(9600, 256, 1024), (31232, 2048, 8192), (9473, 512, 1536)])
def test_swiglu_backward_epilogue_matches_separate_kernels(M, H, I):
    """OMNI_ACT_SWIGLU_BWD64: the dgrad GEMM of down_proj with the SwiGLU backward in its epilogue == dgrad GEMM ->
    omni_swiglu_bwd_blocked.  Same bf16-rounded d(act), same formula; the epilogue's sigmoid uses the hardware reciprocal, so
    single bf16 ulps may differ: at most 0.1 % of the elements, each within 2^-7 relative, everything else bit-identical.
    Then the whole MLP (MlpSwigluFn) against the unfused autograd chain: dx within 1e-2, d(residual) bit-identical."""
    from omni_avsr_b200 import autograd_ops as ag
    from omni_avsr_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(M + 1)
    dy = (torch.randn(M, H, device="cuda", generator=g) * 0.1).bfloat16()
    Wd = (torch.randn(H, I, device="cuda", generator=g) * 0.05).bfloat16()          # down_proj.weight [H, I]
    gu = torch.randn(M, 2 * I, device="cuda", generator=g).bfloat16()               # interleaved gate|up blocks
    WTd = Wd.t().contiguous()
    dact = ops.gemm(dy, WTd, block_n=256)
    want = ops.swiglu_bwd(dact, gu, blk=ops.SWIGLU_BLK)
    got = torch.empty_like(gu)
    launches = ops.LAUNCHES
    ops.gemm(dy, WTd, residual=gu, out=got, act="swiglu_bwd64")
    assert ops.LAUNCHES - launches == 1
    diff = got.view(torch.int16) != want.view(torch.int16)
    assert diff.float().mean().item() <= 1e-3, diff.float().mean().item()
    rel = ((got.float() - want.float()).abs() / want.float().abs().clamp_min(1e-6))[diff]
    assert rel.numel() == 0 or rel.max().item() <= 2 ** -7, rel.max().item()

    if I % 256 == 0 and ag.pair_kernel_shape(M, I):
        x = (torch.randn(M, H, device="cuda", generator=g) * 0.5).bfloat16()
        W = (torch.randn(2 * I, H, device="cuda", generator=g) * 0.05).bfloat16()
        res = torch.randn(M, H, device="cuda", generator=g).bfloat16()
        W_il = ag.interleave_gate_up(W)
        WT_il = W_il.t().contiguous()
        xa, ra = x.clone().requires_grad_(True), res.clone().requires_grad_(True)
        ya = ag.MlpSwigluFn.apply(xa, W_il, WT_il, Wd, WTd, ra)
        xb, rb = x.clone().requires_grad_(True), res.clone().requires_grad_(True)
        yb = ag.frozen_linear(ag.GateUpSwigluFn.apply(xb, W_il, WT_il), Wd, WTd, residual=rb, block_n=256)
        assert _bits(ya.detach(), yb.detach())
        ya.backward(dy)
        yb.backward(dy)
        assert _bits(ra.grad, rb.grad)
        assert ((xa.grad.float() - xb.grad.float()).abs().max() / xb.grad.float().abs().max()).item() <= 1e-2


def test_gelu_backward_epilogue_matches_separate_kernels():
    """OMNI_ACT_GELU_BWD: the dgrad GEMM of fc2 with the GELU backward in its epilogue == dgrad GEMM -> omni_gelu_bwd, bit for
    bit (same erff / __expf formula on the same bf16-rounded d(act)); FfnGeluFn == the unfused autograd chain."""
    from omni_avsr_b200 import autograd_ops as ag
    from omni_avsr_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(9)
    M, E, Fd = 12800, 1024, 4096
    dy = (torch.randn(M, E, device="cuda", generator=g) * 0.1).bfloat16()
    W2 = (torch.randn(E, Fd, device="cuda", generator=g) * 0.03).bfloat16()         # fc2.weight [E, ffn]
    pre = torch.randn(M, Fd, device="cuda", generator=g).bfloat16()
    WT2 = W2.t().contiguous()
    want = ops.gelu_bwd(ops.gemm(dy, WT2, block_n=256), pre)
    got = torch.empty_like(pre)
    ops.gemm(dy, WT2, residual=pre, out=got, act="gelu_bwd")
    assert _bits(got, want)

    x = (torch.randn(M, E, device="cuda", generator=g) * 0.5).bfloat16()
    W1 = (torch.randn(Fd, E, device="cuda", generator=g) * 0.03).bfloat16()
    b1 = (torch.randn(Fd, device="cuda", generator=g) * 0.1).bfloat16()
    b2 = (torch.randn(E, device="cuda", generator=g) * 0.1).bfloat16()
    res = torch.randn(M, E, device="cuda", generator=g).bfloat16()
    WT1 = W1.t().contiguous()
    xa, ra = x.clone().requires_grad_(True), res.clone().requires_grad_(True)
    ya = ag.FfnGeluFn.apply(xa, W1, WT1, b1, W2, WT2, b2, ra)
    xb, rb = x.clone().requires_grad_(True), res.clone().requires_grad_(True)
    yb = ag.frozen_linear(ag.FrozenLinearGeluFn.apply(xb, W1, WT1, b1), W2, WT2, bias=b2, residual=rb, block_n=256)
    assert _bits(ya.detach(), yb.detach())
    ya.backward(dy)
    yb.backward(dy)
    assert _bits(ra.grad, rb.grad) and _bits(xa.grad, xb.grad)


def test_swiglu_epilogue_without_gate_up_output():
    """Inference form of OMNI_ACT_SWIGLU64 (out = NULL): only the activation is written, bit-identical to the training form."""
    from omni_avsr_b200 import autograd_ops as ag
    from omni_avsr_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(11)
    M, H, I = 9473, 512, 1536
    x = (torch.randn(M, H, device="cuda", generator=g) * 0.5).bfloat16()
    W_il = ag.interleave_gate_up((torch.randn(2 * I, H, device="cuda", generator=g) * 0.05).bfloat16())
    gu = torch.empty(M, 2 * I, device="cuda", dtype=torch.bfloat16)
    act = torch.empty(M, I, device="cuda", dtype=torch.bfloat16)
    ops.gemm(x, W_il, out=gu, out2=act, act="swiglu64", block_n=256)
    act2 = torch.full((M, I), float("nan"), device="cuda", dtype=torch.bfloat16)
    ret = ops.gemm(x, W_il, out2=act2, act="swiglu64", block_n=256)
    assert ret is act2 and _bits(act2, act)
