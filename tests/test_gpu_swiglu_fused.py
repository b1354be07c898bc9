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
