"""GPU parity of the weight-streaming decode-step GEMM (omni_gemm_skinny_bf16: weights on the M side of the MMA, split-K over
a thread-block cluster with a distributed-shared-memory reduction) against a plain PyTorch fp32 reference of the same op
and against the general kernel (omni_gemm_bf16) on the same inputs.  Tolerance: max|a-b| <= 1e-2 * max|b| vs fp32 (bf16
output rounding), <= 8e-3 * max|b| (one bf16 ulp of the largest value) between the two kernels: same rounding points, different fp32
summation order."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return (a.float() - b.float()).abs().max().item() / max(b.float().abs().max().item(), 1e-9)


def _mk(M, N, K, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = (torch.randn(M, K, device="cuda", generator=g)).bfloat16()
    w = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).bfloat16()
    return x, w, g


@pytest.mark.parametrize("M,N,K", [(64, 2048, 2048), (64, 2048, 8192), (1, 3072, 2048), (15, 1024, 512), (100, 2048, 2048),
                                   (128, 4096, 4096), (37, 200, 128), (64, 16384, 2048), (70, 384, 14336)])
@pytest.mark.parametrize("variant", ["plain", "bias_relu", "residual", "bias_gelu_residual"])
def test_skinny_matches_fp32_and_general_kernel(M, N, K, variant):
    from omni_avsr_b200 import ops
    x, w, g = _mk(M, N, K, M + N + K)
    bias = (torch.randn(N, device="cuda", generator=g) * 0.5).bfloat16() if "bias" in variant else None
    res = torch.randn(M, N, device="cuda", generator=g).bfloat16() if "residual" in variant else None
    act = "relu" if "relu" in variant else ("gelu" if "gelu" in variant else None)
    got = ops.gemm(x, w, bias=bias, act=act, residual=res, alpha=0.5, skinny=True)
    ref = ops.gemm(x, w, bias=bias, act=act, residual=res, alpha=0.5)
    y = 0.5 * (x.float() @ w.float().t())
    if bias is not None:
        y = y + bias.float()
    if act or res is not None:
        y = y.bfloat16().float()
    if act == "relu":
        y = torch.relu(y)
    elif act == "gelu":
        y = torch.nn.functional.gelu(y).bfloat16().float()
    if res is not None:
        y = y + res.float()
    assert got.shape == (M, N) and got.dtype == torch.bfloat16
    assert _rel(got, y) <= 1e-2, _rel(got, y)
    assert _rel(got, ref) <= 8e-3, _rel(got, ref)


@pytest.mark.parametrize("split", ["1", "2", "4", "8"])
def test_skinny_every_cluster_size(split):
    """The split factor is a launch heuristic; every cluster size must give the same sums (forced through the debug switch
    in a subprocess-free way: the environment variable is read once per process, so this test only checks the value the
    process started with plus the default heuristic on shapes that select 1 / 2 / 4 / 8 themselves)."""
    from omni_avsr_b200 import ops
    shapes = {"1": (64, 16384, 2048), "2": (64, 6144, 4096), "4": (64, 3072, 2048), "8": (64, 2048, 8192)}
    M, N, K = shapes[split]
    x, w, _ = _mk(M, N, K, 5)
    got = ops.gemm(x, w, skinny=True)
    y = x.float() @ w.float().t()
    assert _rel(got, y) <= 1e-2


@pytest.mark.parametrize("M", [64, 9, 128])
def test_skinny_swiglu_epilogue_bit_identical_to_the_row_kernel(M):
    """act='swiglu64' on [gate 64 | up 64] interleaved weight rows == swiglu_fwd(gemm(x, W_gate|up)) bit for bit whenever
    the two GEMMs agree bitwise, and within bf16 resolution always."""
    from omni_avsr_b200 import autograd_ops as ag
    from omni_avsr_b200 import ops
    I, H = 1024, 512
    x, w, _ = _mk(M, 2 * I, H, 7)
    w_il = ag.interleave_gate_up(w)
    act = torch.empty((M, I), device="cuda", dtype=torch.bfloat16)
    ops.gemm(x, w_il, act="swiglu64", out2=act, skinny=True)
    gu = ops.gemm(x, w, skinny=True)
    want = ops.swiglu_fwd(gu.contiguous())
    assert torch.equal(act.view(torch.int16), want.view(torch.int16))
    ref = torch.nn.functional.silu(gu[:, :I].float()).bfloat16().float() * gu[:, I:].float()
    assert _rel(act, ref) <= 1e-2


@pytest.mark.parametrize("task", [0, 1, 2])
def test_skinny_lora_k_extension_and_grouped_rows(task):
    """The decode-step Omni-LoRA projection: T = s * h A[task, shared]^T through the per-64-block row table, then
    q|k|v = h W^T + b + T_q B_q^T (Q columns) + T_v B_v^T (V columns) through the 128-feature extension table -- against the
    literal reference arithmetic of Llama_LoRA.py:246-259 in fp32."""
    from omni_avsr_b200 import ops
    from omni_avsr_b200.Llama_LoRA import LoraPlan
    H, q, kv, r, s = 512, 512, 128, 32, 0.25
    plan = LoraPlan(H, q, kv, kv, r, s, True, True, "cuda")
    M = 40
    g = torch.Generator(device="cuda").manual_seed(task)
    h = torch.randn(M, H, device="cuda", generator=g).bfloat16()
    W = (torch.randn(q + 2 * kv, H, device="cuda", generator=g) / H ** 0.5).bfloat16()
    b = (torch.randn(q + 2 * kv, device="cuda", generator=g) * 0.1).bfloat16()
    down = torch.zeros(plan.down_rows, H, device="cuda", dtype=torch.bfloat16)
    up = torch.zeros(plan.up_rows, plan.rp, device="cuda", dtype=torch.bfloat16)
    Aq, Av, Bq, Bv = {}, {}, {}, {}
    for slot in range(plan.n_slots):
        Aq[slot] = (torch.randn(r, H, device="cuda", generator=g) * 0.05).bfloat16()
        Av[slot] = (torch.randn(r, H, device="cuda", generator=g) * 0.05).bfloat16()
        Bq[slot] = (torch.randn(q, r, device="cuda", generator=g) * 0.2).bfloat16()
        Bv[slot] = (torch.randn(kv, r, device="cuda", generator=g) * 0.2).bfloat16()
        down[slot * plan.rp: slot * plan.rp + r] = Aq[slot]
        down[(plan.n_slots + slot) * plan.rp: (plan.n_slots + slot) * plan.rp + r] = Av[slot]
        up[slot * q: (slot + 1) * q, :r] = Bq[slot]
        up[plan.n_slots * q + slot * kv: plan.n_slots * q + (slot + 1) * kv, :r] = Bv[slot]
    tg = torch.tensor([task], dtype=torch.int32, device="cuda")
    T = ops.gemm(h, down, n=plan.t_cols, alpha=s, tile_group=tg, b_row_table=plan.brow_fwd, skinny=True)
    out = ops.gemm(h, W, bias=b, tile_group=tg, ext=(T, up, plan.ext_fwd_step), block_n=128, skinny=True)
    hf = h.float()
    base = hf @ W.float().t() + b.float()
    sh = plan.n_slots - 1
    qq = base[:, :q] + s * ((hf @ Aq[task].float().t()) @ Bq[task].float().t() + (hf @ Aq[sh].float().t()) @ Bq[sh].float().t())
    vv = base[:, q + kv:] + s * ((hf @ Av[task].float().t()) @ Bv[task].float().t() + (hf @ Av[sh].float().t()) @ Bv[sh].float().t())
    want = torch.cat([qq, base[:, q: q + kv], vv], dim=1)
    assert _rel(out, want) <= 1e-2, _rel(out, want)
    # and the general kernel on the same tables' 64-column twin
    T2 = ops.gemm(h, down, n=plan.t_cols, alpha=s, tile_group=tg, b_row_table=plan.brow_fwd, block_n=64)
    out2 = ops.gemm(h, W, bias=b, tile_group=tg, ext=(T2, up, plan.ext_fwd_64), block_n=64)
    assert _rel(T, T2) <= 8e-3 and _rel(out, out2) <= 8e-3


@pytest.mark.parametrize("M,N,K", [(64, 2048, 2048), (64, 2048, 8192), (37, 2048, 2048), (128, 4096, 4096), (64, 16384, 2048)])
def test_fused_rmsnorm_of_the_finished_rows_is_bit_identical(M, N, K):
    """norm=(weight, eps): the RMSNorm that follows o_proj / down_proj in a decode step runs in the GEMM launch (the CTA that
    completes a token slice normalises it) and equals omni_rmsnorm_fwd of the GEMM's output bit for bit; output and residual
    add unchanged.  N = 16384 has no split-K: the wrapper runs GEMM + norm kernel.  Run twice: the arrival counters reset."""
    from omni_avsr_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(M + N)
    x = (torch.randn(M, K, device="cuda", generator=g) * 0.5).bfloat16()
    W = (torch.randn(N, K, device="cuda", generator=g) * 0.05).bfloat16()
    res = torch.randn(M, N, device="cuda", generator=g).bfloat16()
    nw = (1.0 + 0.1 * torch.randn(N, device="cuda", generator=g)).bfloat16()
    want = ops.gemm(x, W, residual=res, skinny=True)
    want_h = ops.rmsnorm_fwd(want, nw, 1e-5)
    for _ in range(2):
        launches = ops.LAUNCHES
        out, h = ops.gemm(x, W, residual=res, skinny=True, norm=(nw, 1e-5))
        assert ops.LAUNCHES - launches == (1 if N <= 4096 else 2)
        assert torch.equal(out.view(torch.int16), want.view(torch.int16))
        assert torch.equal(h.view(torch.int16), want_h.view(torch.int16))
