"""torch.autograd.Function wrappers over the C-ABI kernels (forward AND backward run our kernels; torch only
owns the device buffers and the autograd tape).  All tensors are 2-D [rows, features] bf16 on CUDA."""
from __future__ import annotations

from typing import Optional

import torch

from . import ops


class FrozenLinearFn(torch.autograd.Function):
    """y = x @ W^T (+bias) (+residual) with a FROZEN weight; backward dx = dy @ W via the pre-transposed copy WT."""

    @staticmethod
    def forward(ctx, x, W, WT, bias, residual, block_n):
        ctx.WT = WT
        ctx.block_n = block_n
        ctx.has_res = residual is not None
        return ops.gemm(x, W, bias=bias, residual=residual, block_n=block_n)

    @staticmethod
    def backward(ctx, dy):
        dy = dy.contiguous()
        dx = ops.gemm(dy, ctx.WT, block_n=ctx.block_n) if ctx.needs_input_grad[0] else None
        return dx, None, None, None, (dy if ctx.has_res else None), None


def skinny_ok(x, residual=None) -> bool:
    """Decode-step shapes (<= 128 token rows, nothing to differentiate, K a multiple of 64): the weight-streaming kernel."""
    return (x.shape[0] <= 128 and x.shape[1] % 64 == 0 and not x.requires_grad and not torch.is_grad_enabled()
            and not (residual is not None and residual.requires_grad))


def frozen_linear(x, W, WT, bias=None, residual=None, block_n=0):
    if not (x.requires_grad or (residual is not None and residual.requires_grad)):
        return ops.gemm(x, W, bias=bias, residual=residual, block_n=block_n)
    return FrozenLinearFn.apply(x, W, WT, bias, residual, block_n)


class TrainableLinearFn(torch.autograd.Function):
    """Projector linear: y = act(x @ W^T + b); W, b trainable.  dx via our GEMM on W^T (transposed per call, the
    weight changes every step); dW = dy^T x runs on the MN-major tcgen05 weight-gradient kernel."""

    @staticmethod
    def forward(ctx, x, W, b, act):
        y = ops.gemm(x, W, bias=b, act=act)
        ctx.save_for_backward(x, W, y if act == "relu" else None)
        ctx.act = act
        return y

    @staticmethod
    def backward(ctx, dy):
        x, W, y = ctx.saved_tensors
        dy = dy.contiguous()
        if ctx.act == "relu":
            dy = dy * (y > 0)
        dx = ops.gemm(dy, ops.transpose(W.detach().contiguous())) if ctx.needs_input_grad[0] else None
        dW = ops.gemm_wgrad(dy, x, mo=W.shape[0], no=W.shape[1])[0]       # dY^T X, both operands token-major
        db = ops.colsum(dy)
        return dx, dW, db, None


class RMSNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, eps):
        y, rstd = ops.rmsnorm_fwd(x, w, eps, want_rstd=True)
        ctx.save_for_backward(x, w, rstd)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w, rstd = ctx.saved_tensors
        return ops.rmsnorm_bwd(dy, x, w, rstd), None, None


def rmsnorm(x, w, eps):
    if not x.requires_grad:
        return ops.rmsnorm_fwd(x, w, eps)
    return RMSNormFn.apply(x, w, eps)


class RMSNormResidualFn(torch.autograd.Function):
    """Pre-norm block entry: returns (x, norm(x)) where the first output is the residual branch.  Autograd would
    otherwise add the two gradients of x (residual path + norm path) with a separate elementwise kernel per block; here
    the backward hands the residual-path gradient to the norm-backward kernel, which adds it on the way out."""

    @staticmethod
    def forward(ctx, x, w, eps):
        y, rstd = ops.rmsnorm_fwd(x, w, eps, want_rstd=True)
        ctx.save_for_backward(x, w, rstd)
        return x.view_as(x), y

    @staticmethod
    def backward(ctx, dres, dy):
        x, w, rstd = ctx.saved_tensors
        if dy is None:
            return dres, None, None
        return ops.rmsnorm_bwd(dy, x, w, rstd, dx_add=dres), None, None


def rmsnorm_residual(x, w, eps):
    """(residual, normed) with the fused gradient add; plain forward when no gradient is needed."""
    if not x.requires_grad:
        return x, ops.rmsnorm_fwd(x, w, eps)
    return RMSNormResidualFn.apply(x, w, eps)


class LayerNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b, eps):
        y, mean, rstd = ops.layernorm_fwd(x, w, b, eps, want_stats=True)
        ctx.save_for_backward(x, w, mean, rstd)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w, mean, rstd = ctx.saved_tensors
        return ops.layernorm_bwd(dy, x, w, mean, rstd), None, None, None


def layernorm(x, w, b, eps):
    if not x.requires_grad:
        return ops.layernorm_fwd(x, w, b, eps)
    return LayerNormFn.apply(x, w, b, eps)


class TrainableLayerNormFn(torch.autograd.Function):
    """LayerNorm with a TRAINABLE affine (the projector's nn.LayerNorm(hidden), modeling_OmniAVSR.py:85,97,111): forward and
    dx on the row kernels; dw = sum_rows(dy * xhat), db = sum_rows(dy) through the column-sum kernel."""

    @staticmethod
    def forward(ctx, x, w, b, eps):
        y, mean, rstd = ops.layernorm_fwd(x, w.detach(), b.detach(), eps, want_stats=True)
        ctx.save_for_backward(x, w, mean, rstd)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w, mean, rstd = ctx.saved_tensors
        dy = dy.contiguous()
        dx = ops.layernorm_bwd(dy, x, w.detach(), mean, rstd) if ctx.needs_input_grad[0] else None
        xhat = ((x.float() - mean[:, None]) * rstd[:, None])
        dw = ops.colsum((dy.float() * xhat).to(torch.bfloat16))
        db = ops.colsum(dy)
        return dx, dw, db, None


class LayerNormResidualFn(torch.autograd.Function):
    """LayerNorm twin of RMSNormResidualFn (AV-HuBERT pre-norm blocks)."""

    @staticmethod
    def forward(ctx, x, w, b, eps):
        y, mean, rstd = ops.layernorm_fwd(x, w, b, eps, want_stats=True)
        ctx.save_for_backward(x, w, mean, rstd)
        return x.view_as(x), y

    @staticmethod
    def backward(ctx, dres, dy):
        x, w, mean, rstd = ctx.saved_tensors
        if dy is None:
            return dres, None, None, None
        return ops.layernorm_bwd(dy, x, w, mean, rstd, dx_add=dres), None, None, None


def layernorm_residual(x, w, b, eps):
    if not x.requires_grad:
        return x, ops.layernorm_fwd(x, w, b, eps)
    return LayerNormResidualFn.apply(x, w, b, eps)


class RopeFn(torch.autograd.Function):
    """RoPE applied in place on the q|k heads of the packed qkv buffer (the input buffer is consumed)."""

    @staticmethod
    def forward(ctx, qkv, cos_t, sin_t, pos, n_heads_total, head_dim, grad_inplace=False):
        ctx.save_for_backward(cos_t, sin_t, pos)
        ctx.meta = (n_heads_total, head_dim)
        ctx.grad_inplace = grad_inplace   # the consumer hands us a gradient buffer it owns (PackedSdpaFn)
        ctx.mark_dirty(qkv)
        ops.rope_(qkv, cos_t, sin_t, pos, n_heads_total, head_dim)
        return qkv

    @staticmethod
    def backward(ctx, d):
        cos_t, sin_t, pos = ctx.saved_tensors
        if not (ctx.grad_inplace and d.is_contiguous()):
            d = d.clone(memory_format=torch.contiguous_format)
        ops.rope_(d, cos_t, sin_t, pos, ctx.meta[0], ctx.meta[1], inverse=True)
        return d, None, None, None, None, None, None


class SwigluFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, gu):
        ctx.save_for_backward(gu)
        return ops.swiglu_fwd(gu)

    @staticmethod
    def backward(ctx, dact):
        (gu,) = ctx.saved_tensors
        return ops.swiglu_bwd(dact, gu)


def swiglu(gu):
    return SwigluFn.apply(gu) if gu.requires_grad else ops.swiglu_fwd(gu)


def interleave_gate_up(w_gu: torch.Tensor, blk: int = ops.SWIGLU_BLK) -> torch.Tensor:
    """[gate I rows ; up I rows] -> [gate blk ; up blk ; gate blk ; ...] (the row order OMNI_ACT_SWIGLU64 expects)."""
    I = w_gu.shape[0] // 2
    g = w_gu[:I].reshape(I // blk, blk, -1)
    u = w_gu[I:].reshape(I // blk, blk, -1)
    return torch.stack((g, u), dim=1).reshape(2 * I, -1).contiguous()


class GateUpSwigluFn(torch.autograd.Function):
    """act = silu(x Wg^T) * (x Wu^T) with FROZEN weights in ONE GEMM launch: the SwiGLU runs in the epilogue of the gate_up
    GEMM (block-interleaved weight rows), which also writes the gate|up tile the backward needs.  Backward:
    d(gate|up) = swiglu_bwd(dact, gate|up) on the interleaved layout, dx = d(gate|up) @ W_il (pre-transposed copy)."""

    @staticmethod
    def forward(ctx, x, W_il, WT_il):
        M, N = x.shape[0], W_il.shape[0]
        gu = torch.empty((M, N), device=x.device, dtype=torch.bfloat16)
        act = torch.empty((M, N // 2), device=x.device, dtype=torch.bfloat16)
        ops.gemm(x, W_il, out=gu, out2=act, act="swiglu64", block_n=256)
        ctx.save_for_backward(gu)
        ctx.WT = WT_il
        return act

    @staticmethod
    def backward(ctx, dact):
        (gu,) = ctx.saved_tensors
        dgu = ops.swiglu_bwd(dact, gu, blk=ops.SWIGLU_BLK)
        return ops.gemm(dgu, ctx.WT, block_n=256), None, None


class MlpSwigluFn(torch.autograd.Function):
    """y = residual + (silu(x Wg^T) * (x Wu^T)) W_down^T with FROZEN weights (LlamaMLP + the decoder layer's residual add,
    Llama_LoRA.py decoder layer): forward = gate_up GEMM with the SwiGLU epilogue, down GEMM with the residual epilogue;
    backward = TWO launches: the dgrad GEMM of down_proj with the SwiGLU-backward epilogue (d(act) never reaches memory, the
    saved gate|up tile is read by the epilogue) and the dgrad GEMM of gate_up."""

    @staticmethod
    def forward(ctx, x, W_il, WT_il, W_down, WT_down, residual):
        M, N = x.shape[0], W_il.shape[0]
        gu = torch.empty((M, N), device=x.device, dtype=torch.bfloat16)
        act = torch.empty((M, N // 2), device=x.device, dtype=torch.bfloat16)
        ops.gemm(x, W_il, out=gu, out2=act, act="swiglu64", block_n=256)
        ctx.save_for_backward(gu)
        ctx.WT, ctx.WTd, ctx.has_res = WT_il, WT_down, residual is not None
        return ops.gemm(act, W_down, residual=residual, block_n=256)

    @staticmethod
    def backward(ctx, dy):
        (gu,) = ctx.saved_tensors
        dy = dy.contiguous()
        dgu = torch.empty_like(gu)
        ops.gemm(dy, ctx.WTd, residual=gu, out=dgu, act="swiglu_bwd64")
        dx = ops.gemm(dgu, ctx.WT, block_n=256) if ctx.needs_input_grad[0] else None
        return dx, None, None, None, None, (dy if ctx.has_res else None)


class FfnGeluFn(torch.autograd.Function):
    """y = residual + gelu(x W1^T + b1) W2^T + b2 with FROZEN weights while something upstream trains (the AV-HuBERT FFN under
    LoRA fine-tuning, wav2vec2.py:1001-1006): forward = fc1 GEMM writing pre-activation + activation, fc2 GEMM with the residual
    epilogue; backward = the dgrad GEMM of fc2 with the GELU-backward epilogue + the dgrad GEMM of fc1."""

    @staticmethod
    def forward(ctx, x, W1, WT1, b1, W2, WT2, b2, residual):
        M, N = x.shape[0], W1.shape[0]
        pre = torch.empty((M, N), device=x.device, dtype=torch.bfloat16)
        act = torch.empty((M, N), device=x.device, dtype=torch.bfloat16)
        ops.gemm(x, W1, bias=b1, out=pre, out2=act, act="gelu_keep", block_n=256)
        ctx.save_for_backward(pre)
        ctx.WT1, ctx.WT2, ctx.has_res = WT1, WT2, residual is not None
        return ops.gemm(act, W2, bias=b2, residual=residual, block_n=256)

    @staticmethod
    def backward(ctx, dy):
        (pre,) = ctx.saved_tensors
        dy = dy.contiguous()
        dpre = torch.empty_like(pre)
        ops.gemm(dy, ctx.WT2, residual=pre, out=dpre, act="gelu_bwd")
        dx = ops.gemm(dpre, ctx.WT1, block_n=256) if ctx.needs_input_grad[0] else None
        return dx, None, None, None, None, None, None, (dy if ctx.has_res else None)


def gate_up_swiglu_supported(M: int, N: int) -> bool:
    """Shapes the CTA-pair kernel takes for the fused epilogue (mirrors the dispatch in csrc/gemm_tcgen05.cu)."""
    return pair_kernel_shape(M, N)


class GeluFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        ctx.save_for_backward(x)
        return ops.gelu_fwd(x)

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        return ops.gelu_bwd(dy, x)


def gelu(x):
    return GeluFn.apply(x) if x.requires_grad else ops.gelu_fwd(x)


class FrozenLinearGeluFn(torch.autograd.Function):
    """gelu(x @ W^T + b) with a FROZEN weight while something upstream trains (AV-HuBERT fc1 under LoRA fine-tuning): ONE
    GEMM launch writes the pre-activation (saved for the backward) and the activation; backward = gelu_bwd + dgrad GEMM."""

    @staticmethod
    def forward(ctx, x, W, WT, bias):
        M, N = x.shape[0], W.shape[0]
        pre = torch.empty((M, N), device=x.device, dtype=torch.bfloat16)
        act = torch.empty((M, N), device=x.device, dtype=torch.bfloat16)
        ops.gemm(x, W, bias=bias, out=pre, out2=act, act="gelu_keep", block_n=256)
        ctx.save_for_backward(pre)
        ctx.WT = WT
        return act

    @staticmethod
    def backward(ctx, dact):
        (pre,) = ctx.saved_tensors
        return ops.gemm(ops.gelu_bwd(dact.contiguous(), pre), ctx.WT, block_n=256), None, None, None


def pair_kernel_shape(M: int, N: int) -> bool:
    """Shapes omni_gemm_bf16 runs on the CTA-pair kernel with 256-column tiles (the only one with the fused epilogues)."""
    return N % 256 == 0 and M > 128 and ((M + 127) // 128) * (N // 256) >= 74


class LoraLinearFn(torch.autograd.Function):
    """Omni-LoRA adapted fused projection (q|k|v of the LLM, or q|k|v of an AV-HuBERT block):

        T   = s * h @ down[sel(task)]^T                   (phase 1, grouped tcgen05 GEMM, N = n_active*2*rp)
        out = h @ W^T (+bias) + T_q @ up_q[sel]^T (Q cols) + T_v @ up_v[sel]^T (V cols)   (phase 2: K-extension)

    Reference math: Llama_LoRA.py:246-259 / Qwen_LoRA.py:557-570 / multihead_attention.py:485-494.
    `plan` carries the static tables (see LoraPlan)."""

    @staticmethod
    def forward(ctx, h, W, WT, bias, down, up, rows, plan):
        tile_group = rows.tile_group
        T = ops.gemm(h, down, n=plan.t_cols, alpha=plan.scaling, tile_group=tile_group, b_row_table=plan.brow_fwd,
                     block_n=64)
        if h.shape[0] <= 128 and not h.requires_grad:
            # one tile of rows: 64-column tiles put 4x the CTAs on the weight stream (N = 3072: 48, not 12)
            out = ops.gemm(h, W, bias=bias, tile_group=tile_group, ext=(T, up, plan.ext_fwd_64), block_n=64)
        else:
            out = ops.gemm(h, W, bias=bias, tile_group=tile_group, ext=(T, up, plan.ext_fwd), block_n=plan.block_n,
                           pair_aligned=getattr(rows, "pair_aligned", False))
        ctx.save_for_backward(h, T, down, up)
        ctx.WT, ctx.plan, ctx.rows = WT, plan, rows
        return out

    @staticmethod
    def backward(ctx, dout):
        h, T, down, up = ctx.saved_tensors
        plan, tile_group = ctx.plan, ctx.rows.tile_group
        dout = dout.contiguous()
        M = h.shape[0]
        # dT' = s * dOut_{q|v} @ up[sel]   (two grouped GEMMs on the transposed up-projections)
        upT_q, upT_v = plan.transposed_up(up)
        dT = torch.empty((M, plan.t_cols), device=h.device, dtype=torch.bfloat16)
        half = plan.t_cols // 2
        ops.gemm(dout[:, : plan.q_cols], upT_q, out=dT[:, :half], n=half, alpha=plan.scaling, tile_group=tile_group,
                 b_row_table=plan.brow_bwd, block_n=64)
        v0 = plan.v_col0
        ops.gemm(dout[:, v0: v0 + plan.v_cols], upT_v, out=dT[:, half:], n=half, alpha=plan.scaling,
                 tile_group=tile_group, b_row_table=plan.brow_bwd, block_n=64)
        # dh = dOut @ W + dT' @ down[sel]   (K-extension again)
        dh = None
        if ctx.needs_input_grad[0]:
            downT = ops.transpose(down.detach().contiguous())
            dh = ops.gemm(dout, ctx.WT, tile_group=tile_group, ext=(dT, downT, plan.ext_bwd), block_n=plan.block_n_bwd,
                          pair_aligned=getattr(ctx.rows, "pair_aligned", False))
        # weight gradients: reductions over the tokens of each task run (MN-major tcgen05 kernel)
        d_down, d_up = plan.wgrads(h, T, dT, dout, ctx.rows.runs)
        return dh, None, None, None, d_down, d_up, None, None


class LmHeadCEFn(torch.autograd.Function):
    """lm_head + fp32 cross-entropy evaluated ONLY on the rows whose shifted label is not ignore_index
    (identical loss to Llama_LoRA.py:372-386: ignored rows contribute nothing to the mean).  The label rows of all
    task segments go through ONE logits GEMM / CE / dgrad GEMM; `seg_sizes` splits the rows back into per-task losses."""

    @staticmethod
    def forward(ctx, hrows, W, WT, targets, row_scale, seg_sizes):
        # logits of the label rows, bf16 (as lm_head produces them before `.float()`)
        V = W.shape[0]
        Vp = (V + 7) // 8 * 8     # row stride must be a multiple of 16 bytes (vector stores, TMA in the backward)
        buf = torch.empty((hrows.shape[0], Vp), device=hrows.device, dtype=torch.bfloat16)
        logits = ops.gemm(hrows, W, out=buf[:, :V], block_n=256 if V >= 256 else 0)
        loss_rows, lse = ops.ce_fwd(logits, targets)
        ctx.save_for_backward(logits, targets, lse, row_scale)
        ctx.WT, ctx.seg_sizes = WT, seg_sizes
        return torch.stack([c.sum() for c in torch.split(loss_rows * row_scale, seg_sizes)])

    @staticmethod
    def backward(ctx, dloss):
        logits, targets, lse, row_scale = ctx.saved_tensors
        per_row = torch.repeat_interleave(dloss.float(), torch.tensor(ctx.seg_sizes, device=dloss.device),
                                          output_size=sum(ctx.seg_sizes))
        scale = (row_scale * per_row).contiguous()
        ops.ce_bwd_(logits, targets, lse, scale)        # in place: logits -> dlogits
        dh = ops.gemm(logits, ctx.WT, block_n=256)
        return dh, None, None, None, None, None
