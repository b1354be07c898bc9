"""Raw (non-autograd) Python entry points over the C ABI.  Each function validates shapes/dtypes,
allocates the output with torch (device memory plumbing only) and launches the sm_100a kernel on the
current CUDA stream."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch

from . import _lib
from ._lib import GemmArgs, PpsArgs, SpliceArgs, check, lib, ptr, require_cuda, stream_ptr

ACT = {None: 0, "none": 0, "relu": 1, "gelu": 2, "swiglu64": 3, "gelu_keep": 4, "prelu_ring": 5, "swiglu_bwd64": 6,
       "gelu_bwd": 7}
SWIGLU_BLK = 64          # gate / up interleave of the OMNI_ACT_SWIGLU64 epilogue
COMPRESS = {"avg-pooling": 0, "avg": 0, "stack": 1}

# number of kernel launches issued through this module (bench.py reports it as gpu_launches)
LAUNCHES = 0


def _count(n: int = 1) -> None:
    global LAUNCHES
    LAUNCHES += n


def _bf16_2d(t: torch.Tensor, name: str) -> torch.Tensor:
    if t.dtype != torch.bfloat16:
        raise TypeError(f"{name} must be bfloat16, got {t.dtype}")
    if t.dim() != 2:
        raise ValueError(f"{name} must be 2-D, got {tuple(t.shape)}")
    if t.stride(1) != 1:
        raise ValueError(f"{name} must have unit stride in the last dim")
    return t


_SKINNY_WS = {}


def _skinny_workspace(device) -> torch.Tensor:
    """Split-K exchange buffer of the weight-streaming GEMM: one per device, zero-filled once (the kernel leaves its
    counters at zero), reused by every launch -- launches on the same stream are ordered, which is how the decode step
    runs (also under CUDA-graph replay: the pointer is stable)."""
    key = (device.type, device.index if device.index is not None else torch.cuda.current_device())
    ws = _SKINNY_WS.get(key)
    if ws is None:
        ws = _SKINNY_WS[key] = torch.zeros(int(lib.omni_gemm_skinny_workspace_bytes()) + 256, device=device, dtype=torch.uint8)
    return ws


def gemm(a: torch.Tensor, b: torch.Tensor, *, bias: Optional[torch.Tensor] = None, act: Optional[str] = None,
         residual: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None, out_dtype=torch.bfloat16,
         alpha: float = 1.0, n: Optional[int] = None, tile_group: Optional[torch.Tensor] = None,
         b_row_table: Optional[torch.Tensor] = None, ext: Optional[tuple] = None, block_n: int = 0,
         pair_aligned: bool = False, out2: Optional[torch.Tensor] = None, skinny: bool = False,
         prelu_ring: Optional[tuple] = None, norm: Optional[tuple] = None) -> torch.Tensor:
    """out[M,N] = epi(alpha * (a[M,K] @ b[rows,K]^T (+ K-extension)))  -- tcgen05 kernel.

    ext = (a2 [M, a2_cols], b2 [b2_rows, b2_cols], ext_table int32 [groups, n_tiles, n_ext, 4]).
    act="swiglu64": b's rows are [gate 64 | up 64] interleaved blocks; out2 [M, N/2] (required) receives
    bf16(bf16(silu(gate)) * up); raises OmniKernelError(OMNI_ERR_UNSUPPORTED) when the shape is not one the CTA-pair
    kernel takes -- the caller then runs the unfused gemm + swiglu_fwd pair.
    act="gelu_keep": out = bf16(a @ b^T + bias) (the pre-activation the backward needs), out2 [M, N] = bf16(gelu(out)).
    skinny=True (M <= 128, K % 64 == 0, bf16 out): the weight-streaming decode-step kernel (omni_gemm_skinny_bf16: weights on
    the M side of the MMA, split-K over a cluster); b_row_table per 64-feature block, ext table per 128-feature tile
    (block_n=128).  act="swiglu64" with out=None writes only out2 = the activation (both kernels).
    norm=(weight [N], eps) with skinny=True: returns (out, rmsnorm(out)) -- the RMSNorm that follows o_proj / down_proj in a
    decode step runs in the same launch (the CTA that completes a token slice normalises it; same bits as rmsnorm_fwd);
    shapes the fused form does not take run the GEMM and the norm kernel.
    act="swiglu_bwd64" / "gelu_bwd" (fused backward epilogues, CTA-pair kernel only): a @ b^T is d(activation) and never
    reaches memory; `residual` carries the tensor saved by the forward (gate|up blocks [M, 2N] / pre-activation [M, N]) and
    `out` (required) receives its gradient ([M, 2N] / [M, N])."""
    if act in ("swiglu_bwd64", "gelu_bwd"):
        return _gemm_fused_bwd(a, b, residual, out, act)
    return _gemm(a, b, bias=bias, act=act, residual=residual, out=out, out_dtype=out_dtype, alpha=alpha, n=n,
                 tile_group=tile_group, b_row_table=b_row_table, ext=ext, block_n=block_n, pair_aligned=pair_aligned, out2=out2,
                 skinny=skinny, prelu_ring=prelu_ring, norm=norm)


def _gemm_fused_bwd(a, b, saved, out, act):
    require_cuda(a, b, saved, out)
    a, b, saved, out = _bf16_2d(a, "a"), _bf16_2d(b, "b"), _bf16_2d(saved, "residual"), _bf16_2d(out, "out")
    M, K = a.shape
    N = b.shape[0]
    width = 2 * N if act == "swiglu_bwd64" else N
    if b.shape[1] != K or saved.shape != (M, width) or out.shape != (M, width):
        raise ValueError(f"{act}: a [M, K], b [N, K], residual / out [M, {width}]")
    g = GemmArgs()
    g.A, g.B, g.out, g.residual = a.data_ptr(), b.data_ptr(), out.data_ptr(), saved.data_ptr()
    g.lda, g.ldb, g.ldo, g.ldr = a.stride(0), b.stride(0), out.stride(0), saved.stride(0)
    g.M, g.N, g.K, g.b_rows = M, N, K, N
    g.block_n, g.act, g.alpha = 256, ACT[act], 1.0
    check(lib.omni_gemm_bf16(C.byref(g), stream_ptr()), f"omni_gemm_bf16 ({act})")
    _count()
    return out


def _gemm(a, b, *, bias, act, residual, out, out_dtype, alpha, n, tile_group, b_row_table, ext, block_n, pair_aligned, out2,
          skinny, prelu_ring, norm=None):
    require_cuda(a, b, bias, residual, out, tile_group, b_row_table)
    a = _bf16_2d(a, "a")
    b = _bf16_2d(b, "b")
    M, K = a.shape
    if b.shape[1] != K:
        raise ValueError(f"K mismatch: a {tuple(a.shape)} vs b {tuple(b.shape)}")
    N = int(n) if n is not None else b.shape[0]
    skinny_swiglu = act == "swiglu64" and out is None      # only out2 = the activation is written (decode step, prefill)
    if skinny_swiglu:
        out = out2                                   # placeholder for the shape checks below; only out2 is written
    elif out is None:
        out = torch.empty((M, N), device=a.device, dtype=out_dtype)
    if skinny_swiglu:
        pass
    elif out.dim() != 2 or out.shape[0] != M or out.shape[1] != N or out.stride(1) != 1:
        raise ValueError("bad out tensor")
    if out.dtype not in (torch.bfloat16, torch.float32):
        raise TypeError("out must be bf16 or fp32")
    g = GemmArgs()
    g.A, g.B, g.out = a.data_ptr(), b.data_ptr(), (None if skinny_swiglu else out.data_ptr())
    g.lda, g.ldb, g.ldo = a.stride(0), b.stride(0), (N if skinny_swiglu else out.stride(0))
    g.M, g.N, g.K = M, N, K
    g.b_rows = b.shape[0]
    if bias is not None:
        if bias.dtype != torch.bfloat16 or (bias.numel() < N and act != "prelu_ring"):
            raise ValueError("bias must be bf16 [N]")
        g.bias = bias.data_ptr()
    if residual is not None:
        residual = _bf16_2d(residual, "residual")
        if residual.shape[0] != M or residual.shape[1] != N:
            raise ValueError("residual shape mismatch")
        g.residual, g.ldr = residual.data_ptr(), residual.stride(0)
    if tile_group is not None:
        if tile_group.dtype != torch.int32 or tile_group.numel() < (M + 127) // 128:
            raise ValueError("tile_group must be int32 [ceil(M/128)]")
        g.tile_group = tile_group.data_ptr()
    if b_row_table is not None:
        if b_row_table.dtype != torch.int32:
            raise ValueError("b_row_table must be int32")
        g.b_row_table = b_row_table.data_ptr()
    if ext is not None:
        a2, b2, table = ext
        require_cuda(a2, b2, table)
        a2 = _bf16_2d(a2, "a2")
        b2 = _bf16_2d(b2, "b2")
        if table.dtype != torch.int32 or table.shape[-1] != 4 or not table.is_contiguous():
            raise ValueError("ext_table must be contiguous int32 [..., n_ext, 4]")
        if a2.shape[0] != M:
            raise ValueError("a2 rows must equal M")
        g.A2, g.B2, g.ext_table = a2.data_ptr(), b2.data_ptr(), table.data_ptr()
        g.lda2, g.ldb2 = a2.stride(0), b2.stride(0)
        g.a2_cols, g.b2_rows, g.b2_cols = a2.shape[1], b2.shape[0], b2.shape[1]
        g.n_ext = table.shape[-2]
    if block_n == 256 and M <= 128 and N < 32768 and ext is None and b_row_table is None:
        # decode-step GEMMs (one 128-row tile of activations): the weight matrix is streamed once, so what matters is
        # how many SMs pull on HBM -- 64-column tiles give 4x the CTAs of 256-column ones (N = 2048: 32 instead of 8)
        block_n = 64 if N < 8192 else 128     # wide outputs: 128 columns halve the activation re-reads per weight byte
    g.block_n = block_n
    g.pair_aligned = 1 if pair_aligned else 0
    g.act = ACT[act]
    if act in ("swiglu64", "gelu_keep"):
        require_cuda(out2)
        n2 = N // 2 if act == "swiglu64" else N
        if out2 is None or out2.dtype != torch.bfloat16 or out2.shape != (M, n2) or out2.stride(1) != 1:
            raise ValueError(f"{act} needs out2: bf16 [M, {n2}]")
        g.out2, g.ldo2 = out2.data_ptr(), out2.stride(0)
    if act == "prelu_ring":
        # ResNet BasicBlock epilogue on ring-padded frames: (slope [C], res_bias [C] or None, H, W, pixels per row, C)
        slope, res_bias, rh, rw, rg, rc = prelu_ring
        require_cuda(slope, res_bias)
        g.slope, g.res_bias = slope.data_ptr(), ptr(res_bias)
        g.ring_h, g.ring_w, g.ring_group, g.ring_c = int(rh), int(rw), int(rg), int(rc)
    g.out_fp32 = 1 if out.dtype == torch.float32 else 0
    g.alpha = float(alpha)
    if skinny:
        ws = _skinny_workspace(a.device)
        g.workspace, g.workspace_bytes = ws.data_ptr(), ws.numel()
        if norm is not None:
            nw, eps = norm
            require_cuda(nw)
            h = torch.empty((M, N), device=a.device, dtype=torch.bfloat16)
            g.norm_weight, g.norm_out, g.norm_ld, g.norm_eps = nw.data_ptr(), h.data_ptr(), N, float(eps)
            rc = lib.omni_gemm_skinny_bf16(C.byref(g), stream_ptr())
            if rc == _lib.OMNI_ERR_UNSUPPORTED:          # no split-K for this shape: GEMM, then the norm kernel
                g.norm_weight, g.norm_out = None, None
                check(lib.omni_gemm_skinny_bf16(C.byref(g), stream_ptr()), "omni_gemm_skinny_bf16")
                _count()
                return out, rmsnorm_fwd(out, nw, eps)
            check(rc, "omni_gemm_skinny_bf16 (+rmsnorm)")
            _count()
            return out, h
        check(lib.omni_gemm_skinny_bf16(C.byref(g), stream_ptr()), "omni_gemm_skinny_bf16")
        _count()
        return out2 if skinny_swiglu else out
    if norm is not None:
        raise ValueError("norm= is a decode-step (skinny=True) epilogue")
    check(lib.omni_gemm_bf16(C.byref(g), stream_ptr()), "omni_gemm_bf16")
    _count()
    return out2 if skinny_swiglu else out


def matryoshka_compress(x: torch.Tensor, n_tok: int, rate: int, mode: str = "avg-pooling") -> torch.Tensor:
    """x [B, T>=n_tok, D] bf16 -> [B, n_tok//rate, D] (avg) / [B, n_tok//rate, rate*D] (stack)."""
    require_cuda(x)
    if x.dtype != torch.bfloat16 or x.dim() != 3 or x.stride(2) != 1 or x.stride(1) != x.shape[2]:
        raise ValueError("x must be bf16 [B, T, D] with contiguous rows")
    B, T, D = x.shape
    if n_tok > T:
        raise ValueError("n_tok exceeds T")
    m = COMPRESS[mode]
    n_out = n_tok // rate
    if m == 0 and n_out == 0:
        # same failure as nn.AvgPool1d in the reference (modeling_OmniAVSR.py:545)
        raise RuntimeError(f"Given input size: ({D}x1x{n_tok}). Calculated output size: ({D}x1x0). Output size is too small")
    out = torch.empty((B, n_out, D if m == 0 else D * rate), device=x.device, dtype=torch.bfloat16)
    if n_out > 0:
        check(lib.omni_matryoshka_compress(x.data_ptr(), out.data_ptr(), B, n_tok, D, x.stride(0), rate, m,
                                           stream_ptr()), "omni_matryoshka_compress")
        _count()
    return out


def matryoshka_compress_bwd(dout: torch.Tensor, n_tok: int, t_full: int, rate: int, mode: str) -> torch.Tensor:
    """Gradient wrt the [B, t_full, D] encoder output (rows >= n_tok get zero)."""
    require_cuda(dout)
    m = COMPRESS[mode]
    B, n_out = dout.shape[0], dout.shape[1]
    D = dout.shape[2] if m == 0 else dout.shape[2] // rate
    dout = dout.contiguous()
    dx = torch.empty((B, t_full, D), device=dout.device, dtype=torch.bfloat16)
    if t_full > n_tok:
        dx[:, n_tok:].zero_()
    check(lib.omni_matryoshka_compress_bwd(dout.data_ptr(), dx.data_ptr(), B, n_tok, D, dx.stride(0), rate, m,
                                           stream_ptr()), "omni_matryoshka_compress_bwd")
    _count()
    return dx


class SpliceLayout:
    """Host-side description of one splice call (mirrors omni_splice_args)."""

    def __init__(self, *, tokens, labels, embed, audio_tok, video_tok, prompts: Sequence[torch.Tensor],
                 marker_ids: Sequence[int], has_bos: bool, task_mask: int = 7, n_audio: Optional[int] = None,
                 n_video: Optional[int] = None):
        """audio_tok / video_tok: projected tokens [B, n, H] (stand-alone splice).  For the fused path
        (pool_project_splice) pass None for both and give the token counts per clip as n_audio / n_video (None = the
        modality is absent): the tokens are then written by the projector epilogue, not read from a tensor."""
        require_cuda(tokens, labels, embed, audio_tok, video_tok, *prompts)
        self.fused = audio_tok is None and video_tok is None and (n_audio is not None or n_video is not None)
        self.has_audio = audio_tok is not None or n_audio is not None
        self.has_video = video_tok is not None or n_video is not None
        self.keep = (tokens, labels, embed, audio_tok, video_tok, tuple(prompts))
        a = SpliceArgs()
        B, L = tokens.shape
        H = embed.shape[1]
        if tokens.dtype != torch.int64 or not tokens.is_contiguous():
            raise ValueError("tokens must be contiguous int64 [B, L]")
        if labels is not None and (labels.dtype != torch.int64 or not labels.is_contiguous() or labels.shape != tokens.shape):
            raise ValueError("labels must be contiguous int64 [B, L]")
        if embed.dtype != torch.bfloat16 or not embed.is_contiguous():
            raise ValueError("embed must be contiguous bf16 [V, H]")
        for name, t in (("audio_tok", audio_tok), ("video_tok", video_tok)):
            if t is not None and (t.dtype != torch.bfloat16 or not t.is_contiguous() or t.shape[0] != B or t.shape[2] != H):
                raise ValueError(f"{name} must be contiguous bf16 [B, n, H]")
        a.tokens, a.labels, a.embed = ptr(tokens) if L > 0 else None, ptr(labels) if L > 0 else None, ptr(embed)
        a.audio_tok, a.video_tok = ptr(audio_tok), ptr(video_tok)
        for t in range(3):
            p = prompts[t]
            if p is not None:
                if p.dtype != torch.bfloat16 or not p.is_contiguous() or p.shape[-1] != H:
                    raise ValueError("prompt must be contiguous bf16 [P, H]")
                a.prompt[t] = p.data_ptr()
                a.prompt_len[t] = p.shape[-2]
        a.B, a.L, a.H = B, L, H
        a.n_a = audio_tok.shape[1] if audio_tok is not None else int(n_audio or 0)
        a.n_v = video_tok.shape[1] if video_tok is not None else int(n_video or 0)
        a.id_audio_sos, a.id_audio_eos, a.id_video_sos, a.id_video_eos = [int(i) for i in marker_ids]
        a.has_bos = 1 if has_bos else 0
        a.task_mask = task_mask
        a.vocab = embed.shape[0]
        self.args = a
        self.B, self.H = B, H
        bos = 1 if has_bos else 0
        self.seq_len = [(bos + (a.n_a + 2 if self.has_audio and t in (0, 2) else 0) +
                         (a.n_v + 2 if self.has_video and t in (1, 2) else 0) + a.prompt_len[t] + (L - bos))
                        if (task_mask >> t) & 1 else 0 for t in range(3)]
        if not self.fused:   # the library's own arithmetic must agree with the mirror above
            for t in range(3):
                if (task_mask >> t) & 1 and int(lib.omni_splice_seq_len(C.byref(a), t)) != self.seq_len[t]:
                    raise RuntimeError("splice sequence length mismatch between the host mirror and the library")


def splice_prompt(layout: SpliceLayout, outs: Sequence[Optional[torch.Tensor]],
                  out_labels: Sequence[Optional[torch.Tensor]], status: Optional[torch.Tensor] = None) -> None:
    if layout.fused:
        raise ValueError("this layout has no token tensors: use pool_project_splice")
    _bind_splice_outputs(layout, outs, out_labels, status)
    check(lib.omni_splice_prompt(C.byref(layout.args), stream_ptr()), "omni_splice_prompt")
    _count()


def _bind_splice_outputs(layout, outs, out_labels, status):
    a = layout.args
    for t in range(3):
        o, l = outs[t], out_labels[t]
        if o is not None:
            require_cuda(o)
            if o.dtype != torch.bfloat16 or not o.is_contiguous() or o.numel() != layout.B * layout.seq_len[t] * layout.H:
                raise ValueError(f"out[{t}] must be contiguous bf16 [B, S_t, H]")
        if l is not None:
            require_cuda(l)
            if l.dtype != torch.int64 or not l.is_contiguous() or l.numel() != layout.B * layout.seq_len[t]:
                raise ValueError(f"out_labels[{t}] must be contiguous int64 [B, S_t]")
        a.out[t] = ptr(o)
        a.out_labels[t] = ptr(l)
    a.status = ptr(status)


class PoolProjectInput:
    """One modality of pool_project_splice: encoder output x [B, T, D] (only the first n_tok rows of a clip are read),
    compression rate, projector weights Linear(K1 -> I) + ReLU + Linear(I -> H)."""

    def __init__(self, x, n_tok: int, rate: int, w1, b1, w2, b2):
        require_cuda(x, w1, b1, w2, b2)
        if x.dtype != torch.bfloat16 or x.dim() != 3 or x.stride(2) != 1 or x.stride(1) != x.shape[2]:
            raise ValueError("x must be bf16 [B, T, D] with contiguous rows")
        if n_tok > x.shape[1]:
            raise ValueError("n_tok exceeds T")
        for t in (w1, b1, w2, b2):
            if t.dtype != torch.bfloat16 or not t.is_contiguous():
                raise ValueError("projector weights must be contiguous bf16")
        self.x, self.n_tok, self.rate = x, int(n_tok), int(rate)
        self.w1, self.b1, self.w2, self.b2 = w1, b1, w2, b2
        self.n = self.n_tok // self.rate


def pool_project_splice(layout: SpliceLayout, outs, out_labels, audio: Optional[PoolProjectInput],
                        video: Optional[PoolProjectInput], mode: str = "avg-pooling", status=None, want_tok: bool = False):
    """Fused compression -> projector MLP -> splice (omni_pool_project_splice): ONE persistent launch writes the three task
    sequences (outs[t]: [B, S_t, H]) and their labels.  Returns per modality (pooled [B*n, K1], hidden [B*n, I],
    tok [B*n, H] or None) -- the side outputs the backward needs."""
    if not layout.fused:
        raise ValueError("pool_project_splice needs a SpliceLayout built with n_audio / n_video (no token tensors)")
    m = COMPRESS[mode]
    _bind_splice_outputs(layout, outs, out_labels, status)
    g = PpsArgs()
    g.splice = layout.args
    keep, res = [], {}
    I = None
    dev = layout.keep[2].device
    for name, inp, n_lay in (("audio", audio, layout.args.n_a if layout.has_audio else None),
                             ("video", video, layout.args.n_v if layout.has_video else None)):
        if (inp is None) != (n_lay is None):
            raise ValueError(f"{name}: the layout and the inputs disagree about the presence of the modality")
        if inp is None:
            res[name] = None
            continue
        B, T, D = inp.x.shape
        if B != layout.B or inp.n != n_lay:
            raise ValueError(f"{name}: batch / token count mismatch with the layout")
        if m == 0 and inp.n == 0:
            raise RuntimeError(f"Given input size: ({D}x1x{inp.n_tok}). Calculated output size: ({D}x1x0). Output size is too small")
        K1 = D if m == 0 else D * inp.rate
        if inp.w1.shape[1] != K1 or inp.w2.shape[1] != inp.w1.shape[0] or inp.w2.shape[0] != layout.H:
            raise ValueError(f"{name}: projector shapes do not match (K1={K1}, H={layout.H})")
        if I is None:
            I = inp.w1.shape[0]
        elif I != inp.w1.shape[0]:
            raise ValueError("both projectors must share the intermediate width")
        M = B * inp.n
        pooled = torch.empty((M, K1), device=dev, dtype=torch.bfloat16)
        hidden = torch.empty((M, I), device=dev, dtype=torch.bfloat16)
        tok = torch.empty((M, layout.H), device=dev, dtype=torch.bfloat16) if want_tok else None
        mm = g.audio if name == "audio" else g.video
        mm.x, mm.x_bs, mm.n_tok, mm.rate, mm.D = inp.x.data_ptr(), inp.x.stride(0), inp.n_tok, inp.rate, D
        mm.w1, mm.b1, mm.w2, mm.b2 = inp.w1.data_ptr(), inp.b1.data_ptr(), inp.w2.data_ptr(), inp.b2.data_ptr()
        mm.pooled, mm.hidden, mm.tok = pooled.data_ptr(), hidden.data_ptr(), ptr(tok)
        res[name] = (pooled, hidden, tok)
        keep.append(inp)
    if I is None:
        raise ValueError("pool_project_splice needs at least one modality")
    g.I, g.mode = I, m
    nbytes = int(lib.omni_pps_workspace_bytes(C.byref(g)))
    ws = torch.empty(max(nbytes, 16), device=dev, dtype=torch.uint8)
    g.workspace, g.workspace_bytes = ws.data_ptr(), nbytes
    check(lib.omni_pool_project_splice(C.byref(g), stream_ptr()), "omni_pool_project_splice")
    _count()
    return res


def splice_prompt_bwd(layout: SpliceLayout, douts: Sequence[Optional[torch.Tensor]], want_audio: bool,
                      want_video: bool):
    a = layout.args
    d = (C.c_void_p * 3)()
    for t in range(3):
        if douts[t] is not None:
            require_cuda(douts[t])
            if not douts[t].is_contiguous() or douts[t].dtype != torch.bfloat16:
                raise ValueError("dout must be contiguous bf16")
            d[t] = douts[t].data_ptr()
    dev = layout.keep[2].device
    da = torch.empty((layout.B, a.n_a, layout.H), device=dev, dtype=torch.bfloat16) if want_audio else None
    dv = torch.empty((layout.B, a.n_v, layout.H), device=dev, dtype=torch.bfloat16) if want_video else None
    saved = (a.audio_tok, a.video_tok)
    if layout.fused:      # the kernel only tests these pointers for presence (it reads dout, not the tokens)
        a.audio_tok = layout.keep[2].data_ptr() if layout.has_audio else None
        a.video_tok = layout.keep[2].data_ptr() if layout.has_video else None
    try:
        check(lib.omni_splice_prompt_bwd(C.byref(a), d, ptr(da), ptr(dv), stream_ptr()), "omni_splice_prompt_bwd")
    finally:
        a.audio_tok, a.video_tok = saved
    _count()
    return da, dv


# ---------------------------------------------------------------------------------------------------
# row kernels
# ---------------------------------------------------------------------------------------------------
def _rows2d(x: torch.Tensor, name: str) -> torch.Tensor:
    require_cuda(x)
    if x.dtype != torch.bfloat16 or x.dim() != 2 or x.stride(1) != 1:
        raise ValueError(f"{name} must be bf16 [rows, H] with unit inner stride")
    return x


def rmsnorm_fwd(x, w, eps: float, want_rstd: bool = False):
    x = _rows2d(x, "x")
    rows, H = x.shape
    y = torch.empty((rows, H), device=x.device, dtype=torch.bfloat16)
    rstd = torch.empty(rows, device=x.device, dtype=torch.float32) if want_rstd else None
    check(lib.omni_rmsnorm_fwd(x.data_ptr(), w.data_ptr(), y.data_ptr(), ptr(rstd), rows, H, x.stride(0), H, eps,
                               stream_ptr()), "omni_rmsnorm_fwd")
    _count()
    return (y, rstd) if want_rstd else y


def rmsnorm_bwd(dy, x, w, rstd, dx_add=None):
    dy, x = dy.contiguous(), x.contiguous()
    rows, H = x.shape
    dx = torch.empty_like(x)
    if dx_add is not None:
        dx_add = dx_add.contiguous()
    check(lib.omni_rmsnorm_bwd(dy.data_ptr(), x.data_ptr(), w.data_ptr(), rstd.data_ptr(), dx.data_ptr(), ptr(dx_add),
                               rows, H, stream_ptr()), "omni_rmsnorm_bwd")
    _count()
    return dx


def layernorm_fwd(x, w, b, eps: float, want_stats: bool = False):
    x = _rows2d(x, "x")
    rows, H = x.shape
    y = torch.empty((rows, H), device=x.device, dtype=torch.bfloat16)
    mean = torch.empty(rows, device=x.device, dtype=torch.float32) if want_stats else None
    rstd = torch.empty(rows, device=x.device, dtype=torch.float32) if want_stats else None
    check(lib.omni_layernorm_fwd(x.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr(), ptr(mean), ptr(rstd), rows, H,
                                 x.stride(0), H, eps, stream_ptr()), "omni_layernorm_fwd")
    _count()
    return (y, mean, rstd) if want_stats else y


def layernorm_bwd(dy, x, w, mean, rstd, dx_add=None):
    dy, x = dy.contiguous(), x.contiguous()
    rows, H = x.shape
    dx = torch.empty_like(x)
    if dx_add is not None:
        dx_add = dx_add.contiguous()
    check(lib.omni_layernorm_bwd(dy.data_ptr(), x.data_ptr(), w.data_ptr(), mean.data_ptr(), rstd.data_ptr(),
                                 dx.data_ptr(), ptr(dx_add), rows, H, stream_ptr()), "omni_layernorm_bwd")
    _count()
    return dx


def rope_(qkv, cos_t, sin_t, pos, n_heads_total: int, head_dim: int, inverse: bool = False):
    """In-place RoPE on the first n_heads_total heads of each row of the packed [rows, ld] buffer."""
    qkv = _rows2d(qkv, "qkv")
    require_cuda(cos_t, sin_t, pos)
    if pos.dtype != torch.int32 or pos.numel() < qkv.shape[0]:
        raise ValueError("pos must be int32 [rows]")
    if cos_t.dtype != torch.bfloat16 or cos_t.shape[-1] != head_dim or not cos_t.is_contiguous():
        raise ValueError("cos/sin tables must be contiguous bf16 [max_pos, head_dim]")
    check(lib.omni_rope(qkv.data_ptr(), cos_t.data_ptr(), sin_t.data_ptr(), pos.data_ptr(), qkv.shape[0],
                        qkv.stride(0), n_heads_total, head_dim, 1 if inverse else 0, stream_ptr()), "omni_rope")
    _count()
    return qkv


def swiglu_fwd(gu):
    gu = _rows2d(gu, "gu")
    if not gu.is_contiguous():
        raise ValueError("gu must be contiguous")
    rows, I2 = gu.shape
    act = torch.empty((rows, I2 // 2), device=gu.device, dtype=torch.bfloat16)
    check(lib.omni_swiglu_fwd(gu.data_ptr(), act.data_ptr(), rows, I2 // 2, stream_ptr()), "omni_swiglu_fwd")
    _count()
    return act


def swiglu_bwd(dact, gu, blk: Optional[int] = None):
    """blk: gate / up column interleave of `gu` (None = the plain [gate | up] halves, 64 = the swiglu64 GEMM layout)."""
    require_cuda(dact, gu)
    dact = dact.contiguous()
    rows, I2 = gu.shape
    dgu = torch.empty_like(gu)
    check(lib.omni_swiglu_bwd_blocked(dact.data_ptr(), gu.data_ptr(), dgu.data_ptr(), rows, I2 // 2,
                                      I2 // 2 if blk is None else blk, stream_ptr()), "omni_swiglu_bwd_blocked")
    _count()
    return dgu


def gelu_fwd(x):
    require_cuda(x)
    x = x.contiguous()
    y = torch.empty_like(x)
    check(lib.omni_gelu_fwd(x.data_ptr(), y.data_ptr(), x.numel(), stream_ptr()), "omni_gelu_fwd")
    _count()
    return y


def gelu_bwd(dy, x):
    dy, x = dy.contiguous(), x.contiguous()
    dx = torch.empty_like(x)
    check(lib.omni_gelu_bwd(dy.data_ptr(), x.data_ptr(), dx.data_ptr(), x.numel(), stream_ptr()), "omni_gelu_bwd")
    _count()
    return dx


def gather_rows(table, idx, status=None):
    require_cuda(table, idx)
    if table.dtype != torch.bfloat16 or table.dim() != 2 or table.stride(1) != 1:
        raise ValueError("table must be bf16 [rows, H]")
    idx = idx.reshape(-1)
    if idx.dtype != torch.int64 or not idx.is_contiguous():
        raise ValueError("idx must be contiguous int64")
    out = torch.empty((idx.numel(), table.shape[1]), device=table.device, dtype=torch.bfloat16)
    check(lib.omni_gather_rows(table.data_ptr(), idx.data_ptr(), out.data_ptr(), idx.numel(), table.shape[1],
                               table.stride(0), table.shape[0], ptr(status), stream_ptr()), "omni_gather_rows")
    _count()
    return out


def scatter_rows(src, idx, out):
    """out[idx[i]] = src[i] (unique idx)."""
    require_cuda(src, idx, out)
    src = src.contiguous()
    check(lib.omni_scatter_rows(src.data_ptr(), idx.data_ptr(), out.data_ptr(), idx.numel(), src.shape[1],
                                out.stride(0), stream_ptr()), "omni_scatter_rows")
    _count()
    return out


def ce_fwd(logits, targets, ignore_index: int = -100):
    """logits bf16 [R, V] (row stride may exceed V), targets int64 [R] -> (loss_rows fp32 [R], lse fp32 [R])."""
    require_cuda(logits, targets)
    R, V = logits.shape
    loss = torch.empty(R, device=logits.device, dtype=torch.float32)
    lse = torch.empty(R, device=logits.device, dtype=torch.float32)
    check(lib.omni_ce_fwd(logits.data_ptr(), targets.data_ptr(), loss.data_ptr(), lse.data_ptr(), R, V,
                          logits.stride(0), ignore_index, stream_ptr()), "omni_ce_fwd")
    _count()
    return loss, lse


def ce_bwd_(logits, targets, lse, scale, ignore_index: int = -100):
    """Overwrites logits with d(loss)/d(logits) * scale[r]."""
    R, V = logits.shape
    check(lib.omni_ce_bwd(logits.data_ptr(), targets.data_ptr(), lse.data_ptr(), scale.data_ptr(), R, V,
                          logits.stride(0), ignore_index, stream_ptr()), "omni_ce_bwd")
    _count()
    return logits


def decode_pick(logits, V: int, unfinished, eos, pad, step_idx, out, alive, embed, x_next):
    """Token bookkeeping of one greedy decode step in one launch (see omni_decode_pick): argmax, pad-after-EOS, output slot,
    unfinished / alive flags, embedding row of the chosen token into x_next[:B]."""
    require_cuda(logits, unfinished, eos, pad, step_idx, out, alive, embed, x_next)
    B = unfinished.shape[0]
    for t in (unfinished, eos, pad, step_idx, out, alive):
        if t.dtype != torch.int64 or not t.is_contiguous():
            raise TypeError("decode_pick: int64 contiguous state tensors")
    if logits.dtype != torch.bfloat16 or embed.dtype != torch.bfloat16 or x_next.dtype != torch.bfloat16:
        raise TypeError("decode_pick: bf16 logits / embed / x_next")
    if logits.shape[0] != B or x_next.shape[0] < B or out.shape[1] != B or embed.shape[1] != x_next.shape[1]:
        raise ValueError("decode_pick: shape mismatch")
    check(lib.omni_decode_pick(logits.data_ptr(), V, logits.stride(0), unfinished.data_ptr(), eos.data_ptr(), pad.data_ptr(),
                               step_idx.data_ptr(), out.data_ptr(), B, alive.data_ptr(), embed.data_ptr(), embed.stride(0),
                               x_next.data_ptr(), x_next.stride(0), embed.shape[1], stream_ptr()), "omni_decode_pick")
    _count()


def decode_advance(step_idx, len_idx, pos):
    require_cuda(step_idx, len_idx, pos)
    if step_idx.dtype != torch.int64 or len_idx.dtype != torch.int64 or pos.dtype != torch.int32:
        raise TypeError("decode_advance: int64 counters, int32 positions")
    check(lib.omni_decode_advance(step_idx.data_ptr(), len_idx.data_ptr(), pos.data_ptr(), pos.numel(), stream_ptr()),
          "omni_decode_advance")
    _count()


def beam_topk_rows(logits, V: int, beam_scores, cand_score, cand_tok):
    """Per logits row: fp32 log-softmax + running beam score of the row's n_cand best tokens, sorted (omni_beam_topk_rows)."""
    require_cuda(logits, beam_scores, cand_score, cand_tok)
    rows, n_cand = cand_score.shape
    if logits.dtype != torch.bfloat16 or logits.stride(1) != 1 or beam_scores.dtype != torch.float32 or \
            cand_score.dtype != torch.float32 or cand_tok.dtype != torch.int32 or not cand_score.is_contiguous() or \
            not cand_tok.is_contiguous() or cand_tok.shape != cand_score.shape or logits.shape[0] != rows or \
            beam_scores.numel() != rows:
        raise ValueError("beam_topk_rows: bf16 logits [rows, >= V], fp32 scores [rows], fp32 / int32 candidates [rows, n_cand]")
    check(lib.omni_beam_topk_rows(logits.data_ptr(), logits.stride(0), rows, V, beam_scores.data_ptr(), n_cand,
                                  cand_score.data_ptr(), cand_tok.data_ptr(), stream_ptr()), "omni_beam_topk_rows")
    _count()


def beam_select(st, embed, x_next):
    """BeamSearchScorer.process of one step on the device (omni_beam_select).  `st` is a decode.BeamState."""
    from ._lib import BeamSelectArgs
    require_cuda(embed, x_next)
    a = BeamSelectArgs()
    for name in ("cand_score", "cand_tok", "beam_scores", "step_idx", "eos", "pad", "seqs", "ind", "hyp_seq", "hyp_len",
                 "hyp_score", "hyp_order", "hyp_count", "hyp_worst", "done", "n_done", "status"):
        setattr(a, name, getattr(st, name).data_ptr())
    a.embed, a.x_next = embed.data_ptr(), x_next.data_ptr()
    a.ld_embed, a.ld_x = embed.stride(0), x_next.stride(0)
    a.B, a.K, a.n_cand, a.V, a.max_new, a.ind_ld, a.H = st.B, st.K, 2 * st.K, st.V, st.max_new, st.ind.shape[2], embed.shape[1]
    import ctypes
    check(lib.omni_beam_select(ctypes.byref(a), stream_ptr()), "omni_beam_select")
    _count()


def argmax_rows(logits):
    require_cuda(logits)
    R, V = logits.shape
    out = torch.empty(R, device=logits.device, dtype=torch.int64)
    check(lib.omni_argmax(logits.data_ptr(), out.data_ptr(), R, V, logits.stride(0), stream_ptr()), "omni_argmax")
    _count()
    return out


def sumsq_(g, acc):
    check(lib.omni_sumsq(g.data_ptr(), g.numel(), acc.data_ptr(), stream_ptr()), "omni_sumsq")
    _count()


def adamw_(p, g, m, v, *, lr, beta1, beta2, eps, weight_decay, step, grad_scale=1.0, max_norm=0.0, sumsq=None):
    require_cuda(p, g, m, v)
    check(lib.omni_adamw(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel(), lr, beta1, beta2, eps,
                         weight_decay, step, grad_scale, max_norm, ptr(sumsq), stream_ptr()), "omni_adamw")
    _count()


def gemm_wgrad(a, b, *, mo: int, no: int, a_col0: int = 0, b_col0: int = 0, ranges=None, out=None,
               out_dtype=torch.bfloat16, alpha: float = 1.0, accumulate: bool = False):
    """out[z][i, j] = alpha * sum_{k in ranges[z]} a[k, a_col0+i] * b[k, b_col0+j]  (tcgen05, MN-major operands).

    a [K, *], b [K, *] token-major bf16; ranges = [(k0, k1), ...] (default: all tokens); out [Z, mo, no] (or [mo, no])."""
    require_cuda(a, b, out)
    a = _bf16_2d(a, "a")
    b = _bf16_2d(b, "b")
    K = a.shape[0]
    if b.shape[0] != K:
        raise ValueError("a and b must have the same number of token rows")
    if ranges is None:
        ranges = [(0, K)]
    Z = len(ranges)
    if Z > _lib.WGRAD_MAX_RANGES:
        raise ValueError("too many token ranges")
    if out is None:
        out = torch.empty((Z, mo, no), device=a.device, dtype=out_dtype)
    o3 = out if out.dim() == 3 else out.unsqueeze(0)
    if o3.shape[0] != Z or o3.shape[1] != mo or o3.shape[2] != no or o3.stride(2) != 1:
        raise ValueError("bad out tensor")
    if o3.dtype not in (torch.bfloat16, torch.float32):
        raise TypeError("out must be bf16 or fp32")
    g = _lib.WgradArgs()
    g.A, g.B, g.out = a.data_ptr(), b.data_ptr(), o3.data_ptr()
    g.lda, g.ldb, g.ldo, g.out_zstride = a.stride(0), b.stride(0), o3.stride(1), o3.stride(0) if Z > 1 else 0
    g.K, g.a_cols, g.b_cols = K, a.shape[1], b.shape[1]
    g.Mo, g.No, g.a_col0, g.b_col0 = mo, no, a_col0, b_col0
    g.n_ranges = Z
    for i, (k0, k1) in enumerate(ranges):
        g.k0[i], g.k1[i] = int(k0), int(k1)
    g.out_fp32 = 1 if o3.dtype == torch.float32 else 0
    g.accumulate = 1 if accumulate else 0
    g.alpha = float(alpha)
    check(lib.omni_gemm_wgrad_bf16(C.byref(g), stream_ptr()), "omni_gemm_wgrad_bf16")
    _count()
    return out


def transpose(x: torch.Tensor) -> torch.Tensor:
    """[R, C] or [Z, R, C] contiguous bf16 -> contiguous [C, R] / [Z, C, R] (tiled shared-memory transpose)."""
    require_cuda(x)
    if x.dtype != torch.bfloat16 or not x.is_contiguous() or x.dim() not in (2, 3):
        raise ValueError("transpose needs a contiguous bf16 [R, C] or [Z, R, C] tensor")
    Z = x.shape[0] if x.dim() == 3 else 1
    R, Cc = x.shape[-2], x.shape[-1]
    out = torch.empty((*x.shape[:-2], Cc, R), device=x.device, dtype=torch.bfloat16)
    check(lib.omni_transpose_bf16(x.data_ptr(), out.data_ptr(), Z, R, Cc, stream_ptr()), "omni_transpose_bf16")
    _count()
    return out


def colsum(x):
    """Column sums (fp32 accumulate) of a bf16 [rows, cols] matrix -> bf16 [cols] (bias gradient)."""
    x = _bf16_2d(x, "x")
    require_cuda(x)
    out = torch.empty(x.shape[1], device=x.device, dtype=torch.bfloat16)
    check(lib.omni_colsum_bf16(x.data_ptr(), out.data_ptr(), x.shape[0], x.shape[1], x.stride(0), stream_ptr()),
          "omni_colsum_bf16")
    _count()
    return out


def _nhwc_rows(x):
    """NCHW-shaped tensor stored channels-last -> (data pointer view as [N*H*W, C])."""
    if x.dim() != 4 or not x.is_contiguous(memory_format=torch.channels_last) or x.dtype != torch.bfloat16:
        raise ValueError("expected a bf16 channels-last [N, C, H, W] tensor")
    return x.shape[0] * x.shape[2] * x.shape[3], x.shape[1]


def prelu_res_(x, slope, residual=None, bias=None, res_bias=None):
    """x <- PReLU((x + bias) (+ residual + res_bias)) in place; x / residual: bf16 channels-last [N, C, H, W]; slope, bias,
    res_bias bf16 [C] (the biases are the folded-BatchNorm shifts of the convolutions that produced x / residual)."""
    require_cuda(x, slope, residual, bias, res_bias)
    rows, Cc = _nhwc_rows(x)
    if residual is not None and (_nhwc_rows(residual) != (rows, Cc)):
        raise ValueError("residual shape mismatch")
    for b in (bias, res_bias):
        if b is not None and (b.dtype != torch.bfloat16 or b.numel() != Cc or not b.is_contiguous()):
            raise ValueError("bias must be contiguous bf16 [C]")
    check(lib.omni_prelu_res(x.data_ptr(), ptr(residual), slope.data_ptr(), ptr(bias), ptr(res_bias), rows, Cc,
                             stream_ptr()), "omni_prelu_res")
    _count()
    return x


def prelu_maxpool3x3s2(x, slope):
    """PReLU then MaxPool2d(3, stride 2, padding 1) on a bf16 channels-last [N, C, H, W] tensor."""
    require_cuda(x, slope)
    _nhwc_rows(x)
    N, Cc, H, W = x.shape
    Ho, Wo = (H + 2 - 3) // 2 + 1, (W + 2 - 3) // 2 + 1
    y = torch.empty((N, Cc, Ho, Wo), device=x.device, dtype=torch.bfloat16, memory_format=torch.channels_last)
    check(lib.omni_prelu_maxpool3x3s2(x.data_ptr(), slope.data_ptr(), y.data_ptr(), N, H, W, Cc, stream_ptr()),
          "omni_prelu_maxpool3x3s2")
    _count()
    return y


def logmel(audio, mel_filters):
    """audio [B, T] (fp32 or bf16, unit inner stride) -> Whisper input features bf16 [B, 80, 3000] (on-device log-mel)."""
    require_cuda(audio, mel_filters)
    if audio.dim() != 2 or audio.stride(1) != 1 or audio.dtype not in (torch.float32, torch.bfloat16):
        raise ValueError("audio must be fp32/bf16 [B, T] with unit inner stride")
    if mel_filters.dtype != torch.float32 or tuple(mel_filters.shape) != (80, 201) or not mel_filters.is_contiguous():
        raise ValueError("mel_filters must be contiguous fp32 [80, 201]")
    B, T = audio.shape
    out = torch.empty((B, 80, 3000), device=audio.device, dtype=torch.bfloat16)
    nbytes = int(lib.omni_logmel_workspace_bytes(B))
    ws = torch.empty(nbytes, device=audio.device, dtype=torch.uint8)
    check(lib.omni_logmel(audio.data_ptr(), 1 if audio.dtype == torch.bfloat16 else 0, audio.stride(0), B, T,
                          mel_filters.data_ptr(), out.data_ptr(), ws.data_ptr(), nbytes, stream_ptr()), "omni_logmel")
    _count(3)
    return out


def conv3d_front(video, wmat, bias):
    """AV-HuBERT front-end Conv3d(1, C, (5,7,7), stride (1,2,2), pad (2,3,3)) on video [B, T, H, W] bf16:
    im2col kernel + tcgen05 GEMM (bias = folded BatchNorm shift).  Returns [B*T, C, Ho, Wo] in channels-last memory."""
    require_cuda(video, wmat, bias)
    if video.dtype != torch.bfloat16 or video.dim() != 4 or not video.is_contiguous():
        raise ValueError("video must be contiguous bf16 [B, T, H, W]")
    if wmat.dtype != torch.bfloat16 or wmat.shape[1] != 256 or not wmat.is_contiguous():
        raise ValueError("wmat must be contiguous bf16 [C, 256]")
    B, T, H, W = video.shape
    Ho, Wo = (H + 6 - 7) // 2 + 1, (W + 6 - 7) // 2 + 1
    Cc = wmat.shape[0]
    out = torch.empty((B * T * Ho * Wo, Cc), device=video.device, dtype=torch.bfloat16)
    per_clip = T * Ho * Wo
    clips = max(1, (4 << 20) // per_clip)          # <= ~4M im2col rows (2 GB) in flight; whole clips per chunk
    cols = torch.empty((min(B, clips) * per_clip, 256), device=video.device, dtype=torch.bfloat16)
    for b0 in range(0, B, clips):
        nb = min(clips, B - b0)
        check(lib.omni_im2col_front3d(video[b0:b0 + nb].data_ptr(), cols.data_ptr(), nb, T, H, W, stream_ptr()),
              "omni_im2col_front3d")
        _count()
        gemm(cols[: nb * per_clip], wmat, bias=bias, out=out[b0 * per_clip:(b0 + nb) * per_clip],
             block_n=64 if Cc <= 64 else 128)
    return out.view(B * T, Ho, Wo, Cc).permute(0, 3, 1, 2)


def front3d_prelu_maxpool(video, wmat5, bias, slope, ring_out: Optional["RingFrames"] = None):
    """AV-HuBERT front-end Conv3d(1, C, (5,7,7), (1,2,2), (2,3,3)) (+ folded BatchNorm bias) + PReLU + MaxPool3d((1,3,3),
    (1,2,2), (0,1,1)) on video [B, T, H, W] bf16.  wmat5 [C, 320]: column dt*64 + ky*7 + kx (49 taps + 15 zero columns
    per temporal tap).  Time-major im2col of the 49 spatial taps, one tcgen05 GEMM whose five K blocks read five
    consecutive rows (overlapping-row TMA view), pooling kernel.  Returns [B*T, C, Hp, Wp] in channels-last memory."""
    require_cuda(video, wmat5, bias, slope)
    if video.dtype != torch.bfloat16 or video.dim() != 4 or not video.is_contiguous():
        raise ValueError("video must be contiguous bf16 [B, T, H, W]")
    if wmat5.dtype != torch.bfloat16 or wmat5.shape[1] != 320 or not wmat5.is_contiguous():
        raise ValueError("wmat5 must be contiguous bf16 [C, 320]")
    B, T, H, W = video.shape
    Ho, Wo = (H + 6 - 7) // 2 + 1, (W + 6 - 7) // 2 + 1
    Hp, Wp = (Ho + 2 - 3) // 2 + 1, (Wo + 2 - 3) // 2 + 1
    Cc = wmat5.shape[0]
    Tp = T + 4
    if ring_out is not None:
        if (ring_out.N, ring_out.H, ring_out.W, ring_out.C) != (B * T, Hp, Wp, Cc):
            raise ValueError("ring_out geometry mismatch")
        y = None
    else:
        y = torch.empty((B * T, Cc, Hp, Wp), device=video.device, dtype=torch.bfloat16, memory_format=torch.channels_last)
    per_clip = Ho * Wo * Tp
    clips = max(1, (16 << 20) // per_clip)          # <= ~16M rows (2 GB of taps + 2 GB of conv output) in flight
    nb_max = min(B, clips)
    cols = torch.empty((nb_max * per_clip + 8, 64), device=video.device, dtype=torch.bfloat16)
    cols[nb_max * per_clip:].zero_()                 # rows read by the last line's scratch outputs
    conv = torch.empty((nb_max * per_clip, Cc), device=video.device, dtype=torch.bfloat16)
    for b0 in range(0, B, clips):
        nb = min(clips, B - b0)
        check(lib.omni_im2col_front2d(video[b0:b0 + nb].data_ptr(), cols.data_ptr(), nb, T, H, W, stream_ptr()),
              "omni_im2col_front2d")
        _count()
        a = cols.as_strided((nb * per_clip, 320), (64, 1))
        gemm(a, wmat5, bias=bias, out=conv[: nb * per_clip], block_n=64 if Cc <= 64 else 128)
        if ring_out is not None:
            dst = ring_out.rows[b0 * T * ring_out.P:]
            check(lib.omni_prelu_maxpool_front_ring(conv.data_ptr(), slope.data_ptr(), dst.data_ptr(), nb, T, Ho, Wo, Cc,
                                                    stream_ptr()), "omni_prelu_maxpool_front_ring")
        else:
            check(lib.omni_prelu_maxpool_front(conv.data_ptr(), slope.data_ptr(), y[b0 * T:].data_ptr(), nb, T, Ho, Wo, Cc,
                                               stream_ptr()), "omni_prelu_maxpool_front")
        _count()
    return ring_out if ring_out is not None else y


# ---------------------------------------------------------------------------------------------------
# ResNet-18 trunk on the tcgen05 GEMM: ring-padded channels-last frames (csrc/resnet_trunk.cu)
# ---------------------------------------------------------------------------------------------------
class RingFrames:
    """N channels-last frames [H + 2, W + 2, C] with a one-pixel zero ring, flattened to rows of C channels.  The storage
    has (W + 3) * C spare elements before and after the rows: the overlapping-row GEMM view of a 3x3 convolution starts one
    padded line + one pixel before row 0 and ends as much after the last row (those reads only feed ring outputs)."""

    def __init__(self, N, H, W, C, device, zero=False):
        self.N, self.H, self.W, self.C = int(N), int(H), int(W), int(C)
        self.P = (H + 2) * (W + 2)
        self.M = self.N * self.P
        self.margin = (W + 3) * C
        n = 2 * self.margin + self.M * C + 8 * C          # (+ 8 pixels: grouped GEMM rows may run past the last frame)
        self.buf = (torch.zeros if zero else torch.empty)(n, device=device, dtype=torch.bfloat16)
        self.rows = self.buf[self.margin: self.margin + self.M * C].view(self.M, C)
        if not zero:
            # the spare elements must be FINITE: the pixel-grouped convolution multiplies them by zero filter blocks
            self.buf[: self.margin].zero_()
            self.buf[self.margin + self.M * C:].zero_()

    _pool = {}

    @classmethod
    def get(cls, N, H, W, C, device):
        """Activation buffer from a small per-geometry ring of persistent buffers (six per geometry: a BasicBlock keeps at
        most four alive -- input, conv1 output, conv2 output, downsample output -- and its successor two more), so that the
        trunk allocates nothing and zero-fills nothing per step.  Only for the frozen, no-grad trunk."""
        key = (int(N), int(H), int(W), int(C), str(device))
        ent = cls._pool.get(key)
        if ent is None:
            # one batch geometry at a time: a different frame count (variable-length batches, train / eval alternation)
            # drops the previous geometry's buffers instead of keeping six buffers per layer for every length ever seen
            if len(cls._pool) > 64 or any(k[0] != key[0] for k in cls._pool):
                cls._pool.clear()
            ent = cls._pool[key] = [[cls(N, H, W, C, device) for _ in range(6)], 0]
        ent[1] = (ent[1] + 1) % 6
        return ent[0][ent[1]]


_CONV_TABLES = {}


def _conv3x3_table(W, C, g, n_out, bn, device):
    """K-extension table of the overlapping-row 3x3 convolution with g output pixels per GEMM row: segments dy = 0, +1 at
    columns (W + 2) C, 2 (W + 2) C of the view, against columns (g + 2) C, 2 (g + 2) C of the filter matrix, in 64-column
    blocks, per N tile."""
    key = (W, C, g, n_out, bn, str(device))
    t = _CONV_TABLES.get(key)
    if t is None:
        nt = (n_out + bn - 1) // bn
        seg = (g + 2) * C
        blocks = seg // 64
        tab = torch.empty((1, nt, 2 * blocks, 4), dtype=torch.int32)
        for i in range(nt):
            j = 0
            for s in (1, 2):
                for b in range(blocks):
                    tab[0, i, j] = torch.tensor([s * (W + 2) * C + 64 * b, i * bn, s * seg + 64 * b, 0])
                    j += 1
        t = _CONV_TABLES[key] = tab.contiguous().to(device)
    return t


def conv3x3_group_weights(w: torch.Tensor, g: int) -> torch.Tensor:
    """[C_out, C_in, 3, 3] filters -> the filter matrix of the overlapping-row GEMM that computes g horizontally consecutive
    output pixels per row: [g * C_out, 3 * (g + 2) * C_in], row p * C_out + co, column dy * (g + 2) C_in + q * C_in + ci holds
    w[co, ci, dy, q - p] (zero where q - p is outside the 3-tap window).  g = 1 is the plain tap-major matrix.  Grouping
    trades (g + 2) / 3 x the flops for a g x wider N: 64- and 128-channel layers reach the 256-wide CTA-pair GEMM tiles."""
    Co, Ci = w.shape[0], w.shape[1]
    out = torch.zeros((g, Co, 3, g + 2, Ci), device=w.device, dtype=w.dtype)
    for p in range(g):
        out[p, :, :, p: p + 3, :] = w.permute(0, 2, 3, 1)          # [Co, ky, kx, Ci] at pixel offsets p .. p + 2
    return out.reshape(g * Co, 3 * (g + 2) * Ci).contiguous()


def conv3x3s1_ring(x: RingFrames, wmat: torch.Tensor, group: int = 1, prelu: Optional[dict] = None) -> RingFrames:
    """3x3 / stride 1 / pad 1 convolution of ring-padded frames as ONE tcgen05 GEMM launch.  wmat = conv3x3_group_weights(w,
    group): [group * C_out, 3 * (group + 2) * C_in].  Without `prelu` the ring rows of the result are garbage: follow with
    prelu_res_ring_.  prelu = dict(slope, bias, residual=None, res_bias=None): the BasicBlock tail (folded-BN shift,
    residual add, PReLU, ring re-zeroing) runs in the GEMM epilogue when the shape is on the CTA-pair kernel, otherwise in
    the separate kernel -- same bits either way."""
    require_cuda(wmat)
    C, g = x.C, group
    Co = wmat.shape[0] // g
    seg = (g + 2) * C
    if wmat.dtype != torch.bfloat16 or wmat.shape[1] != 3 * seg or not wmat.is_contiguous() or g > 8:
        raise ValueError("wmat must be contiguous bf16 [group * C_out, 3 * (group + 2) * C_in]")
    if seg % 64:
        if g != 1:
            raise ValueError("grouped convolution needs (group + 2) * C_in to be a multiple of 64")
        out = conv_s2_ring(x, wmat, 9, stride=1)       # narrow test architectures: gather + plain GEMM
        if prelu is not None:
            prelu_res_ring_(out, prelu["slope"], prelu.get("residual"), bias=prelu["bias"], res_bias=prelu.get("res_bias"))
        return out
    N = g * Co
    bn = 64 if N <= 64 else (128 if N <= 128 else 256)
    Mg = (x.M + g - 1) // g
    a_main = torch.as_strided(x.buf, (Mg, seg), (g * C, 1), 0)
    a_ext = torch.as_strided(x.buf, (Mg, (2 * (x.W + 2) + g + 2) * C), (g * C, 1), 0)
    out = RingFrames.get(x.N, x.H, x.W, Co, x.buf.device)
    out_rows = torch.as_strided(out.buf, (Mg, N), (N, 1), out.margin)
    ext = (a_ext, wmat, _conv3x3_table(x.W, C, g, N, bn, x.buf.device))
    if prelu is not None and bn == 256 and N % 128 == 0 and Co % 64 == 0 and Mg > 128 and ((Mg + 127) // 128) * ((N + 255) // 256) >= 74:
        res = prelu.get("residual")
        res_rows = None if res is None else torch.as_strided(res.buf, (Mg, N), (N, 1), res.margin)
        try:
            gemm(a_main, wmat[:, :seg], ext=ext, out=out_rows, block_n=bn, pair_aligned=True, act="prelu_ring",
                 bias=prelu["bias"], residual=res_rows,
                 prelu_ring=(prelu["slope"], prelu.get("res_bias"), x.H, x.W, g, Co))
            return out
        except _lib.OmniKernelError as e:
            if "unsupported" not in str(e):
                raise
    gemm(a_main, wmat[:, :seg], ext=ext, out=out_rows, block_n=bn, pair_aligned=True)
    if prelu is not None:
        prelu_res_ring_(out, prelu["slope"], prelu.get("residual"), bias=prelu["bias"], res_bias=prelu.get("res_bias"))
    return out


def conv_s2_ring(x: RingFrames, wmat: torch.Tensor, taps: int, stride: int = 2) -> RingFrames:
    """Strided convolution (taps = 9: 3x3 pad 1; taps = 1: the 1x1 downsample) of ring-padded frames: gather kernel ->
    plain GEMM.  The output grid is ring-padded too (ring rows: GEMM of zero rows = 0 before the bias)."""
    require_cuda(wmat)
    Ho, Wo = (x.H - 1) // stride + 1, (x.W - 1) // stride + 1
    out = RingFrames.get(x.N, Ho, Wo, wmat.shape[0], x.buf.device)
    cols = torch.empty((out.M, taps * x.C), device=x.buf.device, dtype=torch.bfloat16)
    check(lib.omni_gather_s2_ring(x.rows.data_ptr(), cols.data_ptr(), x.N, x.H, x.W, x.C, taps, stride, stream_ptr()),
          "omni_gather_s2_ring")
    _count()
    Co = wmat.shape[0]
    gemm(cols, wmat, out=out.rows, block_n=64 if Co <= 64 else (128 if Co <= 128 else 256))
    return out


def prelu_res_ring_(x: RingFrames, slope, residual: Optional[RingFrames] = None, bias=None, res_bias=None) -> RingFrames:
    """x <- PReLU((x + bias) (+ residual + res_bias)) on the interior, zeros on the ring (in place)."""
    require_cuda(slope, bias, res_bias)
    if residual is not None and (residual.M, residual.C) != (x.M, x.C):
        raise ValueError("residual shape mismatch")
    check(lib.omni_prelu_res_ring(x.rows.data_ptr(), None if residual is None else residual.rows.data_ptr(), slope.data_ptr(),
                                  ptr(bias), ptr(res_bias), x.N, x.H, x.W, x.C, stream_ptr()), "omni_prelu_res_ring")
    _count()
    return x


class FrameRows:
    """N plain channels-last frames [H, W, C] WITHOUT a ring, one frame per GEMM row: buf [N, PA * C] with PA >= H W pixel
    slots (the slots past H W pad the last N tile of the convolution that wrote them and are never read)."""

    def __init__(self, N, H, W, C, PA, device):
        self.N, self.H, self.W, self.C, self.PA = int(N), int(H), int(W), int(C), int(PA)
        self.buf = torch.empty((self.N, self.PA * self.C), device=device, dtype=torch.bfloat16)

    @classmethod
    def wrap(cls, buf2d: torch.Tensor, H, W, C):
        """View an existing contiguous [N, H W C] bf16 matrix as frames (PA = H W)."""
        if buf2d.dim() != 2 or buf2d.shape[1] != H * W * C or not buf2d.is_contiguous() or buf2d.dtype != torch.bfloat16:
            raise ValueError("wrap: contiguous bf16 [N, H * W * C]")
        o = cls.__new__(cls)
        o.N, o.H, o.W, o.C, o.PA, o.buf = buf2d.shape[0], int(H), int(W), int(C), int(H) * int(W), buf2d
        return o

    _pool = {}

    @classmethod
    def get(cls, N, H, W, C, PA, device):
        """Same small ring of persistent buffers per geometry as RingFrames.get (frozen, no-grad trunk only)."""
        key = (int(N), int(H), int(W), int(C), int(PA), str(device))
        ent = cls._pool.get(key)
        if ent is None:
            if len(cls._pool) > 64 or any(k[0] != key[0] for k in cls._pool):     # one batch geometry at a time (see RingFrames)
                cls._pool.clear()
            ent = cls._pool[key] = [[cls(N, H, W, C, PA, device) for _ in range(6)], 0]
        ent[1] = (ent[1] + 1) % 6
        return ent[0][ent[1]]


class ConvFramesSpec:
    """Filter-pattern matrix + K-extension table of one table-driven convolution (conv_frames)."""

    def __init__(self, w: torch.Tensor, Hin: int, Win: int, stride: int, in_ring: bool):
        Co, Ci, k, k2 = w.shape
        if k != k2 or k not in (1, 3) or Ci % 64 or not ((Co < 256 and 256 % Co == 0 and Co % 64 == 0) or Co % 256 == 0):
            raise ValueError("conv_frames: 1x1 / 3x3 filters, C_in % 64 == 0, C_out in {64, 128} or a multiple of 256")
        pad = 1 if k == 3 else 0
        self.Hout, self.Wout = (Hin + 2 * pad - k) // stride + 1, (Win + 2 * pad - k) // stride + 1
        self.Hin, self.Win, self.Ci, self.Co, self.in_ring = Hin, Win, Ci, Co, bool(in_ring)
        g = self.g = max(1, 256 // Co)                       # output pixels per 256-wide N tile
        cob = max(1, Co // 256)                              # N tiles per output pixel
        P = self.Hout * self.Wout
        n_groups = (P + g - 1) // g
        self.PA = n_groups * g
        self.N = self.PA * Co
        wp = Win + 2 if in_ring else Win
        self.in_cols = ((Hin + 2) * (Win + 2) if in_ring else None)          # pixel slots of a ring-padded input frame
        patterns, tiles = {}, []
        for t in range(n_groups):
            contrib = {}
            for sl in range(g):
                pix = t * g + sl
                if pix >= P:
                    continue
                oy, ox = divmod(pix, self.Wout)
                for ky in range(k):
                    for kx in range(k):
                        iy, ix = oy * stride + ky - pad, ox * stride + kx - pad
                        if 0 <= iy < Hin and 0 <= ix < Win:      # taps on the zero padding are simply absent
                            pin = (iy + 1) * wp + ix + 1 if in_ring else iy * Win + ix
                            contrib.setdefault(pin, [-1] * g)[sl] = ky * k + kx
            tiles.append([(pin, patterns.setdefault(tuple(contrib[pin]), len(patterns))) for pin in sorted(contrib)])
        kb = Ci // 64
        self.n_ext = max(len(e) for e in tiles) * kb
        tab = torch.zeros((1, n_groups * cob, self.n_ext, 4), dtype=torch.int32)
        tab[..., 1] = -1
        for t, ents in enumerate(tiles):
            for cb in range(cob):
                row = [(pin * Ci + 64 * b, cb * 256, pid * Ci + 64 * b, 0) for pin, pid in ents for b in range(kb)]
                tab[0, t * cob + cb, : len(row)] = torch.tensor(row, dtype=torch.int32)
        b2 = torch.zeros((g * Co, len(patterns) * Ci), device=w.device, dtype=torch.bfloat16)
        for pat, pid in patterns.items():
            for sl, tap in enumerate(pat):
                if tap >= 0:
                    b2[sl * Co: (sl + 1) * Co, pid * Ci: (pid + 1) * Ci] = w[:, :, tap // k, tap % k]
        self.b2 = b2.contiguous()
        self.table = tab.contiguous().to(w.device)
        self.macs_per_frame = sum(len(e) for e in tiles) * cob * 256 * Ci    # executed (zero pattern blocks included)


def conv_frames(x, spec: ConvFramesSpec, prelu: Optional[dict] = None) -> FrameRows:
    """Convolution of the small late-stage grids as ONE tcgen05 GEMM launch whose rows are whole frames: the reduction of an
    N tile (the channels of one output pixel, or of 256 / C_out consecutive ones) is the K-extension list of that tile, one
    64-channel block per contributing input pixel.  No ring, no im2col / gather buffer, no flops on the zero padding.
    x: RingFrames (spec.in_ring) or FrameRows.  prelu as in conv3x3s1_ring (residual: FrameRows of the output geometry)."""
    if isinstance(x, RingFrames):
        if not spec.in_ring or (x.H, x.W, x.C) != (spec.Hin, spec.Win, spec.Ci):
            raise ValueError("conv_frames: input geometry does not match the spec")
        a2 = x.rows.view(x.N, x.P * x.C)
    else:
        if spec.in_ring or (x.H, x.W, x.C) != (spec.Hin, spec.Win, spec.Ci):
            raise ValueError("conv_frames: input geometry does not match the spec")
        a2 = x.buf
    out = FrameRows.get(x.N, spec.Hout, spec.Wout, spec.Co, spec.PA, a2.device)
    g = GemmArgs()
    g.out, g.ldo = out.buf.data_ptr(), out.buf.stride(0)
    g.M, g.N, g.K = x.N, spec.N, 0
    g.A2, g.B2, g.ext_table = a2.data_ptr(), spec.b2.data_ptr(), spec.table.data_ptr()
    g.lda2, g.ldb2 = a2.stride(0), spec.b2.stride(0)
    g.a2_cols, g.b2_rows, g.b2_cols = a2.shape[1], spec.b2.shape[0], spec.b2.shape[1]
    g.n_ext = spec.n_ext
    g.block_n, g.pair_aligned, g.alpha = 256, 1, 1.0
    g.act = ACT[None]
    if prelu is not None:
        res = prelu.get("residual")
        require_cuda(prelu["slope"], prelu["bias"], prelu.get("res_bias"))
        if res is not None:
            if (res.N, res.PA, res.C) != (out.N, out.PA, out.C):
                raise ValueError("conv_frames: residual geometry mismatch")
            g.residual, g.ldr = res.buf.data_ptr(), res.buf.stride(0)
        g.act = ACT["prelu_ring"]
        g.bias, g.slope, g.res_bias = prelu["bias"].data_ptr(), prelu["slope"].data_ptr(), ptr(prelu.get("res_bias"))
        g.ring_h, g.ring_w, g.ring_group, g.ring_c = 0, 0, spec.PA, spec.Co
    check(lib.omni_gemm_bf16(C.byref(g), stream_ptr()), "omni_gemm_bf16 (conv_frames)")
    _count()
    return out


def avgpool_frames(x: FrameRows) -> torch.Tensor:
    out = torch.empty((x.N, x.C), device=x.buf.device, dtype=torch.bfloat16)
    check(lib.omni_avgpool_frames(x.buf.data_ptr(), out.data_ptr(), x.N, x.H * x.W, x.PA, x.C, stream_ptr()),
          "omni_avgpool_frames")
    _count()
    return out


def avgpool_ring(x: RingFrames) -> torch.Tensor:
    out = torch.empty((x.N, x.C), device=x.buf.device, dtype=torch.bfloat16)
    check(lib.omni_avgpool_ring(x.rows.data_ptr(), out.data_ptr(), x.N, x.H, x.W, x.C, stream_ptr()), "omni_avgpool_ring")
    _count()
    return out


def attention_fwd(qkv, out, segments, n_heads: int, n_kv_heads: int, head_dim: int, causal: bool, lse=None,
                  scale: Optional[float] = None):
    """tcgen05 flash-attention forward over the packed q|k|v rows (head_dim 64 / 128).  segments = [(task, B, S, row0)];
    out [M, n_heads*head_dim] bf16 (rows outside the segments untouched).  Raises for unsupported head dims."""
    require_cuda(qkv, out, lse)
    qkv = _bf16_2d(qkv, "qkv")
    out = _bf16_2d(out, "out")
    if head_dim not in (64, 128):
        raise NotImplementedError("attention_fwd: head_dim 64 or 128")
    if scale is None:
        scale = head_dim ** -0.5
    for (_, B, S, row0) in segments:
        check(lib.omni_attention_fwd(qkv.data_ptr(), qkv.shape[0], qkv.stride(0), out.data_ptr(), out.stride(0), ptr(lse),
                                     row0, B, S, n_heads, n_kv_heads, head_dim, 1 if causal else 0, float(scale),
                                     stream_ptr()), "omni_attention_fwd")
        _count()
    return out


def attention_bwd(qkv, out, dout, lse, dqkv, segments, n_heads: int, n_kv_heads: int, head_dim: int, causal: bool,
                  scale: Optional[float] = None):
    """tcgen05 flash-attention backward over the packed rows: fills the segments' rows of dqkv [M, q+2kv] with
    dQ | dK | dV.  out / lse [n_heads, M] are the forward results of attention_fwd."""
    require_cuda(qkv, out, dout, lse, dqkv)
    qkv, out, dout, dqkv = (_bf16_2d(t, n) for t, n in ((qkv, "qkv"), (out, "out"), (dout, "dout"), (dqkv, "dqkv")))
    if head_dim not in (64, 128):
        raise NotImplementedError("attention_bwd: head_dim 64 or 128")
    if lse.dtype != torch.float32 or lse.shape != (n_heads, qkv.shape[0]) or not lse.is_contiguous():
        raise ValueError("attention_bwd: lse must be contiguous fp32 [n_heads, M]")
    if scale is None:
        scale = head_dim ** -0.5
    # scratch of the statistics pre-pass: per segment [n_heads][B][ceil(S / 64)][2][64] fp32
    need = max(int(lib.omni_attention_bwd_scratch_floats(B, S, n_heads)) for (_, B, S, _) in segments)
    delta = torch.empty(need, device=lse.device, dtype=torch.float32)
    for (_, B, S, row0) in segments:
        check(lib.omni_attention_bwd(qkv.data_ptr(), qkv.shape[0], qkv.stride(0), out.data_ptr(), out.stride(0),
                                     dout.data_ptr(), dout.stride(0), lse.data_ptr(), delta.data_ptr(), dqkv.data_ptr(),
                                     dqkv.stride(0), row0, B, S, n_heads, n_kv_heads, head_dim, 1 if causal else 0,
                                     float(scale), stream_ptr()), "omni_attention_bwd")
        _count(3)
    return dqkv


DECODE_ATTN_GROUPS = (1, 2, 3, 4, 5, 6, 7, 8)


def decode_attention(qkv, k_cache, v_cache, len_idx, out, B: int, n_heads: int, n_kv_heads: int, head_dim: int,
                     scale: Optional[float] = None, rope: Optional[tuple] = None, beam: Optional[tuple] = None):
    """Single-token attention over the static KV cache [B, n_kv_heads, max_len, head_dim] (one layer): appends the new
    token's K / V (from the packed q|k|v rows, RoPE applied) at position len_idx[0] (int64, device) and attends over
    positions 0..len_idx[0].  out [>=B, n_heads*head_dim] bf16; rows >= B untouched.
    rope = (cos_t, sin_t) bf16 [>= max_len, head_dim]: the packed rows hold q|k|v BEFORE the rotary embedding, which the
    kernel applies to the q heads and the new key at position len_idx[0] (no separate rope_ launch).
    beam = (ind, prefill_len, K): beam-search rows (B = utterances x K); `ind` int32 [2, B, ind_ld] is the KV-cache
    indirection table written by beam_select, `prefill_len` an int64 device scalar (see omni_decode_attention_beam)."""
    require_cuda(qkv, k_cache, v_cache, len_idx, out)
    qkv = _bf16_2d(qkv, "qkv")
    out = _bf16_2d(out, "out")
    if len_idx.dtype != torch.int64:
        raise TypeError("len_idx must be int64")
    if k_cache.dtype != torch.bfloat16 or not k_cache.is_contiguous() or not v_cache.is_contiguous() \
            or tuple(k_cache.shape[:2]) != (B, n_kv_heads) or k_cache.shape[3] != head_dim or k_cache.shape != v_cache.shape:
        raise ValueError("k_cache / v_cache must be contiguous bf16 [B, n_kv_heads, max_len, head_dim]")
    if scale is None:
        scale = head_dim ** -0.5
    if rope is not None:
        cos_t, sin_t = rope
        require_cuda(cos_t, sin_t)
        if cos_t.dtype != torch.bfloat16 or cos_t.shape[-1] != head_dim or not cos_t.is_contiguous() or \
                not sin_t.is_contiguous() or cos_t.shape != sin_t.shape or cos_t.shape[0] < k_cache.shape[2]:
            raise ValueError("rope tables must be contiguous bf16 [>= max_len, head_dim]")
    if beam is not None:
        ind, prefill_len, K = beam
        require_cuda(ind, prefill_len)
        if ind.dtype != torch.int32 or not ind.is_contiguous() or ind.dim() != 3 or ind.shape[0] != 2 or ind.shape[1] != B \
                or prefill_len.dtype != torch.int64 or B % K:
            raise ValueError("beam = (int32 ind [2, B, ind_ld], int64 prefill_len [1], K) with B a multiple of K")
        cp, sp, tr = (rope[0].data_ptr(), rope[1].data_ptr(), rope[0].shape[0]) if rope is not None else (None, None, 0)
        check(lib.omni_decode_attention_beam(qkv.data_ptr(), qkv.stride(0), k_cache.data_ptr(), v_cache.data_ptr(),
                                             len_idx.data_ptr(), out.data_ptr(), out.stride(0), B, n_heads, n_kv_heads,
                                             head_dim, k_cache.shape[2], float(scale), cp, sp, tr, ind.data_ptr(),
                                             ind.shape[2], prefill_len.data_ptr(), int(K), stream_ptr()),
              "omni_decode_attention_beam")
        _count()
        return out
    if rope is not None:
        check(lib.omni_decode_attention_rope(qkv.data_ptr(), qkv.stride(0), k_cache.data_ptr(), v_cache.data_ptr(),
                                             len_idx.data_ptr(), out.data_ptr(), out.stride(0), B, n_heads, n_kv_heads,
                                             head_dim, k_cache.shape[2], float(scale), cos_t.data_ptr(), sin_t.data_ptr(),
                                             cos_t.shape[0], stream_ptr()), "omni_decode_attention_rope")
        _count()
        return out
    check(lib.omni_decode_attention(qkv.data_ptr(), qkv.stride(0), k_cache.data_ptr(), v_cache.data_ptr(),
                                    len_idx.data_ptr(), out.data_ptr(), out.stride(0), B, n_heads, n_kv_heads, head_dim,
                                    k_cache.shape[2], float(scale), stream_ptr()), "omni_decode_attention")
    _count()
    return out
