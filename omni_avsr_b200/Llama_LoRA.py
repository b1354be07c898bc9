"""Drop-in mirror of the reference's Omni_AVSR/Llama_LoRA.py on top of the sm_100a kernels.

Same public names and call semantics as the reference (file:line refer to /root/reference/Omni_AVSR/Llama_LoRA.py):
  LoRA_config                     :103-110
  LlamaSdpaAttention_lora         :113-316   (LoRA on q/v: shared / task-specific / task-specific+shared)
  LlamaDecoderLayer_lora          :580-655
  LlamaModel_lora                 :446-578
  LlamaForCausalLM_lora           :318-444   (.forward(inputs_embeds, labels, modality) -> .loss/.logits, .generate)

What is different underneath: the frozen base weights are packed ([q;k;v], [gate;up]) and kept with a transposed
copy for the dgrad GEMMs; all trainable adapter weights live in ONE flat bf16 buffer (views are exposed under the
reference's state-dict names); tokens of all tasks are packed into one [rows, H] buffer whose 128-row tiles carry a
task id, so the adapter is selected per tile inside the tcgen05 GEMM and one pass serves ASR+VSR+AVSR.
There is no CPU path: every op below launches a kernel from libomni_avsr.so.
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import torch
from torch import nn

from . import autograd_ops as ag
from . import ops

TASKS = ("audio", "video", "audiovisual")
TASK_ID = {t: i for i, t in enumerate(TASKS)}
IGNORE_INDEX = -100
TILE = 128


@dataclass
class LoRA_config:
    RANK: int
    ALPHA: int = 1
    IS_LLAMA3: bool = False
    IS_LLAMA3_2_3B: bool = False
    IS_TASK_SPECIFIC: bool = False
    SHARED_LORA: bool = False


@dataclass
class LLMArch:
    """Architecture constants of the named checkpoints (no hub access: built from this table, random init)."""
    family: str
    hidden_size: int
    intermediate_size: int
    num_hidden_layers: int
    num_attention_heads: int
    num_key_value_heads: int
    vocab_size: int
    rms_norm_eps: float
    rope_theta: float
    head_dim: int
    rope_scaling: Optional[dict] = None
    attention_bias: bool = False
    tie_word_embeddings: bool = True
    max_position_embeddings: int = 131072
    pad_token_id: Optional[int] = None
    inv_freq_dtype: str = "bf16"   # reference quirk: Lightning bf16-true rounds Llama's inv_freq buffer to bf16
    use_return_dict: bool = True

    @property
    def q_dim(self):
        return self.num_attention_heads * self.head_dim

    @property
    def kv_dim(self):
        return self.num_key_value_heads * self.head_dim


_L3 = dict(low_freq_factor=1.0, high_freq_factor=4.0, original_max_position_embeddings=8192)
ARCHS = {
    "meta-llama/Llama-3.2-1B": dict(family="llama", hidden_size=2048, intermediate_size=8192, num_hidden_layers=16,
                                    num_attention_heads=32, num_key_value_heads=8, vocab_size=128256, rms_norm_eps=1e-5,
                                    rope_theta=500000.0, head_dim=64, rope_scaling=dict(factor=32.0, **_L3),
                                    tie_word_embeddings=True),
    "meta-llama/Llama-3.2-3B": dict(family="llama", hidden_size=3072, intermediate_size=8192, num_hidden_layers=28,
                                    num_attention_heads=24, num_key_value_heads=8, vocab_size=128256, rms_norm_eps=1e-5,
                                    rope_theta=500000.0, head_dim=128, rope_scaling=dict(factor=32.0, **_L3),
                                    tie_word_embeddings=True),
    "meta-llama/Meta-Llama-3.1-8B": dict(family="llama", hidden_size=4096, intermediate_size=14336,
                                         num_hidden_layers=32, num_attention_heads=32, num_key_value_heads=8,
                                         vocab_size=128256, rms_norm_eps=1e-5, rope_theta=500000.0, head_dim=128,
                                         rope_scaling=dict(factor=8.0, **_L3), tie_word_embeddings=False),
    "meta-llama/Meta-Llama-3-8B": dict(family="llama", hidden_size=4096, intermediate_size=14336, num_hidden_layers=32,
                                       num_attention_heads=32, num_key_value_heads=8, vocab_size=128256,
                                       rms_norm_eps=1e-5, rope_theta=500000.0, head_dim=128, rope_scaling=None,
                                       tie_word_embeddings=False, max_position_embeddings=8192),
    "Qwen/Qwen2.5-0.5B": dict(family="qwen2", hidden_size=896, intermediate_size=4864, num_hidden_layers=24,
                              num_attention_heads=14, num_key_value_heads=2, vocab_size=151936, rms_norm_eps=1e-6,
                              rope_theta=1000000.0, head_dim=64, attention_bias=True, tie_word_embeddings=True,
                              max_position_embeddings=32768, inv_freq_dtype="fp32"),
    "Qwen/Qwen2.5-1.5B": dict(family="qwen2", hidden_size=1536, intermediate_size=8960, num_hidden_layers=28,
                              num_attention_heads=12, num_key_value_heads=2, vocab_size=151936, rms_norm_eps=1e-6,
                              rope_theta=1000000.0, head_dim=128, attention_bias=True, tie_word_embeddings=True,
                              max_position_embeddings=32768, inv_freq_dtype="fp32"),
    "Qwen/Qwen2.5-3B": dict(family="qwen2", hidden_size=2048, intermediate_size=11008, num_hidden_layers=36,
                            num_attention_heads=16, num_key_value_heads=2, vocab_size=151936, rms_norm_eps=1e-6,
                            rope_theta=1000000.0, head_dim=128, attention_bias=True, tie_word_embeddings=True,
                            max_position_embeddings=32768, inv_freq_dtype="fp32"),
    "Qwen/Qwen2.5-7B": dict(family="qwen2", hidden_size=3584, intermediate_size=18944, num_hidden_layers=28,
                            num_attention_heads=28, num_key_value_heads=4, vocab_size=152064, rms_norm_eps=1e-6,
                            rope_theta=1000000.0, head_dim=128, attention_bias=True, tie_word_embeddings=False,
                            max_position_embeddings=32768, inv_freq_dtype="fp32"),
    "Qwen/Qwen2.5-14B": dict(family="qwen2", hidden_size=5120, intermediate_size=13824, num_hidden_layers=48,
                             num_attention_heads=40, num_key_value_heads=8, vocab_size=152064, rms_norm_eps=1e-5,
                             rope_theta=1000000.0, head_dim=128, attention_bias=True, tie_word_embeddings=False,
                             max_position_embeddings=32768, inv_freq_dtype="fp32"),
    "Qwen/Qwen2.5-32B": dict(family="qwen2", hidden_size=5120, intermediate_size=27648, num_hidden_layers=64,
                             num_attention_heads=40, num_key_value_heads=8, vocab_size=152064, rms_norm_eps=1e-5,
                             rope_theta=1000000.0, head_dim=128, attention_bias=True, tie_word_embeddings=False,
                             max_position_embeddings=32768, inv_freq_dtype="fp32"),
}


def arch_from_name(name: str, **overrides) -> LLMArch:
    if name not in ARCHS:
        raise KeyError(f"unknown llm_model {name!r}; known: {sorted(ARCHS)}")
    d = dict(ARCHS[name])
    d.update(overrides)
    return LLMArch(**d)


# --------------------------------------------------------------------------------------------------
# flat trainable-parameter store
# --------------------------------------------------------------------------------------------------
class FlatParams:
    """One flat bf16 buffer for every trainable tensor (+ one flat grad buffer).  Parameters handed out are views,
    so the fused clip+AdamW kernel and the NCCL all-reduce run over a single contiguous range."""

    def __init__(self, device, capacity: int):
        self.device = device
        self.data = torch.zeros(capacity, device=device, dtype=torch.bfloat16)
        self.grad = torch.zeros(capacity, device=device, dtype=torch.bfloat16)
        self.used = 0
        self.names: List[Tuple[str, int, Tuple[int, ...]]] = []
        self.params: List[nn.Parameter] = []

    def alloc(self, shape: Sequence[int], name: str = "") -> nn.Parameter:
        n = int(math.prod(shape))
        n_al = (n + 7) // 8 * 8     # keep every tensor 16-byte aligned for TMA / vector loads
        if self.used + n_al > self.data.numel():
            raise RuntimeError("FlatParams capacity exceeded")
        view = self.data[self.used: self.used + n].view(*shape)
        p = nn.Parameter(view, requires_grad=True)
        p.grad = self.grad[self.used: self.used + n].view(*shape)
        self.names.append((name, self.used, tuple(shape)))
        self.params.append(p)
        self.used += n_al
        return p

    def trainable_spans(self) -> List[Tuple[int, int]]:
        """Maximal [start, end) element ranges of the flat buffer made of tensors with requires_grad=True (alignment padding
        between two trainable neighbours included).  torch.optim.AdamW skips parameters without a gradient -- weight
        decay included -- so the fused optimizer kernel must not touch the frozen ranges either."""
        spans: List[Tuple[int, int]] = []
        for (name, off, shape), p in zip(self.names, self.params):
            if not p.requires_grad:
                continue
            n_al = (int(math.prod(shape)) + 7) // 8 * 8
            if spans and spans[-1][1] == off:
                spans[-1] = (spans[-1][0], off + n_al)
            else:
                spans.append((off, off + n_al))
        return spans

    def flat(self):
        return self.data[: self.used], self.grad[: self.used]


# --------------------------------------------------------------------------------------------------
# packed token rows
# --------------------------------------------------------------------------------------------------
class PackedRows:
    """Layout of the token rows of one LLM pass: a list of segments (task id, B, S), each starting on a 128-row
    boundary so that no GEMM tile straddles two tasks."""

    _cache: Dict[tuple, "PackedRows"] = {}

    def __init__(self, segments: Sequence[Tuple[int, int, int]], device, pos_offset: int = 0):
        self.segments = []
        off = 0
        groups, pos = [], []
        # segments start on 256-row boundaries when there are several of them: a CTA PAIR (two 128-row tiles) then never
        # straddles two tasks and the K-extended LoRA GEMM can run on the cta_group::2 kernel
        align = 2 * TILE if len(segments) > 1 else TILE
        self.pair_aligned = True            # one segment = one task for every tile; several = 256-row aligned
        for (task, B, S) in segments:
            n = B * S
            n_pad = (n + align - 1) // align * align
            self.segments.append((task, B, S, off))
            groups += [task] * (n_pad // TILE)
            p = (torch.arange(S, dtype=torch.int32) + pos_offset).repeat(B)
            pos.append(torch.cat([p, torch.zeros(n_pad - n, dtype=torch.int32)]))
            off += n_pad
        self.M = off
        self.max_pos = pos_offset + max(S for _, _, S in segments)
        self.runs = []                      # maximal runs of equal task id: (task, row0, row1)
        for (task, B, S, o) in self.segments:
            end = o + (B * S + align - 1) // align * align
            if self.runs and self.runs[-1][0] == task and self.runs[-1][2] == o:
                self.runs[-1] = (task, self.runs[-1][1], end)
            else:
                self.runs.append((task, o, end))
        self.tile_group = torch.tensor(groups, dtype=torch.int32, device=device)
        self.pos = torch.cat(pos).to(device)
        self.valid_rows = sum(B * S for _, B, S, _ in self.segments)

    @classmethod
    def get(cls, segments, device, pos_offset: int = 0) -> "PackedRows":
        key = (tuple(segments), str(device), pos_offset)
        if key not in cls._cache:
            if len(cls._cache) > 256:
                cls._cache.clear()
            cls._cache[key] = cls(segments, device, pos_offset)
        return cls._cache[key]


# --------------------------------------------------------------------------------------------------
# LoRA plan: static tables of the grouped / K-extended GEMMs
# --------------------------------------------------------------------------------------------------
class LoraPlan:
    """Static description of one Omni-LoRA adapted q|k|v projection.

    slots: S -> 1 adapter; T -> 3 (audio, video, audiovisual); ST -> 4 (+shared).  Per token n_act adapters are
    active (1, or 2 with the shared one).  Packed weights:
        down [(2 * n_slots) * rp, H]  rows = [q slot 0.. | v slot 0..], rp = r rounded up to 64 (pad rows stay 0)
        up   [n_slots * q_cols + n_slots * v_cols, rp]
    T (phase-1 output) [M, 2 * n_act * rp] = [Tq_task | (Tq_shared) | Tv_task | (Tv_shared)].
    """

    def __init__(self, H, q_cols, k_cols, v_cols, r, scaling, task_specific, shared, device):
        self.H, self.q_cols, self.k_cols, self.v_cols, self.r = H, q_cols, k_cols, v_cols, r
        self.scaling = float(scaling)
        self.task_specific, self.shared = bool(task_specific), bool(shared)
        self.n_slots = (3 if task_specific else 1) + (1 if (task_specific and shared) else 0)
        self.n_act = 2 if (task_specific and shared) else 1
        self.rp = (r + 63) // 64 * 64
        self.cb = self.rp // 64                      # 64-wide K blocks per adapter
        self.t_cols = 2 * self.n_act * self.rp
        self.v_col0 = q_cols + k_cols
        N = q_cols + k_cols + v_cols
        self.N = N
        bn = 256
        while bn > 64 and (q_cols % bn or k_cols % bn or v_cols % bn):
            bn //= 2
        if q_cols % bn or k_cols % bn or v_cols % bn:
            raise ValueError("q/k/v widths must be multiples of 64")
        self.block_n = bn
        self.block_n_bwd = 256 if H % 256 == 0 else (128 if H % 128 == 0 else 64)
        self.down_rows = 2 * self.n_slots * self.rp
        self.up_rows = self.n_slots * (q_cols + v_cols)
        self._build(device)

    def slot_of(self, a: int, g: int) -> int:
        """adapter a (0 = task adapter, 1 = shared) for task group g -> slot index."""
        if a == 1:
            return self.n_slots - 1
        return g if self.task_specific else 0

    def _build(self, device):
        rp, cb, ns = self.rp, self.cb, self.n_slots
        # phase 1 (block_n 64): tile nt -> (part q/v, adapter a, chunk c)
        nt1 = self.t_cols // 64
        brow = torch.zeros((3, nt1), dtype=torch.int32)
        for g in range(3):
            for nt in range(nt1):
                part, rem = divmod(nt, self.n_act * cb)
                a, c = divmod(rem, cb)
                brow[g, nt] = (part * ns + self.slot_of(a, g)) * rp + c * 64
        self.brow_fwd = brow.to(device)
        # phase 2: K-extension blocks per (group, n_tile)
        bn = self.block_n
        nt2 = self.N // bn
        n_ext = self.n_act * cb
        ext = torch.full((3, nt2, n_ext, 4), -1, dtype=torch.int32)
        for g in range(3):
            for nt in range(nt2):
                n0 = nt * bn
                if n0 < self.q_cols:
                    part, nn0, width, base = 0, n0, self.q_cols, 0
                elif n0 >= self.v_col0:
                    part, nn0, width, base = 1, n0 - self.v_col0, self.v_cols, ns * self.q_cols
                else:
                    continue
                for a in range(self.n_act):
                    for c in range(cb):
                        ext[g, nt, a * cb + c] = torch.tensor(
                            [(part * self.n_act + a) * rp + c * 64, base + self.slot_of(a, g) * width + nn0, c * 64, 0])
        self.ext_fwd = ext.contiguous().to(device)
        # decode step (weight-streaming kernel, 128-feature tiles); None when q / k / v widths are not multiples of 128
        self.ext_fwd_step = None
        if self.q_cols % 128 == 0 and self.k_cols % 128 == 0 and self.v_cols % 128 == 0:
            self.ext_fwd_step = self.ext_fwd if bn == 128 else self._ext_table(128).to(device)
        self.ext_fwd_64 = self.ext_fwd if bn == 64 else self._ext_table(64).to(device)
        # backward phase 1': dT_part = s * dOut_part @ upT_part, upT_part [n_slots*rp, width]; tile -> (a, c)
        nt1b = self.n_act * cb
        browb = torch.zeros((3, nt1b), dtype=torch.int32)
        for g in range(3):
            for nt in range(nt1b):
                a, c = divmod(nt, cb)
                browb[g, nt] = self.slot_of(a, g) * rp + c * 64
        self.brow_bwd = browb.to(device)
        # backward phase 2': dh = dOut @ W + dT' @ down[sel]; B2 = down^T [H, down_rows]
        bnb = self.block_n_bwd
        nt2b = (self.H + bnb - 1) // bnb
        n_extb = 2 * self.n_act * cb
        extb = torch.full((3, nt2b, n_extb, 4), -1, dtype=torch.int32)
        for g in range(3):
            for nt in range(nt2b):
                j = 0
                for part in range(2):
                    for a in range(self.n_act):
                        for c in range(cb):
                            extb[g, nt, j] = torch.tensor(
                                [(part * self.n_act + a) * rp + c * 64, nt * bnb,
                                 (part * ns + self.slot_of(a, g)) * rp + c * 64, 0])
                            j += 1
        self.ext_bwd = extb.contiguous().to(device)

    def _ext_table(self, bn: int) -> torch.Tensor:
        """K-extension blocks per (group, n_tile) for an N tile width of bn (same content as ext_fwd, other tiling)."""
        rp, cb, ns = self.rp, self.cb, self.n_slots
        nt2 = self.N // bn
        ext = torch.full((3, nt2, self.n_act * cb, 4), -1, dtype=torch.int32)
        for g in range(3):
            for nt in range(nt2):
                n0 = nt * bn
                if n0 < self.q_cols:
                    part, nn0, width, base = 0, n0, self.q_cols, 0
                elif n0 >= self.v_col0:
                    part, nn0, width, base = 1, n0 - self.v_col0, self.v_cols, ns * self.q_cols
                else:
                    continue
                for a in range(self.n_act):
                    for c in range(cb):
                        ext[g, nt, a * cb + c] = torch.tensor(
                            [(part * self.n_act + a) * rp + c * 64, base + self.slot_of(a, g) * width + nn0, c * 64, 0])
        return ext.contiguous()

    def transposed_up(self, up: torch.Tensor):
        ns, rp = self.n_slots, self.rp
        up = up.detach()
        uq = ops.transpose(up[: ns * self.q_cols].view(ns, self.q_cols, rp)).view(ns * rp, self.q_cols)
        uv = ops.transpose(up[ns * self.q_cols:].view(ns, self.v_cols, rp)).view(ns * rp, self.v_cols)
        return uq, uv

    def _scatter_index(self, runs, device):
        key = tuple(g for g, _, _ in runs)
        cache = self.__dict__.setdefault("_idx_cache", {})
        if key not in cache:
            ns, na = self.n_slots, self.n_act
            down = [part * ns + self.slot_of(a, g) for g in key for part in range(2) for a in range(na)]
            up = [self.slot_of(a, g) for g in key for a in range(na)]
            cache[key] = (torch.tensor(down, dtype=torch.int64, device=device),
                          torch.tensor(up, dtype=torch.int64, device=device))
        return cache[key]

    def wgrads(self, h, T, dT, dout, runs):
        """LoRA weight gradients with the MN-major tcgen05 kernel (csrc/gemm_wgrad.cu), one token range per task run:
             d_down[slot] = sum_runs dT'[rows]^T h[rows]          d_up[slot] = sum_runs dOut_part[rows]^T T[rows]
        fp32 partial results per run are scatter-added into their adapter slots (the shared adapter sums over runs)."""
        ns, rp, na = self.n_slots, self.rp, self.n_act
        H = h.shape[1]
        # the outputs are small ([256, H] / [Hq, 128] per run): a run per CTA column leaves most SMs idle, so long runs are
        # cut into 128-row-aligned pieces (more token ranges = more CTAs; the pieces land in the same adapter slot below)
        pieces = max(1, min(8 // max(len(runs), 1), 4))
        if pieces > 1:
            cut = []
            for g, r0, r1 in runs:
                step = ((r1 - r0 + pieces - 1) // pieces + 127) // 128 * 128
                cut += [(g, a, min(a + step, r1)) for a in range(r0, r1, max(step, 128))]
            runs = cut[:8] if len(cut) <= 8 else runs
        Z = len(runs)
        ranges = [(r0, r1) for _, r0, r1 in runs]
        idx_down, idx_up = self._scatter_index(runs, h.device)
        pd = ops.gemm_wgrad(dT, h, mo=self.t_cols, no=H, ranges=ranges, out_dtype=torch.float32)        # [Z, 2*na*rp, H]
        d_down = torch.zeros((2 * ns, rp, H), device=h.device, dtype=torch.float32)
        d_down.index_add_(0, idx_down, pd.view(Z * 2 * na, rp, H))
        ups = []
        for part, (c0, width) in enumerate(((0, self.q_cols), (self.v_col0, self.v_cols))):
            pu = ops.gemm_wgrad(dout, T, mo=width, no=na * rp, a_col0=c0, b_col0=part * na * rp, ranges=ranges,
                                out_dtype=torch.float32)                                                 # [Z, width, na*rp]
            du = torch.zeros((ns, width, rp), device=h.device, dtype=torch.float32)
            du.index_add_(0, idx_up, pu.view(Z, width, na, rp).permute(0, 2, 1, 3).reshape(Z * na, width, rp))
            ups.append(du.view(ns * width, rp))
        return d_down.view(2 * ns * rp, H).to(torch.bfloat16), torch.cat(ups, dim=0).to(torch.bfloat16)


class _LinearView(nn.Module):
    """Reference-named holder of a weight (and bias) that are views into packed storage."""

    def __init__(self, weight: torch.Tensor, bias: Optional[torch.Tensor] = None, trainable: bool = False):
        super().__init__()
        self.weight = weight if isinstance(weight, nn.Parameter) else nn.Parameter(weight, requires_grad=trainable)
        if bias is not None:
            self.bias = bias if isinstance(bias, nn.Parameter) else nn.Parameter(bias, requires_grad=trainable)
        else:
            self.bias = None


def rope_tables(arch: LLMArch, max_pos: int, device):
    """cos/sin bf16 tables [max_pos, head_dim] computed exactly as LlamaRotaryEmbedding does under bf16-true
    (fp32 angles from (optionally bf16-rounded) inverse frequencies, cat(freqs, freqs), cast to bf16)."""
    dim = arch.head_dim
    inv_freq = 1.0 / (arch.rope_theta ** (torch.arange(0, dim, 2, dtype=torch.int64).to(torch.float) / dim))
    if arch.rope_scaling is not None:
        rs = arch.rope_scaling
        factor, lo, hi, old = rs["factor"], rs["low_freq_factor"], rs["high_freq_factor"], rs["original_max_position_embeddings"]
        wavelen = 2 * math.pi / inv_freq
        inv_l = torch.where(wavelen > old / lo, inv_freq / factor, inv_freq)
        smooth = (old / wavelen - lo) / (hi - lo)
        smoothed = (1 - smooth) * inv_l / factor + smooth * inv_l
        is_medium = ~(wavelen < old / hi) * ~(wavelen > old / lo)
        inv_freq = torch.where(is_medium, smoothed, inv_l)
    if arch.inv_freq_dtype == "bf16":
        inv_freq = inv_freq.to(torch.bfloat16)
    pos = torch.arange(max_pos, dtype=torch.float32)
    freqs = (inv_freq.float()[None, :, None] @ pos[None, None, :]).transpose(1, 2)[0]
    emb = torch.cat((freqs, freqs), dim=-1)
    return emb.cos().to(torch.bfloat16).to(device).contiguous(), emb.sin().to(torch.bfloat16).to(device).contiguous()


def _kv_out_dim(arch: LLMArch, lc) -> int:
    """lora_up_V width as the reference derives it from the model-family flags (:143-163)."""
    h = arch.hidden_size
    if getattr(lc, "IS_LLAMA3", False):
        return h // 4
    if getattr(lc, "IS_LLAMA3_2_3B", False):
        return h // 3
    return h


class LlamaSdpaAttention_lora(nn.Module):
    def __init__(self, config: LLMArch, lora_config, layer_idx: Optional[int] = None, flat: Optional[FlatParams] = None,
                 device="cuda", kv_out_dim=None):
        super().__init__()
        self.config, self.lora_config, self.layer_idx = config, lora_config, layer_idx
        a = config
        self.num_heads, self.num_key_value_heads, self.head_dim = a.num_attention_heads, a.num_key_value_heads, a.head_dim
        self.rank = lora_config.RANK
        self.scaling = lora_config.ALPHA / self.rank
        H = a.hidden_size
        r = round(H / self.rank)
        vo = kv_out_dim if kv_out_dim is not None else _kv_out_dim(a, lora_config)
        if vo != a.kv_dim:
            raise ValueError(f"lora_up_V width {vo} (from the LoRA config flags) != num_kv_heads*head_dim {a.kv_dim}: "
                             "the reference would fail at `value_states + V_lora` (Llama_LoRA.py:259)")
        std = 0.02
        # frozen packed base weights
        self.qkv_weight = (torch.randn(a.q_dim + 2 * a.kv_dim, H, device=device) * std).to(torch.bfloat16)
        self.qkv_bias = torch.zeros(a.q_dim + 2 * a.kv_dim, device=device, dtype=torch.bfloat16) if a.attention_bias else None
        q0, k0, v0 = 0, a.q_dim, a.q_dim + a.kv_dim
        qb = self.qkv_bias
        self.q_proj = _LinearView(self.qkv_weight[q0:k0], None if qb is None else qb[q0:k0])
        self.k_proj = _LinearView(self.qkv_weight[k0:v0], None if qb is None else qb[k0:v0])
        self.v_proj = _LinearView(self.qkv_weight[v0:], None if qb is None else qb[v0:])
        self.o_proj = _LinearView((torch.randn(H, a.q_dim, device=device) * std).to(torch.bfloat16))
        # trainable adapters in the flat store
        self.plan = LoraPlan(H, a.q_dim, a.kv_dim, a.kv_dim, r, self.scaling, lora_config.IS_TASK_SPECIFIC,
                             lora_config.SHARED_LORA, device)
        p = self.plan
        own_flat = flat is None
        if own_flat:
            flat = FlatParams(device, (p.down_rows * H + p.up_rows * p.rp) + 64)
        self.lora_down = flat.alloc((p.down_rows, H), f"layers.{layer_idx}.lora_down")
        self.lora_up = flat.alloc((p.up_rows, p.rp), f"layers.{layer_idx}.lora_up")
        self._expose_reference_names(r)
        self.reset_lora_parameters()
        self._wt = None

    # --- reference-named views (state-dict keys of §5.4) -------------------------------------------
    def _slot_views(self, slot: int, r: int):
        p = self.plan
        dq = self.lora_down.data[slot * p.rp: slot * p.rp + r]
        dv = self.lora_down.data[(p.n_slots + slot) * p.rp: (p.n_slots + slot) * p.rp + r]
        uq = self.lora_up.data[slot * p.q_cols: (slot + 1) * p.q_cols, :r]
        ub = p.n_slots * p.q_cols
        uv = self.lora_up.data[ub + slot * p.v_cols: ub + (slot + 1) * p.v_cols, :r]
        return dq, dv, uq, uv

    def _expose_reference_names(self, r: int):
        lc = self.lora_config
        mk = lambda t: _LinearView(t, trainable=False)   # aliases of the packed trainable tensors (state-dict names)
        if lc.IS_TASK_SPECIFIC:
            views = [self._slot_views(i, r) for i in range(3)]
            self.lora_down_Q = nn.ModuleDict({t: mk(views[i][0]) for i, t in enumerate(TASKS)})
            self.lora_down_V = nn.ModuleDict({t: mk(views[i][1]) for i, t in enumerate(TASKS)})
            self.lora_up_Q = nn.ModuleDict({t: mk(views[i][2]) for i, t in enumerate(TASKS)})
            self.lora_up_V = nn.ModuleDict({t: mk(views[i][3]) for i, t in enumerate(TASKS)})
            if lc.SHARED_LORA:
                dq, dv, uq, uv = self._slot_views(3, r)
                self.lora_down_Q_shared, self.lora_down_V_shared = mk(dq), mk(dv)
                self.lora_up_Q_shared, self.lora_up_V_shared = mk(uq), mk(uv)
        else:
            dq, dv, uq, uv = self._slot_views(0, r)
            self.lora_down_Q, self.lora_down_V, self.lora_up_Q, self.lora_up_V = mk(dq), mk(dv), mk(uq), mk(uv)

    def reset_lora_parameters(self, down_std: Optional[float] = None):
        """Reference init (:165-175,:189-192): down = 0, up = kaiming_uniform(a=sqrt(5)); down_std != None gives the
        non-degenerate init used by the parity tests / benchmark."""
        p, r = self.plan, round(self.config.hidden_size / self.rank)
        with torch.no_grad():
            self.lora_down.zero_()
            self.lora_up.zero_()
            for slot in range(p.n_slots):
                dq, dv, uq, uv = self._slot_views(slot, r)
                for u in (uq, uv):
                    bound = 1.0 / math.sqrt(r)     # kaiming_uniform_(a=sqrt(5)) on [out, r] => U(-1/sqrt(r), 1/sqrt(r))
                    u.uniform_(-bound, bound)
                if down_std:
                    dq.normal_(0, down_std)
                    dv.normal_(0, down_std)

    # The packed `lora_down` / `lora_up` tensors are the real (trainable) Parameters; the reference-named views above
    # alias the same storage and are what state_dict() / load_state_dict() see (keys of SURVEY §5.4).
    def state_dict(self, *args, destination=None, prefix="", keep_vars=False):
        sd = super().state_dict(*args, destination=destination, prefix=prefix, keep_vars=keep_vars)
        for k in (prefix + "lora_down", prefix + "lora_up"):
            sd.pop(k, None)
        return sd

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs):
        for k in ("lora_down", "lora_up"):      # packed tensors are filled through the aliased views
            if prefix + k not in state_dict:
                state_dict[prefix + k] = getattr(self, k).data
                self._injected = True
        super()._load_from_state_dict(state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs)
        if getattr(self, "_injected", False):
            for k in ("lora_down", "lora_up"):
                state_dict.pop(prefix + k, None)
            self._injected = False
        self._wt = None

    def transposed(self):
        if self._wt is None:
            self._wt = (self.qkv_weight.t().contiguous(), self.o_proj.weight.data.t().contiguous())
        return self._wt

    def forward(self, h, rows: PackedRows, cos_t, sin_t, kv_cache=None, residual=None, norm=None):
        """h [M, H] normed hidden rows -> residual + o_proj(attn).  norm=(weight, eps) (decode step only): also returns the
        RMSNorm of the result, computed in the o_proj launch."""
        a = self.config
        wt_qkv, wt_o = self.transposed()
        if ag.skinny_ok(h) and self.plan.ext_fwd_step is not None:
            # decode step: both LoRA phases on the weight-streaming kernel (no autograd node: nothing to differentiate)
            p = self.plan
            T = ops.gemm(h, self.lora_down.data, n=p.t_cols, alpha=p.scaling, tile_group=rows.tile_group,
                         b_row_table=p.brow_fwd, skinny=True)
            qkv = ops.gemm(h, self.qkv_weight, bias=self.qkv_bias, tile_group=rows.tile_group,
                           ext=(T, self.lora_up.data, p.ext_fwd_step), block_n=128, skinny=True)
        else:
            qkv = ag.LoraLinearFn.apply(h, self.qkv_weight, wt_qkv, self.qkv_bias, self.lora_down, self.lora_up,
                                        rows, self.plan)
        n_rot = self.num_heads + self.num_key_value_heads
        step = kv_cache is not None and kv_cache.graph_mode and all(S == 1 for (_, _, S, _) in rows.segments)
        if step:
            # decode step: the rotary embedding is applied inside the single-token attention kernel (position = cache length)
            attn = attention_packed(qkv, rows, a, kv_cache, self.layer_idx, rope=(cos_t, sin_t))
        else:
            if qkv.requires_grad:
                qkv = ag.RopeFn.apply(qkv, cos_t, sin_t, rows.pos, n_rot, self.head_dim, True)
            else:
                ops.rope_(qkv, cos_t, sin_t, rows.pos, n_rot, self.head_dim)
            attn = attention_packed(qkv, rows, a, kv_cache, self.layer_idx)
        if ag.skinny_ok(attn, residual):
            return ops.gemm(attn, self.o_proj.weight.data, residual=residual, skinny=True, norm=norm)      # decode step
        assert norm is None
        return ag.frozen_linear(attn, self.o_proj.weight.data, wt_o, residual=residual, block_n=256)


class PackedSdpaFn(torch.autograd.Function):
    """Attention core over packed q|k|v rows -> [M, q_dim], on our tcgen05 flash kernels (csrc/attention.cu,
    csrc/attention_bwd.cu) for head_dim 64 / 128 -- every named architecture.

    Forward: one launch per segment, straight from / into the packed rows, log-sum-exp kept for the backward.
    Backward: dQ | dK | dV written into ONE packed dqkv buffer (autograd's own slice gradients would materialise three
    zero-filled [M, q+2kv] tensors per segment and add them up).  There is no library fallback: other head dims raise."""

    @staticmethod
    def _views(qkv, B, S, off, nh, nkv, hd):
        blk = qkv[off: off + B * S]
        q = blk[:, : nh * hd].view(B, S, nh, hd).transpose(1, 2)
        k = blk[:, nh * hd: (nh + nkv) * hd].view(B, S, nkv, hd).transpose(1, 2)
        v = blk[:, (nh + nkv) * hd:].view(B, S, nkv, hd).transpose(1, 2)
        return q, k, v

    @staticmethod
    def _zero_pad_rows(t, segments):
        cur = 0
        for (_, B, S, off) in segments:
            if off > cur:
                t[cur:off].zero_()
            cur = off + B * S
        if cur < t.shape[0]:
            t[cur:].zero_()

    @staticmethod
    def forward(ctx, qkv, segments, nh, nkv, hd, causal):
        out = torch.empty((qkv.shape[0], nh * hd), device=qkv.device, dtype=torch.bfloat16)
        lse = None
        if ctx.needs_input_grad[0]:
            lse = torch.empty((nh, qkv.shape[0]), device=qkv.device, dtype=torch.float32)
        ops.attention_fwd(qkv, out, segments, nh, nkv, hd, causal, lse=lse)      # raises for head dims other than 64 / 128
        PackedSdpaFn._zero_pad_rows(out, segments)
        if lse is not None:
            ctx.save_for_backward(qkv, out, lse)
        else:
            ctx.save_for_backward(qkv)
        ctx.meta = (segments, nh, nkv, hd, causal)
        return out

    @staticmethod
    def backward(ctx, dout):
        segments, nh, nkv, hd, causal = ctx.meta
        qkv = ctx.saved_tensors[0]
        dqkv = torch.empty_like(qkv)
        _, out, lse = ctx.saved_tensors
        ops.attention_bwd(qkv, out, dout.contiguous(), lse, dqkv, segments, nh, nkv, hd, causal)
        PackedSdpaFn._zero_pad_rows(dqkv, segments)
        return dqkv, None, None, None, None, None


def attention_packed(qkv, rows: PackedRows, a: LLMArch, kv_cache, layer_idx, rope=None):
    """Causal GQA attention per segment of the packed rows.  Training / prefill: the tcgen05 flash kernel on the packed
    buffer (the prefill also fills the static KV cache); decode step: the single-token kernel that appends K / V to the
    cache and attends over it in one launch.  No library attention anywhere: unsupported geometries raise."""
    nh, nkv, hd = a.num_attention_heads, a.num_key_value_heads, a.head_dim
    if kv_cache is None:
        return PackedSdpaFn.apply(qkv, rows.segments, nh, nkv, hd, True)
    out = torch.empty((rows.M, a.q_dim), device=qkv.device, dtype=torch.bfloat16)
    for (task, B, S, off) in rows.segments:
        if S == 1 and kv_cache.graph_mode:
            ops.decode_attention(qkv[off: off + B], kv_cache.k[layer_idx], kv_cache.v[layer_idx], kv_cache.len_idx,
                                 out[off: off + B], B, nh, nkv, hd, rope=rope, beam=kv_cache.beam)
            continue
        if rope is not None:
            raise RuntimeError("fused RoPE is a decode-step feature")
        if kv_cache.len != 0 or kv_cache.graph_mode:
            raise NotImplementedError("multi-token steps on top of a non-empty KV cache are not part of the Omni-AVSR decode "
                                      "path (HF generate: one prefill, then single-token steps)")
        _, k, v = PackedSdpaFn._views(qkv, B, S, off, nh, nkv, hd)
        kv_cache.fill(layer_idx, k, v)
        ops.attention_fwd(qkv, out, [(task, B, S, off)], nh, nkv, hd, True)
    if rows.valid_rows != rows.M:
        PackedSdpaFn._zero_pad_rows(out, rows.segments)
    return out


class KVCache:
    """Static per-layer KV cache [layers, B, kv_heads, max_len, head_dim].

    The prefill writes rows [0, S) (`fill`, python-int offset `len`); decode steps append through the single-token
    attention kernel at the device-side index `len_idx`, so a step captured in a CUDA graph advances without the host."""

    def __init__(self, a: LLMArch, B: int, max_len: int, device):
        if a.head_dim not in (64, 128) or (a.num_attention_heads // a.num_key_value_heads) not in ops.DECODE_ATTN_GROUPS:
            raise NotImplementedError(f"decode attention: head_dim {a.head_dim} / GQA group "
                                      f"{a.num_attention_heads // a.num_key_value_heads} has no kernel (no library fallback)")
        shape = (a.num_hidden_layers, B, a.num_key_value_heads, max_len, a.head_dim)
        self.k = torch.zeros(shape, device=device, dtype=torch.bfloat16)
        self.v = torch.zeros(shape, device=device, dtype=torch.bfloat16)
        self.len = 0
        self.max_len = max_len
        self.graph_mode = False          # True: single-token steps (device-side write index)
        self.len_idx = torch.zeros(1, device=device, dtype=torch.int64)            # device copy of `len`
        # beam search (decode.beam_generate): the prefill runs once per utterance and lands in every `row_stride`-th cache row;
        # `beam` = (indirection table, prefill length on the device, K) resolves the cache row per key position in the
        # single-token attention kernel (no cache gather per step)
        self.row_stride = 1
        self.beam = None

    def fill(self, layer, k, v):
        S = k.shape[2]
        self.k[layer][:: self.row_stride, :, self.len: self.len + S] = k
        self.v[layer][:: self.row_stride, :, self.len: self.len + S] = v

    def advance(self, S):
        self.len += S

    def sync_device_state(self):
        """After the eager prefill: publish `len` to the device-side state used by the decode step."""
        self.len_idx.fill_(self.len)


class LlamaMLP(nn.Module):
    def __init__(self, a: LLMArch, device):
        super().__init__()
        std = 0.02
        I, H = a.intermediate_size, a.hidden_size
        self.gate_up_weight = (torch.randn(2 * I, H, device=device) * std).to(torch.bfloat16)
        self.gate_proj = _LinearView(self.gate_up_weight[:I])
        self.up_proj = _LinearView(self.gate_up_weight[I:])
        self.down_proj = _LinearView((torch.randn(H, I, device=device) * std).to(torch.bfloat16))
        self._wt = None
        self._il = None

    def transposed(self):
        if self._wt is None:
            self._wt = (self.gate_up_weight.t().contiguous(), self.down_proj.weight.data.t().contiguous())
            self._il = None
        return self._wt

    def interleaved(self):
        """gate|up weight with 64-row gate / up blocks interleaved (+ its transpose) for the fused SwiGLU epilogue of the
        training forward; derived copies like `_wt`, rebuilt after a state-dict load."""
        self.transposed()
        if self._il is None:
            w = ag.interleave_gate_up(self.gate_up_weight)
            self._il = (w, w.t().contiguous())
        return self._il

    def forward(self, h, residual, norm=None):
        """residual + down(silu(gate(h)) * up(h)).  norm=(weight, eps) (decode step only): also returns the RMSNorm of the
        result, computed in the down_proj launch."""
        wt_gu, wt_d = self.transposed()
        I = self.gate_up_weight.shape[0] // 2
        if ag.skinny_ok(h, residual) and I % 64 == 0:
            # decode step: weight-streaming kernel, SwiGLU in its epilogue (only the activation is written)
            w_il, _ = self.interleaved()
            act = torch.empty((h.shape[0], I), device=h.device, dtype=torch.bfloat16)
            ops.gemm(h, w_il, act="swiglu64", out2=act, skinny=True)
            return ops.gemm(act, self.down_proj.weight.data, residual=residual, skinny=True, norm=norm)
        assert norm is None, "the fused norm belongs to the decode step"
        if h.requires_grad and I % 256 == 0 and ag.gate_up_swiglu_supported(h.shape[0], 2 * I) \
                and ag.pair_kernel_shape(h.shape[0], I):
            # SwiGLU in the gate_up GEMM's epilogue, its backward in the epilogue of down_proj's dgrad GEMM
            w_il, wt_il = self.interleaved()
            return ag.MlpSwigluFn.apply(h, w_il, wt_il, self.down_proj.weight.data, wt_d, residual)
        if h.requires_grad and I % 64 == 0 and ag.gate_up_swiglu_supported(h.shape[0], 2 * I):
            w_il, wt_il = self.interleaved()
            act = ag.GateUpSwigluFn.apply(h, w_il, wt_il)          # SwiGLU in the gate_up GEMM's epilogue
        elif not h.requires_grad and I % 64 == 0 and ag.gate_up_swiglu_supported(h.shape[0], 2 * I):
            # prefill / evaluation: SwiGLU in the gate_up GEMM's epilogue, the gate|up tile itself is never written
            w_il, _ = self.interleaved()
            act = torch.empty((h.shape[0], I), device=h.device, dtype=torch.bfloat16)
            ops.gemm(h, w_il, act="swiglu64", out2=act, block_n=256)
        else:
            gu = ag.frozen_linear(h, self.gate_up_weight, wt_gu, block_n=256)
            act = ag.swiglu(gu)
        return ag.frozen_linear(act, self.down_proj.weight.data, wt_d, residual=residual, block_n=256)


_FUSED_STEP_NORM = bool(os.environ.get("OMNI_DECODE_FUSED_NORM"))


class _NormWeight(nn.Module):
    def __init__(self, H, eps, device):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(H, device=device, dtype=torch.bfloat16), requires_grad=False)
        self.variance_epsilon = eps

    def forward(self, x):
        return ag.rmsnorm(x, self.weight.data, self.variance_epsilon)

    def with_residual(self, x):
        """(residual branch, normed) -- the backward adds the residual-path gradient inside the norm-backward kernel."""
        return ag.rmsnorm_residual(x, self.weight.data, self.variance_epsilon)


class LlamaDecoderLayer_lora(nn.Module):
    attention_cls = LlamaSdpaAttention_lora

    def __init__(self, config: LLMArch, layer_idx, lora_config, flat=None, device="cuda"):
        super().__init__()
        self.lora_config = lora_config
        self.self_attn = self.attention_cls(config=config, layer_idx=layer_idx, lora_config=lora_config, flat=flat,
                                            device=device)
        self.mlp = LlamaMLP(config, device)
        self.input_layernorm = _NormWeight(config.hidden_size, config.rms_norm_eps, device)
        self.post_attention_layernorm = _NormWeight(config.hidden_size, config.rms_norm_eps, device)

    def forward_step(self, x, h, rows, cos_t, sin_t, kv_cache, next_norm):
        """Decode step: x = residual stream, h = input_layernorm(x) (computed by the previous launch).  Both RMSNorms that
        follow this layer's o_proj / down_proj run inside those GEMM launches; returns (x_out, next_norm(x_out))."""
        post = (self.post_attention_layernorm.weight.data, self.post_attention_layernorm.variance_epsilon)
        x, h = self.self_attn(h, rows, cos_t, sin_t, kv_cache, residual=x, norm=post)
        return self.mlp(h, residual=x, norm=next_norm)

    def forward(self, x, rows, cos_t, sin_t, kv_cache=None):
        res, h = self.input_layernorm.with_residual(x)
        x = self.self_attn(h, rows, cos_t, sin_t, kv_cache, residual=res)     # x + attn  (residual in the epilogue)
        res, h = self.post_attention_layernorm.with_residual(x)
        return self.mlp(h, residual=res)                                     # x + mlp


class _Embedding(nn.Module):
    def __init__(self, V, H, device):
        super().__init__()
        self.weight = nn.Parameter((torch.randn(V, H, device=device) * 0.02).to(torch.bfloat16), requires_grad=False)

    def forward(self, ids):
        shape = ids.shape
        out = ops.gather_rows(self.weight.data, ids.reshape(-1).contiguous())
        return out.view(*shape, -1)


class LlamaModel_lora(nn.Module):
    layer_cls = LlamaDecoderLayer_lora

    def __init__(self, config: LLMArch, lora_config, flat=None, device="cuda"):
        super().__init__()
        self.config, self.lora_config = config, lora_config
        self.embed_tokens = _Embedding(config.vocab_size, config.hidden_size, device)
        self.layers = nn.ModuleList([self.layer_cls(config, i, lora_config, flat, device)
                                     for i in range(config.num_hidden_layers)])
        self.norm = _NormWeight(config.hidden_size, config.rms_norm_eps, device)
        self._rope = None

    def rope(self, need: int):
        if self._rope is None or self._rope[0].shape[0] < need:
            n = max(2048, 1 << (need - 1).bit_length())
            self._rope = rope_tables(self.config, n, self.embed_tokens.weight.device)
        return self._rope

    def forward_packed(self, x, rows: PackedRows, kv_cache=None):
        cos_t, sin_t = self.rope(rows.max_pos)
        I = self.layers[0].mlp.gate_up_weight.shape[0] // 2
        if (_FUSED_STEP_NORM and kv_cache is not None and kv_cache.graph_mode and ag.skinny_ok(x) and I % 64 == 0
                and all(S == 1 for (_, _, S, _) in rows.segments) and self.layers[0].self_attn.plan.ext_fwd_step is not None):
            # decode step with every RMSNorm but the first in the epilogue of the GEMM that produces its input (the CTA that
            # completes a token slice normalises it; bit-identical).  OFF by default: measured SLOWER at B = 64 (Llama-1B
            # 1.50 ms per step against 1.36, Qwen2.5-3B 4.24 against 3.85) -- the fence + arrival atomic of all 128 CTAs and
            # the serial two-pass tail of the last one cost more than the 2.7 us norm kernel they remove under programmatic
            # dependent launch.  OMNI_DECODE_FUSED_NORM=1 turns it on.
            h = self.layers[0].input_layernorm(x)
            for i, layer in enumerate(self.layers):
                nxt = self.layers[i + 1].input_layernorm if i + 1 < len(self.layers) else self.norm
                x, h = layer.forward_step(x, h, rows, cos_t, sin_t, kv_cache, (nxt.weight.data, nxt.variance_epsilon))
            return h
        for layer in self.layers:
            x = layer(x, rows, cos_t, sin_t, kv_cache)
        return self.norm(x)


@dataclass
class CausalLMOutputWithPast:
    loss: Optional[torch.Tensor] = None
    logits: Optional[torch.Tensor] = None
    past_key_values: Optional[object] = None
    hidden_states: Optional[torch.Tensor] = None


class LlamaForCausalLM_lora(nn.Module):
    model_cls = LlamaModel_lora

    def __init__(self, config: LLMArch, lora_config, device="cuda", flat: Optional[FlatParams] = None):
        super().__init__()
        self.config, self.lora_config = config, lora_config
        self.device_ = torch.device(device)
        if self.device_.type != "cuda":
            raise RuntimeError("omni_avsr_b200 is CUDA-only (sm_100a); construct the model on a CUDA device")
        if flat is None:
            flat = FlatParams(device, self.lora_param_count(config, lora_config) + 1024)
        self.flat = flat
        self.model = self.model_cls(config, lora_config, flat, device)
        self.vocab_size = config.vocab_size
        if config.tie_word_embeddings:
            self.lm_head = _LinearView(self.model.embed_tokens.weight)
        else:
            self.lm_head = _LinearView((torch.randn(config.vocab_size, config.hidden_size, device=device) * 0.02)
                                       .to(torch.bfloat16))
        self._head_t = None

    @classmethod
    def from_pretrained(cls, name: str, lora_config, device="cuda", **overrides):
        """No hub/network in this environment: builds the named architecture with random-init weights.
        Real checkpoints are loaded afterwards with load_state_dict (reference key names)."""
        return cls(arch_from_name(name, **overrides), lora_config, device=device)

    @staticmethod
    def lora_param_count(a: LLMArch, lc) -> int:
        r = round(a.hidden_size / lc.RANK)
        rp = (r + 63) // 64 * 64
        ns = (3 if lc.IS_TASK_SPECIFIC else 1) + (1 if (lc.IS_TASK_SPECIFIC and lc.SHARED_LORA) else 0)
        per = 2 * ns * rp * a.hidden_size + ns * (a.q_dim + a.kv_dim) * rp + 16
        return per * a.num_hidden_layers

    # ---- reference API ----------------------------------------------------------------------------
    def resize_token_embeddings(self, n: int):
        """modeling_OmniAVSR.py:214: grows (or shrinks) the embedding matrix; new rows ~ N(0, 0.02)."""
        old = self.model.embed_tokens.weight.data
        if n == old.shape[0]:
            return self.model.embed_tokens
        new = (torch.randn(n, old.shape[1], device=old.device) * 0.02).to(torch.bfloat16)
        k = min(n, old.shape[0])
        new[:k] = old[:k]
        self.model.embed_tokens.weight = nn.Parameter(new, requires_grad=False)
        if self.config.tie_word_embeddings:
            self.lm_head.weight = self.model.embed_tokens.weight
        else:
            oh = self.lm_head.weight.data
            nh = (torch.randn(n, oh.shape[1], device=oh.device) * 0.02).to(torch.bfloat16)
            nh[:k] = oh[:k]
            self.lm_head.weight = nn.Parameter(nh, requires_grad=False)
        self.config.vocab_size = n
        self.vocab_size = n
        self._head_t = None
        return self.model.embed_tokens

    def head_transposed(self):
        """lm_head weight transposed [H, Vp] (Vp = V rounded up to 8 so the TMA row stride is 16-byte aligned)."""
        W = self.lm_head.weight.data
        # `.data` is a fresh tensor object every time: compare storage + version (load_state_dict copies in place), not identity
        key = (W.data_ptr(), tuple(W.shape), self.lm_head.weight._version)
        if self._head_t is None or self._head_t[1] != key:
            V, H = W.shape
            Vp = (V + 7) // 8 * 8
            buf = torch.zeros((H, Vp), device=W.device, dtype=torch.bfloat16)
            buf[:, :V] = W.t()
            self._head_t = (buf[:, :V], key)      # frozen weight: rebuilt only after load_state_dict / resize (they reset it)
        return self._head_t[0]

    def logits_rows(self, hrows):
        """bf16 logits [R, V] (row stride padded to 8) for the given hidden rows."""
        V = self.config.vocab_size
        Vp = (V + 7) // 8 * 8
        buf = torch.empty((hrows.shape[0], Vp), device=hrows.device, dtype=torch.bfloat16)
        return ops.gemm(hrows.contiguous(), self.lm_head.weight.data, out=buf[:, :V], block_n=256)

    def _task_of(self, modality) -> int:
        if self.lora_config.IS_TASK_SPECIFIC:
            # same failure mode as the reference's ModuleDict lookup (Llama_LoRA.py:250)
            return TASK_ID[modality]
        return 0

    def forward(self, input_ids=None, attention_mask=None, position_ids=None, past_key_values=None, inputs_embeds=None,
                labels=None, use_cache=None, output_attentions=None, output_hidden_states=None, return_dict=None,
                cache_position=None, modality=None):
        """Reference-shaped call (:328-398): one task, dense [B, S, H] embeddings."""
        if (input_ids is None) == (inputs_embeds is None):
            raise ValueError("You cannot specify both input_ids and inputs_embeds at the same time, and must specify either one")
        if inputs_embeds is None:
            inputs_embeds = self.model.embed_tokens(input_ids)
        ops.require_cuda(inputs_embeds)
        B, S, H = inputs_embeds.shape
        task = self._task_of(modality)
        past = past_key_values.len if past_key_values is not None else 0
        rows = PackedRows.get([(task, B, S)], inputs_embeds.device, pos_offset=past)
        x = pack_segments([inputs_embeds], rows)
        hid = self.model.forward_packed(x, rows, past_key_values)
        if past_key_values is not None:
            past_key_values.advance(S)
        hid = hid[: B * S]
        loss, logits = None, None
        if labels is not None:
            loss = self.loss_from_hidden(hid, [(B, S, 0)], [labels], [1.0])[0]
        else:
            V = self.config.vocab_size
            logits = self.logits_rows(hid).float().view(B, S, V)            # :372-373 (fp32 logits)
        return CausalLMOutputWithPast(loss=loss, logits=logits, past_key_values=past_key_values,
                                      hidden_states=hid.view(B, S, H))

    def loss_from_hidden(self, hid, segs, labels_list, weights):
        """Shifted CE per segment (:376-386), evaluated on the label rows only: ignored positions contribute nothing
        to CrossEntropyLoss's mean, so the lm_head GEMM skips them.  segs = [(B, S, row_offset)]; returns the
        weighted mean loss of each segment."""
        W = self.lm_head.weight.data
        WT = self.head_transposed()
        idxs, tgts, scales = [], [], []
        for (B, S, off), lab, w in zip(segs, labels_list, weights):
            tgt = lab[:, 1:]
            mask = tgt != IGNORE_INDEX
            b_idx, s_idx = mask.nonzero(as_tuple=True)
            idxs.append(off + b_idx * S + s_idx)
            tgts.append(tgt[mask].contiguous())
            scales.append((w / mask.sum().clamp(min=1).float()).expand(b_idx.numel()).contiguous())
        hrows = GatherRowsFn.apply(hid, torch.cat(idxs).contiguous())      # one gather / one scatter for all tasks
        losses = ag.LmHeadCEFn.apply(hrows, W, WT, torch.cat(tgts).contiguous(), torch.cat(scales).contiguous(),
                                     [int(t.numel()) for t in tgts])          # one logits GEMM + CE for all tasks
        return list(losses.unbind(0))

    @torch.no_grad()
    def generate(self, inputs_embeds=None, max_new_tokens=32, num_beams=1, eos_token_id=None, bos_token_id=None,
                 pad_token_id=None, modality=None, **kw):
        from .decode import beam_generate, greedy_generate
        if num_beams != 1:
            return beam_generate(self, inputs_embeds, max_new_tokens, num_beams, eos_token_id, pad_token_id, modality)
        return greedy_generate(self, inputs_embeds, max_new_tokens, eos_token_id, pad_token_id, modality,
                               trim=kw.get("trim", True))


class GatherRowsFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, idx):
        ctx.save_for_backward(idx)
        ctx.shape = x.shape
        return ops.gather_rows(x, idx.contiguous())

    @staticmethod
    def backward(ctx, d):
        (idx,) = ctx.saved_tensors
        out = torch.zeros(ctx.shape, device=d.device, dtype=d.dtype)
        ops.scatter_rows(d, idx.contiguous(), out)
        return out, None


def pack_segments(embeds: Sequence[torch.Tensor], rows: PackedRows) -> torch.Tensor:
    """Copies dense [B, S, H] blocks into the 128-row-aligned packed buffer (pad rows zero)."""
    H = embeds[0].shape[-1]
    if len(embeds) == 1 and rows.valid_rows == rows.M:
        return embeds[0].reshape(rows.M, H)
    x = torch.zeros((rows.M, H), device=embeds[0].device, dtype=torch.bfloat16)
    parts, cur = [], 0
    for e, (task, B, S, off) in zip(embeds, rows.segments):
        if off > cur:
            parts.append(x[cur:off])
        parts.append(e.reshape(B * S, H))
        cur = off + B * S
    if cur < rows.M:
        parts.append(x[cur:])
    return torch.cat(parts, dim=0)
