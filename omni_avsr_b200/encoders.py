"""Frozen encoders of the Omni-AVSR path on the sm_100a kernels.

* `LogMel` + `WhisperEncoder`  -- replaces WhisperFeatureExtractor (host CPU in the reference, with a D2H/H2D
  round trip every step: Omni_AVSR/modeling_OmniAVSR.py:531-534) and transformers' WhisperEncoder
  (state-dict names kept: conv1, conv2, embed_positions, layers.{i}.self_attn.{q,k,v,out}_proj, ..., layer_norm).
* `AVHubertVideoEncoder` -- AVHubertModel.extract_finetune(source={'video': v, 'audio': None}) of
  av_hubert/avhubert/hubert.py:695-755 (ResEncoder resnet.py:131-169, SubModel hubert.py:318-333,
  TransformerEncoder wav2vec2.py:818-905, layer :977-1006, MultiheadAttention.forward_lora
  multihead_attention.py:485-662), fairseq state-dict names kept, LoRA r = round(1024/16), scale 2
  (modeling_OmniAVSR.py:127-142).

Linear layers, LayerNorm, GELU and the LoRA-fused q|k|v projection run on our kernels (tcgen05 GEMM with bias /
GELU / residual epilogues), the attention core is the tcgen05 flash kernel (forward and backward), the log-mel front end
is our DFT kernel, the Whisper conv stem and the AV-HuBERT Conv3d front-end run on the GEMM.
The ResNet-18 trunk runs on the same GEMM (3x3 stride-1 convolutions as overlapping-row views of ring-padded channels-last
frames, csrc/resnet_trunk.cu), so does the grouped positional convolution: no cuDNN / library convolution on the path.
Encoders run in eval mode (no dropout / layerdrop, BatchNorm running statistics): SURVEY §5.8 / §7.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch
from torch import nn

from . import autograd_ops as ag
from . import ops
from .Llama_LoRA import FlatParams, LoraPlan, PackedSdpaFn, _LinearView

SAMPLE_RATE, N_FFT, HOP, N_MELS, N_SAMPLES = 16000, 400, 160, 80, 480000


# ------------------------------------------------------------------------------------------------
# log-mel (on device)
# ------------------------------------------------------------------------------------------------
def _slaney_mel_filters(n_freq=201, n_mels=80, sr=16000, fmin=0.0, fmax=8000.0) -> torch.Tensor:
    def hz2mel(f):
        f = np.asarray(f, dtype=np.float64)
        return np.where(f >= 1000.0, 15.0 + np.log(np.maximum(f, 1e-10) / 1000.0) * (27.0 / np.log(6.4)), 3.0 * f / 200.0)

    def mel2hz(m):
        m = np.asarray(m, dtype=np.float64)
        return np.where(m >= 15.0, 1000.0 * np.exp((np.log(6.4) / 27.0) * (m - 15.0)), 200.0 * m / 3.0)

    fft_freqs = np.linspace(0, sr // 2, n_freq)
    pts = mel2hz(np.linspace(hz2mel(fmin), hz2mel(fmax), n_mels + 2))
    fdiff = np.diff(pts)
    slopes = pts[None, :] - fft_freqs[:, None]
    fb = np.maximum(0, np.minimum(-slopes[:, :-2] / fdiff[:-1], slopes[:, 2:] / fdiff[1:]))
    fb = fb * (2.0 / (pts[2: n_mels + 2] - pts[:n_mels]))[None, :]
    return torch.from_numpy(fb).to(torch.float32)


class LogMel(nn.Module):
    """[B, T] (bf16/fp32 waveform) -> [B, 80, 3000] bf16 Whisper input features, entirely on the GPU."""

    def __init__(self, device="cuda"):
        super().__init__()
        self.register_buffer("filters", _slaney_mel_filters().to(device).t().contiguous(), persistent=False)

    @torch.no_grad()
    def forward(self, audio: torch.Tensor) -> torch.Tensor:
        return ops.logmel(audio if audio.stride(-1) == 1 else audio.contiguous(), self.filters)


# ------------------------------------------------------------------------------------------------
# Whisper encoder (frozen, inference only)
# ------------------------------------------------------------------------------------------------
@dataclass
class WhisperArch:
    d_model: int = 1024
    encoder_layers: int = 24
    encoder_attention_heads: int = 16
    encoder_ffn_dim: int = 4096
    num_mel_bins: int = 80
    max_source_positions: int = 1500

    @property
    def hidden_size(self):
        return self.d_model


WHISPER_ARCHS = {
    "openai/whisper-medium.en": WhisperArch(), "openai/whisper-medium": WhisperArch(),
    "openai/whisper-small.en": WhisperArch(768, 12, 12, 3072), "openai/whisper-small": WhisperArch(768, 12, 12, 3072),
    "openai/whisper-base.en": WhisperArch(512, 6, 8, 2048), "openai/whisper-tiny.en": WhisperArch(384, 4, 6, 1536),
    "openai/whisper-large-v3": WhisperArch(1280, 32, 20, 5120, 128),
}


def _w(out_f, in_f, device, std=0.02):
    return (torch.randn(out_f, in_f, device=device) * std).to(torch.bfloat16)


def _b(n, device, std=0.0):
    return (torch.randn(n, device=device) * std).to(torch.bfloat16) if std else torch.zeros(n, device=device, dtype=torch.bfloat16)


class _LN(nn.Module):
    def __init__(self, d, device, eps=1e-5):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(d, device=device, dtype=torch.bfloat16), requires_grad=False)
        self.bias = nn.Parameter(torch.zeros(d, device=device, dtype=torch.bfloat16), requires_grad=False)
        self.eps = eps

    def forward(self, x):
        return ag.layernorm(x, self.weight.data, self.bias.data, self.eps)

    def with_residual(self, x):
        return ag.layernorm_residual(x, self.weight.data, self.bias.data, self.eps)


class _PackedSelfAttention(nn.Module):
    """q|k|v packed into one [3D, D] weight (+[3D] bias); reference-named views q_proj/k_proj/v_proj/out_proj."""

    def __init__(self, d, heads, device, k_bias=True):
        super().__init__()
        self.d, self.h, self.hd = d, heads, d // heads
        self.qkv_weight = _w(3 * d, d, device)
        self.qkv_bias = _b(3 * d, device)
        self.q_proj = _LinearView(self.qkv_weight[:d], self.qkv_bias[:d])
        self.k_proj = _LinearView(self.qkv_weight[d:2 * d], self.qkv_bias[d:2 * d] if k_bias else None)
        self.v_proj = _LinearView(self.qkv_weight[2 * d:], self.qkv_bias[2 * d:])
        self.out_proj = _LinearView(_w(d, d, device), _b(d, device))
        self._wt = None

    def transposed(self):
        if self._wt is None:
            self._wt = (self.qkv_weight.t().contiguous(), self.out_proj.weight.data.t().contiguous())
        return self._wt

    def sdpa(self, qkv, B, T):
        # non-causal attention over each clip: tcgen05 flash kernels (csrc/attention.cu, csrc/attention_bwd.cu)
        return PackedSdpaFn.apply(qkv, [(0, B, T, 0)], self.h, self.h, self.hd, False)


class WhisperEncoderLayer(nn.Module):
    def __init__(self, a: WhisperArch, device):
        super().__init__()
        d = a.d_model
        self.self_attn = _PackedSelfAttention(d, a.encoder_attention_heads, device, k_bias=False)
        self.self_attn_layer_norm = _LN(d, device)
        self.fc1 = _LinearView(_w(a.encoder_ffn_dim, d, device), _b(a.encoder_ffn_dim, device))
        self.fc2 = _LinearView(_w(d, a.encoder_ffn_dim, device), _b(d, device))
        self.final_layer_norm = _LN(d, device)

    def forward(self, x, B, T):
        att = self.self_attn
        h = self.self_attn_layer_norm(x)
        qkv = ops.gemm(h, att.qkv_weight, bias=att.qkv_bias, block_n=256)
        o = att.sdpa(qkv, B, T)
        x = ops.gemm(o, att.out_proj.weight.data, bias=att.out_proj.bias.data, residual=x, block_n=256)
        h = self.final_layer_norm(x)
        f = ops.gemm(h, self.fc1.weight.data, bias=self.fc1.bias.data, act="gelu", block_n=256)
        return ops.gemm(f, self.fc2.weight.data, bias=self.fc2.bias.data, residual=x, block_n=256)


@dataclass
class _EncOut:
    last_hidden_state: torch.Tensor


class WhisperEncoder(nn.Module):
    def __init__(self, a: WhisperArch, device="cuda"):
        super().__init__()
        self.config = a
        d = a.d_model
        self.conv1 = nn.Conv1d(a.num_mel_bins, d, 3, padding=1, device=device, dtype=torch.bfloat16)
        self.conv2 = nn.Conv1d(d, d, 3, stride=2, padding=1, device=device, dtype=torch.bfloat16)
        half = d // 2
        inv = torch.exp(-(math.log(10000.0) / (half - 1)) * torch.arange(half))
        t = torch.arange(a.max_source_positions).view(-1, 1) * inv.view(1, -1)
        pos = torch.cat([t.sin(), t.cos()], dim=1)
        self.embed_positions = nn.Embedding(a.max_source_positions, d, device=device, dtype=torch.bfloat16)
        self.embed_positions.weight.data.copy_(pos)
        self.layers = nn.ModuleList([WhisperEncoderLayer(a, device) for _ in range(a.encoder_layers)])
        self.layer_norm = _LN(d, device)
        self.requires_grad_(False)

    def _conv_stem(self, feats):
        """gelu(conv1) -> gelu(conv2, stride 2) -> + positions, with both convolutions on the tcgen05 GEMM: the
        activations are kept channels-last with one zero row of padding on each side of every clip, so the im2col row of
        output t is simply `k` CONSECUTIVE rows starting at row t*stride -- an overlapping-row view (row stride <
        row length) that the TMA tensor map expresses directly; no im2col buffer, no cuDNN.
        Rows that straddle two clips are garbage and land exactly on the padding rows, which are re-zeroed."""
        B, C, L = feats.shape                       # [B, 80, 3000]
        d = self.config.d_model
        if getattr(self, "_wt", None) is None:
            w1 = self.conv1.weight.data.permute(0, 2, 1).reshape(d, 3 * C).contiguous()      # [o, k*C + c]
            w2 = self.conv2.weight.data.permute(0, 2, 1).reshape(d, 3 * d).contiguous()
            self._wt = {"w1": w1, "w2": w2, "pos": {}}
        w = self._wt
        Lp = L + 2
        T = L // 2
        p1 = torch.zeros((B, Lp, C), device=feats.device, dtype=torch.bfloat16)
        p1[:, 1: L + 1] = ops.transpose(feats.contiguous())     # [B, 80, 3000] -> time-major rows (tiled transpose kernel)
        p2 = torch.empty((B, Lp, d), device=feats.device, dtype=torch.bfloat16)
        m1 = B * Lp - 2
        a1 = torch.as_strided(p1, (m1, 3 * C), (C, 1))
        ops.gemm(a1, w["w1"], bias=self.conv1.bias.data, act="gelu", out=p2.view(B * Lp, d)[1: 1 + m1], block_n=256)
        p2[:, 0].zero_()
        p2[:, Lp - 1].zero_()
        Tp = Lp // 2                                # 1501 output slots per clip, the last one is garbage
        if B not in w["pos"]:
            pos = torch.zeros((B, Tp, d), device=feats.device, dtype=torch.bfloat16)
            pos[:, :T] = self.embed_positions.weight.data[:T]
            w["pos"] = {B: pos.view(B * Tp, d)}      # (one batch size cached at a time)
        m2 = B * Tp - 1
        a2 = torch.as_strided(p2, (m2, 3 * d), (2 * d, 1))
        y = torch.empty((B * Tp, d), device=feats.device, dtype=torch.bfloat16)
        ops.gemm(a2, w["w2"], bias=self.conv2.bias.data, act="gelu", residual=w["pos"][B][:m2], out=y[:m2], block_n=256)
        # drop the one garbage slot per clip: row gather with a cached index (vectorised, HBM speed) instead of ATen's strided copy
        if ("rows", B) not in w["pos"]:
            w["pos"][("rows", B)] = (torch.arange(B, device=feats.device)[:, None] * Tp +
                                     torch.arange(T, device=feats.device)[None, :]).reshape(-1).contiguous()
        x = ops.gather_rows(y, w["pos"][("rows", B)])
        return x, B, T

    @classmethod
    def from_pretrained(cls, name: str, device="cuda"):
        if name not in WHISPER_ARCHS:
            raise KeyError(f"unknown audio encoder {name!r}")
        return cls(WHISPER_ARCHS[name], device)

    @torch.no_grad()
    def forward(self, input_features: torch.Tensor) -> _EncOut:
        ops.require_cuda(input_features)
        x, B, T = self._conv_stem(input_features.to(torch.bfloat16))
        d = x.shape[1]
        for layer in self.layers:
            x = layer(x, B, T)
        x = self.layer_norm(x)
        return _EncOut(x.view(B, T, d))


# ------------------------------------------------------------------------------------------------
# AV-HuBERT (video-only), LoRA on q/v of every block
# ------------------------------------------------------------------------------------------------
@dataclass
class AVHubertArch:
    encoder_embed_dim: int = 1024
    encoder_ffn_embed_dim: int = 4096
    encoder_layers: int = 24
    encoder_attention_heads: int = 16
    conv_pos: int = 128
    conv_pos_groups: int = 16
    resnet_widths: tuple = (64, 128, 256, 512)


AVHUBERT_ARCHS = {"large": AVHubertArch(), "base": AVHubertArch(768, 3072, 12, 12)}


class _BasicBlock(nn.Module):
    def __init__(self, inp, planes, stride, downsample):
        super().__init__()
        self.conv1 = nn.Conv2d(inp, planes, 3, stride, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.relu1 = nn.PReLU(planes)
        self.relu2 = nn.PReLU(planes)
        self.conv2 = nn.Conv2d(planes, planes, 3, 1, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.downsample = downsample

    def forward(self, x):
        out = self.relu1(self.bn1(self.conv1(x)))
        out = self.bn2(self.conv2(out))
        res = x if self.downsample is None else self.downsample(x)
        return self.relu2(out + res)


class _Trunk(nn.Module):
    def __init__(self, widths):
        super().__init__()
        inp = widths[0]
        for i, (w, s) in enumerate(zip(widths, (1, 2, 2, 2)), start=1):
            ds = None
            if s != 1 or inp != w:
                ds = nn.Sequential(nn.Conv2d(inp, w, 1, s, bias=False), nn.BatchNorm2d(w))
            setattr(self, f"layer{i}", nn.Sequential(_BasicBlock(inp, w, s, ds), _BasicBlock(w, w, 1, None)))
            inp = w
        self.avgpool = nn.AdaptiveAvgPool2d(1)
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                m.weight.data.normal_(0, math.sqrt(2.0 / (m.kernel_size[0] * m.kernel_size[1] * m.out_channels)))

    def forward(self, x):
        x = self.layer4(self.layer3(self.layer2(self.layer1(x))))
        return self.avgpool(x).flatten(1)


class _ResEncoder(nn.Module):
    """Conv3d front-end + ResNet-18 trunk (resnet.py:131-169), eval-mode BatchNorm folded into the filters; every
    convolution is a tcgen05 GEMM (no library convolution)."""

    def __init__(self, widths):
        super().__init__()
        self.frontend3D = nn.Sequential(
            nn.Conv3d(1, widths[0], (5, 7, 7), (1, 2, 2), (2, 3, 3), bias=False), nn.BatchNorm3d(widths[0]),
            nn.PReLU(widths[0]), nn.MaxPool3d((1, 3, 3), (1, 2, 2), (0, 1, 1)))
        self.trunk = _Trunk(widths)

    # ---- eval-mode fast path: BatchNorm folded into the convolutions, Conv3d(C_in=1) as a channels-last Conv2d ----
    @staticmethod
    def _fold(conv_w, bn):
        scale = bn.weight.float() / torch.sqrt(bn.running_var.float() + bn.eps)
        w = (conv_w.float() * scale.view(-1, *([1] * (conv_w.dim() - 1)))).to(torch.bfloat16)
        b = (bn.bias.float() - bn.running_mean.float() * scale).to(torch.bfloat16)
        return w, b

    def _frames_ok(self, li):
        """Layer li (and every later one) can run as table-driven frame-row GEMMs: 64-multiple input channels, output channels
        that tile the 256-wide CTA-pair N."""
        ok = lambda c: c % 64 == 0 and ((c < 256 and 256 % c == 0) or c % 256 == 0)
        widths = [getattr(self.trunk, f"layer{i}")[0].conv1.out_channels for i in range(1, 5)]
        return all(ok(widths[i - 1]) for i in range(li, 5)) and widths[max(li - 2, 0)] % 64 == 0 and widths[li - 1] >= 64

    def _frames_block(self, key, y, ent, blk):
        """One BasicBlock on the table-driven path (resnet.py:35-74); y: RingFrames (first block after the ring layers) or
        FrameRows."""
        _, w1, b1, s1, w2, b2, ds = ent
        ring = isinstance(y, ops.RingFrames)
        ck = (key, y.H, y.W, ring)
        specs = self._frame_specs.get(ck)
        if specs is None:
            c1 = ops.ConvFramesSpec(w1, y.H, y.W, s1, ring)
            c2 = ops.ConvFramesSpec(w2, c1.Hout, c1.Wout, 1, False)
            cd = None if ds is None else ops.ConvFramesSpec(ds[0], y.H, y.W, s1, ring)
            specs = self._frame_specs[ck] = (c1, c2, cd)
        c1, c2, cd = specs
        o = ops.conv_frames(y, c1, prelu=dict(slope=blk.relu1.weight.data, bias=b1))
        if cd is None:
            return ops.conv_frames(o, c2, prelu=dict(slope=blk.relu2.weight.data, bias=b2, residual=y))
        res = ops.conv_frames(y, cd)
        return ops.conv_frames(o, c2, prelu=dict(slope=blk.relu2.weight.data, bias=b2, residual=res, res_bias=ds[1]))

    def _prepare(self):
        self._frame_specs = {}
        conv3, bn3 = self.frontend3D[0], self.frontend3D[1]
        w3, b3 = self._fold(conv3.weight, bn3)                         # [C, 1, 5, 7, 7]
        C = w3.shape[0]
        wmat = torch.zeros((C, 5, 64), device=w3.device, dtype=torch.bfloat16)
        wmat[:, :, :49] = w3.reshape(C, 5, 49)                         # k = kt*64 + ky*7 + kx, 15 zero columns per kt
        f = {"front": (wmat.view(C, 320), b3)}
        for li in range(1, 5):
            for bi, blk in enumerate(getattr(self.trunk, f"layer{li}")):
                w1, b1 = self._fold(blk.conv1.weight, blk.bn1)
                w2, b2 = self._fold(blk.conv2.weight, blk.bn2)
                ds = None
                if blk.downsample is not None:
                    wd, bd = self._fold(blk.downsample[0].weight, blk.downsample[1])
                    ds = (wd.reshape(wd.shape[0], wd.shape[1]).contiguous(), bd)
                # filter matrices of the overlapping-row GEMM: g output pixels per GEMM row so that every layer presents a
                # 256-wide N to the CTA-pair kernel (64 channels: g = 4, 128: g = 2); the stride-2 convolutions read gathered
                # tap-major rows
                if self._frames_ok(li):
                    # every layer whose widths tile the 256-wide N (all four at the real widths): table-driven convolutions
                    # whose GEMM rows are whole frames
                    # (ops.conv_frames): no ring, no gather buffers, no flops on the zero padding.  Specs are built per
                    # input geometry at the first forward.
                    f[(li, bi)] = ("frames", w1, b1, blk.conv1.stride[0], w2, b2, None if ds is None else (wd, bd))
                    continue
                tap = lambda w: w.permute(0, 2, 3, 1).reshape(w.shape[0], -1).contiguous()
                grp = lambda w: max(1, min(4, 256 // w.shape[0])) if (w.shape[1] * 3) % 64 == 0 else 1
                s1 = blk.conv1.stride[0]
                g1, g2 = (grp(w1) if s1 == 1 else 1), grp(w2)
                m1 = ops.conv3x3_group_weights(w1, g1) if s1 == 1 else tap(w1)
                f[(li, bi)] = (m1, b1, s1, ops.conv3x3_group_weights(w2, g2), b2, ds, g1, g2)
        return f

    def forward(self, x):                       # [B, 1, T, 88, 88] -> [B*T, C]
        if self.training:
            raise RuntimeError("the AV-HuBERT front-end runs with eval semantics (BatchNorm running statistics folded into "
                               "the filters); AVHubertVideoEncoder.train() keeps it in eval mode")
        if getattr(self, "_wt", None) is None:
            self._wt = self._prepare()
        f = self._wt
        B, _, T, Hh, Ww = x.shape
        # Conv3d(1 -> C, (5,7,7), stride (1,2,2)) = time-major im2col of the 49 spatial taps + ONE tcgen05 GEMM whose five
        # K blocks (temporal taps) read five consecutive rows, against the BatchNorm-folded [C, 5*64] filter matrix; the
        # pooling kernel writes the ring-padded channels-last frames the trunk's convolution GEMMs read.
        w, b = f["front"]
        Hp, Wp = ((Hh + 6 - 7) // 2 + 1 + 2 - 3) // 2 + 1, ((Ww + 6 - 7) // 2 + 1 + 2 - 3) // 2 + 1
        key = (B * T, Hp, Wp, w.shape[0])
        if getattr(self, "_ring0", None) is None or self._ring0[0] != key:
            # the ring of this buffer is zeroed once and never written again (the pooling kernel fills the interior only)
            self._ring0 = (key, ops.RingFrames(B * T, Hp, Wp, w.shape[0], x.device, zero=True))
        if isinstance(f[(1, 0)][0], str):
            # layer 1 on the table-driven path too: the pooling kernel writes plain channels-last frames [B T, Hp Wp C]
            yp = ops.front3d_prelu_maxpool(x[:, 0].contiguous(), w, b, self.frontend3D[2].weight.data)
            y = ops.FrameRows.wrap(yp.permute(0, 2, 3, 1).reshape(B * T, Hp * Wp * w.shape[0]), Hp, Wp, w.shape[0])
        else:
            y = ops.front3d_prelu_maxpool(x[:, 0].contiguous(), w, b, self.frontend3D[2].weight.data, ring_out=self._ring0[1])
        for li in range(1, 5):
            for bi, blk in enumerate(getattr(self.trunk, f"layer{li}")):
                if isinstance(f[(li, bi)][0], str):          # "frames": table-driven path
                    y = self._frames_block((li, bi), y, f[(li, bi)], blk)
                    continue
                w1, b1, s1, w2, b2, ds, g1, g2 = f[(li, bi)]
                # the folded-BatchNorm shifts ride in the PReLU kernel (a conv bias would cost one more elementwise pass)
                # (... or in the epilogue of the convolution GEMM itself: bias, residual add, PReLU and ring re-zeroing)
                if s1 == 1:
                    o = ops.conv3x3s1_ring(y, w1, g1, prelu=dict(slope=blk.relu1.weight.data, bias=b1))
                else:
                    o = ops.conv_s2_ring(y, w1, 9)
                    ops.prelu_res_ring_(o, blk.relu1.weight.data, bias=b1)
                if ds is None:
                    y = ops.conv3x3s1_ring(o, w2, g2, prelu=dict(slope=blk.relu2.weight.data, bias=b2, residual=y))
                else:
                    res = ops.conv_s2_ring(y, ds[0], 1)
                    y = ops.conv3x3s1_ring(o, w2, g2, prelu=dict(slope=blk.relu2.weight.data, bias=b2, residual=res,
                                                                  res_bias=ds[1]))
        return ops.avgpool_frames(y) if isinstance(y, ops.FrameRows) else ops.avgpool_ring(y)


class _VideoFeatureExtractor(nn.Module):
    def __init__(self, a: AVHubertArch, device):
        super().__init__()
        self.resnet = _ResEncoder(a.resnet_widths).to(device=device, dtype=torch.bfloat16)
        self.proj = _LinearView(_w(a.encoder_embed_dim, a.resnet_widths[-1], device), _b(a.encoder_embed_dim, device))


class _Rows:
    """Minimal row-layout object for the LoRA kernels when there is a single adapter group."""

    def __init__(self, M):
        self.tile_group = None
        self.runs = [(0, 0, M)]
        self.pair_aligned = True


class AVHAttention_lora(_PackedSelfAttention):
    def __init__(self, a: AVHubertArch, device, flat: Optional[FlatParams], use_lora: bool, layer_idx: int):
        d = a.encoder_embed_dim
        super().__init__(d, a.encoder_attention_heads, device)
        self.use_lora = use_lora
        if use_lora:
            self.rank = 16                                   # modeling_OmniAVSR.py:131
            self.scaling_lora = 2                            # :132
            r = round(d / self.rank)
            self.plan = LoraPlan(d, d, d, d, r, self.scaling_lora, False, False, device)
            p = self.plan
            if flat is None:
                flat = FlatParams(device, p.down_rows * d + p.up_rows * p.rp + 64)
            self.lora_down = flat.alloc((p.down_rows, d), f"avh.{layer_idx}.lora_down")
            self.lora_up = flat.alloc((p.up_rows, p.rp), f"avh.{layer_idx}.lora_up")
            self.lora_down_Q = _LinearView(self.lora_down.data[:r])
            self.lora_down_V = _LinearView(self.lora_down.data[p.rp: p.rp + r])
            self.lora_up_Q = _LinearView(self.lora_up.data[:d, :r])
            self.lora_up_V = _LinearView(self.lora_up.data[d:, :r])
            bound = 1.0 / math.sqrt(r)
            with torch.no_grad():
                self.lora_up_Q.weight.uniform_(-bound, bound)    # kaiming_uniform_(a=sqrt(5)) (:141-142); down stays 0
                self.lora_up_V.weight.uniform_(-bound, bound)

    def state_dict(self, *args, destination=None, prefix="", keep_vars=False):
        sd = super().state_dict(*args, destination=destination, prefix=prefix, keep_vars=keep_vars)
        for k in (prefix + "lora_down", prefix + "lora_up"):
            sd.pop(k, None)
        return sd

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs):
        injected = []
        if self.use_lora:
            for k in ("lora_down", "lora_up"):      # packed tensors are filled through the aliased views
                if prefix + k not in state_dict:
                    state_dict[prefix + k] = getattr(self, k).data
                    injected.append(prefix + k)
        super()._load_from_state_dict(state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs)
        for k in injected:
            state_dict.pop(k, None)
        self._wt = None


class AVHLayer(nn.Module):
    def __init__(self, a: AVHubertArch, device, flat, use_lora, idx):
        super().__init__()
        d = a.encoder_embed_dim
        self.apply_lora = use_lora
        self.self_attn = AVHAttention_lora(a, device, flat, use_lora, idx)
        self.self_attn_layer_norm = _LN(d, device)
        self.fc1 = _LinearView(_w(a.encoder_ffn_embed_dim, d, device), _b(a.encoder_ffn_embed_dim, device))
        self.fc2 = _LinearView(_w(d, a.encoder_ffn_embed_dim, device), _b(d, device))
        self.final_layer_norm = _LN(d, device)
        self._wt = None

    def transposed(self):
        if self._wt is None:
            self._wt = (self.fc1.weight.data.t().contiguous(), self.fc2.weight.data.t().contiguous())
        return self._wt

    def forward(self, x, B, T, rows, first: bool):
        att = self.self_attn
        wt_qkv, wt_o = att.transposed()
        wt1, wt2 = self.transposed()
        res, h = self.self_attn_layer_norm.with_residual(x)
        if att.use_lora and (torch.is_grad_enabled() and att.lora_down.requires_grad):
            qkv = ag.LoraLinearFn.apply(h, att.qkv_weight, wt_qkv, att.qkv_bias, att.lora_down, att.lora_up, rows, att.plan)
        elif att.use_lora:
            Tm = ops.gemm(h, att.lora_down.data, n=att.plan.t_cols, alpha=att.plan.scaling, b_row_table=att.plan.brow_fwd,
                          block_n=64)
            qkv = ops.gemm(h, att.qkv_weight, bias=att.qkv_bias, ext=(Tm, att.lora_up.data, att.plan.ext_fwd),
                           block_n=att.plan.block_n, pair_aligned=True)
        else:
            qkv = ops.gemm(h, att.qkv_weight, bias=att.qkv_bias, block_n=256)
        o = att.sdpa(qkv, B, T)       # q * head_dim^-0.5 (:511) is SDPA's default scale (exact: power of two)
        x = ag.frozen_linear(o, att.out_proj.weight.data, wt_o, bias=att.out_proj.bias.data, residual=res, block_n=256)
        res, h = self.final_layer_norm.with_residual(x)
        if h.requires_grad and ag.pair_kernel_shape(h.shape[0], self.fc1.weight.shape[0]):
            # GELU in the fc1 GEMM's epilogue.  (Its backward as an epilogue of fc2's dgrad GEMM -- ag.FfnGeluFn,
            # OMNI_ACT_GELU_BWD -- is built and tested but NOT used here: with K = 1024 the main loop of a tile is 4096 clk and
            # the erf-based epilogue takes longer, 0.22 ms per launch against 0.09 + 0.09 ms unfused at B = 32.)
            f = ag.FrozenLinearGeluFn.apply(h, self.fc1.weight.data, wt1, self.fc1.bias.data)
        elif h.requires_grad:
            f = ag.frozen_linear(h, self.fc1.weight.data, wt1, bias=self.fc1.bias.data, block_n=256)
            f = ag.gelu(f)
        else:
            f = ops.gemm(h, self.fc1.weight.data, bias=self.fc1.bias.data, act="gelu", block_n=256)
        return ag.frozen_linear(f, self.fc2.weight.data, wt2, bias=self.fc2.bias.data, residual=res, block_n=256)


class _AVHTransformerEncoder(nn.Module):
    def __init__(self, a: AVHubertArch, device, flat, use_lora):
        super().__init__()
        d = a.encoder_embed_dim
        conv = nn.Conv1d(d, d, a.conv_pos, padding=a.conv_pos // 2, groups=a.conv_pos_groups)
        nn.init.normal_(conv.weight, 0, math.sqrt(4.0 / (a.conv_pos * d)))
        nn.init.constant_(conv.bias, 0)
        conv = nn.utils.weight_norm(conv, name="weight", dim=2)      # keys pos_conv.0.{bias,weight_g,weight_v}
        self.pos_conv = nn.Sequential(conv.to(device=device, dtype=torch.bfloat16))
        self.remove = 1 if a.conv_pos % 2 == 0 else 0
        self.layers = nn.ModuleList([AVHLayer(a, device, flat, use_lora, i) for i in range(a.encoder_layers)])
        self.layer_norm = _LN(d, device)

    def _pos_conv_weights(self):
        """Frozen positional convolution (wav2vec2.py:829-841: weight-normed grouped Conv1d + SamePad + GELU) as per-group GEMM
        operands: g * v / ||v|| materialised once, group G's filter as a K-major [gw, k * gw] matrix (K = (tap, in-channel))."""
        if getattr(self, "_wt", None) is None:
            conv = self.pos_conv[0]
            w = torch._weight_norm(conv.weight_v, conv.weight_g, 2)              # [d, gw, k]
            groups, k = conv.groups, conv.kernel_size[0]
            d, gw = w.shape[0], w.shape[1]
            if gw % 8 != 0:
                raise ValueError("positional-conv group width must be a multiple of 8 (TMA row stride)")
            wg = w.view(groups, gw, gw, k).permute(0, 1, 3, 2).reshape(groups, gw, k * gw).contiguous()
            self._wt = (wg, conv.bias.data.contiguous(), k, groups, gw)
        return self._wt

    def forward(self, x, B, T):                  # x [B*T, C]
        d = x.shape[1]
        with torch.no_grad():
            # x + gelu(pos_conv(x)) on the tcgen05 GEMM: the activations of one group, zero-padded per clip and stored
            # group-major, make the im2col row of output t simply the k CONSECUTIVE rows starting at row t (overlapping-row
            # TMA view, row stride gw < row length k * gw); bias + GELU run in the epilogue, which writes the group's gw
            # columns of the [B * Tp, d] result directly.  Rows t >= T of a clip are garbage and never read.
            wg, bias, k, groups, gw = self._pos_conv_weights()
            Tp = T + k
            xp = torch.zeros((B, Tp, d), device=x.device, dtype=torch.bfloat16)
            xp[:, k // 2: k // 2 + T] = x.view(B, T, d)
            xg = xp.view(B * Tp, groups, gw).permute(1, 0, 2).contiguous()          # [groups, B*Tp, gw]
            M = B * Tp - (k - 1)
            y = torch.empty((B * Tp, d), device=x.device, dtype=torch.bfloat16)
            for g in range(groups):
                a = torch.as_strided(xg[g], (M, k * gw), (gw, 1))
                ops.gemm(a, wg[g], bias=bias[g * gw: (g + 1) * gw], act="gelu", out=y[:M, g * gw: (g + 1) * gw], block_n=64)
            x = (x.view(B, T, d) + y.view(B, Tp, d)[:, :T]).reshape(B * T, d)
        rows = _Rows(B * T)
        for i, layer in enumerate(self.layers):
            x = layer(x, B, T, rows, i == 0)
        return self.layer_norm(x)


class AVHubertVideoEncoder(nn.Module):
    """`video_encoder` of AVSR_LLMs: .extract_finetune(source) -> (features [B, T, C], None, layer_outputs)."""

    def __init__(self, a: AVHubertArch = AVHubertArch(), device="cuda", flat: Optional[FlatParams] = None,
                 use_lora: bool = True):
        super().__init__()
        self.arch = a
        d = a.encoder_embed_dim
        self.encoder_embed_dim = d
        self.feature_extractor_video = _VideoFeatureExtractor(a, device)
        self.layer_norm = _LN(2 * d, device)
        self.post_extract_proj = _LinearView(_w(d, 2 * d, device), _b(d, device))
        self.encoder = _AVHTransformerEncoder(a, device, flat, use_lora)
        for p in self.parameters():
            p.requires_grad_(False)
        self.eval()

    def train(self, mode: bool = True):
        """The encoder always runs with eval semantics (no dropout/layerdrop, BatchNorm running statistics): the
        deterministic configuration of SURVEY §5.8; `.train()` on a parent module must not flip it."""
        return super().train(False)

    def lora_parameters(self):
        for layer in self.encoder.layers:
            if layer.self_attn.use_lora:
                yield layer.self_attn.lora_down
                yield layer.self_attn.lora_up

    @staticmethod
    def lora_param_count(a: AVHubertArch) -> int:
        d = a.encoder_embed_dim
        r = round(d / 16)
        rp = (r + 63) // 64 * 64
        return a.encoder_layers * (2 * rp * d + 2 * d * rp + 32)

    def extract_finetune(self, source, padding_mask=None, mask=False, ret_conv=False, output_layer=None):
        video = source["video"]
        if source.get("audio") is not None:
            raise NotImplementedError("the Omni-AVSR path feeds AV-HuBERT with video only (modeling_OmniAVSR.py:463)")
        ops.require_cuda(video)
        B, _, T = video.shape[:3]
        d = self.encoder_embed_dim
        fe = self.feature_extractor_video
        with torch.no_grad():
            f = fe.resnet(video.to(torch.bfloat16))                                   # [B*T, 512]
            f = ops.gemm(f.contiguous(), fe.proj.weight.data, bias=fe.proj.bias.data)  # SubModel.proj
            cat = torch.zeros((B * T, 2 * d), device=video.device, dtype=torch.bfloat16)
            cat[:, d:] = f                                                            # audio half = zeros (hubert.py:709)
            x = ops.layernorm_fwd(cat, self.layer_norm.weight.data, self.layer_norm.bias.data, self.layer_norm.eps)
            x = ops.gemm(x, self.post_extract_proj.weight.data, bias=self.post_extract_proj.bias.data)
        x = self.encoder(x, B, T)
        return x.view(B, T, d), None, []
