"""ORACLE (test infrastructure, not product code) -- CPU restatement of the two frozen encoders.

Audio: WhisperFeatureExtractor + WhisperEncoder of the un-vendored dependency transformers==4.43.1
(requirements.txt:7), called at Omni_AVSR/modeling_OmniAVSR.py:59-60,531-534; the published algorithm
 is restated here and pinned in tests/test_oracle_encoders.py against the installed transformers
WhisperFeatureExtractor / WhisperModel(config).encoder.

Video: AV-HuBERT video-only path, vendored in the reference:
  av_hubert/avhubert/resnet.py:35-74,77-129,131-169 (ResEncoder; pinned by importing that file in the test),
  av_hubert/avhubert/hubert.py:318-333 (SubModel), :695-755 (extract_finetune),
  av_hubert/fairseq/fairseq/models/wav2vec/wav2vec2.py:818-905 (TransformerEncoder), :916-1038 (layer),
  av_hubert/fairseq/fairseq/modules/multihead_attention.py:389-672 (forward_lora), modules/same_pad.py, modules/gelu.py,
  shapes from av_hubert/avhubert/conf/pretrain/large_vox_iter5.yaml:70-101.
Eval-mode semantics (dropout / layerdrop off, BatchNorm running statistics) -- the deterministic configuration the
parity runs use (SURVEY §7 "hard parts").

Parity status: pinned against transformers (log-mel, Whisper encoder), the reference's own resnet.py (imported by
path), the reference's own fairseq MultiheadAttention.forward_lora (tests/golden/make_reference_golden.py ->
tests/test_reference_golden.py) and -- end to end -- the reference's own AVHubertModel.extract_finetune: hubert.py,
wav2vec2.py (TransformerEncoder + TransformerSentenceEncoderLayer with apply_lora), resnet.py and multihead_attention.py
are loaded by path from /root/reference and executed by tests/golden/make_avhubert_golden.py (only fairseq's registry /
dataclass / task machinery is stubbed); tests/test_avhubert_golden.py checks AVHubertVideo below against that run
(identical key layout, outputs equal to 1e-5) and the CUDA path against it within bf16 tolerance.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
import torch
import torch.nn.functional as F
from torch import nn

# ------------------------------------------------------------------------------------------------
# log-mel front end
# ------------------------------------------------------------------------------------------------
SAMPLE_RATE, N_FFT, HOP, N_MELS, N_SAMPLES = 16000, 400, 160, 80, 480000


def _hz_to_mel_slaney(f):
    f = np.asarray(f, dtype=np.float64)
    min_log_hz, min_log_mel, logstep = 1000.0, 15.0, 27.0 / np.log(6.4)
    mels = 3.0 * f / 200.0
    log_region = f >= min_log_hz
    mels = np.where(log_region, min_log_mel + np.log(np.maximum(f, 1e-10) / min_log_hz) * logstep, mels)
    return mels


def _mel_to_hz_slaney(m):
    m = np.asarray(m, dtype=np.float64)
    min_log_hz, min_log_mel, logstep = 1000.0, 15.0, np.log(6.4) / 27.0
    f = 200.0 * m / 3.0
    log_region = m >= min_log_mel
    return np.where(log_region, min_log_hz * np.exp(logstep * (m - min_log_mel)), f)


def mel_filters(n_freq=1 + N_FFT // 2, n_mels=N_MELS, sr=SAMPLE_RATE, fmin=0.0, fmax=8000.0) -> np.ndarray:
    """transformers.audio_utils.mel_filter_bank(norm='slaney', mel_scale='slaney') -> [n_freq, n_mels] float64."""
    fft_freqs = np.linspace(0, sr // 2, n_freq)
    mel_pts = np.linspace(_hz_to_mel_slaney(fmin), _hz_to_mel_slaney(fmax), n_mels + 2)
    filter_freqs = _mel_to_hz_slaney(mel_pts)
    fdiff = np.diff(filter_freqs)
    slopes = np.expand_dims(filter_freqs, 0) - np.expand_dims(fft_freqs, 1)
    down = -slopes[:, :-2] / fdiff[:-1]
    up = slopes[:, 2:] / fdiff[1:]
    fb = np.maximum(0, np.minimum(down, up))
    enorm = 2.0 / (filter_freqs[2: n_mels + 2] - filter_freqs[:n_mels])
    return fb * np.expand_dims(enorm, 0)


def log_mel(audio: torch.Tensor) -> torch.Tensor:
    """audio [B, T] float32 (any T) -> [B, 80, 3000] float32; WhisperFeatureExtractor._torch_extract_fbank_features."""
    B, T = audio.shape
    wav = torch.zeros(B, N_SAMPLES, dtype=torch.float32)
    n = min(T, N_SAMPLES)
    wav[:, :n] = audio[:, :n].float()
    window = torch.hann_window(N_FFT)
    stft = torch.stft(wav, N_FFT, HOP, window=window, return_complex=True)
    mag = stft[..., :-1].abs() ** 2
    fb = torch.from_numpy(mel_filters()).to(torch.float32)
    mel = fb.T @ mag
    log_spec = torch.clamp(mel, min=1e-10).log10()
    mx = log_spec.amax(dim=(1, 2), keepdim=True)
    log_spec = torch.maximum(log_spec, mx - 8.0)
    return (log_spec + 4.0) / 4.0


# ------------------------------------------------------------------------------------------------
# Whisper encoder
# ------------------------------------------------------------------------------------------------
@dataclass
class WhisperCfg:
    d_model: int = 1024
    layers: int = 24
    heads: int = 16
    ffn: int = 4096
    n_mels: int = 80
    max_source_positions: int = 1500


WHISPER = {"openai/whisper-medium.en": WhisperCfg(), "openai/whisper-medium": WhisperCfg(),
           "openai/whisper-small.en": WhisperCfg(768, 12, 12, 3072), "openai/whisper-small": WhisperCfg(768, 12, 12, 3072)}


def sinusoids(length, channels, max_timescale=10000.0):
    log_inc = math.log(max_timescale) / (channels // 2 - 1)
    inv = torch.exp(-log_inc * torch.arange(channels // 2))
    t = torch.arange(length).view(-1, 1) * inv.view(1, -1)
    return torch.cat([t.sin(), t.cos()], dim=1)


class WhisperAttention(nn.Module):
    def __init__(self, d, h):
        super().__init__()
        self.h, self.hd = h, d // h
        self.q_proj = nn.Linear(d, d)
        self.k_proj = nn.Linear(d, d, bias=False)
        self.v_proj = nn.Linear(d, d)
        self.out_proj = nn.Linear(d, d)

    def forward(self, x):
        B, T, D = x.shape
        q = self.q_proj(x).view(B, T, self.h, self.hd).transpose(1, 2)
        k = self.k_proj(x).view(B, T, self.h, self.hd).transpose(1, 2)
        v = self.v_proj(x).view(B, T, self.h, self.hd).transpose(1, 2)
        o = F.scaled_dot_product_attention(q, k, v)
        return self.out_proj(o.transpose(1, 2).reshape(B, T, D))


class WhisperLayer(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.self_attn = WhisperAttention(c.d_model, c.heads)
        self.self_attn_layer_norm = nn.LayerNorm(c.d_model)
        self.fc1 = nn.Linear(c.d_model, c.ffn)
        self.fc2 = nn.Linear(c.ffn, c.d_model)
        self.final_layer_norm = nn.LayerNorm(c.d_model)

    def forward(self, x):
        x = x + self.self_attn(self.self_attn_layer_norm(x))
        return x + self.fc2(F.gelu(self.fc1(self.final_layer_norm(x))))


class WhisperEncoder(nn.Module):
    def __init__(self, c: WhisperCfg):
        super().__init__()
        self.cfg = c
        self.conv1 = nn.Conv1d(c.n_mels, c.d_model, 3, padding=1)
        self.conv2 = nn.Conv1d(c.d_model, c.d_model, 3, stride=2, padding=1)
        self.embed_positions = nn.Embedding(c.max_source_positions, c.d_model)
        self.embed_positions.weight.data = sinusoids(c.max_source_positions, c.d_model)
        self.layers = nn.ModuleList([WhisperLayer(c) for _ in range(c.layers)])
        self.layer_norm = nn.LayerNorm(c.d_model)

    def forward(self, feats):
        x = F.gelu(self.conv1(feats))
        x = F.gelu(self.conv2(x))
        x = x.permute(0, 2, 1)
        x = x + self.embed_positions.weight
        for l in self.layers:
            x = l(x)
        return self.layer_norm(x)


# ------------------------------------------------------------------------------------------------
# AV-HuBERT (video-only)
# ------------------------------------------------------------------------------------------------
@dataclass
class AVHubertCfg:
    embed_dim: int = 1024
    ffn: int = 4096
    layers: int = 24
    heads: int = 16
    conv_pos: int = 128
    conv_pos_groups: int = 16
    lora_rank_factor: int = 16      # modeling_OmniAVSR.py:131 -> r = round(dim / 16)
    lora_scaling: float = 2.0       # :132
    resnet_out: int = 512


class BasicBlock(nn.Module):  # resnet.py:35-74 (prelu variant)
    def __init__(self, inp, planes, stride=1, downsample=None):
        super().__init__()
        self.conv1 = nn.Conv2d(inp, planes, 3, stride, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.relu1 = nn.PReLU(planes)
        self.relu2 = nn.PReLU(planes)
        self.conv2 = nn.Conv2d(planes, planes, 3, 1, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.downsample = downsample

    def forward(self, x):
        residual = x
        out = self.relu1(self.bn1(self.conv1(x)))
        out = self.bn2(self.conv2(out))
        if self.downsample is not None:
            residual = self.downsample(x)
        out = out + residual
        return self.relu2(out)


class ResNet(nn.Module):  # resnet.py:77-129
    def __init__(self, widths=(64, 128, 256, 512)):
        super().__init__()
        self.inplanes = widths[0]
        self.layer1 = self._make(widths[0], 2, 1)
        self.layer2 = self._make(widths[1], 2, 2)
        self.layer3 = self._make(widths[2], 2, 2)
        self.layer4 = self._make(widths[3], 2, 2)
        self.avgpool = nn.AdaptiveAvgPool2d(1)

    def _make(self, planes, blocks, stride):
        ds = None
        if stride != 1 or self.inplanes != planes:
            ds = nn.Sequential(nn.Conv2d(self.inplanes, planes, 1, stride, bias=False), nn.BatchNorm2d(planes))
        layers = [BasicBlock(self.inplanes, planes, stride, ds)]
        self.inplanes = planes
        for _ in range(1, blocks):
            layers.append(BasicBlock(planes, planes))
        return nn.Sequential(*layers)

    def forward(self, x):
        x = self.layer4(self.layer3(self.layer2(self.layer1(x))))
        return self.avgpool(x).view(x.size(0), -1)


class ResEncoder(nn.Module):  # resnet.py:131-169
    def __init__(self, widths=(64, 128, 256, 512)):
        super().__init__()
        self.frontend3D = nn.Sequential(
            nn.Conv3d(1, widths[0], (5, 7, 7), (1, 2, 2), (2, 3, 3), bias=False), nn.BatchNorm3d(widths[0]),
            nn.PReLU(widths[0]), nn.MaxPool3d((1, 3, 3), (1, 2, 2), (0, 1, 1)))
        self.trunk = ResNet(widths)

    def forward(self, x):
        B = x.shape[0]
        x = self.frontend3D(x)
        T = x.shape[2]
        x = x.transpose(1, 2).contiguous()
        x = x.reshape(B * T, *x.shape[2:])
        x = self.trunk(x)
        return x.view(B, T, -1).transpose(1, 2).contiguous()


class SubModel(nn.Module):  # hubert.py:318-333 with sub_encoder_layers = 0
    def __init__(self, cfg: AVHubertCfg, widths):
        super().__init__()
        self.resnet = ResEncoder(widths)
        self.proj = nn.Linear(widths[-1], cfg.embed_dim)

    def forward(self, x):
        x = self.resnet(x)
        return self.proj(x.transpose(1, 2)).transpose(1, 2)


class MHA_lora(nn.Module):  # multihead_attention.py forward_lora :485-494, :511, :619-662
    def __init__(self, cfg: AVHubertCfg):
        super().__init__()
        d = cfg.embed_dim
        self.h, self.hd = cfg.heads, d // cfg.heads
        self.scaling = self.hd ** -0.5
        self.q_proj, self.k_proj, self.v_proj, self.out_proj = (nn.Linear(d, d) for _ in range(4))
        r = round(d / cfg.lora_rank_factor)
        self.scaling_lora = cfg.lora_scaling
        self.lora_down_Q = nn.Linear(d, r, bias=False)
        self.lora_up_Q = nn.Linear(r, d, bias=False)
        self.lora_down_V = nn.Linear(d, r, bias=False)
        self.lora_up_V = nn.Linear(r, d, bias=False)
        nn.init.zeros_(self.lora_down_Q.weight)
        nn.init.zeros_(self.lora_down_V.weight)

    def forward(self, x):  # x [T, B, C]
        T, B, C = x.shape
        q, k, v = self.q_proj(x), self.k_proj(x), self.v_proj(x)
        q = q + self.lora_up_Q(self.lora_down_Q(x)) * self.scaling_lora
        v = v + self.lora_up_V(self.lora_down_V(x)) * self.scaling_lora
        q = q * self.scaling
        q = q.contiguous().view(T, B * self.h, self.hd).transpose(0, 1)
        k = k.contiguous().view(T, B * self.h, self.hd).transpose(0, 1)
        v = v.contiguous().view(T, B * self.h, self.hd).transpose(0, 1)
        w = torch.bmm(q, k.transpose(1, 2))
        w = F.softmax(w.float(), dim=-1).type_as(w)          # utils.softmax -> fp32 softmax, cast back
        a = torch.bmm(w, v)
        a = a.transpose(0, 1).contiguous().view(T, B, C)
        return self.out_proj(a)


class AVHLayer(nn.Module):  # wav2vec2.py:977-1006 (layer_norm_first, apply_lora)
    def __init__(self, cfg):
        super().__init__()
        self.self_attn = MHA_lora(cfg)
        self.self_attn_layer_norm = nn.LayerNorm(cfg.embed_dim)
        self.fc1 = nn.Linear(cfg.embed_dim, cfg.ffn)
        self.fc2 = nn.Linear(cfg.ffn, cfg.embed_dim)
        self.final_layer_norm = nn.LayerNorm(cfg.embed_dim)

    def forward(self, x):
        x = x + self.self_attn(self.self_attn_layer_norm(x))
        h = self.fc1(self.final_layer_norm(x))
        h = F.gelu(h.float()).type_as(h)                      # fairseq gelu
        return x + self.fc2(h)


class AVHEncoder(nn.Module):  # wav2vec2.py:818-905
    def __init__(self, cfg):
        super().__init__()
        d = cfg.embed_dim
        conv = nn.Conv1d(d, d, cfg.conv_pos, padding=cfg.conv_pos // 2, groups=cfg.conv_pos_groups)
        nn.init.normal_(conv.weight, 0, math.sqrt(4.0 / (cfg.conv_pos * d)))
        nn.init.constant_(conv.bias, 0)
        conv = nn.utils.weight_norm(conv, name="weight", dim=2)
        self.pos_conv = nn.Sequential(conv)
        self.remove = 1 if cfg.conv_pos % 2 == 0 else 0
        self.layers = nn.ModuleList([AVHLayer(cfg) for _ in range(cfg.layers)])
        self.layer_norm = nn.LayerNorm(d)

    def forward(self, x):  # [B, T, C]
        xc = self.pos_conv(x.transpose(1, 2))
        if self.remove:
            xc = xc[:, :, : -self.remove]                     # SamePad
        xc = F.gelu(xc)
        x = x + xc.transpose(1, 2)
        x = x.transpose(0, 1)
        for l in self.layers:
            x = l(x)
        return self.layer_norm(x.transpose(0, 1))


class AVHubertVideo(nn.Module):
    """AVHubertModel.extract_finetune(source={'video': v, 'audio': None}) -> features [B, T, C] (hubert.py:695-755)."""

    def __init__(self, cfg: AVHubertCfg = AVHubertCfg(), widths=(64, 128, 256, 512)):
        super().__init__()
        self.cfg = cfg
        d = cfg.embed_dim
        self.feature_extractor_video = SubModel(cfg, widths)
        self.layer_norm = nn.LayerNorm(2 * d)                  # modality_fuse == concat
        self.post_extract_proj = nn.Linear(2 * d, d)
        self.encoder = AVHEncoder(cfg)

    def forward(self, video):  # [B, 1, T, 88, 88]
        fv = self.feature_extractor_video(video)               # [B, C, T]
        fa = fv.new_zeros(fv.size(0), self.cfg.embed_dim, fv.size(-1))
        f = torch.cat([fa, fv], dim=1).transpose(1, 2)
        f = self.layer_norm(f)
        f = self.post_extract_proj(f)
        return self.encoder(f)
