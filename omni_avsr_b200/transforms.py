"""On-device mirror of the reference's datamodule/transforms.py: `VideoTransform` and
`AudioTransform` with the reference's constructor arguments and `__call__` contract, executed by the kernels of
csrc/transforms.cu on the GPU instead of torchvision / torchaudio ops in the dataloader workers.

Reference (file:line in /root/reference/datamodule/transforms.py): FunctionalModule :23-29, AdaptiveTimeMask :32-56,
AddNoise :59-80, VideoTransform :83-104, AudioTransform :107-131.

Random numbers are drawn on the host with EXACTLY the reference's calls in the reference's order (torchvision
`RandomCrop.get_params`: two `torch.randint`; `AdaptiveTimeMask`: one `torch.randint(0, window, (n_mask, 2))` then one
`random.randrange` per mask; `AddNoise`: `random.randint` then `random.choice`), so that with the same seeds the same
crop, masks, noise segment and SNR are applied; only the arithmetic moves to the device.
"""
from __future__ import annotations

import ctypes as C
import random
from typing import List, Optional, Tuple

import torch

from ._lib import check, lib, require_cuda, stream_ptr

MAX_SPANS = 48


def adaptive_time_mask_spans(length: int, window: int, stride: int) -> List[Tuple[int, int]]:
    """AdaptiveTimeMask.forward (:42-56) without touching the data: the [start, end) spans it would zero."""
    n_mask = int((length + stride - 0.1) // stride)
    draws = torch.randint(0, window, size=(n_mask, 2)).tolist()      # column 0 bounds the start, column 1 is the width
    spans = []
    for bound, width in draws:
        if bound >= length:                                           # reference: `length - t <= 0`
            continue
        start = random.randrange(0, length - bound)                   # drawn even when the row is then discarded
        if bound == 0 or width == 0:                                  # `t_start == t_start + t` / empty slice
            continue
        spans.append((start, min(start + width, length)))
    return spans


def _span_array(spans):
    if len(spans) > MAX_SPANS:
        raise ValueError(f"more than {MAX_SPANS} mask spans")
    arr = (C.c_int32 * (2 * max(len(spans), 1)))()
    for i, (a, b) in enumerate(spans):
        arr[2 * i], arr[2 * i + 1] = a, b
    return arr


class VideoTransform:
    """sample: uint8 [T, C, H, W] (C = 1 or 3; host or device) -> [T, 1, 88, 88] on the device, fp32 (bit-exact with the
    reference) or bf16 (`out_dtype=torch.bfloat16`: what Lightning's bf16-true feeds the model)."""

    def __init__(self, subset: str, device="cuda", out_dtype=torch.float32):
        if subset not in ("train", "val", "test"):
            raise ValueError(subset)
        self.subset, self.device, self.out_dtype = subset, torch.device(device), out_dtype

    def __call__(self, sample: torch.Tensor) -> torch.Tensor:
        if sample.dtype != torch.uint8 or sample.dim() != 4:
            raise TypeError("sample must be uint8 [T, C, H, W]")
        T, Cn, H, W = sample.shape
        if H < 88 or W < 88:
            raise ValueError(f"Required crop size (88, 88) is larger than input image size {(H, W)}")
        if self.subset == "train":
            if W == 88 and H == 88:                                   # RandomCrop.get_params
                i = j = 0
            else:
                i = torch.randint(0, H - 88 + 1, size=(1,)).item()
                j = torch.randint(0, W - 88 + 1, size=(1,)).item()
            spans = adaptive_time_mask_spans(T, 10, 25)                # :91
        else:
            i, j = int(round((H - 88) / 2.0)), int(round((W - 88) / 2.0))   # CenterCrop
            spans = []
        x = sample.to(self.device, non_blocking=True).contiguous()
        out = torch.empty((T, 1, 88, 88), device=self.device, dtype=self.out_dtype)
        arr = _span_array(spans)
        check(lib.omni_video_transform(x.data_ptr(), T, Cn, H, W, i, j, arr, len(spans), out.data_ptr(),
                                       1 if self.out_dtype == torch.bfloat16 else 0, stream_ptr()), "omni_video_transform")
        return out


class AudioTransform:
    """sample: float [T, 1] -> [T, 1] fp32 on the device.  `noise`: the babble-noise waveform [1, N] (the reference loads
    babble_noise.wav, :68); required for subset='train' and for an `snr_target`."""

    def __init__(self, subset: str, snr_target=None, noise: Optional[torch.Tensor] = None, device="cuda"):
        if subset not in ("train", "val", "test"):
            raise ValueError(subset)
        self.subset, self.device = subset, torch.device(device)
        self.add_noise = subset == "train" or snr_target is not None
        self.snr_levels = [snr_target] if snr_target else [-5, 0, 5, 10, 15, 20, 999999]     # :66
        if self.add_noise:
            if noise is None or noise.dim() != 2:
                raise ValueError("noise waveform [1, N] required")
            self.noise = noise.to(self.device, torch.float32).contiguous()
        self._ws = torch.empty(int(lib.omni_audio_transform_workspace_bytes()), dtype=torch.uint8, device=self.device)

    def __call__(self, sample: torch.Tensor) -> torch.Tensor:
        if sample.dim() != 2 or sample.shape[1] != 1:
            raise TypeError("sample must be [T, 1]")
        T = sample.shape[0]
        spans = adaptive_time_mask_spans(T, 6400, 16000) if self.subset == "train" else []   # :111
        x = sample.to(self.device, torch.float32, non_blocking=True).contiguous()
        noise_ptr, snr = None, 0.0
        if self.add_noise:
            start_idx = random.randint(0, self.noise.shape[1] - T)                          # :76
            snr = float(random.choice(self.snr_levels))                                       # :78
            seg = self.noise[0, start_idx: start_idx + T]
            noise_ptr = seg.data_ptr()
        out = torch.empty((T, 1), device=self.device, dtype=torch.float32)
        arr = _span_array(spans)
        check(lib.omni_audio_transform(x.data_ptr(), noise_ptr, T, snr, arr, len(spans), out.data_ptr(), self._ws.data_ptr(),
                                       self._ws.numel(), stream_ptr()), "omni_audio_transform")
        return out
