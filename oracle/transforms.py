"""ORACLE (test infrastructure, not product code) -- the reference's datamodule/transforms.py pipelines restated with the
same third-party calls (torchvision 0.26 / torchaudio 2.11 are installed in this image, also on the GPU box), op for op:
VideoTransform :83-104, AudioTransform :107-131, AdaptiveTimeMask :32-56, AddNoise :59-80.
Parity status: PINNED -- tests/golden/make_reference_golden.py executes the reference's own module on seeded inputs;
tests/test_reference_golden.py holds these functions to its outputs bit for bit."""
from __future__ import annotations

import random

import torch
import torchaudio
import torchvision


def time_mask_spans(length, window, stride):
    """The [start, stop) ranges AdaptiveTimeMask (:36-56) zeroes, with its RNG draws in its order: one
    `torch.randint(0, window, (n_mask, 2))` (column 0 bounds the start, column 1 is the span length), then one
    `random.randrange` per usable row."""
    n_mask = int((length + stride - 0.1) // stride)
    draws = torch.randint(0, window, size=(n_mask, 2)).tolist()
    spans = []
    for bound, width in draws:
        if bound >= length:                       # `length - t <= 0`
            continue
        start = random.randrange(0, length - bound)
        if bound == 0:                            # `t_start == t_start + t`
            continue
        spans.append((start, start + width))
    return spans


def adaptive_time_mask(x, window, stride):
    keep = torch.ones(x.size(0), dtype=torch.bool)
    for a, b in time_mask_spans(x.size(0), window, stride):
        keep[a:b] = False
    out = x.clone()
    out[~keep] = 0
    return out


def video_transform(sample, subset):                               # :83-104
    x = sample / 255.0
    if subset == "train":
        x = torchvision.transforms.RandomCrop(88)(x)
        x = torchvision.transforms.Grayscale()(x)
        x = adaptive_time_mask(x, 10, 25)
    else:
        x = torchvision.transforms.CenterCrop(88)(x)
        x = torchvision.transforms.Grayscale()(x)
    return torchvision.transforms.Normalize(0.421, 0.165)(x)


def add_noise(speech, noise, snr_levels):                          # :72-80
    speech = speech.t()
    start_idx = random.randint(0, noise.shape[1] - speech.shape[1])
    noise_segment = noise[:, start_idx: start_idx + speech.shape[1]]
    snr_level = torch.tensor([random.choice(snr_levels)])
    return torchaudio.functional.add_noise(speech, noise_segment, snr_level).t()


def audio_transform(sample, subset, noise=None, snr_target=None):  # :107-131
    x = sample
    if subset == "train":
        x = adaptive_time_mask(x, 6400, 16000)
        x = add_noise(x, noise, [-5, 0, 5, 10, 15, 20, 999999])
    elif snr_target is not None:
        x = add_noise(x, noise, [snr_target])
    return torch.nn.functional.layer_norm(x, x.shape, eps=1e-8)
