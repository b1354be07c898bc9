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


def adaptive_time_mask(x, window, stride):                         # :36-56
    cloned = x.clone()
    length = cloned.size(0)
    n_mask = int((length + stride - 0.1) // stride)
    ts = torch.randint(0, window, size=(n_mask, 2))
    for t, t_end in ts:
        if length - t <= 0:
            continue
        t_start = random.randrange(0, length - t)
        if t_start == t_start + t:
            continue
        t_end += t_start
        cloned[t_start:t_end] = 0
    return cloned


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
