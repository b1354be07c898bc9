"""Synthetic inputs with the reference's batch-dict layout (datamodule/data_module.py:19-79 `collate_LLM`), as
specified in SURVEY.md §8(d) / BASELINE.md §5: 16 kHz audio, 25 fps 88x88 lip ROIs, BOS ... EOS token rows."""
from __future__ import annotations

import torch


def synthetic_batch(B: int, tokenizer, seconds: float = 16.0, text_len: int = 48, seed: int = 1234, device="cpu",
                    dtype=torch.bfloat16, pin: bool = False):
    g = torch.Generator().manual_seed(seed)
    n_samples = int(round(seconds * 16000))
    n_frames = int(round(seconds * 25))
    audio = torch.randn(B, n_samples, generator=g)
    audio = torch.nn.functional.layer_norm(audio, (n_samples,))          # transforms.py:115,124 (utterance layer-norm)
    video = (torch.rand(B, n_frames, 1, 88, 88, generator=g) - 0.421) / 0.165   # transforms.py:95-98
    is_qwen = getattr(tokenizer, "is_qwen", False)
    L = text_len - 1 if is_qwen else text_len
    tokens = torch.randint(0, tokenizer.base_vocab, (B, L), generator=g)
    if not is_qwen:
        tokens[:, 0] = tokenizer.bos_token_id
    tokens[:, -1] = tokenizer.eos_token_id
    batch = {
        "tokens": tokens,
        "labels": tokens.clone(),
        "audio": audio.unsqueeze(-1).to(dtype),
        "lengths": torch.full((B,), n_samples, dtype=torch.int64),
        "video": video.to(dtype),
    }
    if pin:
        batch = {k: v.pin_memory() for k, v in batch.items()}
    if str(device) != "cpu":
        batch = to_device(batch, device)
    return batch


def to_device(batch, device, non_blocking=True):
    out = {}
    for k, v in batch.items():
        if k == "lengths":
            out[k] = v            # stays on the host: only max(lengths) is used, on the host (modeling_OmniAVSR.py:537)
        elif torch.is_tensor(v):
            out[k] = v.to(device, non_blocking=non_blocking)
        else:
            out[k] = v
    return out


def host_bytes(batch) -> int:
    return sum(v.numel() * v.element_size() for k, v in batch.items() if torch.is_tensor(v) and k != "lengths")


class DevicePrefetcher:
    """Double-buffered host -> device feed: the copy of batch k+1 (from pinned host memory, on its own stream) runs under
    the compute of batch k.  `next()` returns the device copy of the batch whose transfer was started by the previous call
    (or starts one if there is none) and immediately starts the transfer of the following batch.

        feed = DevicePrefetcher(lambda k: host_batches[k % n], device)
        for k in range(steps):
            loss = module.train_step(feed.next())
    """

    def __init__(self, host_batch_fn, device):
        self.fn, self.device = host_batch_fn, torch.device(device)
        self.stream = torch.cuda.Stream(device=self.device)
        self.k = 0
        self._pending = None

    def _start(self):
        host = self.fn(self.k)
        self.k += 1
        with torch.cuda.stream(self.stream):
            dev = to_device(host, self.device, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        return dev, ev

    def next(self):
        dev, ev = self._pending if self._pending is not None else self._start()
        torch.cuda.current_stream(self.device).wait_event(ev)
        for v in dev.values():                    # the consumer stream now owns these buffers (caching-allocator safety)
            if torch.is_tensor(v) and v.is_cuda:
                v.record_stream(torch.cuda.current_stream(self.device))
        self._pending = self._start()
        return dev
