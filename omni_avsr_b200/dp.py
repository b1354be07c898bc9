"""Data-parallel host logic (one process per GPU, torch.distributed).

The path shards by utterance: every (utterance, task, rate-pair) cell is an independent forward/backward and the only
exchange is the sum of the trainable gradients (reference: DDPStrategy at train_OmniAVSR.py:46-49 + the W/sum(B)
loss scaling at lightning_OmniAVSR.py:171-173).  These helpers are backend-agnostic (NCCL on the GPUs, gloo in the CPU
tests).

Two things keep the collective off the critical path:
  * `LossScale`: the all_gather of the per-rank batch sizes (one int per rank) is STARTED before the forward and only
    waited for when the loss is scaled, so it never blocks the stream between forward and backward;
  * `GradReducer`: the flat gradient buffer is reduced in two pieces -- the LLM adapters' range as soon as the
    gradient of the LLM input exists (an autograd hook: every LLM layer has finished its backward by then), under the
    backward of the projectors and of the AV-HuBERT encoder, and the remaining range after the backward."""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.distributed as dist


def world() -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_utterances(n_utterances: int, rank: int, world_size: int) -> List[int]:
    """Rank k takes utterances k::W."""
    return list(range(rank, n_utterances, world_size))


class LossScale:
    """W / sum_r B_r  (lightning_OmniAVSR.py:171-173).  start() launches the (asynchronous) gather of the batch sizes,
    value() waits for it.  With one rank it is the host constant 1 / B: no collective, no device work."""

    def __init__(self, local_batch: int, device=None):
        self.rank, self.w = world()
        self.local = int(local_batch)
        self.work = None
        if self.w > 1:
            self.t = torch.tensor([self.local], dtype=torch.int64, device=device)
            self.gathered = torch.zeros(self.w, dtype=torch.int64, device=device)
            self.work = dist.all_gather_into_tensor(self.gathered, self.t, async_op=True)

    def value(self):
        if self.w == 1:
            return 1.0 / self.local
        self.work.wait()
        return self.w / self.gathered.sum().float()


def loss_scale(local_batch: int, device=None):
    """Blocking form (kept for callers that have nothing to overlap with)."""
    return LossScale(local_batch, device).value()


class GradReducer:
    """Sum of the flat trainable-gradient buffer over ranks, in place, in two asynchronous pieces.

        red = GradReducer(flat_grad, split)      # [split, n): gradients that are final once the LLM backward is done
        x.register_hook(red.hook)                # x = packed LLM input: its gradient appears right after LLM layer 0
        loss.backward()
        factor = red.finish()                    # reduces [0, split), waits for both; returns 1 / W

    Without the hook firing (no LLM input gradient, e.g. everything upstream frozen) finish() reduces the whole buffer."""

    def __init__(self, flat_grad: torch.Tensor, split: int):
        self.g, self.split = flat_grad, int(split)
        self.rank, self.w = world()
        self.tail = None

    def hook(self, grad):
        if self.w > 1 and self.tail is None and self.split < self.g.numel():
            self.tail = dist.all_reduce(self.g[self.split:], async_op=True)
        return grad

    def finish(self) -> float:
        if self.w > 1:
            if self.tail is None:
                dist.all_reduce(self.g)
            else:
                head = dist.all_reduce(self.g[: self.split], async_op=True) if self.split > 0 else None
                self.tail.wait()
                if head is not None:
                    head.wait()
            self.tail = None
        return 1.0 / self.w


def allreduce_flat_grad(flat_grad: torch.Tensor) -> float:
    """Sum the flat trainable-gradient buffer over ranks in place; returns the factor (1/W) the optimizer kernel
    applies while reading it (DDP averages gradients)."""
    rank, w = world()
    if w > 1:
        dist.all_reduce(flat_grad)
    return 1.0 / w
