"""Data-parallel host logic (one process per GPU, torch.distributed).

The path shards by utterance: every (utterance, task, rate-pair) cell is an independent forward/backward and the only
exchange is the sum of the trainable gradients (reference: DDPStrategy at train_OmniAVSR.py:46-49 + the W/sum(B)
loss scaling at lightning_OmniAVSR.py:171-173).  Here that is ONE all-reduce of the flat gradient buffer per step.
These helpers are backend-agnostic (NCCL on the GPUs, gloo in the CPU tests)."""
from __future__ import annotations

from typing import List, Tuple

import torch
import torch.distributed as dist


def world() -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_utterances(n_utterances: int, rank: int, world_size: int) -> List[int]:
    """Rank k takes utterances k::W."""
    return list(range(rank, n_utterances, world_size))


def loss_scale(local_batch: int, device=None) -> torch.Tensor:
    """W / sum_r B_r  (lightning_OmniAVSR.py:171-173); an all_gather of one int per rank."""
    rank, w = world()
    t = torch.tensor([local_batch], dtype=torch.int64, device=device)
    if w == 1:
        return (1.0 / t.float())[0]
    gathered = [torch.zeros_like(t) for _ in range(w)]
    dist.all_gather(gathered, t)
    return (w / torch.cat(gathered).sum().float())


def allreduce_flat_grad(flat_grad: torch.Tensor) -> float:
    """Sum the flat trainable-gradient buffer over ranks in place; returns the factor (1/W) the optimizer kernel
    applies while reading it (DDP averages gradients)."""
    rank, w = world()
    if w > 1:
        dist.all_reduce(flat_grad)
    return 1.0 / w
