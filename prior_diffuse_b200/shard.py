"""Utterance sharding for multi-GPU runs: with eval-mode BatchNorm every utterance is
independent end to end, so rank r simply takes a contiguous slice of the batch; nothing is exchanged
on the path.  The only collective is one all_gather of the enhanced waveforms at the end."""
from __future__ import annotations

from typing import List, Tuple

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """contiguous, balanced split: the first n_items % world ranks get one extra item"""
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_utterances(local: torch.Tensor, n_items: int) -> torch.Tensor:
    """local [n_local, L] on every rank -> [n_items, L] on every rank (one all_gather; NCCL on GPUs)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    world, rank = dist.get_world_size(), dist.get_rank()
    sizes = [shard_range(n_items, r, world) for r in range(world)]
    n_max = max(hi - lo for lo, hi in sizes)
    padded = local.new_zeros((n_max,) + tuple(local.shape[1:]))
    padded[:local.shape[0]] = local
    parts: List[torch.Tensor] = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded)
    return torch.cat([parts[r][:hi - lo] for r, (lo, hi) in enumerate(sizes)], dim=0)
