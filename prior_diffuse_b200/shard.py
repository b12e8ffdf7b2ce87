"""Utterance sharding for multi-GPU runs: with eval-mode BatchNorm every utterance is
independent end to end, so rank r simply takes a contiguous slice of the batch; nothing is exchanged
on the path.  The only collective is one all-gather of the enhanced waveforms at the end, into a
preallocated [n_items, L] buffer (no per-step allocation, no concatenation)."""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
import torch.distributed as dist

_BUFFERS: Dict[tuple, Tuple[torch.Tensor, Optional[torch.Tensor]]] = {}


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """contiguous, balanced split: the first n_items % world ranks get one extra item"""
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_utterances(local: torch.Tensor, n_items: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """local [n_local, L] on every rank -> [n_items, L] on every rank: ONE all_gather_into_tensor (NCCL on GPUs).

    ``out`` (optional, [n_items, L]) receives the result; otherwise a buffer cached per (shape, dtype, device) is
    returned, which the next call with the same shape overwrites.  Equal shards are gathered straight into it; an
    uneven split goes through a cached padded staging buffer and one strided copy."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    tail = tuple(local.shape[1:])
    key = (n_items, tail, local.dtype, local.device, world)
    bufs = _BUFFERS.get(key)
    if bufs is None:
        full = torch.empty((n_items,) + tail, dtype=local.dtype, device=local.device)
        n_max = -(-n_items // world)
        stage = None if n_items % world == 0 else torch.zeros((world * n_max,) + tail, dtype=local.dtype, device=local.device)
        bufs = _BUFFERS[key] = (full, stage)
    full, stage = bufs
    dst = out if out is not None else full
    if stage is None:
        dist.all_gather_into_tensor(dst, local.contiguous())
        return dst
    n_max = stage.shape[0] // world
    rank = dist.get_rank()
    mine = stage[rank * n_max:rank * n_max + local.shape[0]]
    mine.copy_(local)
    dist.all_gather_into_tensor(stage, stage[rank * n_max:(rank + 1) * n_max])
    for r in range(world):
        lo, hi = shard_range(n_items, r, world)
        dst[lo:hi].copy_(stage[r * n_max:r * n_max + hi - lo])
    return dst
