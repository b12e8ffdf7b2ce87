"""N > 1 host logic on CPU: world_size-2 gloo run of the shard + gather path."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from prior_diffuse_b200.shard import gather_utterances, shard_range


def test_shard_range_partitions():
    for n in (1, 2, 7, 64, 255, 256):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, n_items, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    full = torch.arange(n_items * 5, dtype=torch.float32).view(n_items, 5)
    lo, hi = shard_range(n_items, rank, world)
    local = full[lo:hi] * 2.0                     # stand-in for enhance() on this rank's utterances
    out = gather_utterances(local, n_items)
    ret[rank] = bool(torch.equal(out, full * 2.0))
    dist.barrier()
    dist.destroy_process_group()


def test_gather_world2_gloo():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    for n_items in (7, 8):
        ret = mp.Manager().dict()
        mp.spawn(_worker, args=(2, port, n_items, ret), nprocs=2, join=True)
        assert ret[0] and ret[1]
