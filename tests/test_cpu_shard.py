"""CPU suite: the N>1 (utterance-sharded) host path over gloo, world_size 2 and 3."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cleanumamba_b200.shard import gather_results, max_over_ranks, shard_batch, shard_bounds


def test_shard_bounds_partition():
    for n in (0, 1, 7, 64, 4096, 4097):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(4, 2, 2)


def _worker(rank, world, port, n_items, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        full = torch.arange(n_items * 3, dtype=torch.float32).view(n_items, 3)
        local = shard_batch(full, rank, world) * 2.0            # stand-in for net(local_batch)
        out = gather_results(local, n_items)
        slow = max_over_ranks(10.0 + rank, torch.device("cpu"))
        if rank == 0:
            ret["ok"] = bool(torch.equal(out, full * 2.0)) and slow == 10.0 + world - 1
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_items", [(2, 7), (3, 8)])
def test_sharded_inference_plumbing_gloo(world, n_items):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(world, port, n_items, ret), nprocs=world, join=True)
    assert ret.get("ok") is True
