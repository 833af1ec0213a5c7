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


def _sync_worker(rank, world, port, ret):
    from cleanumamba_b200.distributed import GradSync
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        flat = torch.arange(1000, dtype=torch.float32) * (rank + 1)          # stand-in for TrainEngine.gflat
        sync = GradSync()
        for lo, hi in ((600, 1000), (250, 600), (0, 250)):                    # decoder, bottleneck, encoder order
            sync.reduce(flat[lo:hi])
        flat.mul_(sync.finish())
        want = torch.arange(1000, dtype=torch.float32) * sum(r + 1 for r in range(world)) / world
        if rank == 0:
            ret["ok"] = bool(torch.allclose(flat, want)) and sync.bytes_reduced == 4000
    finally:
        dist.destroy_process_group()


def test_bucketed_gradient_allreduce_gloo():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ret = mp.Manager().dict()
    mp.spawn(_sync_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret.get("ok") is True


def test_loss_matches_reference_formulas():
    """MR-STFT + L1 restatement against a direct evaluation of the reference formulas (stft_loss.py:16-184)."""
    from cleanumamba_b200.loss import DEFAULT_STFT_CONFIG, MultiResolutionSTFTLoss
    g = torch.Generator().manual_seed(0)
    x, y = torch.randn(2, 4000, generator=g) * 0.1, torch.randn(2, 4000, generator=g) * 0.1
    sc, mag = MultiResolutionSTFTLoss(**DEFAULT_STFT_CONFIG)(x, y)
    sc_ref = mag_ref = 0.0
    for fs, hop, wl in zip(DEFAULT_STFT_CONFIG["fft_sizes"], DEFAULT_STFT_CONFIG["hop_sizes"], DEFAULT_STFT_CONFIG["win_lengths"]):
        w = torch.hann_window(wl)
        mags = []
        for s_ in (x, y):
            st = torch.view_as_real(torch.stft(s_, fs, hop, wl, w, return_complex=True))
            mags.append(torch.sqrt(torch.clamp(st[..., 0] ** 2 + st[..., 1] ** 2, min=1e-7)).transpose(2, 1))
        sc_ref += torch.norm(mags[1] - mags[0], p="fro") / torch.norm(mags[1], p="fro")
        mag_ref += torch.nn.functional.l1_loss(torch.log(mags[1]), torch.log(mags[0]))
    assert torch.allclose(sc, sc_ref * 0.5 / 3) and torch.allclose(mag, mag_ref * 0.5 / 3)
