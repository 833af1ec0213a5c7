"""2-GPU NCCL test of the data-parallel gradient path (reference: apply_gradient_allreduce,
/root/reference/src/training/train_distributed.py:97-149): gradients averaged by the bucketed asynchronous all-reduce over two
ranks must equal the single-rank gradients of the concatenated batch, and rank 0's initial state must reach rank 1.
Skipped on a single-GPU box (run with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_dist.py -m gpu`)."""
import json
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import load_golden

pytestmark = pytest.mark.gpu


def _loss(out, target):
    return torch.nn.functional.l1_loss(out, target) + (out ** 2).mean()


def _worker(rank, world, port, ret):
    from cleanumamba_b200.distributed import apply_gradient_allreduce
    from cleanumamba_b200.network import Net
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        fx = load_golden("e8_pruned_500k")
        net = Net("CleanUMamba", {**json.loads(fx["config"]), "math_mode": "fp32"})
        net.load_pruned_state_dict(fx["state_dict"])
        net = net.to(dev).float().train()
        if rank != 0:           # rank 1 starts from different weights: the broadcast of :107-110 must overwrite them
            with torch.no_grad():
                for p in net.parameters():
                    p.mul_(1.5)
        apply_gradient_allreduce(net)
        T = 4800
        per = 2
        noisy = fx["noisy"][:1, :, :T].repeat(world * per, 1, 1) * (1 + 0.1 * torch.arange(world * per)[:, None, None])
        g = torch.Generator().manual_seed(9)
        target = torch.randn(noisy.shape, generator=g) * 0.05
        lo, hi = rank * per, (rank + 1) * per
        out = net(noisy[lo:hi].clone().to(dev))
        _loss(out, target[lo:hi].to(dev)).backward()
        torch.cuda.synchronize()
        grads = {k: p.grad.detach().cpu() for k, p in net.named_parameters()}
        sync = net._grad_sync
        if rank == 0:
            # single-rank gradients of the concatenated batch, same device, no sync armed
            ref = Net("CleanUMamba", {**json.loads(fx["config"]), "math_mode": "fp32"})
            ref.load_pruned_state_dict(fx["state_dict"])
            ref = ref.to(dev).float().train()
            out = ref(noisy.clone().to(dev))
            _loss(out, target.to(dev)).backward()
            worst = 0.0
            for k, p in ref.named_parameters():
                scale = p.grad.abs().max().item()
                if scale > 1e-12:
                    worst = max(worst, (grads[k] - p.grad.cpu()).abs().max().item() / scale)
            ret["worst"] = worst
            ret["bytes"] = sync.bytes_reduced
            ret["expect_bytes"] = net.train_engine().gflat.numel() * 4
        ret[f"w{rank}"] = float(sum(p.detach().double().sum().item() for p in net.parameters()))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (NCCL)")
def test_nccl_bucketed_allreduce_equals_single_rank_gradients():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret["w0"] == ret["w1"], "rank 1 did not receive rank 0's parameters"
    # mean-reduced loss over equal per-rank batches: average of rank gradients == gradient of the concatenated batch
    assert ret["worst"] < 1e-4, ret["worst"]
    assert ret["bytes"] == ret["expect_bytes"]      # the buckets tile the whole flat gradient buffer
