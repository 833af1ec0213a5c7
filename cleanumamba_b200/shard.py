"""Utterance sharding for multi-GPU inference: every clip / stream is independent, so ranks take
disjoint contiguous slices of the batch and no data-path collective is needed.  ``gather_results`` is host-side
plumbing for callers that want the full batch back on rank 0 (torch.distributed: NCCL on GPUs, gloo in the CPU tests)."""
from typing import List, Optional, Tuple

import torch


def shard_bounds(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced [lo, hi) slice of ``n_items`` for ``rank`` (first ``n_items % world`` ranks get one more)."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world {world}")
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_batch(x: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    lo, hi = shard_bounds(x.shape[0], rank, world)
    return x[lo:hi]


def gather_results(local: torch.Tensor, n_items: int, group=None) -> Optional[torch.Tensor]:
    """Concatenate per-rank outputs (ragged along dim 0) on rank 0; other ranks get None."""
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = [shard_bounds(n_items, r, world) for r in range(world)]
    width = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((width,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    bufs: Optional[List[torch.Tensor]] = [torch.empty_like(pad) for _ in range(world)] if rank == 0 else None
    dist.gather(pad, bufs, dst=0, group=group)
    if rank != 0:
        return None
    return torch.cat([b[: hi - lo] for b, (lo, hi) in zip(bufs, sizes)], dim=0)


def max_over_ranks(value: float, device, group=None) -> float:
    """Timing reduction used by bench.py: the slowest rank defines the step time."""
    import torch.distributed as dist
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
