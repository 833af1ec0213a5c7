"""Data-parallel gradient averaging.

Reference: ``apply_gradient_allreduce`` (/root/reference/src/training/train_distributed.py:97-149) broadcasts the state
from rank 0, then -- once the WHOLE backward has finished -- flattens every gradient into one 165 MB fp32 message,
all-reduces it, divides by the world size and copies it back (no overlap with compute).

Here the backward kernels already write into ONE flat fp32 buffer in packed-weight order (train_engine.TrainEngine.gflat),
so there is no flatten / unflatten, and the buffer is reduced in four contiguous buckets as soon as each is complete --
decoder, bottleneck, the deep encoder levels, the outer encoder levels: the order the backward produces them -- with
``async_op=True`` so NCCL (NVLink 5 / NVSwitch; NVLS in-switch reduction when available) runs on its own stream underneath
the remaining backward kernels.  Only the last bucket is exposed, and it is the outer encoder levels' 1.3 MB: the deep levels
(58 MB of the encoder's 59 MB) are reduced under the backward of the outer levels, whose activations are the largest of the
model (compute-stream stall on NCCL at 2 GPUs: 0.68 -> 0.04 ms per step).  The same object works over gloo for the CPU tests.
"""
from typing import List, Optional

import torch
import torch.distributed as dist


class GradSync:
    """Bucketed asynchronous SUM all-reduce of slices of a flat gradient buffer, averaged at ``finish``."""

    def __init__(self, group=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.works: List = []
        self.bytes_reduced = 0
        self.bucket_bytes: List[int] = []       # bytes of each bucket of the most recent backward, in issue order
        self._stall: List = []                  # (event before the waits, event after): the compute stream's stall on NCCL

    def reduce(self, flat_slice: torch.Tensor) -> None:
        """Start reducing ``flat_slice`` (a contiguous view of the gradient buffer); returns immediately."""
        if self.world == 1 or flat_slice.numel() == 0:
            return
        assert flat_slice.is_contiguous()
        if not self.works:
            self.bucket_bytes = []
        self.bytes_reduced += flat_slice.numel() * flat_slice.element_size()
        self.bucket_bytes.append(flat_slice.numel() * flat_slice.element_size())
        self.works.append(dist.all_reduce(flat_slice, op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def finish(self) -> float:
        """Wait for every bucket (stream-ordered on CUDA) and return the averaging factor 1 / world."""
        timed = bool(self.works) and torch.cuda.is_available() and dist.get_backend(self.group) == "nccl"
        if timed:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        for w in self.works:
            w.wait()
        if timed:
            e1.record()
            self._stall.append((e0, e1))
            del self._stall[:-64]
        self.works.clear()
        return 1.0 / self.world

    def exposed_ms(self) -> float:
        """Mean time per backward the compute stream spent waiting for the all-reduces (CUDA events around the waits of
        ``finish``: the part of the collective NOT hidden behind the backward kernels).  Call after a synchronize; resets."""
        if not self._stall:
            return 0.0
        ms = sum(a.elapsed_time(b) for a, b in self._stall) / len(self._stall)
        self._stall.clear()
        return ms


def apply_gradient_allreduce(module, group=None):
    """Same entry point as the reference's: synchronise the initial state from rank 0 (:107-110) and arm gradient
    averaging for every subsequent backward of ``module`` (a cleanumamba_b200.CleanUMamba)."""
    if not (dist.is_available() and dist.is_initialized()):
        raise RuntimeError("apply_gradient_allreduce: torch.distributed is not initialised")
    for p in module.state_dict().values():
        if torch.is_tensor(p):
            dist.broadcast(p, 0, group=group)
    module._grad_sync = GradSync(group)
    return module


def reduce_tensor(tensor: torch.Tensor, num_gpus: Optional[int] = None) -> torch.Tensor:
    """Mean of a scalar over ranks for logging (train_distributed.py:44-48)."""
    rt = tensor.clone()
    dist.all_reduce(rt, op=dist.ReduceOp.SUM)
    return rt / (num_gpus or dist.get_world_size())
