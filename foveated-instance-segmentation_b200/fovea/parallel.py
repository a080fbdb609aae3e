"""Multi-GPU plumbing of the foveated path: one process per GPU, frames sharded by batch, no data-path collective.

Inference: every frame is independent through S1-S3, so rank r of N simply owns a contiguous
slice of the batch (`shard_range`) and its outputs stay on its GPU.  Training adds exactly one collective per step: the
all-reduce of the saliency + compress network gradients (382,729 parameters = 1.53 MB), which the reference leaves to
DDP's bucketing (`train_deform_semantic.py:395`); `FlatGradBucket` does it as ONE flat NCCL all-reduce over
NVLink/NVSwitch (latency-bound: one launch instead of DDP's per-bucket hooks).  Works with the `gloo` backend on CPU
tensors too, which is how the host-side logic is tested without GPUs.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n_frames: int, rank: int, world: int):
    """Contiguous slice [lo, hi) of `n_frames` owned by `rank`; the first n % world ranks get one extra frame."""
    if world <= 0 or not 0 <= rank < world:
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, extra = divmod(n_frames, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_batch(feed_dict: dict, rank: int, world: int) -> dict:
    """The reference's feed_dict (`train_deform_semantic.py:77`) restricted to this rank's frames."""
    n = next(iter(feed_dict.values())).shape[0]
    lo, hi = shard_range(n, rank, world)
    return {k: v[lo:hi] for k, v in feed_dict.items()}


class FlatGradBucket:
    """One flat all-reduce (mean) for the gradients of a small set of modules (saliency + compress networks)."""

    def __init__(self, modules, group=None):
        self.params = [p for m in modules for p in m.parameters() if p.requires_grad]
        if not self.params:
            raise ValueError("FlatGradBucket: no trainable parameters")
        self.group = group
        p0 = self.params[0]
        self.numel = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(self.numel, device=p0.device, dtype=p0.dtype)
        self.views, off = [], 0
        for p in self.params:
            self.views.append(self.flat[off:off + p.numel()].view_as(p))
            off += p.numel()

    def allreduce(self, async_op: bool = False):
        """Pack grads (missing grads count as zero), all-reduce SUM, divide by world size, unpack in place."""
        for p, v in zip(self.params, self.views):
            if p.grad is None:
                v.zero_()
            else:
                v.copy_(p.grad)
        world = dist.get_world_size(self.group)
        work = dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group, async_op=async_op)
        if async_op:
            return _Pending(self, work, world)
        self._unpack(world)
        return None

    def _unpack(self, world):
        self.flat.div_(world)
        for p, v in zip(self.params, self.views):
            if p.grad is None:
                p.grad = v.clone()
            else:
                p.grad.copy_(v)


class _Pending:
    def __init__(self, bucket, work, world):
        self.bucket, self.work, self.world = bucket, work, world

    def wait(self):
        self.work.wait()
        self.bucket._unpack(self.world)
