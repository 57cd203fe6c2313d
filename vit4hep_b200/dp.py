"""Data parallelism for the CFM-ViT path: one process per GPU, NCCL over NVLink 5 / NVSwitch.

Training (replaces the reference's ``DistributedDataParallel(model.net)``,
experiments/base_experiment.py:161-167): gradients live in one flat fp32 buffer ordered by the time
the backward chain completes them (final layer, blocks depth-1..0, embeddings/conditioning).  The
native backward is issued in stage ranges; as soon as a range is enqueued its slice of the flat
buffer is all-reduced (average) asynchronously, so the collective of bucket k runs on NCCL's stream
while the compute stream executes the backward of bucket k+1.  DDP's Reducer cannot do this for a
single fused autograd node (every gradient would arrive at the end).

Sampling (new: the reference samples on rank 0 only): conditions are split in
contiguous per-rank ranges, each rank integrates its showers with no communication, one all_gather
at the end restores the global order.

Everything here is device-agnostic host logic (`gloo` on CPU tensors in the tests, `nccl` on GPUs).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

__all__ = ["Bucket", "plan_buckets", "GradReducer", "enable_data_parallel", "disable_data_parallel",
           "broadcast_parameters", "shard_range", "sample_sharded", "per_rank_batch"]


@dataclass(frozen=True)
class Bucket:
    stage_begin: int  # first (highest) backward stage of the bucket
    stage_end: int    # last (lowest) backward stage, inclusive
    start: int        # element range [start, stop) of the flat gradient buffer
    stop: int


def plan_buckets(stage_bounds: Sequence[int], min_elems: int) -> List[Bucket]:
    """Group consecutive backward stages into buckets of at least ``min_elems`` gradient elements.

    ``stage_bounds[k]`` is the flat-buffer offset at which stage ``depth + 1 - k`` ends
    (ViT.stage_boundaries()).  A short tail is merged into the bucket before it -- except a tail that is
    stage 0 alone (embeddings / conditioning, 2.5 MB at ds2): the all-reduce of the LAST bucket is the one
    nothing overlaps, so block 0's bucket is issued before stage 0 runs and only the small one is exposed.
    """
    nstages = len(stage_bounds)
    top = nstages - 1  # stage index of the final layer = depth + 1
    buckets: List[Bucket] = []
    start, first = 0, top
    for k, stop in enumerate(stage_bounds):
        stage = top - k
        if stop - start >= min_elems or stage == 0:
            buckets.append(Bucket(first, stage, start, stop))
            start, first = stop, stage - 1
    if len(buckets) >= 2 and buckets[-1].stop - buckets[-1].start < min_elems and buckets[-1].stage_begin != 0:
        a, b = buckets[-2], buckets[-1]
        buckets[-2:] = [Bucket(a.stage_begin, b.stage_end, a.start, b.stop)]
    return buckets


class GradReducer:
    """Averages slices of a flat gradient buffer across a process group, asynchronously."""

    def __init__(self, group: Optional[dist.ProcessGroup] = None, min_bucket_elems: int = 4_000_000):
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised")
        self.group = group
        self.world = dist.get_world_size(group)
        self.min_bucket_elems = int(min_bucket_elems)
        self._avg = dist.get_backend(group) == "nccl"  # gloo has no AVG
        self.launched = 0  # buckets launched by the last backward (tests / gpu_launches accounting)

    def reduce_async(self, flat_slice: torch.Tensor):
        if self._avg:
            return dist.all_reduce(flat_slice, op=dist.ReduceOp.AVG, group=self.group, async_op=True)
        work = dist.all_reduce(flat_slice, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        return _Scaled(work, flat_slice, 1.0 / self.world)

    def run(self, buckets: Sequence[Bucket], flat: torch.Tensor,
            run_stages: Callable[[int, int], None]) -> None:
        """run_stages(stage_begin, stage_end) enqueues that part of the backward chain."""
        works = []
        for b in buckets:
            run_stages(b.stage_begin, b.stage_end)
            works.append(self.reduce_async(flat[b.start:b.stop]))
        self.launched = len(works)
        for w in works:
            w.wait()

    # called by vit._ViTFunction.backward
    def backward(self, module, plan, x, c, dout, ws, ordered, flat) -> None:
        buckets = plan_buckets(module.stage_boundaries(), self.min_bucket_elems)
        self.run(buckets, flat, lambda hi, lo: module._run_backward(plan, x, c, dout, ws, ordered, flat, hi, lo))


class _Scaled:
    def __init__(self, work, tensor, scale):
        self.work, self.tensor, self.scale = work, tensor, scale

    def wait(self):
        self.work.wait()
        self.tensor.mul_(self.scale)


def broadcast_parameters(module: torch.nn.Module, src: int = 0, group=None) -> None:
    """Make every rank start from rank ``src``'s parameters and buffers (what DDP does at wrap time)."""
    with torch.no_grad():
        for t in list(module.parameters()) + list(module.buffers()):
            dist.broadcast(t, src=src, group=group)
    # a collective writes the storage without necessarily bumping Tensor._version: rebuild the bf16 operand copies
    invalidate = getattr(module, "invalidate_weights", None)
    if invalidate is not None:
        invalidate()


def enable_data_parallel(net, group=None, min_bucket_elems: int = 4_000_000, broadcast: bool = True):
    """Switch a vit4hep_b200.ViT to data-parallel training and return it (use instead of wrapping
    ``model.net`` in DistributedDataParallel)."""
    if broadcast:
        broadcast_parameters(net, 0, group)
    net._dp = GradReducer(group, min_bucket_elems)
    return net


def disable_data_parallel(net):
    net._dp = None
    return net


def per_rank_batch(global_batch: int, world: int) -> int:
    """reference experiments/calochallenge/experiment.py:94-98: batchsize // world_size per rank"""
    return global_batch // world


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [begin, end) of ``n`` items owned by ``rank``; the first n % world ranks get one more."""
    base, extra = divmod(n, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


@torch.inference_mode()
def sample_sharded(model, conditions: torch.Tensor, batch_size: int, group=None,
                   gather: bool = True) -> torch.Tensor:
    """Sample one shower per row of ``conditions`` (identical on every rank), sharded over the
    process group.  Returns all showers in input order on every rank (``gather``) or this rank's shard."""
    if dist.is_initialized():
        rank, world = dist.get_rank(group), dist.get_world_size(group)
    else:
        rank, world = 0, 1
    n = conditions.shape[0]
    begin, end = shard_range(n, rank, world)
    parts = [model.sample_batch(conditions[i:min(i + batch_size, end)]) for i in range(begin, end, batch_size)]
    if parts:
        local = torch.cat(parts)
    else:
        probe = model.sample_batch(conditions[:1])
        local = probe[:0]
    if world == 1 or not gather:
        return local
    # ranks may differ by one row: pad to the largest shard, all_gather, trim
    longest = -(-n // world)
    padded = local.new_zeros((longest, *local.shape[1:]))
    padded[: local.shape[0]] = local
    out = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(out, padded, group=group)
    sizes = [shard_range(n, r, world) for r in range(world)]
    return torch.cat([o[: e - b] for o, (b, e) in zip(out, sizes)])
