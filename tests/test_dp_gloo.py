"""Host-side data-parallel logic (vit4hep_b200/dp.py) on CPU: world_size-2 `gloo` process groups.

Covers what replaces the reference's DistributedDataParallel wrap (experiments/base_experiment.py:161-167)
and the new sharded sampling: bucket planning, the bucket-by-bucket averaged
all-reduce interleaved with the staged backward, parameter broadcast, shard ranges, ordered gather."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from vit4hep_b200 import dp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _run(rank, world, port, fn, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ret[rank] = fn(rank, world)
    finally:
        dist.destroy_process_group()


def _spawn(fn, world=2):
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_run, args=(world, _free_port(), fn, ret), nprocs=world, join=True)
    return [ret[r] for r in range(world)]


def test_plan_buckets_groups_stages_in_backward_order():
    # stage bounds of a depth-3 net: final layer, block 2, block 1, block 0, stage 0
    bounds = [10, 110, 210, 310, 400]
    b = dp.plan_buckets(bounds, min_elems=150)
    assert [(x.stage_begin, x.stage_end, x.start, x.stop) for x in b] == [(4, 2, 0, 210), (1, 0, 210, 400)]
    one = dp.plan_buckets(bounds, min_elems=10 ** 9)
    assert len(one) == 1 and (one[0].start, one[0].stop, one[0].stage_begin, one[0].stage_end) == (0, 400, 4, 0)
    # a tail that is not stage 0 alone is merged into the bucket before it
    merged = dp.plan_buckets([10, 110, 210, 230, 240], min_elems=100)
    assert [(x.stage_begin, x.stage_end) for x in merged] == [(4, 3), (2, 0)]
    each = dp.plan_buckets(bounds, min_elems=1)
    assert [x.stop - x.start for x in each] == [10, 100, 100, 100, 90]
    # every element is covered exactly once, in order
    for plan in (b, one, each):
        assert plan[0].start == 0 and plan[-1].stop == 400
        assert all(p.stop == q.start and p.stage_end == q.stage_begin + 1 for p, q in zip(plan, plan[1:]))


def test_shard_range_partitions_in_order():
    for n, world in [(100000, 8), (10, 3), (2, 4), (0, 2)]:
        spans = [dp.shard_range(n, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [e - b for b, e in spans]
        assert max(sizes) - min(sizes) <= 1
    assert dp.per_rank_batch(64, 8) == 8


def _reduce_job(rank, world):
    bounds = [8, 40, 72, 100]
    flat = torch.zeros(100)
    calls = []

    def run_stages(hi, lo):  # the native backward of stages hi..lo fills its slice of the flat buffer
        calls.append((hi, lo))
        top = len(bounds) - 1
        start = 0 if hi == top else bounds[top - hi - 1]
        stop = bounds[top - lo]
        flat[start:stop] = torch.arange(start, stop, dtype=torch.float32) * (rank + 1)

    red = dp.GradReducer(min_bucket_elems=30)
    red.run(dp.plan_buckets(bounds, red.min_bucket_elems), flat, run_stages)
    return flat, calls, red.launched


def test_bucketed_allreduce_averages_while_the_backward_is_staged():
    out = _spawn(_reduce_job)
    want = torch.arange(100, dtype=torch.float32) * 1.5  # mean of x*1 and x*2
    for flat, calls, launched in out:
        assert torch.allclose(flat, want)
        # block 0 closes its own bucket (32 >= 30), so the short stage-0 tail stays separate: only it is exposed
        assert calls == [(3, 2), (1, 1), (0, 0)] and launched == 3


class _FakeModel:
    """sample_batch(cond) -> one deterministic 'shower' per condition row"""

    def sample_batch(self, cond):
        return cond[:, :1].reshape(-1, 1, 1, 1, 1) * torch.ones(1, 1, 2, 2, 2)


def _sample_job(rank, world):
    cond = torch.arange(11, dtype=torch.float32).reshape(-1, 1).repeat(1, 3)
    full = dp.sample_sharded(_FakeModel(), cond, batch_size=2)
    shard = dp.sample_sharded(_FakeModel(), cond, batch_size=4, gather=False)
    return full, shard


def test_sharded_sampling_keeps_the_global_order():
    out = _spawn(_sample_job)
    want = torch.arange(11, dtype=torch.float32).reshape(-1, 1, 1, 1, 1) * torch.ones(1, 1, 2, 2, 2)
    for rank, (full, shard) in enumerate(out):
        assert torch.equal(full, want)
        b, e = dp.shard_range(11, rank, 2)
        assert torch.equal(shard, want[b:e])


def _broadcast_job(rank, world):
    torch.manual_seed(100 + rank)
    lin = torch.nn.Linear(4, 3)
    dp.broadcast_parameters(lin)
    return lin.weight.detach().clone()


def test_broadcast_parameters_makes_ranks_identical():
    w0, w1 = _spawn(_broadcast_job)
    assert torch.equal(w0, w1)
