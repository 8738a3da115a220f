"""Host-side logic of the data-parallel wrapper on CPU: bucket planning and the per-bucket gradient
all-reduce with the gloo backend at world_size 2 (the NCCL path is exercised by bench.py --gpus N)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_plan_buckets_covers_the_buffer_once():
    import vitb200
    from vitb200.ddp import plan_buckets
    buckets, rest = plan_buckets([(100, 300), (300, 520), (600, 900)], 1000)
    assert buckets == [(100, 300), (300, 520), (600, 900)]
    assert rest == [(0, 100), (520, 600), (900, 1000)]
    covered = sorted(buckets + rest)
    assert covered[0][0] == 0 and covered[-1][1] == 1000
    assert all(a[1] == b[0] for a, b in zip(covered, covered[1:]))
    with pytest.raises(ValueError):
        plan_buckets([(0, 10), (5, 20)], 100)
    with pytest.raises(ValueError):
        plan_buckets([(0, 200)], 100)
    assert plan_buckets([], 64) == ([], [(0, 64)])


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import vitb200
        from vitb200.ddp import GradReducer, _BucketTrigger, plan_buckets
        total = 1000
        flat = torch.arange(total, dtype=torch.float32) * (rank + 1)      # rank-dependent "gradients"
        buckets, rest = plan_buckets([(100, 400), (400, 900)], total)
        red = GradReducer(flat, buckets, rest)
        red.start_step()
        # a toy graph: bucket 1's trigger sits downstream of bucket 0's, so it fires first in backward
        x = torch.ones(4, requires_grad=True)
        h0 = _BucketTrigger.apply(x, red, 0) * 2.0
        h1 = _BucketTrigger.apply(h0, red, 1) * 3.0
        h1.sum().backward()
        order = list(red.launched)
        red.finish()
        expect = torch.arange(total, dtype=torch.float32) * (sum(range(1, world + 1)) / world)
        ok = torch.allclose(flat, expect) and order == [1, 0] and torch.allclose(x.grad, torch.full((4,), 6.0))
        # a second step re-arms the reducer
        red.start_step()
        red.finish()
        q.put((rank, bool(ok), order))
    finally:
        dist.destroy_process_group()


def test_grad_reducer_gloo_world_size_2():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res), res
    assert all(order == [1, 0] for _, _, order in res), res
