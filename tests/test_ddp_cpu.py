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


def _ratio(act, reserve_initials=1):
    """ActiveLoss's statistic (res-vit/model.py:80-82): mean keep-probability over the non-reserved tokens."""
    return act[:, reserve_initials:, :].float().mean()


def _active_loss_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import vitb200
        from vitb200 import functional as VF
        g = torch.Generator().manual_seed(5)
        x = torch.randn(4, 9, 3, generator=g)                       # global batch: 4 images, 9 tokens, 3 dynamic layers
        theta = torch.tensor([0.3, -0.2, 0.1], requires_grad=True)  # the replicated "router parameters"
        shard = x[rank * 2:(rank + 1) * 2]
        # the loss itself is a CUDA kernel (vitb_active_loss); what is checked here on gloo is the host-side exchange it
        # is wrapped in: value = global mean, gradient = this shard's (functional.global_mean_shift)
        assert VF.global_mean_needed(True)
        ratio = _ratio(torch.sigmoid(shard * theta))
        loss = (ratio + VF.global_mean_shift(ratio, True) - 0.4) ** 2
        loss.backward()
        grad = theta.grad.clone()
        dist.all_reduce(grad)                                       # what the data-parallel wrapper does: average
        grad /= world
        local = (_ratio(torch.sigmoid(shard * theta.detach())) - 0.4) ** 2
        q.put((rank, float(loss), grad, float(local)))
    finally:
        dist.destroy_process_group()


def test_active_loss_global_batch_semantics_gloo_world_size_2():
    """ActiveLoss(sync_group=...) reproduces the loss AND, after the wrapper's gradient averaging, the gradient of the global
    batch (res-vit/model.py:80-83 is (batch mean - target)^2, not linear in the batch mean); without it every replica sees
    its own shard."""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_active_loss_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    g = torch.Generator().manual_seed(5)
    x = torch.randn(4, 9, 3, generator=g)
    theta = torch.tensor([0.3, -0.2, 0.1], requires_grad=True)
    ref = (_ratio(torch.sigmoid(x * theta)) - 0.4) ** 2              # single process, whole batch
    ref.backward()
    for rank, loss, grad, local in res:
        assert abs(loss - float(ref)) < 1e-7
        assert torch.allclose(grad, theta.grad, atol=1e-7)
    assert abs(res[0][3] - res[1][3]) > 1e-6                        # the per-replica losses do differ from each other
    assert abs(0.5 * (res[0][3] + res[1][3]) - float(ref)) > 1e-8   # and their average is not the global loss


def test_active_loss_is_a_cuda_kernel_and_stays_local_without_a_process_group():
    """The loss is vitb_active_loss (GPU parity: tests/test_resvit_gpu.py); CPU tensors are refused — there is no CPU path —
    and without an initialised process group sync_group=True means local semantics."""
    import pytest
    from vitb200.resvit import ActiveLoss
    from vitb200 import functional as VF
    assert not VF.global_mean_needed(True) and not VF.global_mean_needed(None)
    with pytest.raises(RuntimeError):
        ActiveLoss(0.6, 1)(torch.rand(3, 7, 2))
