"""Data-parallel training across the GPUs of one box: one process per GPU, NCCL over NVLink/NVSwitch.

Replaces the reference's single-process nn.DataParallel (src/train.py:128-129, src/eval.py:42-43,
src/utils.py:44-54), which re-broadcasts the weights and gathers logits to GPU 0 every step.  Here
weights are replicated once, every rank runs the whole step on its shard of the batch, and the only
exchange is the gradient all-reduce (average) — the same mean-over-global-batch semantics as the
reference's gathered loss when shards are equal.

Gradients live in the fused optimizer's ONE flat fp32 buffer (optim.py).  The buffer is cut into
buckets that follow the execution order (one per encoder block plus the embedding / head remainder);
an identity autograd node at each block's input fires when that block's backward has finished and
launches `all_reduce(AVG)` on the block's slice on a side stream, so communication overlaps the rest
of the backward pass.  `optimizer.step()` waits for the side stream through a step pre-hook.
"""
import torch
import torch.distributed as dist
import torch.nn as nn


def plan_buckets(block_ranges, total):
    """block_ranges: [(start, end)] element ranges of the flat buffer owned by the bucket modules, in
    forward order.  Returns (buckets, rest): buckets in the same order, rest = uncovered ranges."""
    buckets = []
    covered = []
    for s, e in block_ranges:
        if not (0 <= s < e <= total):
            raise ValueError("bad bucket range (%d, %d) for buffer of %d" % (s, e, total))
        buckets.append((s, e))
        covered.append((s, e))
    covered.sort()
    rest, pos = [], 0
    for s, e in covered:
        if s < pos:
            raise ValueError("overlapping bucket ranges")
        if s > pos:
            rest.append((pos, s))
        pos = e
    if pos < total:
        rest.append((pos, total))
    return buckets, rest


class GradReducer:
    """Launches the per-bucket all-reduces.  Device-agnostic (gloo on CPU in the tests, NCCL on GPU)."""

    def __init__(self, flat_g, buckets, rest, group=None):
        self.flat_g, self.buckets, self.rest, self.group = flat_g, buckets, rest, group
        self.cuda = flat_g.is_cuda
        self.stream = torch.cuda.Stream(device=flat_g.device) if self.cuda else None
        self.done = set()
        self.launched = []   # order in which buckets were reduced (observable by tests)
        self._works = []

    def _reduce(self, s, e):
        view = self.flat_g[s:e]
        if self.cuda:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.flat_g.device))
            self.stream.wait_event(ev)
            with torch.cuda.stream(self.stream):
                dist.all_reduce(view, op=dist.ReduceOp.AVG, group=self.group)
        else:
            self._works.append(dist.all_reduce(view, op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def start_step(self):
        self.done.clear()
        self.launched.clear()

    def ready(self, i):
        if i in self.done:
            return
        self.done.add(i)
        self.launched.append(i)
        self._reduce(*self.buckets[i])

    def finish(self):
        """Reduce whatever has not been reduced yet, then make the compute stream wait for all of it."""
        for i in range(len(self.buckets)):
            self.ready(i)
        for s, e in self.rest:
            self._reduce(s, e)
        if self.cuda:
            torch.cuda.current_stream(self.flat_g.device).wait_stream(self.stream)
        else:
            for w in self._works:
                w.wait()
            self._works.clear()
            self.flat_g /= dist.get_world_size(self.group)


class _BucketTrigger(torch.autograd.Function):
    """Identity in forward; its backward runs right after the backward of everything downstream of it,
    i.e. when all parameter gradients of the bucket's module are final."""

    @staticmethod
    def forward(ctx, x, reducer, index):
        ctx.reducer, ctx.index = reducer, index
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        ctx.reducer.ready(ctx.index)
        return g, None, None


class DataParallel(nn.Module):
    """module + fused optimizer -> data-parallel replica.  Call like the module; call optimizer.step()
    as usual (its pre-hook waits for the gradient all-reduce)."""

    def __init__(self, module, optimizer, bucket_modules=None, process_group=None, broadcast=True,
                 global_active_loss=False):
        """global_active_loss: Res-ViT's ActiveLoss is (batch mean - target)^2 (res-vit/model.py:80-83), not linear in
        the batch mean; True makes it the loss of the GLOBAL batch (one scalar all-reduce per step, see
        resvit.ActiveLoss) instead of each replica's own shard."""
        super().__init__()
        self.module = module
        if global_active_loss and hasattr(module, "criterion_active"):
            module.criterion_active.sync_group = process_group if process_group is not None else True
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised (launch with torchrun)")
        if len(optimizer._flat) != 1:
            raise NotImplementedError("DataParallel expects a fused optimizer with one parameter group")
        fg = optimizer._flat[0]
        if broadcast:
            dist.broadcast(fg.flat_p, src=0, group=process_group)
            for p in fg.params:
                from . import functional as F
                F.SHADOW.attach(p, F.SHADOW.get(p, False)[0])  # re-cast shadows from the broadcast masters
        if bucket_modules is None:
            bucket_modules = [m for m in module.modules() if type(m).__name__ in ("EncoderBlock", "TransformerBlock")]
        off = {id(p): (o, o + p.numel()) for p, o in zip(fg.params, fg.offsets)}
        ranges = []
        for m in bucket_modules:
            spans = [off[id(p)] for p in m.parameters() if id(p) in off]
            if spans:
                ranges.append((min(s for s, _ in spans), max(e for _, e in spans)))
        buckets, rest = plan_buckets(ranges, fg.total)
        self.reducer = GradReducer(fg.flat_g, buckets, rest, process_group)
        for i, m in enumerate([m for m in bucket_modules if any(id(p) in off for p in m.parameters())]):
            m.register_forward_pre_hook(self._make_hook(i))
        optimizer.register_step_pre_hook(lambda opt, args, kwargs: self.reducer.finish())

    def _make_hook(self, i):
        def hook(mod, args):
            if not torch.is_grad_enabled() or not args or not isinstance(args[0], torch.Tensor):
                return None
            x = args[0]
            if not x.requires_grad:
                x = x.detach().requires_grad_(True)  # first block: make the trigger part of the graph
            return (_BucketTrigger.apply(x, self.reducer, i),) + tuple(args[1:])
        return hook

    def forward(self, *args, **kwargs):
        self.reducer.start_step()
        return self.module(*args, **kwargs)
