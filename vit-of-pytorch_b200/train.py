"""The training step of the reference's loop (src/train.py:16-25: zero_grad, forward, CrossEntropy, backward,
optimizer.step) as ONE replayable CUDA graph — the "next" row N1 of SURVEY.md §8(f).

The reference syncs the device three times per step (`.item()` at src/train.py:30-32) and launches every ATen
kernel from Python; here the ~320 kernels of a ViT-B/16 step are captured once and replayed with a single launch,
inputs are copied into static buffers, the loss stays on the device until the caller reads it, and the learning
rate lives in a device scalar so an LR scheduler keeps working across replays.
"""
import torch

from . import functional as F


class InputPrefetcher:
    """Host -> device input staging that overlaps the copy of batch i+1 with the compute of batch i.

    The reference moves each batch to the device at the top of the step (`batch_data.to(device)`, src/train.py:17-18),
    in line with the compute.  Here two device staging slots are filled from PINNED host tensors on a side stream;
    `get()` makes the current stream wait for the oldest slot and copies it (device to device, ~50 us for a
    128 x 3 x 224 x 224 batch) into the buffers the step reads — for a GraphedTrainStep those are its static capture
    buffers (`into=(step.images, step.labels)`), so the replay then starts without a further copy.

        pre = InputPrefetcher(images_dev, labels_dev, into=(step.images, step.labels))
        pre.start(host_images, host_labels)
        for ...:
            x, y = pre.get(); pre.start(next_host_images, next_host_labels); loss = step(x, y)
    """

    def __init__(self, example_images, example_labels, into=None):
        if not example_images.is_cuda:
            raise RuntimeError("InputPrefetcher stages onto a CUDA (B200) device")
        self.stage = [(torch.empty_like(example_images), torch.empty_like(example_labels)) for _ in range(2)]
        self.cur = into if into is not None else (torch.empty_like(example_images), torch.empty_like(example_labels))
        self.stream = torch.cuda.Stream(device=example_images.device)
        self.ready = [torch.cuda.Event(), torch.cuda.Event()]
        self.free = [torch.cuda.Event(), torch.cuda.Event()]
        self._w = self._r = 0
        self._inflight = 0

    def start(self, images_host, labels_host):
        """Begin copying one batch (pinned host tensors) into the next staging slot; returns immediately."""
        if self._inflight >= 2:
            raise RuntimeError("InputPrefetcher: both staging slots are full — call get() first")
        s = self._w
        self._w ^= 1
        self._inflight += 1
        self.stream.wait_event(self.free[s])          # the slot's previous batch has been copied out
        with torch.cuda.stream(self.stream):
            self.stage[s][0].copy_(images_host, non_blocking=True)
            self.stage[s][1].copy_(labels_host, non_blocking=True)
            self.ready[s].record(self.stream)

    def get(self):
        """(images, labels) on the device for the oldest started batch, ordered on the current stream."""
        if self._inflight == 0:
            raise RuntimeError("InputPrefetcher: get() without a started batch")
        s = self._r
        self._r ^= 1
        self._inflight -= 1
        cur = torch.cuda.current_stream()
        cur.wait_event(self.ready[s])
        self.cur[0].copy_(self.stage[s][0], non_blocking=True)
        self.cur[1].copy_(self.stage[s][1], non_blocking=True)
        self.free[s].record(cur)
        return self.cur


class DeviceMetrics:
    """loss / acc@1 / acc@5 of the reference's loops (src/train.py:27-32,53-65: `accuracy(pred, target, topk=(1, 5))` from
    src/utils.py:28-40 and a MetricTracker that averages the PER-BATCH values with equal weight) accumulated in device
    tensors.  The reference reads three scalars back with `.item()` every step, i.e. three device syncs per step;
    `update` only enqueues a few small kernels, and `result()` syncs once, when the caller wants the numbers.

        meter = DeviceMetrics(device); ...; meter.update(loss, logits, labels) each step; meter.result() per epoch
    """

    def __init__(self, device=None, topk=(1, 5)):
        self.topk = tuple(topk)
        self.sums = torch.zeros(1 + len(self.topk), dtype=torch.float64, device=device)
        self.steps = 0

    def reset(self):
        self.sums.zero_()
        self.steps = 0

    @torch.no_grad()
    def update(self, loss, logits, labels):
        maxk = min(max(self.topk), logits.shape[1])
        _, pred = logits.topk(maxk, 1, True, True)                      # the reference's call, ties included
        correct = pred.eq(labels.view(-1, 1))                            # [B, maxk]
        vals = [loss.detach().double().reshape(())]
        for k in self.topk:
            vals.append(correct[:, :min(k, maxk)].double().sum() * (100.0 / labels.shape[0]))
        self.sums += torch.stack(vals)
        self.steps += 1

    def result(self):
        """{'loss', 'acc1', 'acc5'} averaged over the batches seen since reset() — one device-to-host read."""
        avg = (self.sums / max(self.steps, 1)).tolist()
        out = {'loss': avg[0]}
        for k, v in zip(self.topk, avg[1:]):
            out['acc%d' % k] = v
        return out


class GraphedTrainStep:
    def __init__(self, net, optimizer, example_images, example_labels, loss_fn=None, warmup=3, forward_loss=None,
                 data_parallel=False, process_group=None, broadcast=True, exchange=None):
        """forward_loss(net, images, labels) -> scalar loss overrides the default loss_fn(net(images), labels)
        (Res-ViT: `lambda net, x, y: sum_of(net(x, y)[:3])`, res-vit/train.py:30,51-52).

        data_parallel=True (one process per GPU, torch.distributed initialised, BARE module + fused optimizer with one
        flat buffer): the step all-reduces (AVG) the flat fp32 gradient buffer between backward and optimizer.step() ON THE
        CAPTURE STREAM, so the collective is a node of the same graph — the launch mode is the same at N = 1 and N > 1.
        The exchange is not overlapped with the backward pass on purpose: overlapped NCCL kernels take SMs from the
        persistent one-CTA-per-SM GEMM / attention kernels, whose statically scheduled tiles then wait for them (round 1:
        0.90 efficiency at 8 GPUs with per-block overlapped buckets); one 344 MB all-reduce over NVSwitch costs 1 - 1.5 ms.

        exchange=p2p.NvlinkExchange (whose allocate() the optimizer was given as grad_buffer_factory): the exchange is
        vitb_p2p_allreduce — one kernel of ours over NVLink peer memory / NVSwitch multicast — instead of NCCL's all-reduce."""
        self.forward_loss = forward_loss
        if not example_images.is_cuda:
            raise RuntimeError("GraphedTrainStep needs CUDA (B200) tensors")
        self.net, self.opt = net, optimizer
        self.dist, self.group, self.flat_g = None, process_group, None
        self.exchange = exchange if data_parallel else None
        if data_parallel:
            import torch.distributed as dist
            if not dist.is_initialized():
                raise RuntimeError("torch.distributed is not initialised (launch with torchrun)")
            if len(optimizer._flat) != 1:
                raise NotImplementedError("data_parallel expects a fused optimizer with one parameter group")
            fg = optimizer._flat[0]
            self.dist, self.flat_g = dist, fg.flat_g
            if exchange is not None and (exchange.buf is None or exchange.buf.data_ptr() != fg.flat_g.data_ptr()):
                raise ValueError("exchange: the optimizer's gradient buffer is not the exchange's symmetric buffer "
                                 "(construct the optimizer with grad_buffer_factory=exchange.allocate)")
            if broadcast:
                dist.broadcast(fg.flat_p, src=0, group=process_group)
                for p in fg.params:
                    F.SHADOW.attach(p, F.SHADOW.get(p, False)[0])   # re-cast the bf16 shadows from the broadcast masters
        self.loss_fn = loss_fn if loss_fn is not None else F.cross_entropy
        self.images = example_images.clone()
        self.labels = example_labels.clone()
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):          # warm-up outside capture: lazily built state, first_step flag
            for _ in range(warmup):
                self._eager_step()
        cur.wait_stream(side)
        torch.cuda.synchronize()
        from . import _lib
        n0 = _lib.LAUNCHES[0]
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss = self._eager_step()
        self.launches_per_step = _lib.LAUNCHES[0] - n0

    def _eager_step(self):
        self.opt.zero_grad()
        if self.forward_loss is not None:
            loss = self.forward_loss(self.net, self.images, self.labels)
        else:
            self.logits = self.net(self.images)      # static output of the captured step (for DeviceMetrics)
            loss = self.loss_fn(self.logits, self.labels)
        loss.backward()
        if self.exchange is not None:
            self.exchange.all_reduce_avg()
        elif self.dist is not None:
            self.dist.all_reduce(self.flat_g, op=self.dist.ReduceOp.AVG, group=self.group)
        self.opt.step()
        return loss

    def __call__(self, images, labels):
        """Copies the batch into the static buffers (H2D allowed, non-blocking) and replays the step.
        Returns the device-resident loss tensor of this step (read it with .item() only when needed)."""
        if images.data_ptr() != self.images.data_ptr():
            self.images.copy_(images, non_blocking=True)
        if labels.data_ptr() != self.labels.data_ptr():
            self.labels.copy_(labels, non_blocking=True)
        if hasattr(self.opt, "sync_lr"):
            self.opt.sync_lr()
        self.graph.replay()
        return self.loss


class GraphedDataParallelStep:
    """Data-parallel training step as TWO CUDA graphs with ONE eager collective between them:

        graph A: zero_grad + forward + loss + backward      (no collective inside: every rank captures on its own)
        eager  : all_reduce(AVG) of the fused optimizer's flat fp32 gradient buffer on the current stream
        graph B: optimizer step

    `ddp.DataParallel` overlaps per-block all-reduces with the backward pass but has to be launched kernel by kernel from
    Python (capturing its side-stream all-reduces hung an 8-rank run in round 1); with eight ranks sharing the host cores
    that launch path is what limits scaling once the step itself is fast.  Here the gradient exchange is exposed
    (344 MB for ViT-B/16, about 0.5 ms at the measured 725 GB/s all-reduce bus bandwidth, ~3 % of an 18 ms step) and the
    ~300 launches per step are replayed by the GPU front end.  Same averaged-gradient semantics as the wrapper.

    EXPERIMENTAL: written after round 1's GPU budget was spent, exercised only by `bench.py --graph-ddp` so far.
    Pass the BARE module (not wrapped in ddp.DataParallel) and a fused optimizer with one flat buffer.
    """

    def __init__(self, net, optimizer, example_images, example_labels, loss_fn=None, warmup=3, process_group=None,
                 broadcast=True):
        import torch.distributed as dist
        if not example_images.is_cuda:
            raise RuntimeError("GraphedDataParallelStep needs CUDA (B200) tensors")
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised (launch with torchrun)")
        if len(optimizer._flat) != 1:
            raise NotImplementedError("GraphedDataParallelStep expects a fused optimizer with one parameter group")
        self.dist, self.group = dist, process_group
        self.net, self.opt = net, optimizer
        self.loss_fn = loss_fn if loss_fn is not None else F.cross_entropy
        self.images = example_images.clone()
        self.labels = example_labels.clone()
        fg = optimizer._flat[0]
        self.flat_g = fg.flat_g
        if broadcast:
            dist.broadcast(fg.flat_p, src=0, group=process_group)
            for p in fg.params:
                F.SHADOW.attach(p, F.SHADOW.get(p, False)[0])   # re-cast the bf16 shadows from the broadcast masters
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):          # warm-up outside capture: lazily built state, first_step flag
            for _ in range(warmup):
                self._fwd_bwd()
                self._exchange()
                self.opt.step()
        cur.wait_stream(side)
        torch.cuda.synchronize()
        from . import _lib
        n0 = _lib.LAUNCHES[0]
        self.graph_a = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph_a):
            self.loss = self._fwd_bwd()
        self.graph_b = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph_b):
            self.opt.step()
        self.launches_per_step = _lib.LAUNCHES[0] - n0

    def _fwd_bwd(self):
        self.opt.zero_grad()
        loss = self.loss_fn(self.net(self.images), self.labels)
        loss.backward()
        return loss

    def _exchange(self):
        self.dist.all_reduce(self.flat_g, op=self.dist.ReduceOp.AVG, group=self.group)

    def __call__(self, images, labels):
        if images.data_ptr() != self.images.data_ptr():
            self.images.copy_(images, non_blocking=True)
        if labels.data_ptr() != self.labels.data_ptr():
            self.labels.copy_(labels, non_blocking=True)
        if hasattr(self.opt, "sync_lr"):
            self.opt.sync_lr()
        self.graph_a.replay()
        self._exchange()
        self.graph_b.replay()
        return self.loss
