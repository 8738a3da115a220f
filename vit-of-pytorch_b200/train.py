"""The training step of the reference's loop (src/train.py:16-25: zero_grad, forward, CrossEntropy, backward,
optimizer.step) as ONE replayable CUDA graph — the "next" row N1 of SURVEY.md §8(f).

The reference syncs the device three times per step (`.item()` at src/train.py:30-32) and launches every ATen
kernel from Python; here the ~320 kernels of a ViT-B/16 step are captured once and replayed with a single launch,
inputs are copied into static buffers, the loss stays on the device until the caller reads it, and the learning
rate lives in a device scalar so an LR scheduler keeps working across replays.
"""
import torch

from . import functional as F


class GraphedTrainStep:
    def __init__(self, net, optimizer, example_images, example_labels, loss_fn=None, warmup=3, forward_loss=None):
        """forward_loss(net, images, labels) -> scalar loss overrides the default loss_fn(net(images), labels)
        (Res-ViT: `lambda net, x, y: sum_of(net(x, y)[:3])`, res-vit/train.py:30,51-52)."""
        self.forward_loss = forward_loss
        if not example_images.is_cuda:
            raise RuntimeError("GraphedTrainStep needs CUDA (B200) tensors")
        self.net, self.opt = net, optimizer
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
            loss = self.loss_fn(self.net(self.images), self.labels)
        loss.backward()
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
