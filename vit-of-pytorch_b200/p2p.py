"""Gradient exchange of the data-parallel step over NVLink peer memory: vitb_p2p_allreduce (csrc/vitb_p2p.cu), ONE kernel
per rank and step instead of the NCCL all-reduce — replaces nn.DataParallel's gradient reduction, src/train.py:128-129.

The flat fp32 gradient buffer of the fused optimizer lives in symmetric memory: every rank allocates the same size and
maps every peer's buffer (and, on NVSwitch systems, a multicast address that makes the switch do the sum).  Allocation and
handle exchange come from torch.distributed._symmetric_memory (plumbing, like torch's allocator); the kernel is ours.

    xchg = vitb200.p2p.NvlinkExchange()                                   # after init_process_group("nccl")
    opt = vitb200.optim.FusedSGD(model.parameters(), lr=..., grad_buffer_factory=xchg.allocate)
    step = vitb200.train.GraphedTrainStep(model, opt, images, labels, data_parallel=True, exchange=xchg)
"""
import ctypes as C

import torch

from . import _lib as L


class NvlinkExchange:
    def __init__(self, group=None, use_multicast=True):
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("NvlinkExchange: torch.distributed is not initialised (launch with torchrun)")
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        self.use_multicast = use_multicast
        self.buf = self.pad = None
        self.numel = 0

    def allocate(self, numel, device):
        """The gradient buffer ([numel] fp32, zeroed) in symmetric memory; collective: every rank calls it once, same size."""
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        if self.buf is not None:
            raise RuntimeError("NvlinkExchange.allocate: one buffer per exchange object")
        n4 = (int(numel) + 3) // 4 * 4
        self.buf = symm.empty(n4, dtype=torch.float32, device=device)
        self.buf.zero_()
        self._hdl = symm.rendezvous(self.buf, self.group)
        self.pad = symm.empty(int(L.vitb_p2p_pad_words()), dtype=torch.int32, device=device)
        self.pad.zero_()
        self._pad_hdl = symm.rendezvous(self.pad, self.group)
        torch.cuda.synchronize(device)
        dist.barrier(self.group)                 # every rank's pad is zero before anybody signals into it
        W = self.world
        self._bufs = (C.c_void_p * W)(*[int(p) for p in self._hdl.buffer_ptrs])
        self._pads = (C.c_void_p * W)(*[int(p) for p in self._pad_hdl.buffer_ptrs])
        mc = 0
        if self.use_multicast:
            try:
                mc = int(self._hdl.multicast_ptr or 0)
            except Exception:  # noqa: BLE001 - no multicast object on this fabric: peer loads / stores
                mc = 0
        self.multicast = mc
        self.numel = n4
        return self.buf[:int(numel)]

    @property
    def mode(self):
        return "nvls multicast (the switch reduces)" if self.multicast else "peer loads / stores"

    def all_reduce_avg(self):
        """buffer <- mean over ranks, in place, on the current stream (capturable)."""
        if self.buf is None:
            raise RuntimeError("NvlinkExchange: allocate() has not been called (pass grad_buffer_factory=xchg.allocate to the optimizer)")
        L.check(L._vitb_p2p_allreduce(self._bufs, self._pads, C.c_void_p(self.multicast or None), self.rank, self.world,
                                      self.numel, 1.0 / self.world, L.stream_ptr(self.buf.device)), "vitb_p2p_allreduce")
