"""Times the gradient exchange of the data-parallel step alone: NCCL all-reduce (AVG) of a flat fp32 buffer the size of the
model's gradients (ViT-B/16: 85.88 M values = 343.5 MB; ViT-L/16: 1213.6 MB), CUDA events, max over ranks.  Launch with
torchrun --nproc-per-node N.  Prints one line on rank 0: ms per all-reduce, algorithmic and bus bandwidth."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import nvlink_bytes  # noqa: E402

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
world, rank = dist.get_world_size(), dist.get_rank()
for name, n in (("ViT-B/16", 85_875_556), ("ViT-L/16", 303_400_000), ("Res-ViT trainable", 14_955_896)):
    buf = torch.randn(n, device=dev)
    for _ in range(5):
        dist.all_reduce(buf, op=dist.ReduceOp.AVG)
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    nv0 = nvlink_bytes(local) if rank == 0 else None
    e0.record()
    for _ in range(20):
        dist.all_reduce(buf, op=dist.ReduceOp.AVG)
    e1.record()
    torch.cuda.synchronize()
    nv1 = nvlink_bytes(local) if nv0 is not None else None
    ms = torch.tensor([e0.elapsed_time(e1) / 20], device=dev, dtype=torch.float64)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        t = float(ms) / 1e3
        gb = n * 4 / 1e9
        print("all-reduce %s: %.1f MB fp32 on %d GPUs: %.3f ms, algbw %.0f GB/s, busbw %.0f GB/s"
              % (name, gb * 1e3, world, t * 1e3, gb / t, gb / t * 2 * (world - 1) / world), flush=True)
        if nv1 is not None:
            print("    NVLink counters of rank 0 per all-reduce: tx %.1f MB, rx %.1f MB (NVML NVLINK_THROUGHPUT_DATA)"
                  % ((nv1[0] - nv0[0]) / 20 / 1e6, (nv1[1] - nv0[1]) / 20 / 1e6), flush=True)
        else:
            print("    NVLink counters: not available through NVML on this box", flush=True)
    del buf
# the same exchange with vitb_p2p_allreduce: one kernel of ours over NVLink peer memory (NVSwitch multicast when offered)
import vitb200  # noqa: E402
for use_mc in (True, False):
    try:
        n = 85_875_556
        x = vitb200.p2p.NvlinkExchange(use_multicast=use_mc)
        g = x.allocate(n, dev)
        ref = torch.empty_like(g)
        gen = torch.Generator(device=dev).manual_seed(100 + rank)
        src = torch.randn(n, device=dev, generator=gen)
        g.copy_(src)
        ref.copy_(src)
        dist.all_reduce(ref, op=dist.ReduceOp.AVG)
        x.all_reduce_avg()
        torch.cuda.synchronize()
        err = float((g - ref).abs().max() / ref.abs().max())
        for _ in range(3):
            x.all_reduce_avg()
        torch.cuda.synchronize()
        dist.barrier()
        nv0 = nvlink_bytes(local) if rank == 0 else None
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            x.all_reduce_avg()
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / 20], device=dev, dtype=torch.float64)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        # and replayed from a CUDA graph, as the training step does
        gr = torch.cuda.CUDAGraph()
        s2 = torch.cuda.Stream()
        with torch.cuda.stream(s2):
            with torch.cuda.graph(gr, stream=s2):
                x.all_reduce_avg()
            g.copy_(src)
            gr.replay()
            torch.cuda.synchronize()
            err_g = float((g - ref).abs().max() / ref.abs().max())
        if rank == 0:
            t = float(ms) / 1e3
            gb = n * 4 / 1e9
            print("vitb_p2p_allreduce ViT-B/16 (%s, requested multicast=%s): %.1f MB fp32 on %d GPUs: %.3f ms, algbw %.0f GB/s; "
                  "max |diff| vs NCCL %.2e (eager) %.2e (graph replay)" % (x.mode, use_mc, gb * 1e3, world, t * 1e3, gb / t, err, err_g),
                  flush=True)
        del x, g, ref, src
    except Exception as exc:  # noqa: BLE001
        if rank == 0:
            print("vitb_p2p_allreduce (multicast=%s) failed: %r" % (use_mc, exc), flush=True)
    dist.barrier()
dist.barrier()
torch.cuda.synchronize()
os._exit(0)
