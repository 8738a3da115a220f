"""Column-sum kernel at the shapes of the training step (the query-bias gradient: a [T, 768] slice of the packed
[T, 2304] bf16 gradient buffer; the full [T, 768] fp32 / bf16 tensors), swept over the grid-size knob VITB_COLSUM_WARPS
(resident warps per SM the grid is sized for).  Replayed from a CUDA graph behind a GEMM that evicts nothing, so the input
is L2-warm or cold as in the step (it was just written by the previous kernel).  Writes gpurun_out/colsum_bench.txt."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vitb200  # noqa: E402
from vitb200 import ops  # noqa: E402

T, D = 25216, 768
dq = torch.randn(T, 3 * D, device="cuda").to(torch.bfloat16)
x32 = torch.randn(T, D, device="cuda")
xb = torch.randn(T, D, device="cuda").to(torch.bfloat16)
acc = torch.zeros(D, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def graph_ms(body, reps=10):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        body()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for _ in range(reps):
                body()
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (5 * reps)


def timeit(fn, cold):
    """ms per launch inside a CUDA graph (no host launch latency in the number); cold = behind a 256 MB fill whose own
    time is subtracted."""
    if not cold:
        return graph_ms(fn)
    return graph_ms(lambda: (flush.zero_(), fn())) - graph_ms(lambda: flush.zero_())


lines = []
for v4, warps in [(v, w) for v in (0, 1) for w in (64, 48, 32, 24, 16, 8)]:
    os.environ["VITB_COLSUM_WARPS"] = str(warps)
    os.environ["VITB_COLSUM_V4"] = str(v4)
    row = ["v4=%d warps/SM %2d" % (v4, warps)]
    for name, fn, nbytes in (("dq slice bf16", lambda: ops.colsum(dq[:, :D], acc), T * D * 2),
                             ("dense bf16", lambda: ops.colsum(xb, acc), T * D * 2),
                             ("dense fp32", lambda: ops.colsum(x32, acc), T * D * 4)):
        for cold in (False, True):
            ms = timeit(fn, cold)
            row.append("%s %s %.4f ms (%.0f GB/s)" % (name, "cold" if cold else "warm", ms, nbytes / ms / 1e6))
    lines.append(" | ".join(row))
    print(lines[-1], flush=True)
os.makedirs("gpurun_out", exist_ok=True)
open("gpurun_out/colsum_bench.txt", "w").write("\n".join(lines) + "\n")
