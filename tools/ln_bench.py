"""LayerNorm kernels at the shapes of the training step ([25216, 768]): ms per launch inside a CUDA graph, behind a 256 MB
fill whose own time is subtracted (cold, as in the step: the previous kernels evicted the inputs) and back to back (warm)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vitb200  # noqa: E402
from vitb200 import ops  # noqa: E402

T, D = 25216, 768
x = torch.randn(T, D, device="cuda")
dyb = torch.randn(T, D, device="cuda").to(torch.bfloat16)
dres = torch.randn(T, D, device="cuda")
dres0 = torch.randn(128, D, device="cuda")
gamma = torch.randn(D, device="cuda")
beta = torch.randn(D, device="cuda")
acc = [torch.zeros(D, device="cuda") for _ in range(3)]
_, xn, _, mean, rstd = ops.layernorm_fwd(x, gamma, beta, 1e-6)
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


def both(fn):
    warm = graph_ms(fn)
    cold = graph_ms(lambda: (flush.zero_(), fn())) - graph_ms(lambda: flush.zero_())
    return warm, cold


cases = {
    "ln_fwd fp32 -> bf16 (116 MB)": (lambda: ops.layernorm_fwd(x, gamma, beta, 1e-6), 116.2),
    "ln_bwd bf16 dy + dres -> f32 + bf16, 3 sums (310 MB)": (lambda: ops.layernorm_bwd(
        dyb, x, mean, rstd, gamma, dres=dres, want_f32=True, want_bf16=True, dgamma=acc[0], dbeta=acc[1], dcolsum=acc[2]), 309.9),
    "ln_bwd same, residual on every 197th row (232 MB)": (lambda: ops.layernorm_bwd(
        dyb, x, mean, rstd, gamma, dres=dres0, dres_every=197, want_f32=True, want_bf16=True, dgamma=acc[0], dbeta=acc[1],
        dcolsum=acc[2]), 232.4),
    "ln_bwd no dx (dgamma / dbeta only)": (lambda: ops.layernorm_bwd(
        dyb, x, mean, rstd, gamma, want_f32=False, dgamma=acc[0], dbeta=acc[1]), 116.2),
}
lines = []
for name, (fn, mb) in cases.items():
    w, c = both(fn)
    lines.append("%-58s warm %.4f ms (%.0f GB/s)   cold %.4f ms (%.0f GB/s)" % (name, w, mb / w, c, mb / c))
    print(lines[-1], flush=True)
os.makedirs("gpurun_out", exist_ok=True)
open("gpurun_out/ln_bench.txt", "w").write("\n".join(lines) + "\n")
