"""Is the epilogue cost of the K = 768 GEMMs time that fails to overlap, or energy under the power cap?

Each variant of the ViT-B/16 batch-128 fc1 GEMM (ablation build, masks as in tools/epi_ablate.py) and the cuBLAS bf16 GEMM of
the same shape is replayed from a CUDA graph for about SECONDS seconds while a thread samples the SM clock and the board power
through NVML.  If the variants' CYCLE counts (ms x MHz) agree with the short-loop timings while their clocks differ, the
additivity seen in profiles/epi_ablate_r02b.txt is the power cap; if clocks agree, it is a real serialisation.
Writes gpurun_out/power_probe.txt."""
import ctypes
import importlib.util
import os
import statistics
import sys
import threading
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vitb200  # noqa: E402
from vitb200 import ops  # noqa: E402
import pynvml  # noqa: E402

SECONDS = float(os.environ.get("PROBE_SECONDS", "1.5"))
L = vitb200._lib
_spec = importlib.util.spec_from_file_location("_vitb_build", os.path.join(os.path.dirname(L.LIB_PATH), "build.py"))
_b = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_b)
if not os.path.exists(_b.TOOLS_LIB):
    _b.build_tools()
_tools = ctypes.CDLL(_b.TOOLS_LIB)
_tools.vitb_gemm_diag.argtypes = [ctypes.POINTER(L.GemmParams), ctypes.c_void_p]
_tools.vitb_gemm_diag.restype = ctypes.c_int
_tools.vitb_gemm_diag_mask.argtypes = [ctypes.c_int]
_tools.vitb_gemm_diag_mask.restype = ctypes.c_int

pynvml.nvmlInit()
_h = pynvml.nvmlDeviceGetHandleByIndex(0)


class Sampler(threading.Thread):
    def __init__(self):
        super().__init__(daemon=True)
        self.stop = False
        self.mhz, self.watts = [], []

    def run(self):
        while not self.stop:
            self.mhz.append(pynvml.nvmlDeviceGetClockInfo(_h, pynvml.NVML_CLOCK_SM))
            self.watts.append(pynvml.nvmlDeviceGetPowerUsage(_h) / 1000.0)
            time.sleep(0.01)


def probe(fn, per_graph=50):
    """ms per launch, median SM MHz and watts over a replay loop of about SECONDS seconds."""
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for _ in range(per_graph):
                fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        short = e0.elapsed_time(e1) / per_graph                       # cold-ish, a few ms: what epi_ablate.py measures
        reps = max(4, int(SECONDS * 1000.0 / (short * per_graph)))
        smp = Sampler()
        smp.start()
        e0.record()
        for _ in range(reps):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        smp.stop = True
        smp.join()
    ms = e0.elapsed_time(e1) / (reps * per_graph)
    # the first quarter of the samples is the ramp; judge the steady part
    k = len(smp.mhz) // 4
    return short, ms, statistics.median(smp.mhz[k:]), statistics.median(smp.watts[k:])


T, D, M = 25216, 768, 3072
bf = torch.bfloat16
x = torch.randn(T, D, device="cuda").to(bf)
w1 = torch.randn(M, D, device="cuda").to(bf)
b1 = torch.randn(M, device="cuda")
o_bf = torch.empty(T, M, device="cuda", dtype=bf)
d2 = torch.empty_like(o_bf)
w1t = w1.t()

lines = ["%-44s %9s %9s %7s %7s %12s" % ("variant", "short ms", "long ms", "MHz", "W", "Mcycles/launch")]


def row(name, fn):
    short, ms, mhz, w = probe(fn)
    lines.append("%-44s %9.4f %9.4f %7.0f %7.0f %12.1f" % (name, short, ms, mhz, w, ms * mhz / 1000.0))
    print(lines[-1], flush=True)
    time.sleep(0.5)


row("cuBLAS bf16 [T,768]x[768,3072]", lambda: torch.matmul(x, w1t, out=o_bf))
ops.GEMM_OVERRIDE = _tools.vitb_gemm_diag
try:
    for mask, what in [(32, "bare mainloop"), (219, "+ TMEM loads"), (9, "+ math, no staging / stores"),
                       (1, "+ staging writes, no TMA stores"), (0, "full epilogue")]:
        assert _tools.vitb_gemm_diag_mask(mask) == 0
        row("fc1 GELU+GELU' mask %3d %s" % (mask, what),
            lambda: ops.gemm(x, w1, out=o_bf, bias=b1, epilogue=ops.EPI_GELU_DG, d2=d2))
    for mask, what in [(32, "bare mainloop"), (0, "full epilogue")]:
        assert _tools.vitb_gemm_diag_mask(mask) == 0
        row("fc1 bias only       mask %3d %s" % (mask, what), lambda: ops.gemm(x, w1, out=o_bf, bias=b1))
finally:
    _tools.vitb_gemm_diag_mask(0)
    ops.GEMM_OVERRIDE = None
row("product library: fc1 GELU+GELU'", lambda: ops.gemm(x, w1, out=o_bf, bias=b1, epilogue=ops.EPI_GELU_DG, d2=d2))
os.makedirs("gpurun_out", exist_ok=True)
with open("gpurun_out/power_probe.txt", "w") as fh:
    fh.write("\n".join(lines) + "\n")
