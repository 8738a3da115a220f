"""Where does the time of the epilogue-bound GEMMs go?  Runs the ViT-B/16 batch-128 fc1 forward (GELU + GELU' outputs), the
fc2 dgrad (x gelu' + column sums), the out-projection and the fc2 forward (fp32 + residual) and the plain bf16 fc1 through the
ABLATION build of the kernel (vitb_gemm_diag) with parts of the epilogue switched off:

    1 no TMA store issue   2 no epilogue math   4 no TMEM load   8 no staging-tile writes   16 no side / bias loads
    32 no per-chunk work at all (the bare mainloop)   64 no async-proxy fence   128 no column sums

Outputs are wrong by construction; only the timings mean something.  Writes gpurun_out/epi_ablate.txt."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vitb200  # noqa: E402
from vitb200 import ops  # noqa: E402

L = vitb200._lib
# the ablation build lives in a separate diagnostics library (vit-of-pytorch_b200/build.py --tools), not in the product .so
import ctypes  # noqa: E402
import importlib.util  # noqa: E402
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


def timeit(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


T, D, M = 25216, 768, 3072
bf = torch.bfloat16
x = torch.randn(T, D, device="cuda").to(bf)
w1 = torch.randn(M, D, device="cuda").to(bf)
a = torch.randn(T, M, device="cuda").to(bf)
w2 = torch.randn(D, M, device="cuda").to(bf)
wo = torch.randn(D, D, device="cuda").to(bf)
o_bf = torch.empty(T, M, device="cuda", dtype=bf)
d2 = torch.empty_like(o_bf)
aux = torch.randn(T, M, device="cuda").to(bf)
b1 = torch.randn(M, device="cuda")
b2 = torch.randn(D, device="cuda")
cs = torch.zeros(M, device="cuda")
res = torch.randn(T, D, device="cuda")
o32 = torch.empty(T, D, device="cuda")
cases = {
    "fc1 fwd gelu+gelu' (packed)": lambda: ops.gemm(x, w1, out=o_bf, bias=b1, epilogue=ops.EPI_GELU_DG, d2=d2),
    "fc1 fwd bias only": lambda: ops.gemm(x, w1, out=o_bf, bias=b1),
    "fc2 dgrad x gelu' + colsum (rowmul)": lambda: ops.gemm(x, w2, b_mn=True, out=o_bf, epilogue=ops.EPI_MUL_AUX, aux=aux, colsum=cs),
    "out-proj f32+res (staged)": lambda: ops.gemm(x, wo, b_mn=True, out=o32, bias=b2, residual=res),
    "fc2 fwd f32+res (staged)": lambda: ops.gemm(a, w2, out=o32, bias=b2, residual=res),
}
masks = [0, 32, 1, 1 | 8, 1 | 8 | 64, 2, 4, 16, 128, 1 | 2 | 8 | 16 | 64 | 128]
lines = []
ops.GEMM_OVERRIDE = _tools.vitb_gemm_diag
try:
    for name, fn in cases.items():
        row = []
        for m in masks:
            assert _tools.vitb_gemm_diag_mask(m) == 0
            row.append("%d:%.4f" % (m, timeit(fn)))
        lines.append("%-38s %s" % (name, "  ".join(row)))
        print(lines[-1], flush=True)
finally:
    _tools.vitb_gemm_diag_mask(0)
    ops.GEMM_OVERRIDE = None
os.makedirs("gpurun_out", exist_ok=True)
with open("gpurun_out/epi_ablate.txt", "w") as fh:
    fh.write("mask:ms per launch (ablation build; mask bits in tools/epi_ablate.py)\n" + "\n".join(lines) + "\n")
