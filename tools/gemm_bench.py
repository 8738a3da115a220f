"""Times the tcgen05 GEMM on the ViT-B/16 batch-128 shapes next to torch.matmul (cuBLAS) on a B200.
Writes gpurun_out/gemm_bench.txt.  Diagnostic only."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vitb200  # noqa: E402

lines = []


def log(*a):
    s = " ".join(str(x) for x in a)
    print(s, flush=True)
    lines.append(s)


def timeit(fn, iters=20, warm=5):
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
cases = [
    # name, (A shape, a_mn), (B shape, b_mn), out dtype, extra kwargs, flops
    ("fwd qkv  [T,768]x[2304,768]^T", (T, D), False, (3 * D, D), False, bf, {}),
    ("fwd out  [T,768]x[768,768]^T f32+res", (T, D), False, (D, D), False, torch.float32, {"res": True}),
    ("fwd fc1  [T,768]x[3072,768]^T plain", (T, D), False, (M, D), False, bf, {}),
    ("fwd fc1  [T,768]x[3072,768]^T bias", (T, D), False, (M, D), False, bf, {"bias": True}),
    ("fwd fc1  [T,768]x[3072,768]^T gelu no z", (T, D), False, (M, D), False, bf, {"gelu": True, "noz": True}),
    ("fwd fc1  [T,768]x[3072,768]^T gelu", (T, D), False, (M, D), False, bf, {"gelu": True}),
    ("fwd fc2  [T,3072]x[768,3072]^T f32+res", (T, M), False, (D, M), False, torch.float32, {"res": True}),
    ("dgrad fc2 [T,768]x[768,3072] gelu'", (T, D), False, (D, M), True, bf, {"gelu_bwd": True}),
    ("dgrad fc1 [T,3072]x[3072,768]", (T, M), False, (M, D), True, bf, {}),
    ("wgrad fc1 dW[3072,768]", (T, M), True, (T, D), True, torch.float32, {"acc": True}),
    ("wgrad fc2 dW[768,3072]", (T, D), True, (T, M), True, torch.float32, {"acc": True}),
    ("wgrad qkv dW[2304,768]", (T, 3 * D), True, (T, D), True, torch.float32, {"acc": True}),
    ("wgrad out dW[768,768]", (T, D), True, (T, D), True, torch.float32, {"acc": True}),
]
for name, ash, amn, bsh, bmn, odt, kw in cases:
    A = torch.randn(ash, device="cuda").to(bf)
    B = torch.randn(bsh, device="cuda").to(bf)
    Mm = ash[1] if amn else ash[0]
    K = ash[0] if amn else ash[1]
    Nn = bsh[1] if bmn else bsh[0]
    out = torch.zeros(Mm, Nn, device="cuda", dtype=odt)
    bias = torch.randn(Nn, device="cuda")
    args = dict(a_mn=amn, b_mn=bmn, out=out)
    if kw.get("res"):
        args.update(bias=bias, residual=torch.randn(Mm, Nn, device="cuda"))
    if kw.get("bias"):
        args.update(bias=bias)
    if kw.get("gelu"):
        args.update(bias=bias, epilogue=vitb200.ops.EPI_GELU,
                    d2=None if kw.get("noz") else torch.empty(Mm, Nn, device="cuda", dtype=bf))
    if kw.get("gelu_bwd"):
        args.update(epilogue=vitb200.ops.EPI_GELU_BWD, aux=torch.randn(Mm, Nn, device="cuda").to(bf))
    if kw.get("acc"):
        args.update(accumulate=True)
    ms = timeit(lambda: vitb200.ops.gemm(A, B, **args))
    Af = A.t() if amn else A
    Bf = B if bmn else B.t()
    ms_t = timeit(lambda: torch.matmul(Af, Bf))
    fl = 2.0 * Mm * Nn * K
    log("%-42s ours %.3f ms %.0f TF | cublas(plain) %.3f ms %.0f TF" % (name, ms, fl / ms / 1e9, ms_t, fl / ms_t / 1e9))

os.makedirs("gpurun_out", exist_ok=True)
with open("gpurun_out/gemm_bench.txt", "w") as fh:
    fh.write("\n".join(lines) + "\n")
