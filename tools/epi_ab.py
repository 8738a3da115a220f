"""A/B of epilogue switches on the ViT-B/16 batch-128 fc1 forward (GELU + GELU' outputs):
VITB_EPI_PACKED = 0 / 1 in one process (the library reads the switch at every call).  Diagnostic only."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vitb200  # noqa: E402


def timeit(fn, iters=30, warm=5):
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


def timeit_graph(fn, reps=20):
    """Small kernels: 20 launches captured in one CUDA graph, so the Python / ctypes launch cost is not what is timed."""
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


T, D, M = 25216, 768, 3072
bf = torch.bfloat16
A = torch.randn(T, D, device="cuda").to(bf)
B = torch.randn(M, D, device="cuda").to(bf)
out = torch.empty(T, M, device="cuda", dtype=bf)
d2 = torch.empty_like(out)
bias = torch.randn(M, device="cuda")
fl = 2.0 * T * M * D
lines = []
for rep in range(2):
    for flag in ("0", "1"):
        os.environ["VITB_EPI_PACKED"] = flag
        ms = timeit(lambda: vitb200.ops.gemm(A, B, out=out, bias=bias, epilogue=vitb200.ops.EPI_GELU_DG, d2=d2))
        lines.append("fc1 gelu+gelu' VITB_EPI_PACKED=%s  %.4f ms  %.0f TF" % (flag, ms, fl / ms / 1e9))
        print(lines[-1], flush=True)
dy = torch.randn(T, D, device="cuda").to(bf)
W2 = torch.randn(D, M, device="cuda").to(bf)
aux = torch.randn(T, M, device="cuda").to(bf)
cs = torch.zeros(M, device="cuda")
for rep in range(2):
    for flag in ("0", "1"):
        os.environ["VITB_EPI_ROWMUL"] = flag
        ms = timeit(lambda: vitb200.ops.gemm(dy, W2, b_mn=True, out=out, epilogue=vitb200.ops.EPI_MUL_AUX, aux=aux, colsum=cs))
        lines.append("fc2 dgrad * gelu' + colsum VITB_EPI_ROWMUL=%s  %.4f ms  %.0f TF" % (flag, ms, fl / ms / 1e9))
        print(lines[-1], flush=True)
a2 = torch.randn(T, M, device="cuda").to(bf)
w2 = torch.randn(D, M, device="cuda").to(bf)
res = torch.randn(T, D, device="cuda")
o32 = torch.empty(T, D, device="cuda")
b2 = torch.randn(D, device="cuda")
wo = torch.randn(D, D, device="cuda").to(bf)
for rep in range(2):
    for flag in ("0", "1"):                      # experimental: fp32 + residual epilogue in the register layout
        os.environ["VITB_EPI_ROWRES"] = flag
        ms = timeit(lambda: vitb200.ops.gemm(a2, w2, out=o32, bias=b2, residual=res))
        lines.append("fc2 fwd f32+res VITB_EPI_ROWRES=%s  %.4f ms  %.0f TF" % (flag, ms, fl / ms / 1e9))
        print(lines[-1], flush=True)
        ms = timeit(lambda: vitb200.ops.gemm(A, wo, b_mn=True, out=o32, bias=b2, residual=res))
        lines.append("out-proj f32+res VITB_EPI_ROWRES=%s  %.4f ms  %.0f TF" % (flag, ms, 2.0 * T * D * D / ms / 1e9))
        print(lines[-1], flush=True)
os.environ["VITB_EPI_ROWRES"] = "0"
x = torch.randn(T, 3 * D, device="cuda").to(bf)
o3 = [torch.zeros(D, device="cuda") for _ in range(3)]
ms = timeit_graph(lambda: vitb200.ops.colsum3(x, *o3))
lines.append("colsum3 [25216, 2304] bf16 (graph)  %.4f ms  %.0f GB/s" % (ms, x.numel() * 2 / ms / 1e6))
print(lines[-1], flush=True)
u8 = torch.randint(0, 256, (128, 32, 32, 3), dtype=torch.uint8, device="cuda")
tf = vitb200.DeviceImageTransform((32, 32), 224, device="cuda")
buf = torch.empty(128, 3, 224, 224, device="cuda")
ms = timeit_graph(lambda: tf(u8, out=buf))
lines.append("image_prep 128 x 32x32 -> 224 fp32 (graph)  %.4f ms  %.0f GB/s written" % (ms, buf.numel() * 4 / ms / 1e6))
print(lines[-1], flush=True)
ms = timeit_graph(lambda: tf.patch_columns(u8, 16))
lines.append("image_prep 128 x 32x32 -> bf16 patch operand (graph)  %.4f ms  %.0f GB/s written" % (ms, buf.numel() * 2 / ms / 1e6))
xl = torch.randn(T, D, device="cuda")
gl, bl = torch.randn(D, device="cuda"), torch.randn(D, device="cuda")
dyl = torch.randn(T, D, device="cuda").to(bf)
drl = torch.randn(T, D, device="cuda")
_, _, _, mean, rstd = vitb200.ops.layernorm_fwd(xl, gl, bl, 1e-5, want_bf16=False)
accl = [torch.zeros(D, device="cuda") for _ in range(3)]
ms = timeit_graph(lambda: vitb200.ops.layernorm_bwd(dyl, xl, mean, rstd, gl, dres=drl, want_f32=True, want_bf16=True,
                                                   dgamma=accl[0], dbeta=accl[1], dcolsum=accl[2]))
lines.append("layernorm_bwd [25216, 768] bf16 dy (graph)  %.4f ms  %.0f GB/s" % (ms, T * D * 16 / ms / 1e6))
print(lines[-1], flush=True)
ms = timeit_graph(lambda: vitb200.ops.layernorm_fwd(xl, gl, bl, 1e-5))
lines.append("layernorm_fwd [25216, 768] (graph)  %.4f ms  %.0f GB/s" % (ms, T * D * 6 / ms / 1e6))
print(lines[-1], flush=True)
os.makedirs("gpurun_out", exist_ok=True)
with open("gpurun_out/epi_ab.txt", "w") as fh:
    fh.write("\n".join(lines) + "\n")
