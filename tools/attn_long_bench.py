"""Attention backward at the 384 px token count (577 tokens, head_dim 64: ViT-B/16 and ViT-L/16 fine-tuning at the reference's
default --image-size 384, src/config.py:12): the key-block tcgen05 kernel (vitb_attn_bwd_tc_long) against the fp32 CUDA-core
kernel it replaces for these shapes, and the forward for scale.  CUDA events around 10 launches each."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vitb200  # noqa: E402
from vitb200 import ops  # noqa: E402


def timeit(fn, n=10):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


lines = []
for (B, N, H, dh) in ((32, 577, 12, 64), (16, 577, 16, 64), (64, 257, 12, 64), (32, 257, 16, 80), (8, 730, 16, 80)):
    D = H * dh
    qkv = (torch.randn(B, N, 3 * D, device="cuda") * 0.5).to(torch.bfloat16)
    q, k, v = qkv[:, :, :D], qkv[:, :, D:2 * D], qkv[:, :, 2 * D:]
    do = torch.randn(B, N, D, device="cuda").to(torch.bfloat16)
    o, lse = ops.attn_fwd(q, k, v, H)
    t_f = timeit(lambda: ops.attn_fwd(q, k, v, H))
    t_l = timeit(lambda: ops.attn_bwd(do, q, k, v, o, lse, H))
    try:
        t_s = timeit(lambda: ops.attn_bwd(do, q, k, v, o, lse, H, use_tc=False), n=3)
    except Exception as exc:  # noqa: BLE001
        t_s = float("nan")
        print("simt:", exc)
    flop_b = 10.0 * B * H * N * N * dh
    lines.append("B=%d N=%d H=%d dh=%d: forward %.3f ms | backward tcgen05 key-block kernel %.3f ms (%.0f TFLOP/s) | fp32 CUDA-core "
                 "kernel %.3f ms (%.1fx)" % (B, N, H, dh, t_f, t_l, flop_b / t_l / 1e9, t_s, t_s / t_l))
    print(lines[-1], flush=True)
os.makedirs("gpurun_out", exist_ok=True)
open("gpurun_out/attn_long_bench.txt", "w").write("\n".join(lines) + "\n")
