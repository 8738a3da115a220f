"""Times the tcgen05 attention forward/backward at the ViT-B/16 batch-128 shape.  Diagnostic only."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vitb200  # noqa: E402


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


for (B, N, H) in ((128, 197, 12), (64, 197, 16), (128, 50, 12)):
    D = H * 64
    qkv = (torch.randn(B, N, 3 * D, device="cuda") * 0.5).to(torch.bfloat16)
    q, k, v = qkv[:, :, :D], qkv[:, :, D:2 * D], qkv[:, :, 2 * D:]
    o, lse = vitb200.ops.attn_fwd(q, k, v, H)
    do = torch.randn_like(o)
    dqkv = torch.empty_like(qkv)
    f = timeit(lambda: vitb200.ops.attn_fwd(q, k, v, H))
    b = timeit(lambda: vitb200.ops.attn_bwd(do, q, k, v, o, lse, H, dq=dqkv[:, :, :D], dk=dqkv[:, :, D:2 * D], dv=dqkv[:, :, 2 * D:]))
    fl = 4.0 * B * H * N * N * 64
    print("attn B=%d N=%d H=%d: fwd %.3f ms (%.0f TF)  bwd %.3f ms (%.0f TF)" % (B, N, H, f, fl / f / 1e9, b, 2.5 * fl / b / 1e9), flush=True)
