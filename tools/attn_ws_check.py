"""Step-by-step check of the persistent warp-specialised attention kernels (VITB_ATTN_WS=1) against the one-CTA-per-tile
kernels: forward first, then backward, growing sizes, each printing as it goes (run under `timeout`).  Diagnostic only."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vitb200  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "both"


def rel(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


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


for (B, N, H) in ((1, 16, 1), (1, 128, 1), (1, 197, 1), (2, 197, 3), (3, 50, 12), (2, 256, 3), (40, 130, 12), (128, 197, 12), (64, 197, 16)):
    D = H * 64
    torch.manual_seed(B * 1000 + N)
    qkv = (torch.randn(B, N, 3 * D, device="cuda") * 0.5).to(torch.bfloat16)
    q, k, v = qkv[:, :, :D], qkv[:, :, D:2 * D], qkv[:, :, 2 * D:]
    do = torch.randn(B, N, D, device="cuda").to(torch.bfloat16)
    os.environ["VITB_ATTN_WS"] = "0"
    o0, lse0 = vitb200.ops.attn_fwd(q, k, v, H)
    g0 = torch.empty_like(qkv)
    vitb200.ops.attn_bwd(do, q, k, v, o0, lse0, H, dq=g0[:, :, :D], dk=g0[:, :, D:2 * D], dv=g0[:, :, 2 * D:])
    torch.cuda.synchronize()
    os.environ["VITB_ATTN_WS"] = "1"
    line = "B=%d N=%d H=%d:" % (B, N, H)
    if which in ("fwd", "both"):
        o1, lse1 = vitb200.ops.attn_fwd(q, k, v, H)
        torch.cuda.synchronize()
        line += " fwd o %.2e lse %.2e nan=%d" % (rel(o1, o0), rel(lse1, lse0), int(o1.isnan().any()))
    if which in ("bwd", "both"):
        g1 = torch.full_like(qkv, float("nan"))
        os.environ["VITB_ATTN_WS"] = "0"
        o0, lse0 = vitb200.ops.attn_fwd(q, k, v, H)
        os.environ["VITB_ATTN_WS"] = "1"
        vitb200.ops.attn_bwd(do, q, k, v, o0, lse0, H, dq=g1[:, :, :D], dk=g1[:, :, D:2 * D], dv=g1[:, :, 2 * D:])
        torch.cuda.synchronize()
        line += " bwd dq %.2e dk %.2e dv %.2e nan=%d" % (rel(g1[:, :, :D], g0[:, :, :D]), rel(g1[:, :, D:2 * D], g0[:, :, D:2 * D]),
                                                         rel(g1[:, :, 2 * D:], g0[:, :, 2 * D:]), int(g1.isnan().any()))
    print(line, flush=True)
    if B * H >= 480:
        for flag in ("0", "1"):
            os.environ["VITB_ATTN_WS"] = flag
            msg = "   VITB_ATTN_WS=%s" % flag
            if which in ("fwd", "both"):
                msg += " fwd %.4f ms" % timeit(lambda: vitb200.ops.attn_fwd(q, k, v, H))
            if which in ("bwd", "both"):
                gg = torch.empty_like(qkv)
                msg += " bwd %.4f ms" % timeit(lambda: vitb200.ops.attn_bwd(do, q, k, v, o0, lse0, H, dq=gg[:, :, :D],
                                                                            dk=gg[:, :, D:2 * D], dv=gg[:, :, 2 * D:]))
            print(msg, flush=True)
