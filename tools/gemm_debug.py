"""Diagnostic for the tcgen05 GEMM on a GPU box: per-major error maps that localise descriptor /
swizzle mistakes.  Writes gpurun_out/gemm_debug.txt.  Not part of the product or the tests."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vitb200  # noqa: E402

out_lines = []


def log(*a):
    s = " ".join(str(x) for x in a)
    print(s, flush=True)
    out_lines.append(s)


def run(M, N, K, a_mn, b_mn, dtype=torch.float32):
    g = torch.Generator(device="cuda").manual_seed(0)
    A = torch.randn((K, M) if a_mn else (M, K), generator=g, device="cuda").to(torch.bfloat16)
    B = torch.randn((K, N) if b_mn else (N, K), generator=g, device="cuda").to(torch.bfloat16)
    Af = A.float().t() if a_mn else A.float()
    Bf = B.float().t() if b_mn else B.float()
    ref = Af @ Bf.t()
    try:
        out = vitb200.ops.gemm(A, B, a_mn=a_mn, b_mn=b_mn, out_dtype=dtype)
        torch.cuda.synchronize()
    except Exception as e:  # noqa: BLE001
        log("EXC", M, N, K, a_mn, b_mn, repr(e))
        return False
    err = (out.float() - ref).abs()
    rel = float((out.float() - ref).norm() / ref.norm())
    log("gemm M=%d N=%d K=%d a_mn=%d b_mn=%d rel=%.3e max=%.3e nan=%d" % (M, N, K, a_mn, b_mn, rel, float(err.max()), int(torch.isnan(out).sum())))
    if rel > 1e-3:
        # error by row%8 x col-block-of-16, to expose swizzle / descriptor stride problems
        e = err[: (M // 8) * 8, : (N // 16) * 16]
        byrow = e.view(-1, 8, e.shape[1]).mean(dim=(0, 2))
        bycol = e.view(e.shape[0], -1, 16).mean(dim=(0, 2))
        log("  err by row%8:", [round(float(x), 3) for x in byrow])
        log("  err by col/16 (first 16):", [round(float(x), 3) for x in bycol[:16]])
        byrow32 = e[: (M // 32) * 32].view(-1, 32, e.shape[1]).mean(dim=(0, 2))
        log("  err by row%32:", [round(float(x), 2) for x in byrow32])
        # does a K-subset explain the output?  (k-advance errors)
        for kk in range(0, min(K, 64), 16):
            part = Af[:, kk:kk + 16] @ Bf[:, kk:kk + 16].t()
            log("  corr with k-slice", kk, float((out.float() * part).sum() / (part.norm() * out.float().norm() + 1e-9)))
    return rel < 1e-3


ok = True
for a_mn in (False, True):
    for b_mn in (False, True):
        for shp in ((128, 256, 16), (128, 256, 64), (128, 128, 64), (256, 512, 128), (512, 768, 768)):
            ok &= run(*shp, a_mn, b_mn)
ok &= run(512, 512, 256, False, False, torch.bfloat16)
log("ALL_OK" if ok else "SOME_FAILED")
os.makedirs("gpurun_out", exist_ok=True)
with open("gpurun_out/gemm_debug.txt", "w") as fh:
    fh.write("\n".join(out_lines) + "\n")
