"""One-shot GPU check of a kernel under development against the path it would replace (diagnostic; appends to
gpurun_out/dev_check.txt).  Current subject: the CTA-pair (cta_group::2) weight-gradient GEMM, switched with
VITB_GEMM_PAIR=1, against the single-CTA kernel and a torch fp32 reference, plus timing at the ViT-B/16 shapes.
Every case is wrapped so that one failure does not hide the others."""
import os
import sys
import traceback

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vitb200  # noqa: E402

ops = vitb200.ops
lines = []


def log(*a):
    s = " ".join(str(x) for x in a)
    print(s, flush=True)
    lines.append(s)


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


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


def check_pair():
    bf = torch.bfloat16
    g = torch.Generator(device="cuda").manual_seed(1)

    def rn(*s, scale=1.0):
        return (torch.randn(*s, device="cuda", generator=g) * scale).to(bf)

    # (M, N, K): dW[M,N] += A[K,M]^T B[K,N]
    for (M, N, K) in ((256, 256, 64), (256, 256, 640), (512, 256, 300), (768, 768, 2048), (3072, 768, 1000),
                      (768, 3072, 25216), (304, 520, 777)):
        try:
            A = rn(K, M, scale=0.5)
            B = rn(K, N, scale=0.5)
            ref = A.float().t() @ B.float() + 1.0
            outs = {}
            for flag in ("0", "1"):
                os.environ["VITB_GEMM_PAIR"] = flag
                out = torch.ones(M, N, device="cuda")
                ops.gemm(A, B, a_mn=True, b_mn=True, out=out, accumulate=True)
                torch.cuda.synchronize()
                outs[flag] = out
            e1, e0 = rel(outs["1"], ref), rel(outs["0"], ref)
            ok = e1 < 2e-5 and bool(torch.isfinite(outs["1"]).all())
            log("pair wgrad M=%4d N=%4d K=%5d  pair vs ref %.2e | single vs ref %.2e | pair vs single %.2e  %s"
                % (M, N, K, e1, e0, rel(outs["1"], outs["0"]), "OK" if ok else "FAIL"))
        except Exception:  # noqa: BLE001
            log("pair wgrad %d %d %d EXC" % (M, N, K), traceback.format_exc()[-500:])
            return
    T, D, Mh = 25216, 768, 3072
    for name, (M, N) in (("wgrad fc1 dW[3072,768]", (Mh, D)), ("wgrad fc2 dW[768,3072]", (D, Mh)),
                         ("wgrad q   dW[768,768]", (D, D)), ("wgrad qkv dW[768,2304]", (D, 3 * D))):
        try:
            A = rn(T, M)
            B = rn(T, N)
            out = torch.zeros(M, N, device="cuda")
            t = {}
            for flag in ("0", "1"):
                os.environ["VITB_GEMM_PAIR"] = flag
                t[flag] = timeit(lambda: ops.gemm(A, B, a_mn=True, b_mn=True, out=out, accumulate=True))
            tt = timeit(lambda: torch.matmul(A.t(), B))
            fl = 2.0 * M * N * T
            log("pair time %-26s single %.3f ms %.0f TF | pair %.3f ms %.0f TF | cublas %.3f ms %.0f TF"
                % (name, t["0"], fl / t["0"] / 1e9, t["1"], fl / t["1"] / 1e9, tt, fl / tt / 1e9))
        except Exception:  # noqa: BLE001
            log("pair time %s EXC" % name, traceback.format_exc()[-500:])


try:
    check_pair()
finally:
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/dev_check.txt", "a") as fh:
        fh.write("\n".join(lines) + "\n")
