"""One-shot GPU check of the round-1 kernel changes (diagnostic; appends to gpurun_out/dev_check.txt).

  * attn_bwd generation 2 (VITB_ATTN_BWD=2) against generation 1 and a torch fp32 reference, plus timing
  * GEMM bf16 epilogues through TMA stores (VITB_GEMM_TMA_STORE=1) against the staged path, plus timing
Every case is wrapped so that one failure does not hide the others.
"""
import math
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


def attn_ref(q, k, v, H):
    B, N, HD = q.shape
    dh = HD // H
    qh = q.float().view(B, N, H, dh).permute(0, 2, 1, 3)
    kh = k.float().view(B, N, H, dh).permute(0, 2, 1, 3)
    vh = v.float().view(B, N, H, dh).permute(0, 2, 1, 3)
    s = qh @ kh.transpose(-1, -2) / math.sqrt(dh)
    return (torch.softmax(s, -1) @ vh).permute(0, 2, 1, 3).reshape(B, N, HD)


def check_attn():
    H, dh = 12, 64
    D = H * dh
    for N in (197, 50, 256, 16, 130, 128, 129):
        try:
            B = 3
            g = torch.Generator(device="cuda").manual_seed(N)
            qkv = torch.randn(B, N, 3 * D, device="cuda", generator=g).to(torch.bfloat16)
            q, k, v = qkv[:, :, :D], qkv[:, :, D:2 * D], qkv[:, :, 2 * D:]
            o, lse = ops.attn_fwd(q, k, v, H)
            do = torch.randn(B, N, D, device="cuda", generator=g).to(torch.bfloat16)
            qf, kf, vf = (t.float().detach().requires_grad_(True) for t in (q, k, v))
            attn_ref(qf, kf, vf, H).backward(do.float())
            res = {}
            for gen in ("1", "2"):
                os.environ["VITB_ATTN_BWD"] = gen
                dqkv = torch.full_like(qkv, float("nan"))
                ops.attn_bwd(do, q, k, v, o, lse, H, dq=dqkv[:, :, :D], dk=dqkv[:, :, D:2 * D], dv=dqkv[:, :, 2 * D:])
                torch.cuda.synchronize()
                res[gen] = dqkv
            e = [rel(res["2"][:, :, i * D:(i + 1) * D], r.grad) for i, r in enumerate((qf, kf, vf))]
            e1 = [rel(res["1"][:, :, i * D:(i + 1) * D], r.grad) for i, r in enumerate((qf, kf, vf))]
            same = rel(res["2"], res["1"])
            ok = all(x < 1.5e-2 for x in e) and bool(torch.isfinite(res["2"].float()).all())
            log("attn_bwd N=%3d  gen2 vs ref dq/dk/dv %.2e %.2e %.2e | gen1 %.2e %.2e %.2e | gen2 vs gen1 %.2e  %s"
                % (N, e[0], e[1], e[2], e1[0], e1[1], e1[2], same, "OK" if ok else "FAIL"))
        except Exception:  # noqa: BLE001
            log("attn_bwd N=%d EXC" % N, traceback.format_exc()[-400:])
    for (B, N, Hh) in ((128, 197, 12), (64, 197, 16), (128, 50, 12)):
        try:
            Dd = Hh * 64
            qkv = (torch.randn(B, N, 3 * Dd, device="cuda") * 0.5).to(torch.bfloat16)
            q, k, v = qkv[:, :, :Dd], qkv[:, :, Dd:2 * Dd], qkv[:, :, 2 * Dd:]
            o, lse = ops.attn_fwd(q, k, v, Hh)
            do = torch.randn_like(o)
            dqkv = torch.empty_like(qkv)
            f = timeit(lambda: ops.attn_fwd(q, k, v, Hh))
            t = {}
            for gen in ("1", "2"):
                os.environ["VITB_ATTN_BWD"] = gen
                t[gen] = timeit(lambda: ops.attn_bwd(do, q, k, v, o, lse, Hh, dq=dqkv[:, :, :Dd], dk=dqkv[:, :, Dd:2 * Dd],
                                                     dv=dqkv[:, :, 2 * Dd:]))
            fl = 4.0 * B * Hh * N * N * 64
            log("attn time B=%d N=%d H=%d: fwd %.3f ms (%.0f TF) bwd gen1 %.3f ms (%.0f TF) gen2 %.3f ms (%.0f TF)"
                % (B, N, Hh, f, fl / f / 1e9, t["1"], 2.5 * fl / t["1"] / 1e9, t["2"], 2.5 * fl / t["2"] / 1e9))
        except Exception:  # noqa: BLE001
            log("attn time EXC", traceback.format_exc()[-400:])


def check_gemm():
    bf = torch.bfloat16
    g = torch.Generator(device="cuda").manual_seed(1)

    def rn(*s, scale=1.0):
        return (torch.randn(*s, device="cuda", generator=g) * scale).to(bf)

    cases = [
        ("plain+bias M=1000 N=768 K=768 (row tail)", 1000, 768, 768, {}),
        ("plain+bias strided out (qkv slice)", 640, 768, 768, {"slice": True}),
        ("gelu+z M=777 N=3072 K=768", 777, 3072, 768, {"gelu": True}),
        ("gelu no z N=200 (column tail, BN=128)", 300, 200, 128, {"gelu": True, "noz": True}),
        ("plain no bias b_mn K=3072", 512, 768, 3072, {"bmn": True, "nobias": True}),
    ]
    for name, M, N, K, kw in cases:
        try:
            A = rn(M, K, scale=0.5)
            Bm = rn(K, N, scale=0.05) if kw.get("bmn") else rn(N, K, scale=0.05)
            bias = None if kw.get("nobias") else torch.randn(N, device="cuda", generator=g)
            outs = {}
            for flag in ("0", "1"):
                os.environ["VITB_GEMM_TMA_STORE"] = flag
                if kw.get("slice"):
                    big = torch.full((M, 3 * N), 7.0, device="cuda", dtype=bf)
                    out = big[:, N:2 * N]
                else:
                    big = None
                    out = torch.full((M, N), float("nan"), device="cuda", dtype=bf)
                args = dict(b_mn=bool(kw.get("bmn")), out=out, bias=bias)
                z = None
                if kw.get("gelu"):
                    z = None if kw.get("noz") else torch.full((M, N), float("nan"), device="cuda", dtype=bf)
                    args.update(epilogue=ops.EPI_GELU, d2=z)
                ops.gemm(A, Bm, **args)
                torch.cuda.synchronize()
                outs[flag] = (out.clone(), None if z is None else z.clone(), None if big is None else big.clone())
            ref = A.float() @ (Bm.float() if kw.get("bmn") else Bm.float().t())
            if bias is not None:
                ref = ref + bias
            zref = ref
            if kw.get("gelu"):
                ref = torch.nn.functional.gelu(ref)
            e_new, e_old = rel(outs["1"][0], ref), rel(outs["0"][0], ref)
            same = float((outs["1"][0].float() - outs["0"][0].float()).abs().max())
            msg = "gemm %-44s new vs ref %.2e | old vs ref %.2e | max|new-old| %.3g" % (name, e_new, e_old, same)
            ok = e_new < 1e-2 and bool(torch.isfinite(outs["1"][0].float()).all())
            if outs["1"][1] is not None:
                ez = rel(outs["1"][1], zref)
                msg += " | z vs ref %.2e" % ez
                ok = ok and ez < 1e-2
            if outs["1"][2] is not None:
                untouched = bool((outs["1"][2][:, :N] == 7.0).all() and (outs["1"][2][:, 2 * N:] == 7.0).all())
                msg += " | neighbours untouched %s" % untouched
                ok = ok and untouched
            log(msg, "OK" if ok else "FAIL")
        except Exception:  # noqa: BLE001
            log("gemm %s EXC" % name, traceback.format_exc()[-400:])
    # timing at the c2 shapes
    T, D, Mh = 25216, 768, 3072
    tcases = [
        ("fwd q    [T,768]x[768,768] bias (b_mn)", (T, D), (D, D), True, {"bias": True}),
        ("fwd fc1  [T,768]x[3072,768]^T bias", (T, D), (Mh, D), False, {"bias": True}),
        ("fwd fc1  gelu no z", (T, D), (Mh, D), False, {"gelu": True, "noz": True}),
        ("fwd fc1  gelu + z", (T, D), (Mh, D), False, {"gelu": True}),
        ("dgrad fc1 [T,3072]x[3072,768] (b_mn)", (T, Mh), (Mh, D), True, {}),
        ("dgrad out [T,768]x[768,768]^T", (T, D), (D, D), False, {}),
        ("dgrad fc2 gelu' + colsum (staged path)", (T, D), (D, Mh), True, {"gelu_bwd": True}),
        ("fwd out  f32+res (staged path)", (T, D), (D, D), True, {"res": True}),
    ]
    for name, ash, bsh, bmn, kw in tcases:
        try:
            A = rn(*ash)
            Bm = rn(*bsh)
            Nn = bsh[1] if bmn else bsh[0]
            K = ash[1]
            odt = torch.float32 if kw.get("res") else bf
            out = torch.empty(ash[0], Nn, device="cuda", dtype=odt)
            bias = torch.randn(Nn, device="cuda")
            args = dict(b_mn=bmn, out=out)
            if kw.get("bias"):
                args.update(bias=bias)
            if kw.get("res"):
                args.update(bias=bias, residual=torch.randn(ash[0], Nn, device="cuda"))
            if kw.get("gelu"):
                args.update(bias=bias, epilogue=ops.EPI_GELU,
                            d2=None if kw.get("noz") else torch.empty(ash[0], Nn, device="cuda", dtype=bf))
            if kw.get("gelu_bwd"):
                args.update(epilogue=ops.EPI_GELU_BWD, aux=rn(ash[0], Nn), colsum=torch.zeros(Nn, device="cuda"))
            t = {}
            for flag in ("0", "1"):
                os.environ["VITB_GEMM_TMA_STORE"] = flag
                t[flag] = timeit(lambda: ops.gemm(A, Bm, **args))
            fl = 2.0 * ash[0] * Nn * K
            log("gemm time %-42s staged %.3f ms %.0f TF | tma-store %.3f ms %.0f TF"
                % (name, t["0"], fl / t["0"] / 1e9, t["1"], fl / t["1"] / 1e9))
        except Exception:  # noqa: BLE001
            log("gemm time %s EXC" % name, traceback.format_exc()[-400:])


which = sys.argv[1:] or ["attn", "gemm"]
try:
    if "attn" in which:
        check_attn()
    if "gemm" in which:
        check_gemm()
finally:
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/dev_check.txt", "a") as fh:
        fh.write("\n".join(lines) + "\n")
