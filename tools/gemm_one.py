"""Runs one GEMM configuration a few times (for ncu captures).  usage: gemm_one.py <case>"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vitb200  # noqa: E402

case = sys.argv[1] if len(sys.argv) > 1 else "fc1_plain"
T, D, M = 25216, 768, 3072
bf = torch.bfloat16
A = torch.randn(T, D, device="cuda").to(bf)
if case in ("fc1_plain", "fc1_gelu"):
    B = torch.randn(M, D, device="cuda").to(bf)
    out = torch.empty(T, M, device="cuda", dtype=bf)
    kw = {}
    if case == "fc1_gelu":
        kw = dict(bias=torch.randn(M, device="cuda"), epilogue=vitb200.ops.EPI_GELU, d2=torch.empty_like(out))
    for _ in range(4):
        vitb200.ops.gemm(A, B, out=out, **kw)
elif case == "out_res":
    B = torch.randn(D, D, device="cuda").to(bf)
    out = torch.empty(T, D, device="cuda")
    res = torch.randn(T, D, device="cuda")
    for _ in range(4):
        vitb200.ops.gemm(A, B, out=out, bias=torch.randn(D, device="cuda"), residual=res)
elif case == "dgrad_fc1":
    A = torch.randn(T, M, device="cuda").to(bf)
    B = torch.randn(M, D, device="cuda").to(bf)
    out = torch.empty(T, D, device="cuda", dtype=bf)
    for _ in range(4):
        vitb200.ops.gemm(A, B, b_mn=True, out=out)
elif case == "fc1_gelu_dg":      # the fused block's forward: D = gelu(z), D2 = gelu'(z), both through TMA stores
    B = torch.randn(M, D, device="cuda").to(bf)
    out = torch.empty(T, M, device="cuda", dtype=bf)
    for _ in range(4):
        vitb200.ops.gemm(A, B, out=out, bias=torch.randn(M, device="cuda"), epilogue=vitb200.ops.EPI_GELU_DG,
                         d2=torch.empty_like(out))
elif case == "dgrad_fc2_mul":    # the fused block's backward: dz = (dy W2) * gelu'(z) + column sums
    B = torch.randn(D, M, device="cuda").to(bf)
    out = torch.empty(T, M, device="cuda", dtype=bf)
    aux = torch.randn(T, M, device="cuda").to(bf)
    cs = torch.zeros(M, device="cuda")
    for _ in range(4):
        vitb200.ops.gemm(A, B, b_mn=True, out=out, epilogue=vitb200.ops.EPI_MUL_AUX, aux=aux, colsum=cs)
elif case == "wgrad_fc1_pair":   # dW1[3072,768] += dz^T hn on CTA pairs
    A = torch.randn(T, M, device="cuda").to(bf)
    B = torch.randn(T, D, device="cuda").to(bf)
    out = torch.zeros(M, D, device="cuda")
    for _ in range(4):
        vitb200.ops.gemm(A, B, a_mn=True, b_mn=True, out=out, accumulate=True)
torch.cuda.synchronize()
print("done", case)
