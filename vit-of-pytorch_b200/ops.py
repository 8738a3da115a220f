"""Tensor-level wrappers over the C ABI (no autograd here — see functional.py).

Every function takes torch CUDA tensors, checks layout, and launches on torch's current stream.
Outputs are allocated with torch (the library never allocates device memory).
"""
import ctypes as C

import torch

from . import _lib as L
from ._lib import EPI_GELU, EPI_GELU_BWD, EPI_NONE  # noqa: F401


def _as_list(x):
    return list(x) if isinstance(x, (list, tuple)) else [x]


def _check_2d_rowmajor(t, what):
    if t.dim() != 2 or t.stride(1) != 1:
        raise L.VitbError("%s must be a 2-D tensor with unit inner stride, got shape %s stride %s"
                          % (what, tuple(t.shape), tuple(t.stride())))


def gemm(A, B, *, a_mn=False, b_mn=False, out=None, out_dtype=torch.bfloat16, bias=None,
         row_bias=None, row_bias_group=0, epilogue=EPI_NONE, d2=None, residual=None, aux=None,
         accumulate=False, split_k=0, row_remap_group=0, out_rows=None):
    """D[M,N] = epilogue(sum_i A_i B_i^T) on the tcgen05 GEMM (contract: include/vitb200.h).

    a_mn=False: A_i is [M,K_i]; a_mn=True: A_i is stored [K_i,M].  Same for B with N.
    """
    As, Bs = _as_list(A), _as_list(B)
    if len(As) != len(Bs) or not 1 <= len(As) <= 3:
        raise L.VitbError("gemm: need 1..3 matching (A,B) segments")
    L.require_cuda(*As, *Bs, out, bias, row_bias, d2, residual, aux)
    p = L.GemmParams()
    p.struct_bytes = C.sizeof(L.GemmParams)
    M = N = None
    for i, (a, b) in enumerate(zip(As, Bs)):
        if a.dtype != torch.bfloat16 or b.dtype != torch.bfloat16:
            raise L.VitbError("gemm: operands must be bf16")
        _check_2d_rowmajor(a, "A[%d]" % i)
        _check_2d_rowmajor(b, "B[%d]" % i)
        m, ka = (a.shape[1], a.shape[0]) if a_mn else (a.shape[0], a.shape[1])
        n, kb = (b.shape[1], b.shape[0]) if b_mn else (b.shape[0], b.shape[1])
        if ka != kb:
            raise L.VitbError("gemm: K mismatch in segment %d: %d vs %d" % (i, ka, kb))
        if M is None:
            M, N = m, n
        elif (M, N) != (m, n):
            raise L.VitbError("gemm: segment %d has M,N=%d,%d, expected %d,%d" % (i, m, n, M, N))
        p.A[i] = a.data_ptr()
        p.B[i] = b.data_ptr()
        p.lda[i] = a.stride(0)
        p.ldb[i] = b.stride(0)
        p.K[i] = ka
    p.M, p.N, p.num_segments = M, N, len(As)
    p.a_mn_major, p.b_mn_major = int(a_mn), int(b_mn)
    p.split_k = split_k
    p.epilogue = epilogue
    if out is None:
        rows = out_rows if out_rows is not None else M
        out = torch.empty((rows, N), dtype=out_dtype, device=As[0].device)
    _check_2d_rowmajor(out, "out")
    if out.shape[1] != N:
        raise L.VitbError("gemm: out has %d columns, expected %d" % (out.shape[1], N))
    p.D, p.ldd, p.d_dtype = out.data_ptr(), out.stride(0), L.dtype_code(out)
    p.accumulate = int(accumulate)
    if d2 is not None:
        _check_2d_rowmajor(d2, "d2")
        p.D2, p.ldd2 = d2.data_ptr(), d2.stride(0)
    if bias is not None:
        if bias.dtype != torch.float32 or bias.numel() != N or not bias.is_contiguous():
            raise L.VitbError("gemm: bias must be contiguous fp32 [N]")
        p.bias = bias.data_ptr()
    if row_bias is not None:
        if row_bias.dtype != torch.float32 or not row_bias.is_contiguous() or row_bias.shape[-1] != N:
            raise L.VitbError("gemm: row_bias must be contiguous fp32 [*, N]")
        p.row_bias, p.row_bias_group = row_bias.data_ptr(), row_bias_group
    p.row_remap_group = row_remap_group
    if residual is not None:
        _check_2d_rowmajor(residual, "residual")
        p.residual, p.ldr, p.r_dtype = residual.data_ptr(), residual.stride(0), L.dtype_code(residual)
    if aux is not None:
        _check_2d_rowmajor(aux, "aux")
        if aux.dtype != torch.bfloat16:
            raise L.VitbError("gemm: aux must be bf16")
        p.aux, p.ldaux = aux.data_ptr(), aux.stride(0)
    L.check(L._vitb_gemm(C.byref(p), L.stream_ptr(out.device)), "vitb_gemm")
    return out
