"""Tensor-level wrappers over the C ABI (no autograd here — see functional.py).

Every function takes torch CUDA tensors, checks layout, and launches on torch's current stream.
Outputs are allocated with torch (the library never allocates device memory).
"""
import ctypes as C
import os

import torch

from . import _lib as L
from ._lib import EPI_GELU, EPI_GELU_BWD, EPI_GELU_DG, EPI_MUL_AUX, EPI_NONE  # noqa: F401


# When set to a list, every gemm() call appends (start_event, end_event, flops): bench.py uses it to time
# the dominant kernel live with CUDA events on the launching stream.
PROFILE_GEMM = None


def _as_list(x):
    return list(x) if isinstance(x, (list, tuple)) else [x]


def _check_2d_rowmajor(t, what):
    if t.dim() != 2 or t.stride(1) != 1:
        raise L.VitbError("%s must be a 2-D tensor with unit inner stride, got shape %s stride %s"
                          % (what, tuple(t.shape), tuple(t.stride())))


GEMM_OVERRIDE = None    # tools/epi_ablate.py sets this to the ablation build's entry point (a separate tools library)


def gemm(A, B, *, a_mn=False, b_mn=False, out=None, out_dtype=torch.bfloat16, bias=None,
         row_bias=None, row_bias_group=0, epilogue=EPI_NONE, d2=None, residual=None, aux=None,
         accumulate=False, split_k=0, row_remap_group=0, out_rows=None, colsum=None, m_dev=None):
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
    # column groups (merged q|k|v): B_i given as [G, K, Ng] (b_mn) / [G, Ng, K] — G matrices at a uniform distance
    groups = 1
    if Bs[0].dim() == 3:
        groups = Bs[0].shape[0]
        for b in Bs:
            if b.dim() != 3 or b.shape[0] != groups or b.stride(0) != Bs[0].stride(0) or b.stride(2) != 1:
                raise L.VitbError("gemm: grouped B segments must be [G, ., .] with one group stride and unit inner stride")
        p.n_groups, p.b_group_stride = groups, Bs[0].stride(0)
        Bs = [b[0] for b in Bs]
    for i, (a, b) in enumerate(zip(As, Bs)):
        if a.dtype != torch.bfloat16 or b.dtype != torch.bfloat16:
            raise L.VitbError("gemm: operands must be bf16")
        _check_2d_rowmajor(a, "A[%d]" % i)
        _check_2d_rowmajor(b, "B[%d]" % i)
        m, ka = (a.shape[1], a.shape[0]) if a_mn else (a.shape[0], a.shape[1])
        n, kb = (b.shape[1], b.shape[0]) if b_mn else (b.shape[0], b.shape[1])
        n *= groups
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
    if out.dim() == 3:      # grouped output [G, M, Ng]: the G weight gradients of a merged projection
        if groups == 1 and out.shape[0] > 1 and N % out.shape[0] == 0:
            # B is one [., N] matrix (the packed dq|dk|dv buffer): its column groups are simply N / G apart
            groups = out.shape[0]
            p.n_groups = groups
            p.b_group_stride = (N // groups) if b_mn else (N // groups) * Bs[0].stride(0)
        if out.shape[0] != groups or groups * out.shape[2] != N or out.stride(2) != 1:
            raise L.VitbError("gemm: a grouped output must be [n_groups, M, N / n_groups] with unit inner stride")
        p.d_group_stride = out.stride(0)
        p.D, p.ldd, p.d_dtype = out.data_ptr(), out.stride(1), L.dtype_code(out)
    else:
        _check_2d_rowmajor(out, "out")
        if out.shape[1] != N:
            raise L.VitbError("gemm: out has %d columns, expected %d" % (out.shape[1], N))
        p.D, p.ldd, p.d_dtype = out.data_ptr(), out.stride(0), L.dtype_code(out)
    p.accumulate = int(accumulate)
    if d2 is not None:
        _check_2d_rowmajor(d2, "d2")
        if d2.dtype != out.dtype:
            raise L.VitbError("gemm: d2 must have the dtype of the output")
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
        if aux.dtype != out.dtype:
            raise L.VitbError("gemm: aux must have the dtype of the output")
        p.aux, p.ldaux = aux.data_ptr(), aux.stride(0)
    if colsum is not None:
        if colsum.dtype != torch.float32 or colsum.numel() != N or not colsum.is_contiguous():
            raise L.VitbError("gemm: colsum must be contiguous fp32 [N]")
        p.colsum = colsum.data_ptr()
    if m_dev is not None:       # device scalar: rows of A / D that hold data (Res-ViT token compaction)
        if m_dev.dtype != torch.int32 or m_dev.numel() != 1 or not m_dev.is_cuda:
            raise L.VitbError("gemm: m_dev must be an int32 CUDA scalar")
        p.m_dev = m_dev.data_ptr()
    if PROFILE_GEMM is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        L.check(L._vitb_gemm(C.byref(p), L.stream_ptr(out.device)), "vitb_gemm")
        e1.record()
        ksum = sum(int(p.K[i]) for i in range(len(As)))
        esz = out.element_size()
        nbytes = 2.0 * (M + N) * ksum + M * N * esz * (1 + (d2 is not None) + (aux is not None)) + \
            (M * N * residual.element_size() if residual is not None else 0)
        PROFILE_GEMM.append((e0, e1, 2.0 * M * N * ksum, nbytes))
        return out
    if GEMM_OVERRIDE is not None:      # tools/epi_ablate.py only: the ablation build (results wrong by construction)
        L.check(GEMM_OVERRIDE(C.byref(p), L.stream_ptr(out.device)), "vitb_gemm_diag")
        return out
    L.check(L._vitb_gemm(C.byref(p), L.stream_ptr(out.device)), "vitb_gemm")
    return out


# --------------------------------------------------------------------------------------------------
# LayerNorm
# --------------------------------------------------------------------------------------------------
def layernorm_fwd(x, gamma, beta, eps, *, want_f32=False, want_bf16=True, want_lo=False):
    """x: [rows, D] (row stride free, fp32 or bf16).  Returns (y_f32|None, y_bf16|None, y_lo|None, mean, rstd)."""
    L.require_cuda(x, gamma, beta)
    _check_2d_rowmajor(x, "x")
    rows, D = x.shape
    dev = x.device
    yf = torch.empty((rows, D), dtype=torch.float32, device=dev) if want_f32 else None
    yh = torch.empty((rows, D), dtype=torch.bfloat16, device=dev) if want_bf16 else None
    yl = torch.empty((rows, D), dtype=torch.bfloat16, device=dev) if want_lo else None
    mean = torch.empty(rows, dtype=torch.float32, device=dev)
    rstd = torch.empty(rows, dtype=torch.float32, device=dev)
    L.check(L._vitb_layernorm_fwd(L.ptr(x), L.dtype_code(x), x.stride(0), rows, D, L.ptr(gamma), L.ptr(beta),
                                  float(eps), L.ptr(yf), L.ptr(yh), L.ptr(yl), L.ptr(mean), L.ptr(rstd),
                                  L.stream_ptr(dev)), "vitb_layernorm_fwd")
    return yf, yh, yl, mean, rstd


def layernorm_bwd(dy, x, mean, rstd, gamma, *, dres=None, want_f32=True, want_bf16=False, want_lo=False,
                  dgamma=None, dbeta=None, dcolsum=None, dx_out=None, dres_every=1):
    """dx = LN'(dy) + dres.  dgamma/dbeta/dcolsum ([D] fp32) are accumulated in place when given.  dres_every = n > 1:
    dres has rows / n rows and row i of it belongs to row i * n (the other rows have no residual-branch gradient)."""
    L.require_cuda(dy, x, mean, rstd, gamma, dres)
    _check_2d_rowmajor(x, "x")
    rows, D = x.shape
    if not dy.is_contiguous() or tuple(dy.shape) != (rows, D):
        raise L.VitbError("layernorm_bwd: dy must be contiguous [rows, D]")
    if x.dtype != torch.float32:
        raise L.VitbError("layernorm_bwd: x must be fp32")
    dev = x.device
    dxf = dx_out if dx_out is not None else (torch.empty((rows, D), dtype=torch.float32, device=dev) if want_f32 else None)
    dxh = torch.empty((rows, D), dtype=torch.bfloat16, device=dev) if want_bf16 else None
    dxl = torch.empty((rows, D), dtype=torch.bfloat16, device=dev) if want_lo else None
    if dres is not None:
        _check_2d_rowmajor(dres, "dres")
    if dres is not None and dres_every > 1 and dres.shape[0] * dres_every < rows:
        raise L.VitbError("layernorm_bwd: dres has %d rows, needs %d" % (dres.shape[0], (rows + dres_every - 1) // dres_every))
    L.check(L._vitb_layernorm_bwd_sparse_res(L.ptr(dy), L.dtype_code(dy), L.ptr(x), x.stride(0), L.ptr(mean), L.ptr(rstd),
                                             L.ptr(gamma), rows, D, L.ptr(dres), dres.stride(0) if dres is not None else 0,
                                             int(dres_every), L.ptr(dxf), dxf.stride(0) if dxf is not None else 0,
                                             L.ptr(dxh), L.ptr(dxl), L.ptr(dgamma), L.ptr(dbeta), L.ptr(dcolsum),
                                             L.stream_ptr(dev)),
            "vitb_layernorm_bwd")
    return dxf, dxh, dxl


# --------------------------------------------------------------------------------------------------
# attention
# --------------------------------------------------------------------------------------------------
def _head_strides(t, H, dh, what):
    """t: [B, N, H*dh]-shaped view (last dim contiguous, any row/batch stride)."""
    if t.dim() != 3 or t.stride(2) != 1 or t.shape[2] != H * dh:
        raise L.VitbError("%s must be a [B, N, H*dh] view with unit inner stride" % what)
    return t.stride(0), t.stride(1)


def attn_supported_tc(dh, Nq, Nk, dtype):
    """tcgen05 forward AND backward (head_dim 64, <= 256 tokens)."""
    return dtype == torch.bfloat16 and bool(L.vitb_attn_supported_tc(dh, Nq, Nk))


def attn_bwd_supported_any(dh, Nq, Nk, dtype):
    """tcgen05 backward for these shapes: the <= 256-token kernels or the key-block kernel (head_dim 64, any count)."""
    if dtype != torch.bfloat16:
        return False
    if L.vitb_attn_supported_tc(dh, Nq, Nk):
        return True
    # VITB_ATTN_BWD_LONG=0: these shapes go back to the fp32 CUDA-core backward (A/B runs, profiles/attn_long_bwd_r02.txt)
    return os.environ.get("VITB_ATTN_BWD_LONG", "1") != "0" and bool(L.vitb_attn_bwd_long_supported(dh, Nq, Nk))


def attn_fwd_supported_tc(dh, Nq, Nk, dtype):
    """tcgen05 forward: the above plus wide heads (64 < head_dim <= 128, <= 320 tokens: ViT-H/14)."""
    return dtype == torch.bfloat16 and bool(L.vitb_attn_fwd_supported_tc(dh, Nq, Nk))


def _attn_ws_enabled():
    """VITB_ATTN_WS (read per call): 1 (default) = the persistent warp-specialised attention kernels (measured on B200,
    profiles/attn_ws_r02.txt: forward 0.081 -> 0.051 ms, backward 0.183 -> 0.114 ms at B=128, N=197, H=12);
    0 = one CTA per tile (vitb_attention_tc.cu), kept for A/B runs."""
    return os.environ.get("VITB_ATTN_WS", "1") != "0"


def _attn_params(q, k, v, o, lse, H):
    B, Nq, HD = q.shape
    Nk = k.shape[1]
    dh = HD // H
    p = L.AttnParams()
    p.struct_bytes = C.sizeof(L.AttnParams)
    p.dtype = L.dtype_code(q)
    p.B, p.H, p.Nq, p.Nk, p.head_dim = B, H, Nq, Nk, dh
    p.q, p.k, p.v, p.o, p.lse = q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), lse.data_ptr()
    p.q_batch_stride, p.q_row_stride = _head_strides(q, H, dh, "q")
    p.k_batch_stride, p.k_row_stride = _head_strides(k, H, dh, "k")
    p.v_batch_stride, p.v_row_stride = _head_strides(v, H, dh, "v")
    p.o_batch_stride, p.o_row_stride = _head_strides(o, H, dh, "o")
    return p


def attn_fwd(q, k, v, H, *, use_tc=None):
    """q: [B,Nq,H*dh], k/v: [B,Nk,H*dh] views (same dtype).  Returns (o [B,Nq,H*dh], lse [B,H,Nq])."""
    L.require_cuda(q, k, v)
    B, Nq, HD = q.shape
    dh = HD // H
    if use_tc is None:
        use_tc = attn_fwd_supported_tc(dh, Nq, k.shape[1], q.dtype)
    o = torch.empty((B, Nq, HD), dtype=q.dtype, device=q.device)
    lse = torch.empty((B, H, Nq), dtype=torch.float32, device=q.device)
    p = _attn_params(q, k, v, o, lse, H)
    fn = L._vitb_attn_fwd_tc if use_tc else L._vitb_attn_fwd_simt
    if use_tc and _attn_ws_enabled() and L.vitb_attn_ws_supported(0, dh, Nq, k.shape[1]):
        fn = L._vitb_attn_fwd_ws        # persistent warp-specialised kernel
    L.check(fn(C.byref(p), L.stream_ptr(q.device)), "vitb_attn_fwd")
    return o, lse


def attn_bwd(dout, q, k, v, o, lse, H, *, use_tc=None, dq=None, dk=None, dv=None):
    """Returns (dq, dk, dv).  use_tc (default: whenever a tcgen05 kernel takes the shape — bf16, Nq == Nk, head_dim 64..128):
    bf16 gradients from the persistent kernel (<= 256 tokens, head_dim 64) or the key-block kernel (anything else);
    otherwise the fp32 CUDA-core kernel with fp32 gradients.  dq/dk/dv may be strided views."""
    L.require_cuda(dout, q, k, v, o, lse)
    B, Nq, HD = q.shape
    Nk = k.shape[1]
    dh = HD // H
    if use_tc is None:
        use_tc = attn_bwd_supported_any(dh, Nq, Nk, q.dtype)
    gdt = torch.bfloat16 if use_tc else torch.float32
    if dq is None:
        dq = torch.zeros((B, Nq, HD), dtype=gdt, device=q.device) if not use_tc else torch.empty((B, Nq, HD), dtype=gdt, device=q.device)
    elif not use_tc:
        dq.zero_()
    if dk is None:
        dk = torch.empty((B, Nk, HD), dtype=gdt, device=q.device)
    if dv is None:
        dv = torch.empty((B, Nk, HD), dtype=gdt, device=q.device)
    for t in (dq, dk, dv):
        if t.dtype != gdt:
            raise L.VitbError("attn_bwd: gradient buffers must be %s" % gdt)
    p = _attn_params(q, k, v, o, lse, H)
    p.dout = dout.data_ptr()
    p.do_batch_stride, p.do_row_stride = _head_strides(dout, H, dh, "dout")
    p.dq, p.dk, p.dv = dq.data_ptr(), dk.data_ptr(), dv.data_ptr()
    p.dq_batch_stride, p.dq_row_stride = _head_strides(dq, H, dh, "dq")
    p.dk_batch_stride, p.dk_row_stride = _head_strides(dk, H, dh, "dk")
    p.dv_batch_stride, p.dv_row_stride = _head_strides(dv, H, dh, "dv")
    if use_tc and not L.vitb_attn_supported_tc(dh, Nq, Nk):
        # more than 256 tokens: key blocks on separate CTAs, dQ summed in an fp32 scratch buffer
        dq_acc = torch.zeros((B, Nq, HD), dtype=torch.float32, device=q.device)
        L.check(L._vitb_attn_bwd_tc_long(C.byref(p), L.ptr(dq_acc), L.stream_ptr(q.device)), "vitb_attn_bwd_tc_long")
        return dq, dk, dv
    fn = L._vitb_attn_bwd_tc if use_tc else L._vitb_attn_bwd_simt
    if use_tc and _attn_ws_enabled() and L.vitb_attn_ws_supported(1, dh, Nq, Nk):
        fn = L._vitb_attn_bwd_ws        # persistent warp-specialised kernel
    L.check(fn(C.byref(p), L.stream_ptr(q.device)), "vitb_attn_bwd")
    return dq, dk, dv


def attn_q1_supported(dh, Nk, dtype):
    """Single-query attention with bf16 gradients (the class-token-only last block): bf16, head_dim 64, <= 256 keys."""
    return dtype == torch.bfloat16 and bool(L.vitb_attn_q1_supported(dh, Nk))


def attn_q1_bwd(dout, q, k, v, o, lse, H, *, dq=None, dk=None, dv=None):
    """Backward of attn_fwd for ONE query per image (q, o, dout: [B, 1, H*64] bf16): bf16 dq [B,1,HD] and dk / dv
    [B, Nk, HD] (rows written whole; strided views of a packed buffer are fine)."""
    L.require_cuda(dout, q, k, v, o, lse)
    B, Nq, HD = q.shape
    Nk = k.shape[1]
    dh = HD // H
    dev = q.device
    dq = torch.empty((B, 1, HD), dtype=torch.bfloat16, device=dev) if dq is None else dq
    dk = torch.empty((B, Nk, HD), dtype=torch.bfloat16, device=dev) if dk is None else dk
    dv = torch.empty((B, Nk, HD), dtype=torch.bfloat16, device=dev) if dv is None else dv
    for t in (dq, dk, dv, dout, o):
        if t.dtype != torch.bfloat16:
            raise L.VitbError("attn_q1_bwd: bf16 tensors only")
    p = _attn_params(q, k, v, o, lse, H)
    p.dout = dout.data_ptr()
    p.do_batch_stride, p.do_row_stride = _head_strides(dout, H, dh, "dout")
    p.dq, p.dk, p.dv = dq.data_ptr(), dk.data_ptr(), dv.data_ptr()
    p.dq_batch_stride, p.dq_row_stride = _head_strides(dq, H, dh, "dq")
    p.dk_batch_stride, p.dk_row_stride = _head_strides(dk, H, dh, "dk")
    p.dv_batch_stride, p.dv_row_stride = _head_strides(dv, H, dh, "dv")
    L.check(L._vitb_attn_q1_bwd(C.byref(p), L.stream_ptr(dev)), "vitb_attn_q1_bwd")
    return dq, dk, dv


# --------------------------------------------------------------------------------------------------
# operand preparation / embedding stage
# --------------------------------------------------------------------------------------------------
def cast_split(x, *, want_lo=False, hi=None, lo=None):
    L.require_cuda(x)
    if x.dtype != torch.float32 or not x.is_contiguous():
        raise L.VitbError("cast_split: x must be contiguous fp32")
    if hi is None:
        hi = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    if want_lo and lo is None:
        lo = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    L.check(L._vitb_cast_split(L.ptr(x), x.numel(), L.ptr(hi), L.ptr(lo), L.stream_ptr(x.device)), "vitb_cast_split")
    return hi, lo


def im2col(img, P, *, want_lo=False):
    L.require_cuda(img)
    if img.dtype != torch.float32 or not img.is_contiguous() or img.dim() != 4:
        raise L.VitbError("im2col: img must be contiguous fp32 [B,C,H,W]")
    B, Cin, H, W = img.shape
    K = Cin * P * P
    ldk = (K + 7) // 8 * 8
    rows = B * (H // P) * (W // P)
    hi = torch.empty((rows, ldk), dtype=torch.bfloat16, device=img.device)
    lo = torch.empty((rows, ldk), dtype=torch.bfloat16, device=img.device) if want_lo else None
    L.check(L._vitb_im2col(L.ptr(img), B, Cin, H, W, P, ldk, L.ptr(hi), L.ptr(lo), L.stream_ptr(img.device)), "vitb_im2col")
    return hi, lo


def resize_tables(in_size, out_size):
    """Host tables of one resize axis (Pillow 8-bit bilinear): (ksize, bounds int32 [out,2], coeffs int32 [out,ksize])
    as CPU tensors — vitb_resize_tables_host is host code inside the library, no device work."""
    import ctypes as C
    k = C.c_int(0)
    L.check_host(L._vitb_resize_tables_host(in_size, out_size, None, None, 0, C.byref(k)), "vitb_resize_tables_host")
    bounds = torch.zeros((out_size, 2), dtype=torch.int32)
    coeffs = torch.zeros((out_size, k.value), dtype=torch.int32)
    L.check_host(L._vitb_resize_tables_host(in_size, out_size, C.c_void_p(bounds.data_ptr()), C.c_void_p(coeffs.data_ptr()),
                                            coeffs.numel(), C.byref(k)), "vitb_resize_tables_host")
    return k.value, bounds, coeffs


def image_prep(src, out_hw, xtab, ytab, lut, *, flip=None, want_img=True, P=None, want_lo=False, want_u8=False,
               out_img=None, cols=None):
    """src uint8 [B,H,W,C] (device) -> (img fp32 [B,C,oh,ow] | None, hi | None, lo | None, u8 | None).

    xtab / ytab: (bounds, coeffs) device int32 tensors from resize_tables, or None when that axis keeps its size.
    cols=(hi, lo): persistent, caller-zeroed patch-operand buffers to write into (their padding columns stay zero)."""
    L.require_cuda(src, lut, flip)
    if src.dtype != torch.uint8 or src.dim() != 4 or not src.is_contiguous():
        raise L.VitbError("image_prep: src must be contiguous uint8 [B,H,W,C]")
    B, H, W, Cn = src.shape
    oh, ow = out_hw
    if lut.dtype != torch.float32 or tuple(lut.shape) != (Cn, 256) or not lut.is_contiguous():
        raise L.VitbError("image_prep: lut must be contiguous fp32 [C,256]")
    if flip is not None and (flip.dtype != torch.uint8 or flip.numel() != B or not flip.is_contiguous()):
        raise L.VitbError("image_prep: flip must be contiguous uint8 [B]")
    for name, tab, n_out in (("x", xtab, ow), ("y", ytab, oh)):
        if tab is not None:
            b, c = tab
            L.require_cuda(b, c)
            if (b.dtype != torch.int32 or c.dtype != torch.int32 or tuple(b.shape) != (n_out, 2) or c.shape[0] != n_out
                    or not b.is_contiguous() or not c.is_contiguous()):
                raise L.VitbError("image_prep: %s tables must be contiguous int32 [%d,2] / [%d,k]" % (name, n_out, n_out))
    dev = src.device
    img = None
    if want_img:
        img = out_img if out_img is not None else torch.empty((B, Cn, oh, ow), dtype=torch.float32, device=dev)
        if img.dtype != torch.float32 or tuple(img.shape) != (B, Cn, oh, ow) or not img.is_contiguous():
            raise L.VitbError("image_prep: out_img must be contiguous fp32 [B,C,oh,ow]")
    hi = lo = None
    ldk = 0
    if P:
        K = Cn * P * P
        ldk = (K + 7) // 8 * 8
        rows = B * (oh // P) * (ow // P)
        if cols is not None:
            hi, lo = cols
            if tuple(hi.shape) != (rows, ldk) or hi.dtype != torch.bfloat16 or (want_lo and lo is None):
                raise L.VitbError("image_prep: cols buffers must be bf16 [%d,%d]" % (rows, ldk))
        else:
            alloc = torch.zeros if ldk != K else torch.empty
            hi = alloc((rows, ldk), dtype=torch.bfloat16, device=dev)
            lo = alloc((rows, ldk), dtype=torch.bfloat16, device=dev) if want_lo else None
        if not want_lo:
            lo = None
    u8 = torch.empty((B, oh, ow, Cn), dtype=torch.uint8, device=dev) if want_u8 else None
    xb, xc = xtab if xtab is not None else (None, None)
    yb, yc = ytab if ytab is not None else (None, None)
    L.check(L._vitb_image_prep(L.ptr(src), B, H, W, Cn, oh, ow, L.ptr(xb), L.ptr(xc), xc.shape[1] if xc is not None else 0,
                               L.ptr(yb), L.ptr(yc), yc.shape[1] if yc is not None else 0, L.ptr(flip), L.ptr(lut),
                               L.ptr(img), P or 0, ldk, L.ptr(hi), L.ptr(lo), L.ptr(u8), L.stream_ptr(dev)),
            "vitb_image_prep")
    return img, hi, lo, u8


def cls_rows(x, cls, pos):
    """x: [B,N,D] fp32 contiguous; writes x[:,0,:] = cls + pos[0]."""
    L.require_cuda(x, cls, pos)
    B, N, D = x.shape
    L.check(L._vitb_cls_rows(L.ptr(x), B, N, D, L.ptr(cls), L.ptr(pos), L.stream_ptr(x.device)), "vitb_cls_rows")


def embed_bwd(dx, *, dpos=None, dcls=None, dbias=None, want_patch=True, want_lo=False):
    L.require_cuda(dx)
    if dx.dtype != torch.float32 or not dx.is_contiguous() or dx.dim() != 3:
        raise L.VitbError("embed_bwd: dx must be contiguous fp32 [B,N,D]")
    B, N, D = dx.shape
    hi = torch.empty((B * (N - 1), D), dtype=torch.bfloat16, device=dx.device) if want_patch else None
    lo = torch.empty((B * (N - 1), D), dtype=torch.bfloat16, device=dx.device) if (want_patch and want_lo) else None
    L.check(L._vitb_embed_bwd(L.ptr(dx), B, N, D, L.ptr(dpos), L.ptr(dcls), L.ptr(dbias), L.ptr(hi), L.ptr(lo),
                              L.stream_ptr(dx.device)), "vitb_embed_bwd")
    return hi, lo


def gelu_bwd(dy, z):
    """dy * gelu'(z) elementwise (same dtype, contiguous)."""
    L.require_cuda(dy, z)
    if dy.dtype != z.dtype or not dy.is_contiguous() or not z.is_contiguous() or dy.shape != z.shape:
        raise L.VitbError("gelu_bwd: dy and z must be contiguous with equal shape and dtype")
    out = torch.empty_like(dy)
    L.check(L._vitb_gelu_bwd(L.ptr(dy), L.ptr(z), L.ptr(out), dy.numel(), L.dtype_code(dy), L.stream_ptr(dy.device)),
            "vitb_gelu_bwd")
    return out


def dropout_fwd(x, p, seed, state, *, residual=None):
    """(y, mask): y = [residual +] keep * x / (1 - p).  x: contiguous fp32 / bf16; residual: contiguous fp32 of x's shape;
    state: uint64-as-int64 CUDA tensor [2], zero-initialised once per dropout site (the draw counter lives there)."""
    L.require_cuda(x, residual, state)
    if not x.is_contiguous() or (residual is not None and (not residual.is_contiguous() or residual.dtype != torch.float32
                                                           or residual.shape != x.shape)):
        raise L.VitbError("dropout_fwd: x (and the fp32 residual of the same shape) must be contiguous")
    if state.dtype != torch.int64 or state.numel() != 2:
        raise L.VitbError("dropout_fwd: state must be an int64 CUDA tensor with two elements")
    y = torch.empty(x.shape, dtype=torch.float32 if residual is not None else x.dtype, device=x.device)
    mask = torch.empty(x.shape, dtype=torch.uint8, device=x.device)
    L.check(L._vitb_dropout_fwd(L.ptr(x), L.dtype_code(x), L.ptr(residual), L.ptr(y), L.ptr(mask), x.numel(), float(p),
                                int(seed) & 0xFFFFFFFFFFFFFFFF, L.ptr(state), L.stream_ptr(x.device)), "vitb_dropout_fwd")
    return y, mask


def dropout_bwd(dy, mask, p, out_dtype):
    """dx = dy * mask / (1 - p) in `out_dtype`."""
    L.require_cuda(dy, mask)
    if not dy.is_contiguous() or dy.shape != mask.shape:
        dy = dy.contiguous()
    dx = torch.empty(dy.shape, dtype=out_dtype, device=dy.device)
    L.check(L._vitb_dropout_bwd(L.ptr(dy), L.dtype_code(dy), L.ptr(mask), L.ptr(dx), L.dtype_code(dx), dy.numel(), float(p),
                                L.stream_ptr(dy.device)), "vitb_dropout_bwd")
    return dx


def colsum(x, out):
    """out[c] += sum_r x[r,c]."""
    L.require_cuda(x, out)
    _check_2d_rowmajor(x, "x")
    L.check(L._vitb_colsum(L.ptr(x), L.dtype_code(x), x.shape[0], x.shape[1], x.stride(0), L.ptr(out),
                           L.stream_ptr(x.device)), "vitb_colsum")
    return out


def colsum3(x, out0, out1, out2):
    """x: packed [rows, 3*S]; out_i[S] += column sums of segment i (q/k/v bias gradients in one pass)."""
    L.require_cuda(x, out0, out1, out2)
    _check_2d_rowmajor(x, "x")
    S = x.shape[1] // 3
    L.check(L._vitb_colsum3(L.ptr(x), L.dtype_code(x), x.shape[0], S, x.stride(0), L.ptr(out0), L.ptr(out1), L.ptr(out2),
                            L.stream_ptr(x.device)), "vitb_colsum3")


def cross_entropy(logits, labels, *, want_grad=True):
    L.require_cuda(logits, labels)
    if logits.dtype != torch.float32 or not logits.is_contiguous() or labels.dtype != torch.int64:
        raise L.VitbError("cross_entropy: logits fp32 contiguous [B,C], labels int64 [B]")
    B, Ccls = logits.shape
    loss = torch.empty((), dtype=torch.float32, device=logits.device)
    dl = torch.empty_like(logits) if want_grad else None
    L.check(L._vitb_cross_entropy(L.ptr(logits), L.ptr(labels), B, Ccls, L.ptr(loss), L.ptr(dl),
                                  L.stream_ptr(logits.device)), "vitb_cross_entropy")
    return loss, dl


def _check_hyper(h):
    if h is not None and (h.dtype != torch.float32 or h.numel() != 4 or not h.is_contiguous() or not h.is_cuda):
        raise L.VitbError("hyper_dev must be a contiguous fp32 CUDA tensor of 4 values")


def sgd_momentum(p, g, m, lr, momentum, *, dampening=0.0, weight_decay=0.0, nesterov=False, first_step=False,
                 shadow_hi=None, shadow_lo=None, hyper_dev=None):
    """hyper_dev: optional device tensor of 4 floats {lr, momentum, dampening, weight_decay} overriding the scalars."""
    L.require_cuda(p, g, m)
    _check_hyper(hyper_dev)
    L.check(L._vitb_sgd_momentum(L.ptr(p), L.ptr(g), L.ptr(m), p.numel(), float(lr), L.ptr(hyper_dev), float(momentum),
                                 float(dampening), float(weight_decay), int(nesterov), int(first_step),
                                 L.ptr(shadow_hi), L.ptr(shadow_lo), L.stream_ptr(p.device)), "vitb_sgd_momentum")


def adamw(p, g, m, v, lr, beta1, beta2, eps, weight_decay, step, *, grad_scale=None, shadow_hi=None, shadow_lo=None,
          hyper_dev=None, step_dev=None):
    """hyper_dev: optional device tensor of 4 floats {lr, beta1, beta2, weight_decay} overriding the scalars."""
    L.require_cuda(p, g, m, v)
    _check_hyper(hyper_dev)
    if step_dev is not None and step_dev.dtype != torch.int32:
        raise L.VitbError("adamw: step_dev must be an int32 device scalar")
    L.check(L._vitb_adamw(L.ptr(p), L.ptr(g), L.ptr(m), L.ptr(v), p.numel(), float(lr), L.ptr(hyper_dev), float(beta1),
                          float(beta2), float(eps), float(weight_decay), int(step), L.ptr(step_dev), L.ptr(grad_scale),
                          L.ptr(shadow_hi), L.ptr(shadow_lo), L.stream_ptr(p.device)), "vitb_adamw")


def adamw_segments(p, g, m, v, eps, hyper_dev, seg_end, seg_flag, flags, seg_step, *, grad_scale=None, shadow_hi=None,
                   shadow_lo=None):
    """AdamW over a flat buffer whose parameters can be skipped one by one (torch's `grad is None`): seg_end int64 [S]
    ascending end offsets, seg_flag int32 [S] (index into `flags`, -1 = always live), flags int32 [F] (non-zero = got a
    gradient this step), seg_step fp32 [S] per-parameter step counts (advanced here for the live ones)."""
    L.require_cuda(p, g, m, v, seg_end, seg_flag, flags, seg_step)
    _check_hyper(hyper_dev)
    if hyper_dev is None:
        raise L.VitbError("adamw_segments: hyper_dev is required")
    if (seg_end.dtype != torch.int64 or seg_flag.dtype != torch.int32 or flags.dtype != torch.int32 or
            seg_step.dtype != torch.float32 or seg_flag.numel() != seg_end.numel() or seg_step.numel() != seg_end.numel()):
        raise L.VitbError("adamw_segments: seg_end int64 [S], seg_flag int32 [S], flags int32 [F], seg_step fp32 [S]")
    L.check(L._vitb_adamw_segments(L.ptr(p), L.ptr(g), L.ptr(m), L.ptr(v), p.numel(), L.ptr(hyper_dev), float(eps),
                                   L.ptr(grad_scale), L.ptr(shadow_hi), L.ptr(shadow_lo), L.ptr(seg_end), L.ptr(seg_flag),
                                   L.ptr(flags), L.ptr(seg_step), seg_end.numel(), L.stream_ptr(p.device)),
            "vitb_adamw_segments")


def sumsq(x, out):
    L.require_cuda(x, out)
    L.check(L._vitb_sumsq(L.ptr(x), x.numel(), L.ptr(out), L.stream_ptr(x.device)), "vitb_sumsq")


def clip_coef(sumsq_t, max_norm, coef, norm_out=None):
    L.check(L._vitb_clip_coef(L.ptr(sumsq_t), float(max_norm), L.ptr(coef), L.ptr(norm_out),
                              L.stream_ptr(coef.device)), "vitb_clip_coef")


# --------------------------------------------------------------------------------------------------
# Res-ViT token compaction (inference): the row list and its length never leave the device
# --------------------------------------------------------------------------------------------------
def _member_mask(member_ids):
    mask = 0
    for i in member_ids:
        mask |= 1 << int(i)
    return mask


def compact_rows(index, member_ids):
    """index [..., 1] fp32 packed router indices -> (rows int32 [T], count int32 [1]): the rows t whose index is in
    member_ids, in unspecified order; only rows[:count] is meaningful."""
    L.require_cuda(index)
    idx = index.detach().reshape(-1).float().contiguous()
    T = idx.numel()
    rows = torch.empty(T, dtype=torch.int32, device=idx.device)
    count = torch.zeros(1, dtype=torch.int32, device=idx.device)
    L.check(L._vitb_compact_rows(L.ptr(idx), _member_mask(member_ids), T, L.ptr(rows), L.ptr(count), L.stream_ptr(idx.device)),
            "vitb_compact_rows")
    return rows, count


def gather_rows(src, rows, count):
    """dst[i] = src[rows[i]] for i < count; dst has src's capacity ([T, cols]), rows >= count are left uninitialised."""
    L.require_cuda(src, rows, count)
    _check_2d_rowmajor(src, "src")
    dst = torch.empty((rows.numel(), src.shape[1]), dtype=src.dtype, device=src.device)
    L.check(L._vitb_gather_rows(L.ptr(src), src.stride(0), L.dtype_code(src), L.ptr(rows), L.ptr(count), rows.numel(),
                                src.shape[1], L.ptr(dst), dst.stride(0), L.stream_ptr(src.device)), "vitb_gather_rows")
    return dst


def scatter_rows(src, rows, count, dst):
    """dst[rows[i]] = src[i] for i < count (in place)."""
    L.require_cuda(src, rows, count, dst)
    _check_2d_rowmajor(src, "src")
    _check_2d_rowmajor(dst, "dst")
    if src.dtype != dst.dtype or src.shape[1] != dst.shape[1]:
        raise L.VitbError("scatter_rows: src / dst must share dtype and width")
    L.check(L._vitb_scatter_rows(L.ptr(src), src.stride(0), L.dtype_code(src), L.ptr(rows), L.ptr(count), rows.numel(),
                                 src.shape[1], L.ptr(dst), dst.stride(0), L.stream_ptr(src.device)), "vitb_scatter_rows")
    return dst


# --------------------------------------------------------------------------------------------------
# Res-ViT scalar losses
# --------------------------------------------------------------------------------------------------
def distill_loss(student, teacher, loss_acc, want_grad):
    """*loss_acc += mean((student - teacher)^2) over two [rows, cols] (row-strided) slices of one dtype; returns the
    gradient with respect to `student` ([rows, cols] fp32) when want_grad."""
    L.require_cuda(student, teacher, loss_acc)
    if student.dim() != 2 or student.shape != teacher.shape or student.stride(1) != 1 or teacher.stride(1) != 1 or \
            student.dtype != teacher.dtype:
        raise L.VitbError("distill_loss: student / teacher must be [rows, cols] views of one dtype with unit inner stride")
    rows, cols = student.shape
    ds = torch.empty((rows, cols), dtype=torch.float32, device=student.device) if want_grad else None
    L.check(L._vitb_distill_loss(L.ptr(student), student.stride(0), L.ptr(teacher), teacher.stride(0), L.dtype_code(student),
                                 rows, cols, L.ptr(loss_acc), L.ptr(ds), L.stream_ptr(student.device)), "vitb_distill_loss")
    return ds


def active_loss(probs, reserve_initials, target, *, shift=None, want_grad=False):
    """probs [B, N, L] fp32 -> (ratio scalar, loss scalar, d_probs or None); see include/vitb200.h."""
    L.require_cuda(probs, shift)
    if probs.dim() != 3 or probs.dtype != torch.float32 or not probs.is_contiguous():
        raise L.VitbError("active_loss: probs must be contiguous fp32 [B, N, L]")
    B, N, Lr = probs.shape
    ratio = torch.empty((), dtype=torch.float32, device=probs.device)
    loss = torch.empty((), dtype=torch.float32, device=probs.device)
    dp = torch.empty_like(probs) if want_grad else None
    L.check(L._vitb_active_loss(L.ptr(probs), B, N, Lr, int(reserve_initials), float(target), L.ptr(shift), L.ptr(ratio),
                                L.ptr(loss), L.ptr(dp), L.stream_ptr(probs.device)), "vitb_active_loss")
    return ratio, loss, dp


# --------------------------------------------------------------------------------------------------
# Res-ViT routing
# --------------------------------------------------------------------------------------------------
def router_decide_fwd(logits, noise, N, block_size, reserve_initials, training, tau=1.0):
    """logits [T, bs, 2] fp32 contiguous -> (soft, hard, ysoft|None, indices [T], entropy_sum scalar)."""
    L.require_cuda(logits, noise)
    if logits.dtype != torch.float32 or not logits.is_contiguous():
        raise L.VitbError("router_decide_fwd: logits must be contiguous fp32")
    T = logits.numel() // (2 * block_size)
    dev = logits.device
    soft = torch.empty_like(logits)
    hard = torch.empty_like(logits)
    ysoft = torch.empty_like(logits) if training else None
    idx = torch.empty(T, dtype=torch.float32, device=dev)
    ent = torch.zeros((), dtype=torch.float32, device=dev)
    if training and (noise is None or noise.dtype != torch.float32 or not noise.is_contiguous() or noise.numel() != logits.numel()):
        raise L.VitbError("router_decide_fwd: training needs fp32 noise shaped like logits")
    L.check(L._vitb_router_decide_fwd(L.ptr(logits), L.ptr(noise), T, N, block_size, reserve_initials, int(training),
                                      float(tau), L.ptr(soft), L.ptr(hard), L.ptr(ysoft), L.ptr(idx), L.ptr(ent),
                                      L.stream_ptr(dev)), "vitb_router_decide_fwd")
    return soft, hard, ysoft, idx, ent


def router_decide_bwd(soft, ysoft, d_soft, d_hard, d_entropy, entropy_scale, N, block_size, reserve_initials, training,
                      tau=1.0):
    T = soft.numel() // (2 * block_size)
    dl = torch.empty_like(soft)
    L.check(L._vitb_router_decide_bwd(L.ptr(soft), L.ptr(ysoft), L.ptr(d_soft), L.ptr(d_hard), L.ptr(d_entropy),
                                      float(entropy_scale), T, N, block_size, reserve_initials, int(training), float(tau),
                                      L.ptr(dl), L.stream_ptr(soft.device)), "vitb_router_decide_bwd")
    return dl


def token_mean_fwd(x, reserve_initials):
    """x [B,N,C] (f32/bf16 contiguous) -> [B,C] fp32 mean over tokens n >= reserve_initials."""
    L.require_cuda(x)
    B, N, Cc = x.shape
    out = torch.empty((B, Cc), dtype=torch.float32, device=x.device)
    L.check(L._vitb_token_mean_fwd(L.ptr(x), L.dtype_code(x), B, N, Cc, reserve_initials, L.ptr(out),
                                   L.stream_ptr(x.device)), "vitb_token_mean_fwd")
    return out


def token_mean_bwd(dg, B, N, reserve_initials, dtype):
    Cc = dg.shape[-1]
    dx = torch.empty((B, N, Cc), dtype=dtype, device=dg.device)
    L.check(L._vitb_token_mean_bwd(L.ptr(dg), L.dtype_code(dx), B, N, Cc, reserve_initials, L.ptr(dx),
                                   L.stream_ptr(dg.device)), "vitb_token_mean_bwd")
    return dx


def select_rows(a, b, index, member_mask, any_flag=None):
    """out[t,:] = member(index[t]) ? a[t,:] : b[t,:] on 2-D contiguous tensors; a or b may be None (zeros).  any_flag
    (int32 CUDA scalar, optional) is set to 1 when at least one row is a member (never cleared here)."""
    ref = a if a is not None else b
    L.require_cuda(ref, index)
    if not ref.is_contiguous() or (a is not None and b is not None and (a.shape != b.shape or a.dtype != b.dtype or not b.is_contiguous())):
        raise L.VitbError("select_rows: operands must be contiguous with equal shape and dtype")
    rows = index.numel()
    cols = ref.numel() // rows
    out = torch.empty_like(ref)
    if any_flag is not None and (any_flag.dtype != torch.int32 or any_flag.numel() != 1 or not any_flag.is_cuda):
        raise L.VitbError("select_rows: any_flag must be an int32 CUDA scalar")
    L.check(L._vitb_select_rows_flag(L.ptr(a), L.ptr(b), L.ptr(index), int(member_mask) & 0xFFFFFFFF, rows, cols,
                                     L.dtype_code(ref), L.ptr(out), L.ptr(any_flag), L.stream_ptr(ref.device)),
            "vitb_select_rows")
    return out
