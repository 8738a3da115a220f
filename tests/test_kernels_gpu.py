"""GPU parity of the non-GEMM kernels against plain torch fp32 references of the same op.
Tolerances are stated per test: fp32 paths 1e-5 rel-L2 (accumulation order only); bf16 outputs 4e-3
(one bf16 rounding); bf16 attention 1e-2 (P and dS are rounded to bf16 before the second MMA)."""
import math
import os

import pytest
import torch

from conftest import rel_l2

pytestmark = pytest.mark.gpu


def _randn(shape, seed, scale=1.0, dtype=torch.float32):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(shape, generator=g, device="cuda") * scale).to(dtype)


# ---------------------------------------------------------------------------------------------- LN
@pytest.mark.parametrize("D", [128, 384, 768, 1024, 1280])
def test_layernorm_fwd(D):
    import vitb200
    rows = 1000
    x = _randn((rows, D), 1, 3.0) + 0.5
    g, b = _randn((D,), 2), _randn((D,), 3)
    yf, yh, yl, mean, rstd = vitb200.ops.layernorm_fwd(x, g, b, 1e-5, want_f32=True, want_bf16=True, want_lo=True)
    ref = torch.nn.functional.layer_norm(x, (D,), g, b, 1e-5)
    assert rel_l2(yf, ref) < 1e-6
    assert rel_l2(yh, ref) < 4e-3
    assert rel_l2(yh.float() + yl.float(), ref) < 2e-5
    assert rel_l2(mean, x.mean(1)) < 1e-5
    assert rel_l2(rstd, (x.var(1, unbiased=False) + 1e-5).rsqrt()) < 1e-5


def test_layernorm_fwd_strided_rows():
    import vitb200
    B, N, D = 16, 197, 768
    x = _randn((B, N, D), 4)
    g, b = _randn((D,), 5), _randn((D,), 6)
    cls = x.view(B, N * D)[:, :D]  # row 0 of every image, stride N*D
    yf, _, _, _, _ = vitb200.ops.layernorm_fwd(cls, g, b, 1e-5, want_f32=True, want_bf16=False)
    assert rel_l2(yf, torch.nn.functional.layer_norm(x[:, 0], (D,), g, b, 1e-5)) < 1e-6


@pytest.mark.parametrize("D,dy_dtype", [(256, torch.float32), (768, torch.bfloat16), (768, torch.float32),
                                        (1024, torch.bfloat16), (1280, torch.float32)])
def test_layernorm_bwd(D, dy_dtype):
    import vitb200
    rows = 2500
    x = (_randn((rows, D), 7, 2.0) + 0.3).requires_grad_(True)
    g = _randn((D,), 8).requires_grad_(True)
    b = _randn((D,), 9).requires_grad_(True)
    dy = _randn((rows, D), 10, 1.0, dy_dtype)
    dres = _randn((rows, D), 11)
    y = torch.nn.functional.layer_norm(x, (D,), g, b, 1e-5)
    y.backward(dy.float())
    _, _, _, mean, rstd = vitb200.ops.layernorm_fwd(x.detach(), g.detach(), b.detach(), 1e-5, want_bf16=False)
    dgamma = torch.zeros(D, device="cuda"); dbeta = torch.zeros(D, device="cuda"); dcol = torch.zeros(D, device="cuda")
    dxf, dxh, dxl = vitb200.ops.layernorm_bwd(dy, x.detach(), mean, rstd, g.detach(), dres=dres, want_f32=True,
                                              want_bf16=True, want_lo=True, dgamma=dgamma, dbeta=dbeta, dcolsum=dcol)
    ref_dx = x.grad + dres
    assert rel_l2(dxf, ref_dx) < 1e-5
    assert rel_l2(dxh, ref_dx) < 4e-3
    assert rel_l2(dxh.float() + dxl.float(), ref_dx) < 3e-5
    assert rel_l2(dgamma, g.grad) < 1e-4
    assert rel_l2(dbeta, b.grad) < 1e-4
    assert rel_l2(dcol, ref_dx.sum(0)) < 1e-4


@pytest.mark.parametrize("rows,D,dy_dtype,extras", [(25216, 768, torch.bfloat16, True), (25216, 768, torch.float32, False),
                                                     (5, 768, torch.bfloat16, True), (12608, 1024, torch.bfloat16, True),
                                                     (4000, 384, torch.bfloat16, False)])
def test_layernorm_bwd_many_rows_per_warp(rows, D, dy_dtype, extras):
    """Full-size token counts (many rows per warp, the per-warp shared-memory partial sums carry across rows), tiny row
    counts (most warps idle), and the variants without residual gradient / column sums."""
    import vitb200
    x = (_randn((rows, D), 17, 2.0) + 0.3).requires_grad_(True)
    g = _randn((D,), 18).requires_grad_(True)
    b = _randn((D,), 19).requires_grad_(True)
    dy = _randn((rows, D), 20, 1.0, dy_dtype)
    dres = _randn((rows, D), 21) if extras else None
    torch.nn.functional.layer_norm(x, (D,), g, b, 1e-5).backward(dy.float())
    _, _, _, mean, rstd = vitb200.ops.layernorm_fwd(x.detach(), g.detach(), b.detach(), 1e-5, want_bf16=False)
    dgamma = torch.ones(D, device="cuda"); dbeta = torch.ones(D, device="cuda")
    dcol = torch.ones(D, device="cuda") if extras else None
    dxf, dxh, _ = vitb200.ops.layernorm_bwd(dy, x.detach(), mean, rstd, g.detach(), dres=dres, want_f32=True, want_bf16=True,
                                            dgamma=dgamma, dbeta=dbeta, dcolsum=dcol)
    ref_dx = x.grad + (dres if extras else 0)
    assert rel_l2(dxf, ref_dx) < 1e-5
    assert rel_l2(dxh, ref_dx) < 4e-3
    assert rel_l2(dgamma, g.grad + 1) < 1e-4          # accumulated (+=) into the caller's buffers
    assert rel_l2(dbeta, b.grad + 1) < 1e-4
    if extras:
        assert rel_l2(dcol, ref_dx.sum(0) + 1) < 1e-4


# ---------------------------------------------------------------------------------------- attention
def _attn_ref(q, k, v, H):
    B, Nq, HD = q.shape
    dh = HD // H
    qh = q.float().view(B, Nq, H, dh).permute(0, 2, 1, 3)
    kh = k.float().view(B, -1, H, dh).permute(0, 2, 1, 3)
    vh = v.float().view(B, -1, H, dh).permute(0, 2, 1, 3)
    s = qh @ kh.transpose(-1, -2) / math.sqrt(dh)
    p = torch.softmax(s, -1)
    o = (p @ vh).permute(0, 2, 1, 3).reshape(B, Nq, HD)
    return o, torch.logsumexp(s, -1)


@pytest.mark.parametrize("N", [197, 50, 256, 16, 130])
def test_attn_tc_fwd_bwd_packed_qkv(N):
    import vitb200
    B, H, dh = 3, 12, 64
    D = H * dh
    qkv = _randn((B, N, 3 * D), 20 + N, 1.0, torch.bfloat16)
    q, k, v = qkv[:, :, :D], qkv[:, :, D:2 * D], qkv[:, :, 2 * D:]
    assert vitb200.ops.attn_supported_tc(dh, N, N, torch.bfloat16)
    o, lse = vitb200.ops.attn_fwd(q, k, v, H)
    qf, kf, vf = (t.float().detach().requires_grad_(True) for t in (q, k, v))
    ro, rlse = _attn_ref(qf, kf, vf, H)
    assert rel_l2(o, ro) < 1e-2
    assert rel_l2(lse, rlse) < 1e-4
    do = _randn((B, N, D), 99, 1.0, torch.bfloat16)
    ro.backward(do.float())
    dqkv = torch.empty_like(qkv)
    dq, dk, dv = vitb200.ops.attn_bwd(do, q, k, v, o, lse, H, dq=dqkv[:, :, :D], dk=dqkv[:, :, D:2 * D],
                                      dv=dqkv[:, :, 2 * D:])
    torch.cuda.synchronize()
    assert rel_l2(dv, vf.grad) < 1.5e-2
    assert rel_l2(dq, qf.grad) < 1.5e-2
    assert rel_l2(dk, kf.grad) < 1.5e-2


@pytest.mark.parametrize("N,B,H", [(197, 3, 12), (50, 3, 12), (256, 2, 3), (16, 2, 2), (129, 2, 2), (128, 5, 4), (130, 40, 12),
                                   (197, 128, 12)])
def test_attn_ws_persistent_kernels(N, B, H, monkeypatch):
    """VITB_ATTN_WS=1: the persistent warp-specialised forward / backward (vitb_attention_ws.cu) against the torch reference
    (small cases) and against the one-CTA-per-tile kernels (all cases, incl. several items per resident CTA)."""
    import vitb200
    dh = 64
    D = H * dh
    qkv = _randn((B, N, 3 * D), 60 + N, 1.0 if B * H <= 64 else 0.5, torch.bfloat16)
    q, k, v = qkv[:, :, :D], qkv[:, :, D:2 * D], qkv[:, :, 2 * D:]
    do = _randn((B, N, D), 97, 1.0, torch.bfloat16)
    res = {}
    for flag in ("0", "1"):
        monkeypatch.setenv("VITB_ATTN_WS", flag)
        o, lse = vitb200.ops.attn_fwd(q, k, v, H)
        dqkv = torch.full_like(qkv, float("nan"))
        vitb200.ops.attn_bwd(do, q, k, v, o, lse, H, dq=dqkv[:, :, :D], dk=dqkv[:, :, D:2 * D], dv=dqkv[:, :, 2 * D:])
        torch.cuda.synchronize()
        assert not bool(o.isnan().any()) and not bool(dqkv.isnan().any()), flag
        res[flag] = (o, lse, dqkv)
    assert rel_l2(res["1"][0], res["0"][0]) < 4e-3
    assert rel_l2(res["1"][1], res["0"][1]) < 1e-5
    assert rel_l2(res["1"][2], res["0"][2]) < 6e-3
    if B * H <= 64:
        qf, kf, vf = (t.float().detach().requires_grad_(True) for t in (q, k, v))
        ro, rlse = _attn_ref(qf, kf, vf, H)
        ro.backward(do.float())
        o, lse, dqkv = res["1"]
        assert rel_l2(o, ro) < 1e-2
        assert rel_l2(lse, rlse) < 1e-4
        assert rel_l2(dqkv[:, :, :D], qf.grad) < 1.5e-2
        assert rel_l2(dqkv[:, :, D:2 * D], kf.grad) < 1.5e-2
        assert rel_l2(dqkv[:, :, 2 * D:], vf.grad) < 1.5e-2


@pytest.mark.parametrize("dh,N,H,simt", [(80, 257, 16, True), (128, 130, 3, True), (96, 300, 2, False),
                                         (80, 50, 4, True), (112, 256, 2, False),
                                         (64, 577, 12, False), (80, 730, 2, False), (64, 321, 2, False), (128, 640, 1, False)])
def test_attn_tc_general_shapes_forward(dh, N, H, simt):
    """tcgen05 forward for 64 <= head_dim <= 128 and any token count: ViT-H/14 at 224 px (head_dim 80, 257 tokens:
    two MMAs per S row block, a partially used second head-dim chunk) and the 384 px evaluation resolution
    (577 tokens for */16, 730 for h14: several key blocks, online softmax).  The backward of these shapes runs on
    vitb_attn_bwd_tc_long (next test); here the CUDA-core backward is checked against the same forward."""
    import vitb200
    B = 2
    D = H * dh
    qkv = _randn((B, N, 3 * D), 300 + N + dh, 1.0, torch.bfloat16)
    q, k, v = qkv[:, :, :D], qkv[:, :, D:2 * D], qkv[:, :, 2 * D:]
    assert vitb200.ops.attn_fwd_supported_tc(dh, N, N, torch.bfloat16)
    assert not vitb200.ops.attn_supported_tc(dh, N, N, torch.bfloat16)
    o, lse = vitb200.ops.attn_fwd(q, k, v, H)
    assert o.dtype == torch.bfloat16
    ro, rlse = _attn_ref(q, k, v, H)
    torch.cuda.synchronize()
    assert rel_l2(o, ro) < 1e-2
    assert rel_l2(lse, rlse) < 1e-4
    if not simt:      # larger than the CUDA-core kernel's shared-memory budget: forward only
        return
    o_s, lse_s = vitb200.ops.attn_fwd(q, k, v, H, use_tc=False)
    assert rel_l2(o, o_s) < 1e-2 and rel_l2(lse, lse_s) < 1e-4
    # the SIMT backward consumes this forward's o / lse
    do = _randn((B, N, D), 7, 1.0, torch.bfloat16)
    qf, kf, vf = (t.float().detach().requires_grad_(True) for t in (q, k, v))
    _attn_ref(qf, kf, vf, H)[0].backward(do.float())
    dq, dk, dv = vitb200.ops.attn_bwd(do, q, k, v, o, lse, H, use_tc=False)
    assert rel_l2(dq, qf.grad) < 4e-2 and rel_l2(dk, kf.grad) < 4e-2 and rel_l2(dv, vf.grad) < 4e-2


@pytest.mark.parametrize("dh,N,B,H", [(64, 577, 2, 12), (64, 257, 3, 2), (64, 300, 2, 3), (64, 730, 1, 2), (64, 512, 2, 1),
                                      (64, 197, 2, 2), (64, 40, 2, 1),
                                      (80, 257, 2, 16), (80, 730, 1, 2), (80, 50, 2, 3), (96, 300, 2, 2), (112, 256, 1, 2),
                                      (128, 130, 2, 3)])
def test_attn_tc_backward_any_token_count(dh, N, B, H):
    """tcgen05 backward for ANY number of tokens (vitb_attn_bwd_tc_long).  head_dim 64: 577 tokens is the 384 px
    resolution of */16 (src/config.py:12); key blocks of 256 run on separate CTAs and add their shares of dQ into an fp32
    buffer; 257 and 300 leave a nearly empty second block, 512 two full ones, 197 and 40 a single block.  64 < head_dim
    <= 128: ViT-H/14 (head_dim 80, 257 tokens at 224 px, 730 at 384 px; src/config.py:95-104) — the head dimension spans
    two column chunks, key blocks of 128.  Compared with the fp32 reference at the bf16 attention bar (P and dS are
    rounded to bf16 before the second MMAs), packed q|k|v and packed gradient buffer as the model uses them."""
    import vitb200
    from vitb200 import _lib as L
    import ctypes as C
    D = H * dh
    qkv = _randn((B, N, 3 * D), 900 + N, 1.0, torch.bfloat16)
    q, k, v = qkv[:, :, :D], qkv[:, :, D:2 * D], qkv[:, :, 2 * D:]
    o, lse = vitb200.ops.attn_fwd(q, k, v, H)
    do = _randn((B, N, D), 11, 1.0, torch.bfloat16)
    qf, kf, vf = (t.float().detach().requires_grad_(True) for t in (q, k, v))
    _attn_ref(qf, kf, vf, H)[0].backward(do.float())
    assert vitb200.ops.attn_bwd_supported_any(dh, N, N, torch.bfloat16)
    dqkv = torch.full((B, N, 3 * D), float("nan"), dtype=torch.bfloat16, device="cuda")
    dq, dk, dv = dqkv[:, :, :D], dqkv[:, :, D:2 * D], dqkv[:, :, 2 * D:]
    # straight through the C ABI, so that <= 256 tokens take the key-block kernel too
    p = vitb200.ops._attn_params(q, k, v, o, lse, H)
    p.dout = do.data_ptr()
    p.do_batch_stride, p.do_row_stride = do.stride(0), do.stride(1)
    p.dq, p.dk, p.dv = dq.data_ptr(), dk.data_ptr(), dv.data_ptr()
    p.dq_batch_stride = p.dk_batch_stride = p.dv_batch_stride = dqkv.stride(0)
    p.dq_row_stride = p.dk_row_stride = p.dv_row_stride = dqkv.stride(1)
    acc = torch.zeros((B, N, D), dtype=torch.float32, device="cuda")
    L.check(L._vitb_attn_bwd_tc_long(C.byref(p), L.ptr(acc), L.stream_ptr(q.device)), "vitb_attn_bwd_tc_long")
    torch.cuda.synchronize()
    assert not torch.isnan(dqkv.float()).any()
    assert rel_l2(dq, qf.grad) < 4e-2 and rel_l2(dk, kf.grad) < 4e-2 and rel_l2(dv, vf.grad) < 4e-2, \
        (rel_l2(dq, qf.grad), rel_l2(dk, kf.grad), rel_l2(dv, vf.grad))
    if N > 256 or dh != 64:     # and through ops.attn_bwd, which picks this kernel for these shapes
        dq2, dk2, dv2 = vitb200.ops.attn_bwd(do, q, k, v, o, lse, H)
        assert dq2.dtype == torch.bfloat16
        assert rel_l2(dq2, qf.grad) < 4e-2 and rel_l2(dk2, kf.grad) < 4e-2 and rel_l2(dv2, vf.grad) < 4e-2


def test_attn_tc_c2_size_runs_and_matches_on_a_slice():
    import vitb200
    B, N, H, dh = 128, 197, 12, 64
    D = H * dh
    qkv = _randn((B, N, 3 * D), 5, 0.5, torch.bfloat16)
    q, k, v = qkv[:, :, :D], qkv[:, :, D:2 * D], qkv[:, :, 2 * D:]
    o, lse = vitb200.ops.attn_fwd(q, k, v, H)
    ro, _ = _attn_ref(q[-2:], k[-2:], v[-2:], H)
    assert rel_l2(o[-2:], ro) < 1e-2


@pytest.mark.parametrize("dtype,dh,N,tol", [(torch.float32, 64, 197, 2e-5), (torch.float32, 80, 257, 2e-5),
                                            (torch.bfloat16, 80, 257, 1e-2), (torch.float32, 32, 50, 2e-5)])
def test_attn_simt_fwd_bwd(dtype, dh, N, tol):
    import vitb200
    B, H = 2, 4
    D = H * dh
    q, k, v = (_randn((B, N, D), s, 1.0, dtype) for s in (1, 2, 3))
    o, lse = vitb200.ops.attn_fwd(q, k, v, H, use_tc=False)
    qf, kf, vf = (t.float().detach().requires_grad_(True) for t in (q, k, v))
    ro, rlse = _attn_ref(qf, kf, vf, H)
    assert rel_l2(o, ro) < tol
    assert rel_l2(lse, rlse) < 1e-5
    do = _randn((B, N, D), 4, 1.0, dtype)
    ro.backward(do.float())
    dq, dk, dv = vitb200.ops.attn_bwd(do, q, k, v, o, lse, H, use_tc=False)
    btol = tol * (4 if dtype == torch.bfloat16 else 2)
    assert rel_l2(dq, qf.grad) < btol and rel_l2(dk, kf.grad) < btol and rel_l2(dv, vf.grad) < btol


def test_attn_simt_asymmetric_queries():
    import vitb200
    B, H, dh, Nq, Nk = 2, 4, 64, 77, 197
    D = H * dh
    q = _randn((B, Nq, D), 1); k = _randn((B, Nk, D), 2); v = _randn((B, Nk, D), 3)
    o, _ = vitb200.ops.attn_fwd(q, k, v, H, use_tc=False)
    ro, _ = _attn_ref(q, k, v, H)
    assert rel_l2(o, ro) < 2e-5


@pytest.mark.parametrize("dtype,dh,Nk,tol", [(torch.bfloat16, 64, 197, 1e-2), (torch.float32, 64, 197, 2e-5),
                                             (torch.bfloat16, 80, 257, 1e-2), (torch.bfloat16, 64, 577, 1e-2),
                                             (torch.float32, 128, 50, 2e-5)])
def test_attn_single_query_fwd_bwd(dtype, dh, Nk, tol):
    """One query per (image, head) — the class-token-only last block (EncoderBlock.forward_row0) — runs on the
    dedicated attn_q1 kernels (any key count, K / V read once from HBM)."""
    import vitb200
    B, H = 3, 4
    D = H * dh
    q = _randn((B, 1, D), 1, 1.0, dtype); k = _randn((B, Nk, D), 2, 1.0, dtype); v = _randn((B, Nk, D), 3, 1.0, dtype)
    o, lse = vitb200.ops.attn_fwd(q, k, v, H, use_tc=False)
    qf, kf, vf = (t.float().detach().requires_grad_(True) for t in (q, k, v))
    ro, rlse = _attn_ref(qf, kf, vf, H)
    assert o.shape == (B, 1, D) and lse.shape == (B, H, 1)
    assert rel_l2(o, ro) < tol
    assert rel_l2(lse, rlse) < 1e-5
    do = _randn((B, 1, D), 4, 1.0, dtype)
    ro.backward(do.float())
    dq, dk, dv = vitb200.ops.attn_bwd(do, q, k, v, o, lse, H, use_tc=False)
    torch.cuda.synchronize()
    btol = tol * (4 if dtype == torch.bfloat16 else 2)
    assert rel_l2(dq, qf.grad) < btol and rel_l2(dk, kf.grad) < btol and rel_l2(dv, vf.grad) < btol


@pytest.mark.parametrize("Nk,B,H", [(197, 5, 12), (17, 3, 2), (256, 2, 4), (1, 2, 1)])
def test_attn_single_query_bf16_gradients_into_a_packed_buffer(Nk, B, H):
    """vitb_attn_q1_bwd: one query per image, bf16 dk | dv written through strides into a packed [B, Nk, 2D] buffer (what
    the fused last block hands to its grouped weight-gradient GEMM), bf16 dq; forward through the same bandwidth-bound
    kernel family.  Tolerance: bf16 rounding of the outputs (4e-3) on top of the bf16 attention bar."""
    import vitb200
    D = H * 64
    assert vitb200.ops.attn_q1_supported(64, Nk, torch.bfloat16)
    q = _randn((B, 1, D), 1, 1.0, torch.bfloat16)
    kv = _randn((B, Nk, 2 * D), 2, 1.0, torch.bfloat16)
    k, v = kv[:, :, :D], kv[:, :, D:]
    o, lse = vitb200.ops.attn_fwd(q, k, v, H, use_tc=False)
    qf, kf, vf = (t.float().detach().requires_grad_(True) for t in (q, k, v))
    ro, rlse = _attn_ref(qf, kf, vf, H)
    assert rel_l2(o, ro) < 1e-2 and rel_l2(lse, rlse) < 1e-5
    do = _randn((B, 1, D), 4, 1.0, torch.bfloat16)
    ro.backward(do.float())
    dkv = torch.full((B, Nk, 2 * D), float("nan"), dtype=torch.bfloat16, device="cuda")
    dq, dk, dv = vitb200.ops.attn_q1_bwd(do, q, k, v, o, lse, H, dk=dkv[:, :, :D], dv=dkv[:, :, D:])
    torch.cuda.synchronize()
    assert dq.dtype == torch.bfloat16 and not torch.isnan(dkv.float()).any()
    assert rel_l2(dq, qf.grad) < 4e-2 and rel_l2(dkv[:, :, :D], kf.grad) < 4e-2 and rel_l2(dkv[:, :, D:], vf.grad) < 4e-2


def test_layernorm_bwd_residual_gradient_on_every_nth_row():
    """dres_every = N: only rows 0, N, 2N, ... carry a residual-branch gradient (row i of a [rows / N, D] tensor)."""
    import vitb200
    B, N, D = 7, 5, 256
    rows = B * N
    x = _randn((rows, D), 1); dy = _randn((rows, D), 2).to(torch.bfloat16); gamma = _randn((D,), 3) + 1.0
    dres0 = _randn((B, D), 4)
    _, _, _, mean, rstd = vitb200.ops.layernorm_fwd(x, gamma, torch.zeros_like(gamma), 1e-6)
    dense = torch.zeros(rows, D, device="cuda")
    dense[::N] = dres0
    cs_a, cs_b = torch.zeros(D, device="cuda"), torch.zeros(D, device="cuda")
    dx_a, dxb_a, _ = vitb200.ops.layernorm_bwd(dy, x, mean, rstd, gamma, dres=dense, want_bf16=True, dcolsum=cs_a)
    dx_b, dxb_b, _ = vitb200.ops.layernorm_bwd(dy, x, mean, rstd, gamma, dres=dres0, dres_every=N, want_bf16=True, dcolsum=cs_b)
    torch.cuda.synchronize()
    assert torch.equal(dx_a, dx_b) and torch.equal(dxb_a, dxb_b)
    assert rel_l2(cs_b, cs_a) < 1e-6


@pytest.mark.parametrize("dtype,n,p", [(torch.float32, 1 << 20, 0.1), (torch.bfloat16, 1 << 20, 0.1),
                                       (torch.float32, 1000003, 0.5), (torch.bfloat16, 7, 0.25)])
def test_dropout_kernel_statistics_mask_consistency_and_fresh_draws(dtype, n, p):
    """nn.Dropout (src/model.py:12,35-36,111) as vitb_dropout_fwd / _bwd.  The mask stream is the library's own Philox
    stream, so parity with torch is statistical: keep rate within 5 sigma of 1 - p, kept values scaled by exactly
    1 / (1 - p) (bf16: one rounding), dropped ones exactly zero, backward = dy * mask / (1 - p), the residual form adds in
    fp32, and two calls (as two replays of a captured step would) draw different masks."""
    import vitb200
    x = (_randn((n,), 1).abs() + 1.0).to(dtype)    # no zeros, so the mask is visible in y
    state = torch.zeros(2, dtype=torch.int64, device="cuda")
    y, mask = vitb200.ops.dropout_fwd(x, p, 1234, state)
    y2, mask2 = vitb200.ops.dropout_fwd(x, p, 1234, state)
    torch.cuda.synchronize()
    assert int(state[0]) == 2 and int(state[1]) == 0
    keep = mask.float().mean().item()
    if n > 1000:
        assert abs(keep - (1 - p)) < 5 * math.sqrt(p * (1 - p) / n), keep
        assert (mask != mask2).float().mean().item() > 0.5 * 2 * p * (1 - p)      # independent draws differ on 2p(1-p)
        # no visible structure between neighbours: the keep rate of every fourth element matches too
        for j in range(4):
            assert abs(mask[j::4].float().mean().item() - (1 - p)) < 6 * math.sqrt(p * (1 - p) / (n / 4))
    ref = torch.where(mask.bool(), x.float() / (1 - p), torch.zeros((), device="cuda"))
    assert rel_l2(y, ref) < (4e-3 if dtype == torch.bfloat16 else 1e-6)
    assert bool(((y == 0) == (mask == 0)).all())
    dy = _randn((n,), 2, 1.0, dtype)
    dx = vitb200.ops.dropout_bwd(dy, mask, p, dtype)
    refd = torch.where(mask.bool(), dy.float() / (1 - p), torch.zeros((), device="cuda"))
    assert rel_l2(dx, refd) < (4e-3 if dtype == torch.bfloat16 else 1e-6)
    if n % 4 == 0:
        res = _randn((n,), 3)
        st2 = torch.zeros(2, dtype=torch.int64, device="cuda")
        yr, mr = vitb200.ops.dropout_fwd(x, p, 99, st2, residual=res)
        refr = res + torch.where(mr.bool(), x.float() / (1 - p), torch.zeros((), device="cuda"))
        assert yr.dtype == torch.float32 and rel_l2(yr, refr) < 1e-6


# ---------------------------------------------------------------------------------- elementwise etc.
def test_cast_split():
    import vitb200
    x = _randn((1000, 771), 1)
    hi, lo = vitb200.ops.cast_split(x, want_lo=True)
    assert torch.equal(hi, x.to(torch.bfloat16))
    assert rel_l2(hi.float() + lo.float(), x) < 1e-5


@pytest.mark.parametrize("P", [16, 32, 14])
def test_im2col_matches_conv2d(P):
    import vitb200
    B, D = 3, 64
    img = _randn((B, 3, 224, 224), 2)
    w = _randn((D, 3, P, P), 3, 0.05)
    hi, lo = vitb200.ops.im2col(img, P, want_lo=True)
    K = 3 * P * P
    cols = hi.float() + lo.float()
    assert cols.shape[1] % 8 == 0 and cols.shape[1] >= K
    assert (cols[:, K:] == 0).all()
    # fp64 reference: torch's fp32 conv2d / matmul may use TF32 on this GPU
    ref = torch.nn.functional.conv2d(img.double(), w.double(), stride=P).permute(0, 2, 3, 1).reshape(-1, D)
    out = cols[:, :K].double() @ w.double().reshape(D, K).t()
    assert rel_l2(out, ref) < 2e-5


def test_cls_rows_and_embed_bwd():
    import vitb200
    B, N, D = 5, 50, 256
    x = torch.zeros(B, N, D, device="cuda")
    cls, pos = _randn((D,), 1), _randn((N, D), 2)
    vitb200.ops.cls_rows(x, cls, pos)
    assert torch.allclose(x[:, 0], (cls + pos[0]).expand(B, D)) and x[:, 1:].abs().max() == 0
    dx = _randn((B, N, D), 3)
    dpos = torch.zeros(N, D, device="cuda"); dcls = torch.zeros(D, device="cuda"); dbias = torch.zeros(D, device="cuda")
    hi, lo = vitb200.ops.embed_bwd(dx, dpos=dpos, dcls=dcls, dbias=dbias, want_lo=True)
    assert rel_l2(dpos, dx.sum(0)) < 1e-6
    assert rel_l2(dcls, dx[:, 0].sum(0)) < 1e-6
    assert rel_l2(dbias, dx[:, 1:].sum((0, 1))) < 1e-5
    assert rel_l2(hi.float() + lo.float(), dx[:, 1:].reshape(-1, D)) < 1e-5


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_colsum(dtype):
    import vitb200
    x = _randn((5000, 2304), 1, 1.0, dtype)
    out = torch.ones(2304, device="cuda")
    vitb200.ops.colsum(x, out)
    assert rel_l2(out, x.float().sum(0) + 1) < 1e-4


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("rows,cols", [(25216, 768), (197, 3072), (1, 8), (333, 100), (4097, 36), (130, 2304)])
def test_colsum_shapes(dtype, rows, cols):
    """16-byte / 8-byte column groups, widths that are not a multiple of 8, single rows, ragged row slices."""
    import vitb200
    x = _randn((rows, cols), 3, 1.0, dtype)
    out = torch.full((cols,), 2.0, device="cuda")
    vitb200.ops.colsum(x, out)
    assert rel_l2(out, x.float().sum(0) + 2) < 1e-4


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_colsum_strided_view_and_three_segments(dtype):
    import vitb200
    big = _randn((1000, 3 * 768 + 64), 4, 1.0, dtype)
    x = big[:, 8:8 + 3 * 768]                      # row stride > width, base offset of 8 elements
    outs = [torch.zeros(768, device="cuda") for _ in range(3)]
    vitb200.ops.colsum3(x, *outs)
    ref = x.float().sum(0)
    for i in range(3):
        assert rel_l2(outs[i], ref[i * 768:(i + 1) * 768]) < 1e-4
    x4 = big[:, 4:4 + 3 * 100]                     # segments of 100 columns: the 4-column path, 8-byte alignment only
    o = [torch.zeros(100, device="cuda") for _ in range(3)]
    if dtype == torch.bfloat16:
        vitb200.ops.colsum3(x4, *o)
        r4 = x4.float().sum(0)
        for i in range(3):
            assert rel_l2(o[i], r4[i * 100:(i + 1) * 100]) < 1e-4


def test_cross_entropy():
    import vitb200
    B, Ccls = 128, 100
    logits = _randn((B, Ccls), 1, 3.0).requires_grad_(True)
    labels = torch.randint(0, Ccls, (B,), device="cuda")
    ref = torch.nn.functional.cross_entropy(logits, labels)
    ref.backward()
    loss, dl = vitb200.ops.cross_entropy(logits.detach(), labels)
    assert abs(float(loss) - float(ref)) < 1e-5 * abs(float(ref))
    assert rel_l2(dl, logits.grad) < 1e-5


def test_sgd_momentum_matches_torch():
    import vitb200
    n = 4 * 1000 + 4
    p0 = _randn((n,), 1)
    p_ref = p0.clone().requires_grad_(True)
    opt = torch.optim.SGD([p_ref], lr=0.03, momentum=0.9, weight_decay=1e-4)
    p = p0.clone(); m = torch.zeros_like(p)
    hi = torch.empty(n, dtype=torch.bfloat16, device="cuda")
    for step in range(3):
        g = _randn((n,), 10 + step)
        p_ref.grad = g.clone()
        opt.step()
        vitb200.ops.sgd_momentum(p, g, m, 0.03, 0.9, weight_decay=1e-4, first_step=(step == 0), shadow_hi=hi)
    assert rel_l2(p, p_ref.detach()) < 1e-6
    assert torch.equal(hi, p.to(torch.bfloat16))


def test_adamw_and_clip_match_torch():
    import vitb200
    n = 5003
    p0 = _randn((n,), 1)
    p_ref = p0.clone().requires_grad_(True)
    opt = torch.optim.AdamW([p_ref], lr=1e-3, weight_decay=0.05)
    p = p0.clone(); m = torch.zeros_like(p); v = torch.zeros_like(p)
    ss = torch.zeros((), device="cuda"); coef = torch.zeros((), device="cuda"); nrm = torch.zeros((), device="cuda")
    for step in range(1, 4):
        g = _randn((n,), 10 + step, 2.0)
        p_ref.grad = g.clone()
        ref_norm = torch.nn.utils.clip_grad_norm_([p_ref], 1.0)
        opt.step()
        ss.zero_()
        vitb200.ops.sumsq(g, ss)
        vitb200.ops.clip_coef(ss, 1.0, coef, nrm)
        assert abs(float(nrm) - float(ref_norm)) < 1e-4 * float(ref_norm)
        vitb200.ops.adamw(p, g, m, v, 1e-3, 0.9, 0.999, 1e-8, 0.05, step, grad_scale=coef)
    assert rel_l2(p, p_ref.detach()) < 1e-6
