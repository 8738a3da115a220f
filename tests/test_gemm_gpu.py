"""GPU parity of the tcgen05 GEMM (vitb_gemm) against a plain torch fp32 matmul on the same
bf16-rounded operands.  Tolerances: fp32 outputs differ only by accumulation order (rel-L2 1e-5);
bf16 outputs add one rounding (2^-9 relative per element -> rel-L2 4e-3)."""
import pytest
import torch

from conftest import rel_l2

pytestmark = pytest.mark.gpu


def _mk(shape, seed, scale=1.0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(shape, generator=g, device="cuda") * scale).to(torch.bfloat16)


def _operands(M, N, K, a_mn, b_mn, seed=0):
    A = _mk((K, M) if a_mn else (M, K), seed)
    B = _mk((K, N) if b_mn else (N, K), seed + 1)
    Af = A.float().t() if a_mn else A.float()
    Bf = B.float().t() if b_mn else B.float()
    return A, B, Af @ Bf.t()


@pytest.mark.parametrize("a_mn", [False, True])
@pytest.mark.parametrize("b_mn", [False, True])
@pytest.mark.parametrize("shape", [(128, 256, 64), (256, 512, 256), (384, 128, 192), (592, 104, 72),
                                   (1000, 776, 328)])
def test_gemm_majors_fp32_out(shape, a_mn, b_mn):
    import vitb200
    M, N, K = shape
    A, B, ref = _operands(M, N, K, a_mn, b_mn)
    out = vitb200.ops.gemm(A, B, a_mn=a_mn, b_mn=b_mn, out_dtype=torch.float32)
    torch.cuda.synchronize()
    assert rel_l2(out, ref) < 1e-5, (shape, a_mn, b_mn, rel_l2(out, ref))


def test_gemm_c2_qkv_bias_bf16():
    import vitb200
    M, N, K = 25216, 2304, 768
    A, B, ref = _operands(M, N, K, False, False, seed=3)
    bias = torch.randn(N, device="cuda")
    out = vitb200.ops.gemm(A, B, bias=bias)
    torch.cuda.synchronize()
    assert out.dtype == torch.bfloat16
    assert rel_l2(out, ref + bias) < 4e-3


def test_gemm_gelu_epilogue_and_preact():
    import vitb200
    M, N, K = 640, 3072, 768
    A, B, ref = _operands(M, N, K, False, False, seed=5)
    A = (A.float() * 0.05).to(torch.bfloat16)
    ref = A.float() @ B.float().t()
    bias = torch.randn(N, device="cuda") * 0.1
    z = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
    out = vitb200.ops.gemm(A, B, bias=bias, epilogue=vitb200.ops.EPI_GELU, d2=z)
    torch.cuda.synchronize()
    zr = ref + bias
    assert rel_l2(z, zr) < 4e-3
    assert rel_l2(out, torch.nn.functional.gelu(zr)) < 4e-3


def test_gemm_gelu_bwd_epilogue():
    import vitb200
    M, N, K = 512, 1024, 256
    A, B, ref = _operands(M, N, K, False, True, seed=7)
    z = _mk((M, N), 11)
    out = vitb200.ops.gemm(A, B, b_mn=True, epilogue=vitb200.ops.EPI_GELU_BWD, aux=z)
    torch.cuda.synchronize()
    zf = z.float().requires_grad_(True)
    torch.nn.functional.gelu(zf).backward(ref)
    assert rel_l2(out, zf.grad) < 4e-3


@pytest.mark.parametrize("r_dtype", [torch.float32, torch.bfloat16])
def test_gemm_bias_residual_fp32_out(r_dtype):
    import vitb200
    M, N, K = 788, 768, 3072
    A, B, ref = _operands(M, N, K, False, False, seed=9)
    bias = torch.randn(N, device="cuda")
    res = torch.randn(M, N, device="cuda").to(r_dtype)
    out = vitb200.ops.gemm(A, B, bias=bias, residual=res, out_dtype=torch.float32)
    torch.cuda.synchronize()
    assert rel_l2(out, ref + bias + res.float()) < 1e-5


def test_gemm_patch_embed_row_remap():
    import vitb200
    Bsz, npatch, D, K = 4, 196, 768, 768
    A, B, ref = _operands(Bsz * npatch, D, K, False, False, seed=13)
    bias = torch.randn(D, device="cuda")
    pos = torch.randn(npatch + 1, D, device="cuda")
    out = torch.zeros(Bsz * (npatch + 1), D, device="cuda")
    vitb200.ops.gemm(A, B, bias=bias, residual=pos, row_remap_group=npatch, out=out)
    torch.cuda.synchronize()
    exp = torch.zeros(Bsz, npatch + 1, D, device="cuda")
    exp[:, 1:] = (ref + bias).view(Bsz, npatch, D) + pos[1:]
    assert rel_l2(out, exp.view(-1, D)) < 1e-5
    assert out.view(Bsz, npatch + 1, D)[:, 0].abs().max() == 0  # class-token rows untouched


@pytest.mark.parametrize("split_k", [0, 1, 3, 8])
def test_gemm_wgrad_accumulate_split_k(split_k):
    import vitb200
    T, Nout, Kin = 5000, 768, 384
    dY = _mk((T, Nout), 21, 0.1)
    X = _mk((T, Kin), 22)
    ref = dY.float().t() @ X.float()
    acc = torch.ones(Nout, Kin, device="cuda")
    vitb200.ops.gemm(dY, X, a_mn=True, b_mn=True, out=acc, accumulate=True, split_k=split_k)
    torch.cuda.synchronize()
    assert rel_l2(acc, ref + 1.0) < 2e-5


def test_gemm_three_segments_bf16x3():
    """fp32 parity mode: A = Ah + Al, B = Bh + Bl (bf16 pieces); AhBh + AhBl + AlBh ~ fp32 product."""
    import vitb200
    M, N, K = 300, 520, 200
    g = torch.Generator(device="cuda").manual_seed(31)
    A = torch.randn(M, K, generator=g, device="cuda")
    B = torch.randn(N, K, generator=g, device="cuda")
    Ah = A.to(torch.bfloat16); Al = (A - Ah.float()).to(torch.bfloat16)
    Bh = B.to(torch.bfloat16); Bl = (B - Bh.float()).to(torch.bfloat16)
    out = vitb200.ops.gemm([Ah, Ah, Al], [Bh, Bl, Bh], out_dtype=torch.float32)
    torch.cuda.synchronize()
    ref = A.double() @ B.double().t()
    assert rel_l2(out, ref) < 2e-5


def test_gemm_lora_segment_small_k():
    import vitb200
    M, N, K, r = 394, 768, 768, 8
    A, B, ref = _operands(M, N, K, False, False, seed=41)
    t = _mk((M, r), 42)
    Bl = _mk((N, r), 43)
    out = vitb200.ops.gemm([A, t], [B, Bl], out_dtype=torch.float32)
    torch.cuda.synchronize()
    assert rel_l2(out, ref + t.float() @ Bl.float().t()) < 1e-5


def test_gemm_row_bias():
    import vitb200
    Bsz, Ntok, N, K = 6, 50, 512, 512
    A, B, ref = _operands(Bsz * Ntok, N, K, False, False, seed=51)
    rb = torch.randn(Bsz, N, device="cuda")
    out = vitb200.ops.gemm(A, B, row_bias=rb, row_bias_group=Ntok, out_dtype=torch.float32)
    torch.cuda.synchronize()
    assert rel_l2(out, ref + rb.repeat_interleave(Ntok, 0)) < 1e-5


def test_gemm_rejects_cpu_tensors():
    import vitb200
    with pytest.raises(RuntimeError):
        vitb200.ops.gemm(torch.zeros(8, 8, dtype=torch.bfloat16), torch.zeros(8, 8, dtype=torch.bfloat16))


def test_gemm_gelu_bwd_with_colsum_bf16():
    """fc2-dgrad epilogue as the fused block uses it: dz = (dy W2) * gelu'(z), db1 += column sums of dz."""
    import vitb200
    M, N, K = 1000, 3072, 768
    A, B, ref = _operands(M, N, K, False, True, seed=61)
    A = (A.float() * 0.05).to(torch.bfloat16)
    ref = A.float() @ B.float()
    z = _mk((M, N), 62)
    cs = torch.ones(N, device="cuda")
    out = vitb200.ops.gemm(A, B, b_mn=True, epilogue=vitb200.ops.EPI_GELU_BWD, aux=z, colsum=cs)
    torch.cuda.synchronize()
    zf = z.float().requires_grad_(True)
    torch.nn.functional.gelu(zf).backward(ref)
    assert rel_l2(out, zf.grad) < 4e-3
    assert rel_l2(cs, zf.grad.sum(0) + 1.0) < 4e-3


@pytest.mark.parametrize("M,N", [(777, 3072), (300, 200)])
def test_gemm_gelu_dg_then_mul_aux_equals_gelu_backward(M, N):
    """bf16 MLP as the fused block runs it: the forward epilogue stores gelu'(z) (VITB_EPI_GELU_DG), the fc2
    dgrad epilogue multiplies by it (VITB_EPI_MUL_AUX) and accumulates the fc1 bias gradient."""
    import vitb200
    K = 768
    A, B, _ = _operands(M, N, K, False, False, seed=71)
    A = (A.float() * 0.05).to(torch.bfloat16)
    bias = torch.randn(N, device="cuda") * 0.1
    zr = (A.float() @ B.float().t() + bias).requires_grad_(True)
    dg = torch.full((M, N), float("nan"), dtype=torch.bfloat16, device="cuda")
    out = vitb200.ops.gemm(A, B, bias=bias, epilogue=vitb200.ops.EPI_GELU_DG, d2=dg)
    torch.cuda.synchronize()
    g = torch.nn.functional.gelu(zr)
    g.backward(torch.ones_like(g))
    assert rel_l2(out, g.detach()) < 4e-3
    assert rel_l2(dg, zr.grad) < 4e-3            # gelu'(z) itself, rounded to bf16
    # backward: dz = (dy W2) * dg, column sums -> bias gradient
    Kb = 256
    dy, W2, ref = _operands(M, N, Kb, False, True, seed=72)
    cs = torch.ones(N, device="cuda")
    dz = vitb200.ops.gemm(dy, W2, b_mn=True, epilogue=vitb200.ops.EPI_MUL_AUX, aux=dg, colsum=cs if N % 4 == 0 else None)
    torch.cuda.synchronize()
    want = ref * dg.float()
    assert rel_l2(dz, want) < 4e-3
    if N % 4 == 0:
        assert rel_l2(cs, want.sum(0) + 1.0) < 4e-3


@pytest.mark.parametrize("M,N,with_bias", [(777, 3072, True), (300, 200, True), (1000, 768, False), (25216, 3072, True)])
def test_gemm_gelu_dg_packed_pair_epilogue(M, N, with_bias, monkeypatch):
    """VITB_EPI_PACKED=1: the GELU + GELU' epilogue on packed fp32 pairs (FFMA2), bias staged in shared memory.
    Same bar as the scalar epilogue against the torch reference, and the two epilogues agree with each other to
    bf16 rounding (gelu = z * Phi(z) on the packed path, |z| * half_erf + z/2 on the scalar one)."""
    import vitb200
    K = 768
    A, B, _ = _operands(M, N, K, False, False, seed=91)
    A = (A.float() * 0.05).to(torch.bfloat16)
    bias = torch.randn(N, device="cuda") * 0.1 if with_bias else None
    z = A.float() @ B.float().t()
    if with_bias:
        z = z + bias
    zr = z.requires_grad_(True)
    g = torch.nn.functional.gelu(zr)
    g.backward(torch.ones_like(g))
    outs = {}
    for flag in ("0", "1"):
        monkeypatch.setenv("VITB_EPI_PACKED", flag)
        dg = torch.full((M + 3, N), float("nan"), dtype=torch.bfloat16, device="cuda")
        out = torch.full((M + 3, N), 7.0, dtype=torch.bfloat16, device="cuda")
        vitb200.ops.gemm(A, B, bias=bias, epilogue=vitb200.ops.EPI_GELU_DG, d2=dg[:M], out=out[:M])
        torch.cuda.synchronize()
        assert rel_l2(out[:M], g.detach()) < 4e-3, flag
        assert rel_l2(dg[:M], zr.grad) < 4e-3, flag
        assert bool((out[M:] == 7.0).all()) and bool(dg[M:].isnan().all()), flag     # rows past M untouched
        outs[flag] = (out[:M].float(), dg[:M].float())
    assert rel_l2(outs["1"][0], outs["0"][0]) < 2e-3 and rel_l2(outs["1"][1], outs["0"][1]) < 2e-3


@pytest.mark.parametrize("M,N,with_colsum", [(777, 3072, True), (300, 200, True), (1000, 768, False), (25216, 3072, True)])
def test_gemm_mul_aux_register_layout_epilogue(M, N, with_colsum, monkeypatch):
    """VITB_EPI_ROWMUL=1: v *= aux in the TMEM register layout (aux rows prefetched a chunk ahead, TMA stores, column
    sums by a warp transpose-reduce) against the same reference and the staged epilogue."""
    import vitb200
    K = 256
    dy, W2, ref = _operands(M, N, K, False, True, seed=93)
    aux = torch.full((M + 5, N), float("nan"), dtype=torch.bfloat16, device="cuda")     # rows past M must not be read into the sums
    aux[:M] = _mk((M, N), 94)
    want = ref * aux[:M].float()
    res = {}
    for flag in ("0", "1"):
        monkeypatch.setenv("VITB_EPI_ROWMUL", flag)
        cs = torch.ones(N, device="cuda") if with_colsum else None
        out = torch.full((M + 3, N), 7.0, dtype=torch.bfloat16, device="cuda")
        vitb200.ops.gemm(dy, W2, b_mn=True, epilogue=vitb200.ops.EPI_MUL_AUX, aux=aux[:M], colsum=cs, out=out[:M])
        torch.cuda.synchronize()
        assert rel_l2(out[:M], want) < 4e-3, flag
        assert bool((out[M:] == 7.0).all()), flag
        if with_colsum:
            assert rel_l2(cs, want.sum(0) + 1.0) < 4e-3, flag
        res[flag] = out[:M].clone()
    assert torch.equal(res["0"], res["1"])          # same fp32 product, same rounding: bit-identical tiles


def test_gemm_bf16_tma_store_respects_row_and_column_tails_and_strided_outputs():
    """bf16 outputs leave through 32x32 TMA-store tiles: rows >= M / columns >= N are clipped by the tensor map,
    and a column slice of a wider buffer (the packed q|k|v projection output) keeps its neighbours intact."""
    import vitb200
    M, N, K = 1000, 768, 768
    A, B, ref = _operands(M, N, K, False, False, seed=81)
    bias = torch.randn(N, device="cuda")
    big = torch.full((M + 8, 3 * N), 7.0, dtype=torch.bfloat16, device="cuda")
    vitb200.ops.gemm(A, B, bias=bias, out=big[:M, N:2 * N])
    torch.cuda.synchronize()
    assert rel_l2(big[:M, N:2 * N], ref + bias) < 4e-3
    assert bool((big[:M, :N] == 7.0).all()) and bool((big[:M, 2 * N:] == 7.0).all()) and bool((big[M:] == 7.0).all())
    M2, N2, K2 = 300, 200, 128                      # BN = 128 tiles, 200 = 6 full 32-column chunks + 8 columns
    A, B, ref = _operands(M2, N2, K2, False, False, seed=82)
    out = torch.full((M2 + 4, N2), 7.0, dtype=torch.bfloat16, device="cuda")
    vitb200.ops.gemm(A, B, out=out[:M2])
    torch.cuda.synchronize()
    assert rel_l2(out[:M2], ref) < 4e-3 and bool((out[M2:] == 7.0).all())


@pytest.mark.parametrize("M,N,K", [(256, 256, 64), (512, 256, 300), (768, 768, 2048), (3072, 768, 1000), (304, 520, 777)])
def test_gemm_wgrad_on_cta_pairs(M, N, K):
    """Weight gradients with M, N >= 256 run on the cta_group::2 pair kernel (256 x 256 tiles, automatic split-K,
    fp32 red.global accumulation); VITB_GEMM_PAIR=0 keeps the single-CTA kernel, both must agree with torch."""
    import os
    import vitb200
    A = _mk((K, M), 91, 0.5)
    B = _mk((K, N), 92, 0.5)
    ref = A.float().t() @ B.float() + 1.0
    outs = []
    for flag in ("1", "0"):
        os.environ["VITB_GEMM_PAIR"] = flag
        try:
            out = torch.ones(M, N, device="cuda")
            vitb200.ops.gemm(A, B, a_mn=True, b_mn=True, out=out, accumulate=True)
            torch.cuda.synchronize()
        finally:
            os.environ.pop("VITB_GEMM_PAIR", None)
        assert rel_l2(out, ref) < 2e-5
        outs.append(out)
    assert rel_l2(outs[0], outs[1]) < 1e-6


@pytest.mark.parametrize("b_mn", [True, False])
@pytest.mark.parametrize("M,Ng,K,gap", [(200, 256, 192, 0), (1000, 768, 768, 768), (25216, 768, 768, 64)])
def test_gemm_column_groups_merged_qkv(M, Ng, K, gap, b_mn):
    """n_groups = 3: q | k | v as ONE GEMM against three weight matrices that only sit at a uniform distance (`gap` extra
    elements between them, as the biases do in a parameter-ordered flat buffer) — src/model.py:86-88."""
    import vitb200
    G = 3
    A = _mk((M, K), 11)
    flat = _mk((G * (K * Ng + gap),), 12)
    Bst = torch.as_strided(flat, (G, K, Ng) if b_mn else (G, Ng, K), (K * Ng + gap, Ng if b_mn else K, 1))
    bias = torch.randn(G * Ng, device="cuda")
    out = vitb200.ops.gemm(A, Bst, b_mn=b_mn, bias=bias)
    torch.cuda.synchronize()
    assert out.shape == (M, G * Ng)
    for g in range(G):
        Bf = Bst[g].float() if b_mn else Bst[g].float().t()
        ref = A.float() @ Bf + bias[g * Ng:(g + 1) * Ng]
        assert rel_l2(out[:, g * Ng:(g + 1) * Ng], ref) < 4e-3, g


@pytest.mark.parametrize("pair", ["1", "0"])
@pytest.mark.parametrize("T,Kin,Ng,gap", [(1000, 256, 256, 256), (25216, 768, 768, 0), (4000, 768, 768, 768)])
def test_gemm_grouped_output_merged_qkv_wgrad(T, Kin, Ng, gap, pair, monkeypatch):
    """The three weight gradients dW_g[Kin, Ng] += X^T dY[:, g] of a merged projection as ONE GEMM whose output groups sit
    at a uniform distance (views of a flat gradient buffer), on CTA pairs and on the single-CTA kernel."""
    import vitb200
    monkeypatch.setenv("VITB_GEMM_PAIR", pair)
    G = 3
    X = _mk((T, Kin), 21, 0.5)
    dY = _mk((T, G * Ng), 22, 0.5)
    flat = torch.full((G * (Kin * Ng + gap),), 0.25, device="cuda")
    out = torch.as_strided(flat, (G, Kin, Ng), (Kin * Ng + gap, Ng, 1))
    vitb200.ops.gemm(X, dY, a_mn=True, b_mn=True, out=out, accumulate=True)
    torch.cuda.synchronize()
    for g in range(G):
        ref = X.float().t() @ dY[:, g * Ng:(g + 1) * Ng].float() + 0.25
        assert rel_l2(out[g], ref) < 2e-5, g
    if gap:     # nothing was written between the groups
        between = torch.as_strided(flat, (G, gap), (Kin * Ng + gap, 1), Kin * Ng)
        assert bool((between == 0.25).all())
