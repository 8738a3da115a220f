"""GPU parity of the Res-ViT modules (router, LoRA-fused projections, approximators, Transformer) against
the golden vectors of the unmodified reference and the oracle (oracle/resvit_oracle.py).
north_star bars: router token indices bit-exact; fp32 mode logits/gradients within rel 1e-4; bf16 logits 2e-2."""
import os
import sys
from types import SimpleNamespace

import pytest
import torch

from conftest import grad_close, rel_l2

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import resvit_oracle  # noqa: E402

pytestmark = pytest.mark.gpu
GOLD = os.path.join(ROOT, "tests", "golden", "resvit_tiny.pt")


def _build(g):
    import vitb200
    from vitb200 import resvit
    kw = dict(g["args"])
    kw["device"] = "cuda"
    m = resvit.Transformer(resvit.ModelArgs(**kw))
    m.load_state_dict(g["state_dict"])
    return m.cuda()


def _replay_noise(m, g):
    """The Gumbel samples the reference drew (seeded) — re-drawn by the oracle with the same seed."""
    args = SimpleNamespace(**g["args"])
    log = []
    torch.manual_seed(g["gumbel_seed"])
    with torch.no_grad():
        resvit_oracle.resvit_forward(g["state_dict"], args, g["img"], g["labels"], training=True, noise_log=log)
    it = iter(log)
    for layer in m.layers:
        if hasattr(layer, "router"):
            layer.router.noise_fn = lambda logits, it=it: next(it).to(logits.device)


@pytest.mark.parametrize("variant", ["bs2", "bs1"])
def test_resvit_train_fp32_mode_matches_reference_golden(variant):
    import vitb200
    g = torch.load(GOLD)[variant]
    m = _build(g).train()
    _replay_noise(m, g)
    t = g["train"]
    with vitb200.precision("fp32"):
        c, a, d, e, metric = m(g["img"].cuda(), g["labels"].cuda())
        (1.0 * c + 2.0 * a + 0.5 * d + 0.1 * e).backward()
    torch.cuda.synchronize()
    acts = torch.cat([w.float() for w in m.acts], -1).cpu()
    assert torch.equal(acts, t["acts"]), "router keep/skip decisions must be bit-exact"
    assert rel_l2(m.logits.cpu(), t["logits"]) < 1e-4
    for got, key in ((c, "c"), (a, "a"), (d, "d"), (e, "e")):
        assert abs(float(got) - float(t[key])) < 1e-4 * max(1.0, abs(float(t[key]))), key
    assert abs(float(metric["non_low_rank_ratio"]) - t["metric"]) < 1e-6
    named = dict(m.named_parameters())
    assert sorted(k for k, p in named.items() if p.requires_grad) == t["trainable"]
    for k, ref in t["grads"].items():
        assert named[k].grad is not None, k
        assert grad_close(named[k].grad.cpu(), ref, 1e-4, atol=1e-8), (k, rel_l2(named[k].grad.cpu(), ref))
    for k, p in named.items():
        if not p.requires_grad:
            assert p.grad is None, k


@pytest.mark.parametrize("variant", ["bs2", "bs1"])
def test_resvit_eval_indices_bit_exact_and_logits(variant):
    import vitb200
    g = torch.load(GOLD)[variant]
    m = _build(g).eval()
    args = SimpleNamespace(**g["args"])
    with torch.no_grad():
        ref = resvit_oracle.resvit_forward(g["state_dict"], args, g["img"], g["labels"], training=False)
        with vitb200.precision("fp32"):
            c, a, d, e, metric = m(g["img"].cuda(), g["labels"].cuda())
            idx32 = {bid: None for bid in ref["indices"]}
            logits32 = m.logits.cpu()
            acts32 = torch.cat([w.float() for w in m.acts], -1).cpu()
        with vitb200.precision("bf16"):
            m(g["img"].cuda(), g["labels"].cuda())
            logits16 = m.logits.cpu()
            acts16 = torch.cat([w.float() for w in m.acts], -1).cpu()
    assert torch.equal(acts32, g["eval"]["acts"])
    assert rel_l2(logits32, g["eval"]["logits"]) < 1e-4
    assert abs(float(e) - float(g["eval"]["e"])) < 1e-4
    # bf16: decisions may only flip where the two router logits are within bf16 noise of a tie
    agree = float((acts16 == g["eval"]["acts"]).float().mean())
    assert agree > 0.97, agree
    # images are independent of each other: every image whose decisions all agree must meet the 2e-2 bar
    same = (acts16 == g["eval"]["acts"]).flatten(1).all(1)
    assert int(same.sum()) >= 1, "no image kept all of its decisions"
    assert rel_l2(logits16[same], g["eval"]["logits"][same]) < 2e-2


def test_router_module_standalone_indices_and_gradients():
    import vitb200
    from vitb200 import resvit
    torch.manual_seed(0)
    r = resvit.RouterModule(256, 128, 1, 1e-5, block_size=2, use_lora=False)
    with torch.no_grad():
        r.out_conv[-1].weight.normal_(0, 0.5)
        r.out_conv[-1].bias.zero_()
    sd = {"router." + k: v.detach().clone() for k, v in r.state_dict().items()}
    args = SimpleNamespace(block_size=2, dynamic_reserve_initials=1, norm_eps=1e-5)
    x = torch.randn(3, 50, 256)
    log = []
    torch.manual_seed(11)
    xo = x.clone().requires_grad_(True)
    leaf = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    hard, idx, ent, soft = resvit_oracle.router(xo, leaf, "router.", args, True, log)
    wts = torch.linspace(0.5, 1.5, hard.numel()).view_as(hard)
    ((hard * wts).sum() + 3.0 * ent + (soft * wts.flip(0)).sum()).backward()
    r = r.cuda().train()
    r.noise_fn = lambda logits: log[0].to(logits.device)
    with vitb200.precision("fp32"):
        xc = x.cuda().requires_grad_(True)
        h2, i2, e2, s2 = r(xc)
        ((h2 * wts.cuda()).sum() + 3.0 * e2 + (s2 * wts.flip(0).cuda()).sum()).backward()
    assert torch.equal(i2.cpu(), idx.detach()), "packed router indices must be bit-exact"
    assert torch.equal(h2.detach().cpu(), hard.detach().round())
    assert rel_l2(s2.detach().cpu(), soft.detach()) < 1e-4
    assert abs(float(e2) - float(ent)) < 1e-5
    assert rel_l2(xc.grad.cpu(), xo.grad) < 1e-4
    for k, p in r.named_parameters():
        assert grad_close(p.grad.cpu(), leaf["router." + k].grad, 1e-4, atol=1e-8), k
    r.eval()
    with torch.no_grad(), vitb200.precision("fp32"):
        h3, i3, _, _ = r(x.cuda())
        ho, io, _, _ = resvit_oracle.router(x, sd, "router.", args, False, None)
    assert torch.equal(i3.cpu(), io)


def test_lora_module_and_attention_with_lora_match_oracle():
    import vitb200
    from vitb200 import resvit
    torch.manual_seed(1)
    args = resvit.ModelArgs(dim=256, n_heads=4, n_kv_heads=4, use_lora=True, lora_rank=8)
    att = resvit.Attention(args)
    with torch.no_grad():
        for n, p in att.named_parameters():
            if "lora" in n:
                p.mul_(20.0)        # make the rank-8 update visible next to the base projection
    sd = {"a." + k: v.detach().clone() for k, v in att.state_dict().items()}
    x = torch.randn(2, 50, 256)
    ref = resvit_oracle.attention(x, x, sd, "a.", 4, True)
    att = att.cuda()
    with torch.no_grad():
        with vitb200.precision("fp32"):
            y32 = att(x.cuda()).cpu()
            lo = att.lora_q(x.cuda()).cpu()
        with vitb200.precision("bf16"):
            y16 = att(x.cuda()).float().cpu()
            y_asym = att(x.cuda()[:, :20], x.cuda()).float().cpu()
    assert rel_l2(y32, ref) < 1e-4
    assert rel_l2(y16, ref) < 2e-2
    assert rel_l2(lo, (x @ sd["a.lora_q.lora_A.weight"].t()) @ sd["a.lora_q.lora_B.weight"].t()) < 1e-4
    assert rel_l2(y_asym, resvit_oracle.attention(x[:, :20], x, sd, "a.", 4, True)) < 2e-2


# ---------------------------------------------------------------------------------------------------
# BASELINE.json configs[4] geometry (D = 768, 197 tokens, router hidden 512, approximator rank 256, LoRA rank 8)
# ---------------------------------------------------------------------------------------------------
GOLD_C5 = os.path.join(ROOT, "tests", "golden", "resvit_c5.pt")


def _build_c5(g):
    """Re-creates the golden run's weights: the product's constructors reproduce the reference state_dict bit for bit
    under the same seed (tests/test_resvit_oracle.py); the fingerprints stored with the golden vectors are checked."""
    from vitb200 import resvit
    from oracle.make_golden_c5 import c5_edit_
    torch.manual_seed(g["seed"])
    m = resvit.Transformer(resvit.ModelArgs(**g["args"]))
    c5_edit_(m, g["edit_seed"])
    for k, v in m.state_dict().items():
        fp = g["weights_fp"][k]
        assert tuple(v.shape) == fp["shape"] and abs(float(v.double().sum()) - fp["sum"]) <= 1e-6 * max(1.0, fp["abs"]), k
    return m.cuda()


def _replay_c5_noise(m, g):
    it = iter(g["train"]["noise"])
    for layer in m.layers:
        if hasattr(layer, "router"):
            layer.router.noise_fn = lambda logits, it=it: next(it).to(logits.device)


def _fp_close(t, fp, rtol, atol=1e-8):
    """gradient vs its stored fingerprint: norm and the first 16 values."""
    n = float(t.double().norm())
    if abs(n - fp["norm"]) > rtol * fp["norm"] + atol * (t.numel() ** 0.5):
        return False
    head = t.detach().flatten()[:16].float().cpu()
    return float((head - fp["head"]).norm()) <= rtol * max(float(fp["head"].norm()), fp["norm"] / (t.numel() ** 0.5)) * 4 + atol


@pytest.mark.parametrize("variant", ["bs1", "bs2"])
def test_resvit_c5_geometry_train_fp32_mode(variant):
    import vitb200
    from conftest import grad_close
    from oracle import make_golden_c5
    g = torch.load(GOLD_C5)[variant]
    m = _build_c5(g).train()
    _replay_c5_noise(m, g)
    img, labels = make_golden_c5.c5_inputs()
    t = g["train"]
    with vitb200.precision("fp32"):
        c, a, d, e, metric = m(img.cuda(), labels.cuda())
        (1.0 * c + 2.0 * a + 0.5 * d + 0.1 * e).backward()
    torch.cuda.synchronize()
    acts = torch.cat([w.float() for w in m.acts], -1).cpu()
    assert torch.equal(acts, t["acts"]), "router keep/skip decisions must be bit-exact"
    assert rel_l2(m.logits.cpu(), t["logits"]) < 1e-4
    for got, key in ((c, "c"), (a, "a"), (d, "d"), (e, "e")):
        assert abs(float(got) - float(t[key])) < 1e-4 * max(1.0, abs(float(t[key]))), key
    assert abs(float(metric["non_low_rank_ratio"]) - t["metric"]) < 1e-6
    named = dict(m.named_parameters())
    assert sorted(k for k, p in named.items() if p.requires_grad) == t["trainable"]
    for k, fp in t["grads_fp"].items():
        assert named[k].grad is not None, k
        assert _fp_close(named[k].grad, fp, 2e-4), (k, float(named[k].grad.double().norm()), fp["norm"])
    for k, ref in t["grads"].items():
        assert grad_close(named[k].grad.cpu(), ref, 2e-4, atol=1e-8), (k, rel_l2(named[k].grad.cpu(), ref))


@pytest.mark.parametrize("variant", ["bs1", "bs2"])
def test_resvit_c5_geometry_bf16_train_and_eval(variant):
    """bf16 mode at the benchmarked geometry: decisions may flip only where the two router logits are within bf16 noise
    of a tie; every image that kept all its decisions meets the 2e-2 logits bar (train with the replayed Gumbel draw,
    and eval); the losses follow."""
    import vitb200
    from oracle import make_golden_c5
    g = torch.load(GOLD_C5)[variant]
    m = _build_c5(g)
    img, labels = make_golden_c5.c5_inputs()
    for mode in ("train", "eval"):
        t = g[mode]
        m.train(mode == "train")
        if mode == "train":
            _replay_c5_noise(m, g)
        with vitb200.precision("bf16"):
            if mode == "train":
                c, a, d, e, metric = m(img.cuda(), labels.cuda())
                (1.0 * c + 2.0 * a + 0.5 * d + 0.1 * e).backward()
            else:
                with torch.no_grad():
                    c, a, d, e, metric = m(img.cuda(), labels.cuda())
        torch.cuda.synchronize()
        acts = torch.cat([w.float() for w in m.acts], -1).cpu()
        agree = float((acts == t["acts"]).float().mean())
        assert agree > 0.97, (mode, agree)
        same = (acts == t["acts"]).flatten(1).all(1)
        if int(same.sum()) > 0:
            assert rel_l2(m.logits.float().cpu()[same], t["logits"][same]) < 2e-2, mode
        assert abs(float(metric["non_low_rank_ratio"]) - t["metric"]) < 0.02
        if agree == 1.0:
            assert abs(float(c) - float(t["c"])) < 2e-2 * max(1.0, abs(float(t["c"]))), mode
        if mode == "train":
            for k, p in m.named_parameters():
                assert (p.grad is not None) == p.requires_grad, k
                if p.grad is not None:
                    assert bool(torch.isfinite(p.grad).all()), k
            m.zero_grad(set_to_none=True)


def test_resvit_loss_kernels_match_the_reference_formulas():
    """vitb_distill_loss / vitb_active_loss (value and gradient from one kernel) against res-vit/model.py:40-59, :61-85."""
    from vitb200 import resvit
    g = torch.Generator().manual_seed(8)
    # DistillLoss on strided class-token rows, fp32 and bf16
    for dt, tol in ((torch.float32, 1e-6), (torch.bfloat16, 1e-6)):
        s_all = torch.randn(5, 7, 64, generator=g).to(dt)
        t_all = torch.randn(5, 7, 64, generator=g).to(dt)
        sr = s_all.float().clone().requires_grad_(True)
        want = torch.nn.functional.mse_loss(sr[:, 0, :], t_all.float()[:, 0, :])
        want.backward()
        sc = s_all.cuda().requires_grad_(True)
        got = resvit.DistillLoss()(sc[:, 0, :], t_all.cuda()[:, 0, :])
        (3.0 * got).backward()
        assert abs(float(got) - float(want)) <= tol * max(1.0, abs(float(want)))
        assert rel_l2(sc.grad.float().cpu(), 3.0 * sr.grad) < (1e-6 if dt == torch.float32 else 4e-3)
    # ActiveLoss: masked mean over the non-reserved tokens, squared distance to the target
    a = torch.rand(3, 9, 4, generator=g)
    ar = a.clone().requires_grad_(True)
    want = (ar[:, 1:, :].mean() - 0.4) ** 2
    want.backward()
    ac = a.cuda().requires_grad_(True)
    crit = resvit.ActiveLoss(0.4, 1)
    got = crit(ac)
    got.backward()
    assert abs(float(got) - float(want)) < 1e-7
    assert rel_l2(ac.grad.cpu(), ar.grad) < 1e-6
    assert abs(float(crit.metric(ac)["non_low_rank_ratio"]) - float(a[:, 1:, :].mean())) < 1e-6


def test_row_compaction_kernels():
    """vitb_compact_rows / gather / scatter and the GEMM's device row count (vitb_gemm_params.m_dev)."""
    import vitb200
    from vitb200 import ops
    g = torch.Generator().manual_seed(21)
    T, D = 1000, 256
    index = torch.randint(0, 4, (T, 1), generator=g).float().cuda()
    rows, count = ops.compact_rows(index, [1, 3])
    torch.cuda.synchronize()
    want = torch.nonzero((index.view(-1) == 1) | (index.view(-1) == 3)).view(-1).int()
    n = int(count)
    assert n == want.numel()
    assert torch.equal(rows[:n].sort().values.cpu(), want.cpu())
    for dt in (torch.float32, torch.bfloat16):
        src = torch.randn(T, D, generator=g).to(dt).cuda()
        got = ops.gather_rows(src, rows, count)
        assert torch.equal(got[:n], src[rows[:n].long()])
        dst = torch.zeros(T, D, dtype=dt, device="cuda")
        ops.scatter_rows(got, rows, count, dst)
        ref = torch.zeros_like(dst)
        ref[rows[:n].long()] = src[rows[:n].long()]
        assert torch.equal(dst, ref)
    # a GEMM over `count` rows equals the dense GEMM on those rows; tiles past the count are skipped (sentinel survives)
    A = (torch.randn(T, D, generator=g) * 0.5).to(torch.bfloat16).cuda()
    W = (torch.randn(512, D, generator=g) * 0.1).to(torch.bfloat16).cuda()
    bias = torch.randn(512, generator=g).cuda()
    dense = ops.gemm(A, W, bias=bias)
    out = torch.full((T, 512), 7.0, dtype=torch.bfloat16, device="cuda")
    ops.gemm(A, W, bias=bias, out=out, m_dev=count)
    torch.cuda.synchronize()
    assert torch.equal(out[:n], dense[:n])
    first_skipped = (n + 127) // 128 * 128
    assert bool((out[first_skipped:] == 7.0).all())


@pytest.mark.parametrize("variant", ["bs2", "bs1"])
def test_resvit_eval_token_compaction_equals_dense_select(variant, monkeypatch):
    """Inference with device-side token compaction (output projection + MLP on the active rows only) gives the logits of
    the dense-select path: the skipped rows' results are never used (res-vit/model.py:503-524)."""
    import vitb200
    g = torch.load(GOLD)[variant]
    m = _build(g).eval()
    res = {}
    for flag in ("0", "1"):
        monkeypatch.setenv("VITB_RESVIT_COMPACT", flag)
        with torch.no_grad(), vitb200.precision("bf16"):
            m(g["img"].cuda(), g["labels"].cuda())
        torch.cuda.synchronize()
        res[flag] = (m.logits.float().cpu(), torch.cat([w.float() for w in m.acts], -1).cpu())
    assert torch.equal(res["1"][1], res["0"][1])
    assert rel_l2(res["1"][0], res["0"][0]) < 1e-5
