"""GPU tests of the training-step plumbing: the CUDA-graph step equals the eager step, the LR scheduler
keeps driving a captured graph, and the Res-ViT fine-tune step (FusedAdamW + on-device grad clipping) matches
torch.optim.AdamW + clip_grad_norm_ applied to the oracle's gradients."""
import os
import sys
from types import SimpleNamespace

import pytest
import torch

from conftest import grad_close, rel_l2

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import resvit_oracle, vit_init  # noqa: E402

pytestmark = pytest.mark.gpu


def _tiny(seed=0):
    import vitb200
    cfg = dict(image_size=(64, 64), patch_size=(16, 16), emb_dim=128, mlp_dim=256, num_heads=2, num_layers=3, num_classes=16)
    sd = vit_init.reference_state_dict(cfg, seed=seed, scaled=True)
    m = vitb200.VisionTransformer(dropout_rate=0.0, attn_dropout_rate=0.0, **cfg)
    m.load_state_dict(sd)
    return m.cuda().train()


def test_graphed_step_matches_eager_and_follows_lr_schedule():
    import vitb200
    g = torch.Generator().manual_seed(3)
    img = torch.randn(8, 3, 64, 64, generator=g).cuda()
    lab = torch.randint(0, 16, (8,), generator=g).cuda()
    # eager trajectory: 6 steps, StepLR halves... (x0.1 every 2 steps)
    m = _tiny()
    opt = vitb200.optim.FusedSGD(m.parameters(), lr=0.05, momentum=0.9)
    sched = torch.optim.lr_scheduler.StepLR(opt, step_size=2, gamma=0.1)
    eager = []
    for i in range(6):
        opt.zero_grad()
        loss = vitb200.functional.cross_entropy(m(img), lab)
        loss.backward()
        opt.step()
        sched.step()
        eager.append(float(loss))
    w_eager = {k: v.detach().clone() for k, v in m.state_dict().items()}
    # graph trajectory: the constructor runs ONE eager warm-up step (step 0; capture itself executes nothing),
    # then 5 replays are steps 1..5; the scheduler keeps driving the captured step through the device-side lr
    m = _tiny()
    opt = vitb200.optim.FusedSGD(m.parameters(), lr=0.05, momentum=0.9)
    sched = torch.optim.lr_scheduler.StepLR(opt, step_size=2, gamma=0.1)
    step = vitb200.train.GraphedTrainStep(m, opt, img, lab, warmup=1)
    sched.step()
    graph = []
    for i in range(5):
        graph.append(float(step(img, lab)))
        sched.step()
    torch.cuda.synchronize()
    for a, b in zip(eager[1:], graph):
        assert abs(a - b) <= 1e-2 * max(1.0, abs(a)), (eager, graph)
    for k, v in m.state_dict().items():
        assert rel_l2(v, w_eager[k]) < 2e-2 or float((v - w_eager[k]).abs().max()) < 1e-3, k


def test_resvit_adamw_step_matches_torch_adamw_on_oracle_grads():
    import vitb200
    from vitb200 import resvit
    g = torch.load(os.path.join(ROOT, "tests", "golden", "resvit_tiny.pt"))["bs1"]
    kw = dict(g["args"]); kw["device"] = "cuda"
    m = resvit.Transformer(resvit.ModelArgs(**kw))
    m.load_state_dict(g["state_dict"])
    m = m.cuda().train()
    args = SimpleNamespace(**g["args"])
    log = []
    torch.manual_seed(g["gumbel_seed"])
    leaf = {k: v.clone().requires_grad_(k in g["train"]["trainable"]) for k, v in g["state_dict"].items()}
    out = resvit_oracle.resvit_forward(leaf, args, g["img"], g["labels"], training=True, noise_log=log)
    (out["c_loss"] + out["a_loss"] + out["d_loss"]).backward()
    # torch's AdamW skips parameters whose grad is None (approximators whose key did not occur in the batch);
    # compare the parameters the reference step actually touches
    train_keys = [k for k in g["train"]["trainable"] if leaf[k].grad is not None]
    ref_params = [leaf[k].detach().clone().requires_grad_(True) for k in train_keys]
    for p, k in zip(ref_params, train_keys):
        p.grad = leaf[k].grad.clone()
    ref_opt = torch.optim.AdamW(ref_params, lr=1e-3, weight_decay=0.05)
    torch.nn.utils.clip_grad_norm_(ref_params, 1.0)
    ref_opt.step()
    it = iter(log)
    for layer in m.layers:
        if hasattr(layer, "router"):
            layer.router.noise_fn = lambda logits, it=it: next(it).to(logits.device)
    named = dict(m.named_parameters())
    opt = vitb200.optim.FusedAdamW([named[k] for k in train_keys], lr=1e-3, weight_decay=0.05, max_grad_norm=1.0)
    with vitb200.precision("fp32"):
        opt.zero_grad()
        c, a, d, e, _ = m(g["img"].cuda(), g["labels"].cuda())
        (c + a + d).backward()
        opt.step()
    torch.cuda.synchronize()
    for p, k in zip(ref_params, train_keys):
        delta_ref = p.detach() - g["state_dict"][k]
        delta = named[k].detach().cpu() - g["state_dict"][k]
        assert grad_close(delta, delta_ref, 5e-3, atol=2e-5), (k, rel_l2(delta, delta_ref))


def test_input_prefetcher_delivers_batches_in_order_into_the_graph_buffers():
    """Batches staged from pinned host memory on the side stream arrive in order, and with `into=` they land in the
    static buffers a GraphedTrainStep captured (so the replay needs no further copy)."""
    import vitb200
    g = torch.Generator().manual_seed(11)
    m = _tiny()
    opt = vitb200.optim.FusedSGD(m.parameters(), lr=0.01, momentum=0.9)
    img = torch.randn(8, 3, 64, 64, generator=g).cuda()
    lab = torch.randint(0, 16, (8,), generator=g).cuda()
    step = vitb200.train.GraphedTrainStep(m, opt, img, lab, warmup=1)
    pre = vitb200.train.InputPrefetcher(img, lab, into=(step.images, step.labels))
    host = [(torch.randn(8, 3, 64, 64, generator=g).pin_memory(), torch.randint(0, 16, (8,), generator=g).pin_memory())
            for _ in range(5)]
    with pytest.raises(RuntimeError):
        pre.get()
    pre.start(*host[0])
    losses = []
    for i in range(5):
        x, y = pre.get()
        assert x.data_ptr() == step.images.data_ptr()
        if i + 1 < 5:
            pre.start(*host[i + 1])
        losses.append(step(x, y))
        torch.cuda.synchronize()
        assert torch.equal(step.images.cpu(), host[i][0]) and torch.equal(step.labels.cpu(), host[i][1])
    assert all(torch.isfinite(l) for l in losses)
    # without `into` the prefetcher owns the buffers it hands out
    pre2 = vitb200.train.InputPrefetcher(img, lab)
    pre2.start(*host[3]); pre2.start(*host[4])
    with pytest.raises(RuntimeError):
        pre2.start(*host[0])
    a, _ = pre2.get()
    torch.cuda.synchronize()
    assert torch.equal(a.cpu(), host[3][0])
    b, _ = pre2.get()
    torch.cuda.synchronize()
    assert torch.equal(b.cpu(), host[4][0])


def test_graphed_step_follows_onecycle_momentum():
    """OneCycleLR (the reference's schedule, src/train.py:159-163) cycles the MOMENTUM as well as the learning rate; the
    captured step reads both from device scalars, so replayed training equals eager training."""
    import vitb200
    g = torch.Generator().manual_seed(5)
    img = torch.randn(8, 3, 64, 64, generator=g).cuda()
    lab = torch.randint(0, 16, (8,), generator=g).cuda()

    def make():
        m = _tiny()
        opt = vitb200.optim.FusedSGD(m.parameters(), lr=0.05, momentum=0.9)
        sched = torch.optim.lr_scheduler.OneCycleLR(opt, max_lr=0.05, total_steps=8, pct_start=0.4)
        return m, opt, sched

    m, opt, sched = make()
    moms = []
    for i in range(6):
        moms.append(opt.param_groups[0]["momentum"])
        opt.zero_grad()
        vitb200.functional.cross_entropy(m(img), lab).backward()
        opt.step()
        sched.step()
    assert max(moms) - min(moms) > 0.05, "the schedule is expected to move the momentum"
    w_eager = {k: v.detach().clone() for k, v in m.state_dict().items()}
    m, opt, sched = make()
    step = vitb200.train.GraphedTrainStep(m, opt, img, lab, warmup=1)
    sched.step()
    for i in range(5):
        step(img, lab)
        sched.step()
    torch.cuda.synchronize()
    for k, v in m.state_dict().items():   # (the key bias only ever receives rounding noise: softmax is invariant to it)
        assert grad_close(v, w_eager[k], 2e-3, atol=1e-6), (k, rel_l2(v, w_eager[k]))


def test_backward_side_channel_hits_and_survives_a_second_consumer(monkeypatch):
    """The bf16 copy / column sums that a fused block leaves for the block upstream are used when that block receives
    exactly the tensor (3 blocks -> 2 hits), and are rejected when autograd has added a second consumer's gradient into
    the buffer (an auxiliary loss on a block's input): gradients then still equal the run without the side channel."""
    import vitb200
    from vitb200 import functional as F
    g = torch.Generator().manual_seed(9)
    img = torch.randn(4, 3, 64, 64, generator=g).cuda()
    lab = torch.randint(0, 16, (4,), generator=g).cuda()

    def run(aux, side):
        monkeypatch.setenv("VITB_GRAD_SIDE", side)
        m = _tiny(seed=1)
        feats = []
        if aux:
            m.transformer.encoder_layers[1].register_forward_hook(lambda mod, inp, out: feats.append(out))
        F.SIDE_STATS[0] = F.SIDE_STATS[1] = 0
        loss = vitb200.functional.cross_entropy(m(img), lab)
        if aux:
            loss = loss + 0.01 * feats[0].float().pow(2).mean()
        loss.backward()
        torch.cuda.synchronize()
        return {k: p.grad.detach().clone() for k, p in m.named_parameters()}, tuple(F.SIDE_STATS)

    # the last block computes only the class-token row (no fused node), so 2 fused blocks -> 1 hand-over
    g_on, stats = run(False, "1")
    assert stats[0] >= 1, stats
    g_off, _ = run(False, "0")
    for k in g_on:
        assert grad_close(g_on[k], g_off[k], 2e-2, atol=1e-6), k
    g_aux_on, stats_aux = run(True, "1")
    g_aux_off, _ = run(True, "0")
    for k in g_aux_on:
        assert grad_close(g_aux_on[k], g_aux_off[k], 2e-2, atol=1e-6), (k, stats_aux)


def test_fused_optimizer_state_dict_round_trip_and_zero_grad_misuse():
    import vitb200
    g = torch.Generator().manual_seed(4)
    img = torch.randn(4, 3, 64, 64, generator=g).cuda()
    lab = torch.randint(0, 16, (4,), generator=g).cuda()

    def steps(m, opt, n):
        for _ in range(n):
            opt.zero_grad()
            vitb200.functional.cross_entropy(m(img), lab).backward()
            opt.step()

    for make in (lambda ps: vitb200.optim.FusedSGD(ps, lr=0.05, momentum=0.9),
                 lambda ps: vitb200.optim.FusedAdamW(ps, lr=1e-3, weight_decay=0.05, max_grad_norm=1.0)):
        m1 = _tiny()
        o1 = make(m1.parameters())
        steps(m1, o1, 2)
        sd_m = {k: v.detach().clone() for k, v in m1.state_dict().items()}
        sd_o = o1.state_dict()
        assert len(sd_o["state"]) == len(list(m1.parameters()))
        steps(m1, o1, 2)
        m2 = _tiny()
        o2 = make(m2.parameters())
        m2.load_state_dict(sd_m)
        o2.load_state_dict(sd_o)
        steps(m2, o2, 2)
        torch.cuda.synchronize()
        for (k, a), (_, b) in zip(m1.state_dict().items(), m2.state_dict().items()):
            assert rel_l2(b, a) < 1e-5, k
    # model.zero_grad() hides the flat gradient buffer without zeroing it: step() must refuse
    m = _tiny()
    opt = vitb200.optim.FusedSGD(m.parameters(), lr=0.05, momentum=0.9)
    steps(m, opt, 1)
    m.zero_grad(set_to_none=True)
    vitb200.functional.cross_entropy(m(img), lab).backward()
    with pytest.raises(RuntimeError, match="optimizer.zero_grad"):
        opt.step()


@pytest.mark.parametrize("width,heads", [(128, 2), (256, 4)])
def test_fused_last_block_and_direct_bias_sums_match_the_composed_path(width, heads, monkeypatch):
    """The class-token-only last block as ONE autograd node (functional._EncoderBlockRow0: k | v as a grouped GEMM, bf16
    single-query attention backward into the packed buffer, LayerNorm backward with a residual gradient on rows 0 only) and
    the column sums that a block's LayerNorm backward adds straight into the upstream block's fc2 bias gradient must give the
    logits and gradients of the composed path (VITB_ROW0_FUSED=0, VITB_GRAD_SIDE=0) — with the fused optimizer's flat
    gradient buffer in place, which is what enables both, and also when an auxiliary loss adds a second consumer to a
    block's output (the hand-over is then rejected and the share added in advance is taken back out)."""
    import vitb200
    from vitb200 import functional as F
    cfg = dict(image_size=(64, 64), patch_size=(16, 16), emb_dim=width, mlp_dim=2 * width, num_heads=heads, num_layers=4,
               num_classes=16)
    sd = vit_init.reference_state_dict(cfg, seed=2, scaled=True)
    g = torch.Generator().manual_seed(11)
    img = torch.randn(6, 3, 64, 64, generator=g).cuda()
    lab = torch.randint(0, 16, (6,), generator=g).cuda()

    def run(fused, side, aux):
        monkeypatch.setenv("VITB_ROW0_FUSED", fused)
        monkeypatch.setenv("VITB_GRAD_SIDE", side)
        m = vitb200.VisionTransformer(dropout_rate=0.0, attn_dropout_rate=0.0, **cfg)
        m.load_state_dict(sd)
        m = m.cuda().train()
        opt = vitb200.optim.FusedSGD(m.parameters(), lr=0.01, momentum=0.9)
        opt.zero_grad()
        feats = []
        if aux:
            m.transformer.encoder_layers[1].register_forward_hook(lambda mod, inp, out: feats.append(out))
        logits = m(img)
        loss = vitb200.functional.cross_entropy(logits, lab)
        if aux:
            loss = loss + 0.05 * feats[0].float().pow(2).mean()
        loss.backward()
        torch.cuda.synchronize()
        return logits.detach().clone(), {k: p.grad.detach().clone() for k, p in m.named_parameters()}

    for aux in (False, True):
        lo_ref, g_ref = run("0", "0", aux)
        lo, gr = run("1", "1", aux)
        assert rel_l2(lo, lo_ref) < 5e-3, (aux, rel_l2(lo, lo_ref))
        for k in g_ref:
            assert grad_close(gr[k], g_ref[k], 2e-2, atol=1e-6), (aux, k, rel_l2(gr[k], g_ref[k]))


def test_fused_adamw_skips_parameters_without_a_gradient_like_torch():
    """torch.optim.AdamW leaves a parameter whose .grad is None alone: no weight decay, no moment decay, and ITS step
    count (the bias corrections) stands still.  FusedAdamW.register_skippable + the device flag reproduce that on the
    flat buffers: three steps, the middle one without a gradient for parameter B."""
    import vitb200
    g = torch.Generator().manual_seed(4)
    shapes = [(7, 130), (64, 33), (5,)]
    init = [torch.randn(s, generator=g) for s in shapes]
    grads = [[torch.randn(s, generator=g) * 0.1 for s in shapes] for _ in range(3)]
    ref_p = [t.clone().requires_grad_(True) for t in init]
    ref = torch.optim.AdamW(ref_p, lr=3e-3, betas=(0.9, 0.95), weight_decay=0.1)
    ps = [torch.nn.Parameter(t.clone().cuda()) for t in init]
    opt = vitb200.optim.FusedAdamW(ps, lr=3e-3, betas=(0.9, 0.95), weight_decay=0.1)
    flag = opt.register_skippable([ps[1]])
    for it in range(3):
        live_b = it != 1
        for p, gr, j in zip(ref_p, grads[it], range(3)):
            p.grad = None if (j == 1 and not live_b) else gr.clone()
        ref.step()
        opt.zero_grad()
        assert int(flag) == 0
        for p, gr, j in zip(ps, grads[it], range(3)):
            if j == 1 and not live_b:
                continue                      # the flat gradient stays zero, the flag stays 0
            p.grad.copy_(gr.cuda())
        if live_b:
            flag.fill_(1)
        opt.step()
        torch.cuda.synchronize()
        for p, r in zip(ps, ref_p):
            assert rel_l2(p.detach().cpu(), r.detach()) < 1e-6, it
    sd = opt.state_dict()["state"]
    assert float(sd[0]["step"]) == 3 and float(sd[1]["step"]) == 2 and float(sd[2]["step"]) == 3
    assert rel_l2(sd[1]["exp_avg"].cpu(), ref.state[ref_p[1]]["exp_avg"]) < 1e-6


def test_resvit_approximators_that_saw_no_token_are_left_alone_by_adamw():
    """res-vit/model.py:363-367 runs an approximator only for keys that occur in the batch, so the others have no gradient
    and torch's AdamW (res-vit/train.py:272-277) does not touch them.  With resvit.bind_optimizer the selection kernel
    reports which approximators saw a token and FusedAdamW skips the rest.  block_size 2 (keys 0, 1, 2), two steps: in the
    first only keys 0 and 3 occur, in the second keys 0 and 2; the reference loop is restated in plain torch."""
    import vitb200
    from vitb200 import resvit
    dim, rank, T = 128, 32, 96
    g = torch.Generator().manual_seed(8)
    mod = resvit.BlockPathApproximators(dim, rank, 2)
    for p in mod.parameters():
        p.data = torch.randn(p.shape, generator=g) * 0.05
    init = {k: v.detach().clone() for k, v in mod.named_parameters()}
    mod = mod.cuda().train()
    ref_p = {k: v.clone().requires_grad_(True) for k, v in init.items()}
    ref_opt = torch.optim.AdamW(list(ref_p.values()), lr=1e-2, weight_decay=0.1)
    opt = vitb200.optim.FusedAdamW(list(mod.parameters()), lr=1e-2, weight_decay=0.1)
    assert resvit.bind_optimizer(mod, opt) == 3
    named = dict(mod.named_parameters())
    xs = [torch.randn(2, T // 2, dim, generator=g) for _ in range(2)]
    idxs = [torch.tensor([0, 3] * (T // 2)).float().view(2, T // 2, 1), torch.tensor([2, 0, 0] * (T // 3)).float().view(2, T // 2, 1)]
    for step, (x, idx) in enumerate(zip(xs, idxs)):
        # reference semantics (res-vit/model.py:349-368) in plain torch, fp32
        for p in ref_p.values():
            p.grad = None
        out = x.clone()
        for key in (0, 1, 2):
            sub = idx.squeeze(-1) == key
            if sub.any():
                dw, uw = ref_p["approximators.%d.down_proj.weight" % key], ref_p["approximators.%d.up_proj.weight" % key]
                out = out.clone()
                out[sub] = out[sub] @ dw.t() @ uw.t() + out[sub]
        out.pow(2).mean().backward()
        ref_opt.step()
        with vitb200.precision("fp32"):
            opt.zero_grad()
            y = mod(x.cuda(), idx.cuda(), [0, 1, 2])
            y.float().pow(2).mean().backward()
            opt.step()
        torch.cuda.synchronize()
        absent = [k for k in (0, 1, 2) if not bool((idx == k).any())]
        assert absent, step
        for k, p in named.items():
            key = int(k.split(".")[1])
            if step == 0 and key in absent:
                assert torch.equal(p.detach().cpu(), init[k]), (step, k)       # never touched: bit-identical
            delta_ref = ref_p[k].detach() - init[k]
            delta = p.detach().cpu() - init[k]
            assert grad_close(delta, delta_ref, 5e-3, atol=1e-6), (step, k, rel_l2(delta, delta_ref))
    steps = {k: float(v["step"]) for k, v in zip(named, opt.state_dict()["state"].values())}
    assert steps["approximators.0.down_proj.weight"] == 2 and steps["approximators.1.down_proj.weight"] == 0 \
        and steps["approximators.2.up_proj.weight"] == 1, steps
