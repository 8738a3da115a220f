"""Host-side pieces of the training-step row that need no GPU: the device-resident metric accumulator
against the reference's `accuracy` + MetricTracker arithmetic (src/utils.py:28-40,79-100; src/train.py:27-32)."""
import importlib.util
import os

import pytest
import torch

import vitb200

REF_UTILS = "/root/reference/src/utils.py"


def _ref_accuracy(output, target, topk=(1,)):
    """src/utils.py:28-40, restated (the module imports swanlab / pandas at load time)."""
    maxk = max(topk)
    batch_size = target.size(0)
    _, pred = output.topk(maxk, 1, True, True)
    pred = pred.t()
    correct = pred.eq(target.view(1, -1).expand_as(pred))
    return [correct[:k].reshape(-1).float().sum(0) / batch_size * 100.0 for k in topk]


def test_device_metrics_equal_reference_tracker():
    g = torch.Generator().manual_seed(0)
    meter = vitb200.train.DeviceMetrics(device="cpu")
    tot = {"loss": 0.0, "acc1": 0.0, "acc5": 0.0}
    batches = [(128, 100), (128, 100), (37, 100), (5, 100)]
    for B, C in batches:
        logits = torch.randn(B, C, generator=g)
        logits[0, :7] = 3.0                                           # ties inside the top-5: same topk call, same answer
        labels = torch.randint(0, C, (B,), generator=g)
        loss = torch.nn.functional.cross_entropy(logits, labels)
        meter.update(loss, logits, labels)
        a1, a5 = _ref_accuracy(logits, labels, topk=(1, 5))
        tot["loss"] += loss.item(); tot["acc1"] += a1.item(); tot["acc5"] += a5.item()
    res = meter.result()
    for k in tot:                                                      # MetricTracker: plain mean of the per-batch values
        assert abs(res[k] - tot[k] / len(batches)) < 1e-5, k
    meter.reset()
    assert meter.steps == 0 and float(meter.sums.abs().sum()) == 0.0


def test_device_metrics_with_fewer_classes_than_k():
    meter = vitb200.train.DeviceMetrics(device="cpu")
    logits = torch.tensor([[0.1, 0.9, 0.0], [0.8, 0.1, 0.1]])
    labels = torch.tensor([1, 2])
    meter.update(torch.tensor(0.5), logits, labels)
    res = meter.result()
    assert res["acc1"] == 50.0 and res["acc5"] == 100.0 and res["loss"] == 0.5


@pytest.mark.skipif(not os.path.exists(REF_UTILS), reason="reference not mounted")
def test_restated_accuracy_is_the_reference_source():
    src = open(REF_UTILS).read()
    assert "_, pred = output.topk(maxk, 1, True, True)" in src and "correct_k / batch_size * 100.0" in src


def test_packed_order_keeps_packs_together_and_uniform_stack_detects_strides():
    """functional.mark_packed / packed_order: the q, k, v weights of a block end up back to back in the fused optimizers'
    flat buffers (order of everything else unchanged); _uniform_stack turns same-shape views at one distance into ONE
    [G, ...] strided view (what the grouped GEMM reads) and refuses anything else."""
    from vitb200 import functional as F
    a, b, c, d, e = (torch.nn.Parameter(torch.zeros(4, 3)) for _ in range(5))
    F.mark_packed(b, d, e)
    order = F.packed_order([a, b, c, d, e])
    assert [id(t) for t in order] == [id(t) for t in (a, b, d, e, c)]
    assert [id(t) for t in F.packed_order([a, b, c, d])] == [id(t) for t in (a, b, c, d)]     # an incomplete pack is left alone
    flat = torch.arange(64, dtype=torch.float32)
    v = [flat[0:12].view(4, 3), flat[16:28].view(4, 3), flat[32:44].view(4, 3)]
    st = F._uniform_stack(v)
    assert st is not None and st.shape == (3, 4, 3) and st.stride(0) == 16
    assert torch.equal(st[2], v[2]) and st.data_ptr() == flat.data_ptr()
    assert F._uniform_stack([flat[0:12].view(4, 3), flat[16:28].view(4, 3), flat[36:48].view(4, 3)]) is None   # uneven distance
    assert F._uniform_stack([flat[0:12].view(4, 3), torch.zeros(4, 3)]) is None                               # different storage
    assert F._uniform_stack([flat[0:12].view(4, 3), flat[6:18].view(4, 3)]) is None                           # overlapping


def test_bench_flop_tables_match_the_survey():
    """bench.py's per-image FLOP counts: SURVEY.md App. A gives 35.126 GFLOP forward for ViT-B/16 at 224 px, C = 100
    (105.379 for forward + backward); the executed count is smaller by the last block's skipped rows; ViT-H/14 has 257
    tokens, 384 px gives 577."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("_bench", os.path.join(os.path.dirname(os.path.dirname(__file__)), "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    fwd, ex = bench.vit_gflop("b16", 224, 100)
    assert abs(3 * fwd - bench.GFLOP_PER_IMG_TRAIN) < 2e-3 and abs(fwd - 35.126) < 1e-3
    assert 0.92 * fwd < ex < 0.94 * fwd
    f384, _ = bench.vit_gflop("b16", 384, 100)
    assert 3.0 * fwd < f384 < 3.4 * fwd          # 577 / 197 = 2.93 x the tokens, attention grows quadratically
    assert set(bench.CONFIGS) >= {"c2", "c3", "c4", "c5"}
    assert bench.CONFIGS["c2"]["batch"] == 128 and bench.CONFIGS["c3"]["batch"] == 64 and bench.CONFIGS["c4"]["batch"] == 256


def test_nvlink_exchange_and_dropout_sites_fail_loudly_or_stay_lazy_without_a_gpu():
    """p2p.NvlinkExchange needs an initialised process group (it is a collective); dropout sites hold no device state
    until a module is actually trained on a device, and never appear in a state dict."""
    with pytest.raises(RuntimeError):
        vitb200.p2p.NvlinkExchange()
    m = vitb200.VisionTransformer(image_size=(32, 32), patch_size=(16, 16), emb_dim=128, mlp_dim=256, num_heads=2, num_layers=2,
                                  num_classes=10)                        # ctor default dropout_rate = 0.1
    sites = [mod._site for mod in m.modules() if hasattr(mod, "_site")]
    assert len(sites) >= 3 and all(s._state is None for s in sites)
    assert len({s.index for s in sites}) == len(sites)
    assert not any("site" in k or "drop" in k for k in m.state_dict())
    assert vitb200.functional.dropout(torch.ones(3), 0.1, False, sites[0]) is not None      # identity when not training
