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
