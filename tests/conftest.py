import importlib.util
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run on the GPU box with -m gpu)")
    # make sure the C-ABI library exists (nvcc cross-compiles without a GPU)
    lib = os.path.join(ROOT, "vit-of-pytorch_b200", "libvitb200.so")
    if not os.path.exists(lib):
        spec = importlib.util.spec_from_file_location("_vitb_build", os.path.join(ROOT, "vit-of-pytorch_b200", "build.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        mod.build()


def pytest_collection_modifyitems(config, items):
    import torch

    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def rel_l2(a, b):
    a = a.double().flatten()
    b = b.double().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def grad_close(a, b, rtol, atol=1e-7):
    """||a-b|| <= rtol*||b|| + atol*sqrt(numel): mathematically-zero gradients (e.g. the key bias, to which
    softmax is invariant) carry only rounding noise, so a pure relative test is meaningless for them."""
    a = a.double().flatten()
    b = b.double().flatten()
    return float((a - b).norm()) <= rtol * float(b.norm()) + atol * (b.numel() ** 0.5)
