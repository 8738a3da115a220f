"""Pins oracle/image_prep_oracle.py (the input transform in front of the encoder).

Golden vectors: tests/golden/image_prep.npz, produced by oracle/make_golden_prep.py running the reference's own
loader classes (src/data_loaders.py:62-124) through the installed torchvision / Pillow.  Bar: BIT-EXACT — the
resize is integer arithmetic and the normalisation is a function of the resized byte."""
import os

import numpy as np
import pytest

from oracle import image_prep_oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden", "image_prep.npz")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


@pytest.mark.parametrize("case,src,size", [("cifar_train", "cifar_in", 224), ("cifar_eval", "cifar_in", 64),
                                           ("inet_train", "inet_in", (64, 64))])
def test_oracle_matches_reference_loader_output(gold, case, src, size):
    order = gold[case + "_order"]
    flip = gold[case + "_flip"] if case + "_flip" in gold.files else None
    x = gold[src][order]
    _, out = O.image_prep(x, size, flip=flip)
    ref = gold[case + "_out"]
    assert out.dtype == np.float32 and out.shape == ref.shape
    assert np.array_equal(out, ref)


def test_golden_flips_are_exercised(gold):
    assert gold["cifar_train_flip"].any() or gold["inet_train_flip"].any()
    assert not (gold["cifar_train_flip"].all() and gold["inet_train_flip"].all())


def test_tables_upsampling_window_is_three_taps():
    ksize, bounds, coeffs = O.resample_tables(32, 224)
    assert ksize == 3 and bounds.shape == (224, 2) and coeffs.shape == (224, 3)
    assert (bounds[:, 1] >= 1).all() and (bounds[:, 1] <= 3).all()
    assert (bounds[:, 0] + bounds[:, 1] <= 32).all()
    s = coeffs.sum(1)
    assert (np.abs(s - (1 << O.PRECISION_BITS)) <= 2).all()      # fixed-point weights sum to one (rounding)


def test_identity_and_constant_images():
    rng = np.random.default_rng(1)
    x = rng.integers(0, 256, (2, 17, 23, 3), dtype=np.uint8)
    assert np.array_equal(O.resize_bilinear_u8(x, 17, 23), x)                  # neither pass runs
    c = np.full((1, 9, 9, 3), 200, dtype=np.uint8)
    assert (O.resize_bilinear_u8(c, 50, 31) == 200).all()


def test_resize_target_matches_torchvision_rule():
    assert O.resize_target(32, 32, 224) == (224, 224)
    assert O.resize_target(375, 500, 224) == (224, 298)
    assert O.resize_target(500, 333, 224) == (336, 224)
    assert O.resize_target(75, 100, (64, 64)) == (64, 64)


def test_against_live_torchvision_when_importable():
    tv = pytest.importorskip("torchvision")
    pil = pytest.importorskip("PIL.Image")
    import torch
    from torchvision.transforms import transforms

    rng = np.random.default_rng(5)
    for (h, w, size) in [(32, 32, 224), (32, 32, 384), (40, 30, 56), (7, 5, (96, 96)), (130, 97, (64, 80))]:
        x = rng.integers(0, 256, (2, h, w, 3), dtype=np.uint8)
        tf = transforms.Compose([transforms.Resize(size), transforms.ToTensor(),
                                 transforms.Normalize([0.5, 0.5, 0.5], [0.5, 0.5, 0.5])])
        ref = torch.stack([tf(pil.fromarray(im)) for im in x]).numpy()
        _, out = O.image_prep(x, size)
        assert np.array_equal(out, ref), (h, w, size)


def test_patch_columns_is_the_conv_flattening():
    import torch

    rng = np.random.default_rng(2)
    x = rng.standard_normal((2, 3, 32, 48)).astype(np.float32)
    w = rng.standard_normal((5, 3, 16, 16)).astype(np.float32)
    cols = O.patch_columns(x, 16, ldk=776)
    assert cols.shape == (2 * 2 * 3, 776) and (cols[:, 768:] == 0).all()
    y = cols[:, :768] @ w.reshape(5, -1).T
    ref = torch.nn.functional.conv2d(torch.from_numpy(x), torch.from_numpy(w), stride=16)
    ref = ref.permute(0, 2, 3, 1).reshape(-1, 5).numpy()
    assert np.allclose(y, ref, rtol=1e-4, atol=1e-4)


# ---- host side of the product's input pipeline (no device work): tables, draws, target sizes -------------------------
def test_product_host_helpers_match_oracle_and_torchvision():
    import importlib

    import torch

    ip = importlib.import_module("vit-of-pytorch_b200.input_pipeline")
    for mean, std in [((0.5, 0.5, 0.5), (0.5, 0.5, 0.5)), ((0.485, 0.456, 0.406), (0.229, 0.224, 0.225))]:
        assert np.array_equal(ip.normalize_lut(mean, std).numpy(), O.normalize_lut(mean, std))
    for h, w, size in [(32, 32, 224), (375, 500, 224), (500, 333, 224), (75, 100, (64, 64)), (7, 5, 3)]:
        assert ip.resize_target(h, w, size) == O.resize_target(h, w, size)
    with pytest.raises(ValueError):
        ip.normalize_lut((0.5,), (0.0,))
    with pytest.raises(RuntimeError):
        ip.DeviceImageTransform((32, 32), 224, device="cpu")        # no CPU path


def test_draw_flips_replays_torchvision_random_horizontal_flip():
    tvt = pytest.importorskip("torchvision.transforms")
    pil = pytest.importorskip("PIL.Image")
    import importlib

    import torch

    ip = importlib.import_module("vit-of-pytorch_b200.input_pipeline")
    img = np.zeros((4, 6, 3), dtype=np.uint8)
    img[:, 0] = 255                                                   # a marker column: lands on the right when flipped
    flipper = tvt.RandomHorizontalFlip()
    torch.manual_seed(123)
    seen = [int(np.array(flipper(pil.fromarray(img)))[0, -1, 0] == 255) for _ in range(40)]
    torch.manual_seed(123)
    drawn = ip.draw_flips(40).tolist()
    assert drawn == seen and 0 < sum(seen) < 40


class _OracleTransform:
    """CPU stand-in for DeviceImageTransform in the loader tests: the oracle's transform on CPU tensors."""

    def __init__(self, size):
        self.size = size

    def __call__(self, u8, flip=None):
        import torch
        _, out = O.image_prep(u8.numpy(), self.size, flip=None if flip is None else flip.numpy())
        return torch.from_numpy(out)


@pytest.mark.parametrize("case,src,split,size,bs", [("cifar_train", "cifar_in", "train", 224, 3), ("cifar_eval", "cifar_in", "val", 64, 4)])
def test_batch_loader_reproduces_the_reference_loader(gold, case, src, split, size, bs):
    """Order (seeded shuffle), flips (global-RNG draws, replayed with the seed the golden script used) and batching of
    DeviceBatchLoader against the batches the reference's own CIFAR100DataLoader produced (oracle/make_golden_prep.py)."""
    import importlib

    import torch

    ip = importlib.import_module("vit-of-pytorch_b200.input_pipeline")
    data = gold[src]
    loader = ip.DeviceBatchLoader(data, np.arange(len(data)), split=split, image_size=size, batch_size=bs, seed=42, device="cpu",
                                  transform=_OracleTransform(size))
    torch.manual_seed(7)                                              # FLIP_SEED of oracle/make_golden_prep.py
    xs, ys = zip(*list(loader))
    assert [x.shape[0] for x in xs] == [min(bs, len(data) - i) for i in range(0, len(data), bs)]
    assert np.array_equal(torch.cat(ys).numpy(), gold[case + "_order"])
    assert np.array_equal(torch.cat(xs).numpy(), gold[case + "_out"])


def test_batch_loader_order_follows_torch_dataloader_across_epochs():
    import importlib

    import torch
    from torch.utils.data import DataLoader, TensorDataset

    ip = importlib.import_module("vit-of-pytorch_b200.input_pipeline")
    n, bs = 37, 5
    g = torch.Generator()
    g.manual_seed(42)
    ref = DataLoader(TensorDataset(torch.arange(n)), batch_size=bs, shuffle=True, generator=g, num_workers=0)
    loader = ip.DeviceBatchLoader(np.zeros((n, 2, 2, 3), dtype=np.uint8), np.arange(n), split="train", image_size=2, batch_size=bs,
                                  seed=42, device="cpu", transform=lambda u8, flip=None: u8)
    for _ in range(3):
        want = [b[0].tolist() for b in ref]
        got = [y.tolist() for _, y in loader]
        assert got == want
    assert len(loader) == len(ref)
