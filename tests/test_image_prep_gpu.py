"""GPU parity of the device input transform through the C ABI (vitb_image_prep), BIT-EXACT against the
oracle (Pillow's 8-bit bilinear resample + torchvision ToTensor / Normalize) and against the golden vectors produced
by the reference's own loader classes (oracle/make_golden_prep.py)."""
import os

import numpy as np
import pytest
import torch

import vitb200
from oracle import image_prep_oracle as O
from conftest import rel_l2

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "image_prep.npz")
DEV = "cuda"


def bf16_bits(x):
    return torch.from_numpy(np.ascontiguousarray(x)).to(torch.bfloat16).view(torch.int16)


@pytest.mark.parametrize("case,src,size", [("cifar_train", "cifar_in", 224), ("cifar_eval", "cifar_in", 64),
                                           ("inet_train", "inet_in", (64, 64))])
def test_matches_reference_loader_output(case, src, size):
    g = np.load(GOLD)
    order = g[case + "_order"]
    flip = torch.from_numpy(g[case + "_flip"].astype(np.uint8)) if case + "_flip" in g.files else None
    x = torch.from_numpy(g[src][order]).to(DEV)
    tf = vitb200.DeviceImageTransform(x.shape[1:3], size, device=DEV)
    out = tf(x, flip=flip)
    assert torch.equal(out.cpu(), torch.from_numpy(g[case + "_out"]))


@pytest.mark.parametrize("H,W,out_hw,P", [
    (32, 32, (224, 224), 16), (32, 32, (224, 224), 14), (32, 32, (384, 384), 16), (32, 32, (32, 32), 16),
    (40, 32, (40, 96), 8), (40, 32, (100, 32), 4), (75, 100, (64, 64), 16), (333, 500, (61, 45), 7),
    (375, 500, (224, 224), 16), (9, 700, (300, 10), 5), (1, 1, (17, 19), 4),
])
def test_matches_oracle_all_outputs(H, W, out_hw, P):
    rng = np.random.default_rng(H * 1000 + W)
    B = 5
    x = rng.integers(0, 256, (B, H, W, 3), dtype=np.uint8)
    flip = np.array([1, 0, 1, 1, 0], dtype=np.uint8)
    ref_u8, ref = O.image_prep(x, out_hw, flip=flip)
    tf = vitb200.DeviceImageTransform((H, W), out_hw, device=DEV)
    xd, fd = torch.from_numpy(x).to(DEV), torch.from_numpy(flip)
    assert torch.equal(tf(xd, flip=fd).cpu(), torch.from_numpy(ref))
    assert torch.equal(tf.resized_bytes(xd, flip=fd).cpu(), torch.from_numpy(ref_u8))
    with vitb200.precision("fp32"):
        pc = tf.patch_columns(xd, P, flip=fd)
    cols = O.patch_columns(ref, P, pc.hi.shape[1])
    t = torch.from_numpy(cols)
    assert torch.equal(pc.hi.cpu().view(torch.int16), bf16_bits(cols))
    assert torch.equal(pc.lo.cpu().view(torch.int16), (t - t.to(torch.bfloat16).float()).to(torch.bfloat16).view(torch.int16))


def test_full_batch_properties():
    """BASELINE config size (128 CIFAR images -> 224 px): size-independent properties instead of a CPU oracle pass:
    flipping twice is the identity, a flipped batch is the mirror of the unflipped one, constant images stay constant,
    and the fp32 output is the table lookup of the byte output."""
    g = torch.Generator().manual_seed(0)
    x = torch.randint(0, 256, (128, 32, 32, 3), dtype=torch.uint8, generator=g)
    x[0] = 200
    xd = x.to(DEV)
    tf = vitb200.DeviceImageTransform((32, 32), 224, device=DEV)
    plain = tf(xd)
    ones = torch.ones(128, dtype=torch.uint8)
    assert torch.equal(tf(xd, flip=ones), plain.flip(-1))
    mixed = vitb200.draw_flips(128, generator=torch.Generator().manual_seed(3))
    out = tf(xd, flip=mixed)
    sel = mixed.bool().to(DEV)
    assert torch.equal(out[sel], plain[sel].flip(-1)) and torch.equal(out[~sel], plain[~sel])
    assert (plain[0] == tf.lut[0, 200]).all()
    u8 = tf.resized_bytes(xd)
    lut = tf.lut
    look = torch.stack([lut[c][u8[..., c].long()] for c in range(3)], dim=1)
    assert torch.equal(look, plain)
    # first 4 images against the oracle
    _, ref = O.image_prep(x[:4].numpy(), 224)
    assert torch.equal(plain[:4].cpu(), torch.from_numpy(ref))


@pytest.mark.parametrize("mode", ["bf16", "fp32"])
def test_model_on_patch_columns_equals_model_on_loader_batch(mode):
    """VisionTransformer fed the PatchColumns of a uint8 batch computes bit-identical logits and gradients to the same
    model fed the fp32 batch the reference's loader would have produced (the im2col of that batch IS the columns)."""
    torch.manual_seed(0)
    net = vitb200.VisionTransformer(image_size=(64, 64), patch_size=(16, 16), emb_dim=128, mlp_dim=256, num_heads=2,
                                    num_layers=2, num_classes=10, dropout_rate=0.0).to(DEV)
    with torch.no_grad():
        for n, p in net.named_parameters():
            if "attn" in n or "pos_embedding" in n:
                p.mul_(0.02)
    rng = np.random.default_rng(9)
    x = rng.integers(0, 256, (6, 32, 32, 3), dtype=np.uint8)
    flip = np.array([0, 1, 0, 1, 1, 0], dtype=np.uint8)
    _, ref_img = O.image_prep(x, 64, flip=flip)
    labels = torch.arange(6, device=DEV) % 10
    tf = vitb200.DeviceImageTransform((32, 32), 64, device=DEV)
    xd, fd = torch.from_numpy(x).to(DEV), torch.from_numpy(flip)
    with vitb200.precision(mode):
        la = net(torch.from_numpy(ref_img).to(DEV))
        vitb200.functional.cross_entropy(la, labels).backward()
        ga = net.embedding.weight.grad.clone()
        net.zero_grad()
        lb = net(tf.patch_columns(xd, 16, flip=fd))
        vitb200.functional.cross_entropy(lb, labels).backward()
        gb = net.embedding.weight.grad.clone()
    assert torch.equal(la, lb)
    assert rel_l2(gb, ga) < 1e-5          # wgrad accumulates with fp32 atomics (split-K): the order differs run to run


def test_bad_arguments_raise():
    tf = vitb200.DeviceImageTransform((32, 32), 224, device=DEV)
    with pytest.raises(RuntimeError):
        tf(torch.zeros((2, 32, 32, 3), dtype=torch.uint8))                       # CPU tensor: no CPU path
    with pytest.raises(RuntimeError):
        tf(torch.zeros((2, 16, 32, 3), dtype=torch.uint8, device=DEV))           # wrong source size
    with pytest.raises(RuntimeError):
        tf(torch.zeros((2, 32, 32, 3), dtype=torch.float32, device=DEV))         # not bytes


@pytest.mark.parametrize("patch", [None, 16])
def test_batch_loader_on_device_matches_reference_loader(patch):
    """DeviceBatchLoader end to end on the GPU (pinned staging, device transform) against the reference loader's batches."""
    g = np.load(GOLD)
    data = g["cifar_in"]
    loader = vitb200.DeviceBatchLoader(data, np.arange(len(data)), split="train", image_size=224, batch_size=3, seed=42,
                                       device=DEV, patch=patch)
    torch.manual_seed(7)
    xs, ys = zip(*list(loader))
    assert np.array_equal(torch.cat(ys).cpu().numpy(), g["cifar_train_order"])
    want = g["cifar_train_out"]
    if patch is None:
        assert torch.equal(torch.cat(xs).cpu(), torch.from_numpy(want))
    else:
        hi = torch.cat([pc.hi for pc in xs]).cpu()
        cols = O.patch_columns(want, patch, hi.shape[1])
        assert torch.equal(hi.view(torch.int16), bf16_bits(cols))
