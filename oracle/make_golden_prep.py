"""Generates tests/golden/image_prep.npz by running the reference's OWN loader classes.

Run in the build container:  python oracle/make_golden_prep.py

/root/reference/src/data_loaders.py is imported unmodified; only the torchvision DATASET classes it
names (CIFAR100 / ImageFolder need files or a download) are replaced in its namespace by in-memory
stand-ins that do what torchvision's datasets do with a sample — `Image.fromarray(data[i])`, then
`self.transform(img)` — so the transform pipeline, the shuffling generator and the batching are the
reference's code, executed by the installed torchvision / Pillow.  The label of sample i is i, which
records the order the loader produced.  RandomHorizontalFlip draws `torch.rand(1) < 0.5` from the global
torch RNG once per image (num_workers = 0 keeps the draws in this process): the script re-seeds and
replays the draws to store the flip mask as an input of the transform.

Cases (all uint8 HWC inputs from numpy's PCG64, seeds below):
  cifar_train  CIFAR100DataLoader(split='train', image_size=224): 32x32 -> 224, shuffled, flips
  cifar_eval   CIFAR100DataLoader(split='val',   image_size=64) : 32x32 -> 64, in order, no flips
  inet_train   ImageNetDataLoader(split='train', image_size=64) : 75x100 -> (64,64) (downsampling
               windows wider than 3 taps), shuffled, flips
"""
import importlib.util
import os
import sys

import numpy as np
import torch
from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("VITB_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(ROOT, "tests", "golden", "image_prep.npz")
FLIP_SEED = 7


class _MemoryDataset(torch.utils.data.Dataset):
    """What torchvision.datasets.CIFAR100 / ImageFolder do per sample, over an in-memory uint8 array."""
    data = None

    def __init__(self, root=None, train=True, transform=None, download=False):
        self.transform = transform

    def __len__(self):
        return len(self.data)

    def __getitem__(self, index):
        img = Image.fromarray(self.data[index])
        if self.transform is not None:
            img = self.transform(img)
        return img, index


def load_reference_loaders():
    spec = importlib.util.spec_from_file_location("ref_data_loaders", os.path.join(REF, "src", "data_loaders.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def run(loader_cls, attr, data, split, image_size, batch_size):
    ref = load_reference_loaders()
    ds = type("Stub", (_MemoryDataset,), {"data": data})
    setattr(ref, attr, ds)
    loader = loader_cls(ref)(data_dir="/nonexistent", split=split, image_size=image_size, batch_size=batch_size,
                             num_workers=0, seed=42)
    torch.manual_seed(FLIP_SEED)
    outs, order = [], []
    for x, idx in loader:
        outs.append(x.numpy())
        order.append(idx.numpy())
    out = np.concatenate(outs)
    order = np.concatenate(order)
    torch.manual_seed(FLIP_SEED)
    if split == "train":
        flips = np.array([bool(torch.rand(1) < 0.5) for _ in range(len(order))])
    else:
        flips = np.zeros(len(order), dtype=bool)
    return out, order.astype(np.int64), flips


def main():
    rng = np.random.default_rng(20261018)
    cifar = rng.integers(0, 256, (4, 32, 32, 3), dtype=np.uint8)
    cifar[0, :, :, :] = np.linspace(0, 255, 32, dtype=np.uint8)[None, :, None]       # a smooth ramp
    cifar[1, ::2] = 255
    cifar[1, 1::2] = 0                                                               # saturating stripes
    inet = rng.integers(0, 256, (4, 75, 100, 3), dtype=np.uint8)
    pack = {}
    out, order, flips = run(lambda r: r.CIFAR100DataLoader, "CIFAR100", cifar, "train", 224, 3)
    pack.update(cifar_in=cifar, cifar_train_out=out, cifar_train_order=order, cifar_train_flip=flips)
    out, order, flips = run(lambda r: r.CIFAR100DataLoader, "CIFAR100", cifar, "val", 64, 4)
    pack.update(cifar_eval_out=out, cifar_eval_order=order)
    out, order, flips = run(lambda r: r.ImageNetDataLoader, "ImageFolder", inet, "train", 64, 2)
    pack.update(inet_in=inet, inet_train_out=out, inet_train_order=order, inet_train_flip=flips)
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    np.savez_compressed(OUT, **pack)
    for k, v in pack.items():
        print(k, v.shape, v.dtype)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    sys.exit(main())
