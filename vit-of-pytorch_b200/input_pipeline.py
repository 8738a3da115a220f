"""The input transform of the reference's loaders, on the device.

Reference: src/data_loaders.py:36-48 / :69-82 (CIFAR: `Resize(image_size)`, `RandomHorizontalFlip()` in the train
split, `ToTensor()`, `Normalize([.5]*3, [.5]*3)`) and :102-114 (ImageNet: `Resize((S, S))`).  There each sample is
upsampled by Pillow on a CPU worker (`num_workers=1` in src/config.py) and the batch crosses PCIe as fp32 — 602 KB
per 224 px image; at the ~7 k images/s the encoder step runs here, one worker falls short by two orders of magnitude.
Here the decoded uint8 batch (3 KB per CIFAR image) is staged to the device and ONE kernel (`vitb_image_prep`)
produces, bit-exactly, what the reference's loader would have produced — either the fp32 NCHW batch the module API
takes, or directly the bf16 patch operand of the patch-embedding GEMM (`PatchColumns`), in which case the fp32 image
never exists in HBM.

The flip decision is an INPUT (uint8 [B]): `draw_flips` draws it the way `RandomHorizontalFlip` does
(`torch.rand(1) < p`, once per image, from the given generator), so a test can replay the reference's draws.
JPEG decoding, dataset indexing and shuffling stay on the host: they are not part of this row.
"""
import torch

from . import _lib as L
from . import functional as F
from . import ops


def draw_flips(batch, p=0.5, generator=None):
    """uint8 [batch] on the CPU: one `torch.rand(1) < p` per image, in order (torchvision RandomHorizontalFlip.forward)."""
    return torch.tensor([bool(torch.rand(1, generator=generator) < p) for _ in range(batch)], dtype=torch.uint8)


def normalize_lut(mean, std):
    """fp32 [C,256]: ToTensor (`byte.to(float32).div(255)`) then Normalize (`sub_(mean).div_(std)`), evaluated with the
    same torch CPU ops torchvision uses, so the table holds exactly the floats the reference's loader produces."""
    v = torch.arange(256, dtype=torch.uint8).to(torch.float32).div(255)
    m = torch.as_tensor(mean, dtype=torch.float32)
    s = torch.as_tensor(std, dtype=torch.float32)
    if (s == 0).any():
        raise ValueError("std evaluated to zero")            # torchvision raises here too
    t = v.repeat(m.numel(), 1)
    return t.sub_(m[:, None]).div_(s[:, None]).contiguous()


def resize_target(h, w, size):
    """torchvision `Resize`: an int matches the shorter side (long side int(size*long/short)); a pair is (h, w)."""
    if isinstance(size, (tuple, list)):
        return int(size[0]), int(size[1])
    if h <= w:
        return int(size), int(size * w / h)
    return int(size * h / w), int(size)


class DeviceImageTransform:
    """`Compose([Resize(image_size), RandomHorizontalFlip(), ToTensor(), Normalize(mean, std)])` for uint8 batches of one
    source size, as one device kernel.

        tf = DeviceImageTransform((32, 32), 224, device="cuda")
        x = tf(u8_batch, flip=flips)                       # fp32 [B,3,224,224], equals the reference loader's batch
        logits = net(tf.patch_columns(u8_batch, 16, flip=flips))   # or skip the fp32 image altogether
    """

    def __init__(self, in_hw, image_size, mean=(0.5, 0.5, 0.5), std=(0.5, 0.5, 0.5), device="cuda"):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise L.VitbError("DeviceImageTransform runs only on a B200 (sm_100a) CUDA device — there is no CPU path")
        self.in_hw = (int(in_hw[0]), int(in_hw[1]))
        self.out_hw = resize_target(self.in_hw[0], self.in_hw[1], image_size)
        self.lut = normalize_lut(mean, std).to(self.device)
        self.xtab = self._tables(self.in_hw[1], self.out_hw[1])
        self.ytab = self._tables(self.in_hw[0], self.out_hw[0])

    def _tables(self, n_in, n_out):
        if n_in == n_out:
            return None                                      # Pillow skips the pass: no rounding happens on this axis
        _, bounds, coeffs = ops.resize_tables(n_in, n_out)
        return bounds.to(self.device), coeffs.to(self.device)

    def _check(self, u8, flip):
        if tuple(u8.shape[1:3]) != self.in_hw or u8.shape[3] != self.lut.shape[0]:
            raise L.VitbError("expected uint8 [B,%d,%d,%d], got %s" % (self.in_hw + (self.lut.shape[0], tuple(u8.shape))))
        if flip is not None:
            flip = flip.to(device=self.device, dtype=torch.uint8, non_blocking=True).contiguous()
        return flip

    def __call__(self, u8, flip=None, out=None):
        """uint8 [B,H,W,C] (device) -> fp32 [B,C,S_h,S_w]."""
        flip = self._check(u8, flip)
        img, _, _, _ = ops.image_prep(u8, self.out_hw, self.xtab, self.ytab, self.lut, flip=flip, out_img=out)
        return img

    def resized_bytes(self, u8, flip=None):
        """uint8 [B,S_h,S_w,C]: the resized (and flipped) bytes, i.e. the PIL image the reference hands to ToTensor."""
        flip = self._check(u8, flip)
        return ops.image_prep(u8, self.out_hw, self.xtab, self.ytab, self.lut, flip=flip, want_img=False, want_u8=True)[3]

    def patch_columns(self, u8, patch, flip=None):
        """uint8 batch -> PatchColumns: the bf16 operand of the patch-embedding GEMM (plus its low half in fp32 mode),
        accepted by VisionTransformer.forward / patch_embed in place of the fp32 image."""
        flip = self._check(u8, flip)
        fp32 = F.get_precision() == "fp32"
        B, Cn = u8.shape[0], u8.shape[3]
        oh, ow = self.out_hw
        _, hi, lo, _ = ops.image_prep(u8, self.out_hw, self.xtab, self.ytab, self.lut, flip=flip, want_img=False,
                                      P=patch, want_lo=fp32)
        return F.PatchColumns(hi, lo, B, oh // patch, ow // patch, patch, Cn)


class DeviceBatchLoader:
    """The reference's `CIFAR10DataLoader` / `CIFAR100DataLoader` (src/data_loaders.py:32-92) over an in-memory uint8
    dataset, with the per-sample transform moved to the device.

        loader = DeviceBatchLoader(cifar.data, cifar.targets, split='train', image_size=224, batch_size=128, seed=42)
        for images, labels in loader:      # images fp32 [B,3,224,224] on the device — what the reference loader yields

    What is kept from the reference: `split='train'` shuffles with `torch.Generator().manual_seed(seed)` exactly as
    `DataLoader(shuffle=True, generator=generator)` does (one `random_()` draw for the iterator's base seed, the epoch's
    `randperm`, and the sampler's trailing `randperm` — so the sample order is the reference's, epoch after epoch) and flips
    with one `torch.rand(1) < 0.5` per image from the global RNG in batch order (what `RandomHorizontalFlip` draws with
    `num_workers=0`); other splits keep the dataset order and never flip; the last batch may be short (`drop_last=False`).
    What changes: the host only gathers the batch's uint8 images into pinned memory (3 KB per CIFAR image); resize, flip,
    ToTensor and Normalize run in `vitb_image_prep`.  `patch=P` yields `PatchColumns` (the patch-embedding GEMM operand)
    instead of the fp32 batch.  `transform=` replaces the device transform (tests inject a CPU stand-in to check the order
    and flip logic without a GPU).
    """

    def __init__(self, images_u8, labels, split='train', image_size=224, batch_size=16, seed=42, device="cuda",
                 mean=(0.5, 0.5, 0.5), std=(0.5, 0.5, 0.5), patch=None, transform=None):
        self.images = torch.as_tensor(images_u8)
        if self.images.dtype != torch.uint8 or self.images.dim() != 4:
            raise ValueError("images_u8 must be a uint8 [N,H,W,C] array (what torchvision's CIFAR datasets hold in .data)")
        self.labels = torch.as_tensor(labels, dtype=torch.int64)
        if self.labels.shape[0] != self.images.shape[0]:
            raise ValueError("labels and images disagree on the number of samples")
        self.train = split == 'train'
        self.batch_size = int(batch_size)
        self.patch = patch
        self.generator = torch.Generator()
        self.generator.manual_seed(seed)
        self.device = torch.device(device)
        if transform is None:
            transform = DeviceImageTransform(self.images.shape[1:3], image_size, mean, std, device=self.device)
        self.transform = transform
        self._pinned = None

    def __len__(self):
        return (self.images.shape[0] + self.batch_size - 1) // self.batch_size

    def epoch_order(self):
        """Sample order of the next epoch, advancing the generator exactly as the reference's DataLoader would."""
        n = self.images.shape[0]
        if not self.train:
            return torch.arange(n)
        torch.empty((), dtype=torch.int64).random_(generator=self.generator)      # _BaseDataLoaderIter: base seed
        order = torch.randperm(n, generator=self.generator)                        # RandomSampler: the epoch's permutation
        torch.randperm(n, generator=self.generator)                                # ... and its (empty) remainder draw
        return order

    def _stage(self, idx):
        if self.device.type != "cuda":
            return self.images[idx]
        B = idx.numel()
        if self._pinned is None or self._pinned.shape[0] < B:
            self._pinned = torch.empty((max(B, self.batch_size),) + tuple(self.images.shape[1:]), dtype=torch.uint8).pin_memory()
        buf = self._pinned[:B]
        torch.index_select(self.images, 0, idx, out=buf)
        dev = buf.to(self.device, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()     # the pinned buffer is reused by the next batch
        return dev

    def __iter__(self):
        order = self.epoch_order()
        for i in range(0, order.numel(), self.batch_size):
            idx = order[i:i + self.batch_size]
            u8 = self._stage(idx)
            flip = draw_flips(idx.numel()) if self.train else None
            labels = self.labels[idx]
            if self.device.type == "cuda":
                labels = labels.to(self.device, non_blocking=True)
            if self.patch is not None:
                yield self.transform.patch_columns(u8, self.patch, flip=flip), labels
            else:
                yield self.transform(u8, flip=flip), labels
