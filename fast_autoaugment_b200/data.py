"""Reference ``FastAutoAugment/data.py`` surface on the CUDA path.

* ``Augmentation(policies)``   - same constructor and ``__call__(PIL) -> PIL`` as the reference
  (``data.py:253-264``), plus the batched entry ``augment_batch`` the training loop should use.
* ``CutoutDefault(length)``    - same callable as the reference (``data.py:228-250``).
* ``GpuAugmentedLoader``       - what ``get_dataloaders`` (``data.py:37-225``) hands to
  ``train.py:47`` / ``search.py:101``: an iterable of ``(data, label)`` whose ``data`` is already
  the augmented, normalised CUDA tensor, so the caller's ``.cuda()`` (``train.py:49``) is a
  no-op.  The raw uint8 dataset lives on the device; no CUDA is touched in worker processes
  because there are none.
"""
from __future__ import annotations

import math

import numpy as np
import PIL.Image
import torch

from . import _lib, archive
from .engine import (CIFAR_MEAN, CIFAR_STD, IMAGENET_MEAN, IMAGENET_STD, CompiledPolicy, TailSpec,
                     augment_batch, make_rng)


class Augmentation(object):
    """Drop-in for reference ``Augmentation`` (data.py:253-264)."""

    def __init__(self, policies):
        self.policies = policies
        self._compiled = None

    @property
    def compiled(self) -> CompiledPolicy:
        if self._compiled is None:
            self._compiled = CompiledPolicy(self.policies)
        return self._compiled

    def __call__(self, img):
        """PIL RGB image in, new PIL image out; consumes ``random`` / ``numpy.random`` exactly
        like the reference, so identical seeds give identical pixels."""
        if not torch.cuda.is_available():
            raise _lib.FaaRuntimeError("fast_autoaugment_b200 needs a CUDA device (no CPU fallback)")
        arr = np.ascontiguousarray(np.asarray(img.convert("RGB")))
        h, w = arr.shape[:2]
        samples, boxes = self.compiled.sample_parity(1, h, w, TailSpec.raw_u8())
        y = augment_batch(self.compiled, torch.from_numpy(arr[None].copy()).cuda(), TailSpec.raw_u8(), samples, boxes)
        return PIL.Image.fromarray(y[0].cpu().numpy())

    def augment_batch(self, batch_u8, tail: TailSpec | None = None, seed=None, first_index=0, parity=False,
                      out=None):
        """uint8 [B,H,W,3] CUDA tensor -> augmented batch.  ``parity=True`` replays the global
        generators like the reference's per-sample loop; otherwise the kernel draws with Philox
        keyed by (``seed``, ``first_index`` + i)."""
        tail = tail or TailSpec.raw_u8()
        b, h, w, _ = batch_u8.shape
        if parity:
            samples, boxes = self.compiled.sample_parity(b, h, w, tail)
            return augment_batch(self.compiled, batch_u8, tail, samples, boxes, out=out)
        if seed is None:
            seed = int(torch.randint(0, 2 ** 62, (1,)).item())
        return augment_batch(self.compiled, batch_u8, tail, rng=make_rng(seed, first_index, tail), out=out)


class CutoutDefault(object):
    """Drop-in for reference ``CutoutDefault`` (data.py:228-250): zero box on a CHW tensor, in
    place, centre drawn from ``numpy.random`` (y first).  In the batched path the same box is
    applied inside the fused kernel (``TailSpec.cutout``)."""

    def __init__(self, length):
        self.length = length

    def __call__(self, img):
        h, w = img.size(1), img.size(2)
        y = np.random.randint(h)
        x = np.random.randint(w)
        half = self.length // 2
        y1, y2 = int(np.clip(y - half, 0, h)), int(np.clip(y + half, 0, h))
        x1, x2 = int(np.clip(x - half, 0, w)), int(np.clip(x + half, 0, w))
        img[:, y1:y2, x1:x2] = 0.0
        return img


def policy_by_conf_name(aug):
    """conf['aug'] -> policy list, with the reference's error behaviour (data.py:86-109)."""
    if isinstance(aug, list):
        return aug
    if aug in archive.BY_CONF_NAME:
        return archive.BY_CONF_NAME[aug]()
    if aug in ("default",):
        return None
    raise ValueError("not found augmentations. %s" % aug)


class GpuAugmentedLoader:
    """Iterable of ``(augmented CUDA batch, labels)`` over a device-resident uint8 dataset.

    images: uint8 [N,H,W,3] (numpy or tensor; moved to ``device`` once), labels: int64 [N].
    Mirrors ``DataLoader(..., shuffle=True, drop_last=True)`` of reference data.py:214-216 for
    the train loader; ``rank``/``world_size`` shard indices like ``DistributedSampler``
    (data.py:205-212).  ``set_epoch`` reseeds the shuffle (train.py:252).
    """

    def __init__(self, images, labels, batch, policies, tail: TailSpec, device="cuda", shuffle=True,
                 drop_last=True, seed=0, rank=0, world_size=1, parity=False):
        if not torch.cuda.is_available():
            raise _lib.FaaRuntimeError("fast_autoaugment_b200 needs a CUDA device (no CPU fallback)")
        self.device = torch.device(device)
        self.images = torch.as_tensor(images).to(self.device).contiguous()
        self.labels = torch.as_tensor(labels).to(self.device)
        self.batch, self.tail, self.shuffle, self.drop_last = batch, tail, shuffle, drop_last
        self.seed, self.epoch, self.rank, self.world_size, self.parity = seed, 0, rank, world_size, parity
        self.aug = Augmentation(policies) if policies is not None else Augmentation([[("Invert", 0.0, 0.0)]])
        n = self.images.shape[0]
        self.per_rank = n // world_size if drop_last else math.ceil(n / world_size)

    def set_epoch(self, epoch):
        self.epoch = epoch

    def __len__(self):
        return self.per_rank // self.batch if self.drop_last else math.ceil(self.per_rank / self.batch)

    def __iter__(self):
        n = self.images.shape[0]
        g = torch.Generator(device="cpu")
        g.manual_seed(self.seed + self.epoch)
        order = torch.randperm(n, generator=g) if self.shuffle else torch.arange(n)
        order = order[self.rank::self.world_size][: self.per_rank].to(self.device)
        for k in range(len(self)):
            idx = order[k * self.batch:(k + 1) * self.batch]
            raw = self.images.index_select(0, idx)
            data = self.aug.augment_batch(raw, self.tail, seed=self.seed * 1000003 + self.epoch,
                                          first_index=(self.rank * len(self) + k) * self.batch, parity=self.parity)
            yield data, self.labels.index_select(0, idx)


def get_dataloaders(dataset, batch, images, labels, aug="fa_reduced_cifar10", cutout=16, out_dtype=torch.float16,
                    test_images=None, test_labels=None, rank=0, world_size=1, seed=0):
    """Device-resident counterpart of reference ``get_dataloaders`` (data.py:37-225) for datasets
    already in memory as uint8 arrays (there are no dataset files on the build / GPU boxes).
    Returns ``(train_sampler, trainloader, validloader, testloader)`` like the reference; the
    train loader doubles as its own sampler (``set_epoch``)."""
    if "cifar" in dataset or "svhn" in dataset:
        tail = TailSpec((32, 32), 4, True, CIFAR_MEAN, CIFAR_STD, cutout, out_dtype)
        test_tail = TailSpec(None, 0, False, CIFAR_MEAN, CIFAR_STD, 0, out_dtype)
    elif "imagenet" in dataset:
        tail = TailSpec(None, 0, True, IMAGENET_MEAN, IMAGENET_STD, cutout, out_dtype)
        test_tail = TailSpec(None, 0, False, IMAGENET_MEAN, IMAGENET_STD, 0, out_dtype)
    else:
        raise ValueError("dataset=%s" % dataset)
    policies = policy_by_conf_name(aug)
    train = GpuAugmentedLoader(images, labels, batch, policies, tail, seed=seed, rank=rank, world_size=world_size)
    valid = GpuAugmentedLoader(images[:0], labels[:0], batch, policies, tail, shuffle=False, drop_last=False)
    test = None
    if test_images is not None:
        test = GpuAugmentedLoader(test_images, test_labels, batch, None, test_tail, shuffle=False, drop_last=False)
    return train, train, valid, test
