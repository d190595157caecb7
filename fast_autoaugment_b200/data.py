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
import os

import numpy as np
import PIL.Image
import torch
from torch.utils.data import Sampler, SubsetRandomSampler

from . import _lib, archive
from .conf import Config as C
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


_IMAGENET_PCA = {                                            # reference data.py:25-32
    "eigval": [0.2175, 0.0188, 0.0045],
    "eigvec": [[-0.5675, 0.7192, 0.4009], [-0.5808, -0.0045, -0.8140], [-0.5836, -0.6948, 0.4203]],
}


class Lighting(object):
    """Drop-in for reference ``Lighting`` (augmentations.py:197-215): AlexNet-style PCA noise on a CHW float tensor,
    between ToTensor and Normalize in the ImageNet chain (data.py:70-72).  In the batched path the per-image offsets
    ``sample_rgb(n)`` go to ``augment_batch(..., lighting_rgb=...)``, which folds them into per-image normalisation tables."""

    def __init__(self, alphastd, eigval=_IMAGENET_PCA["eigval"], eigvec=_IMAGENET_PCA["eigvec"]):
        self.alphastd = alphastd
        self.eigval = torch.Tensor(eigval)
        self.eigvec = torch.Tensor(eigvec)

    def _rgb(self, like):
        alpha = like.new().resize_(3).normal_(0, self.alphastd)
        return self.eigvec.type_as(like).clone().mul(alpha.view(1, 3).expand(3, 3)) \
            .mul(self.eigval.view(1, 3).expand(3, 3)).sum(1).squeeze()

    def __call__(self, img):
        if self.alphastd == 0:
            return img
        rgb = self._rgb(img)
        return img.add(rgb.view(3, 1, 1).expand_as(img))

    def sample_rgb(self, n):
        """[n, 3] fp32 offsets, drawn from torch's CPU generator exactly like n per-image calls of the reference."""
        if self.alphastd == 0:
            return torch.zeros(n, 3)
        like = torch.empty(0)
        return torch.stack([self._rgb(like) for _ in range(n)])


class ColorJitter(object):
    """torchvision ``ColorJitter(brightness, contrast, saturation)`` of the ImageNet train chain (reference data.py:65-69)
    for uint8 NHWC CUDA batches: per image a random order (``torch.randperm(4)``) of Brightness / Contrast / Color
    ``ImageEnhance`` blends with factors uniform in ``[max(0, 1 - x), 1 + x]`` - the arithmetic of the policy ops of the
    same names (C ABI ``faa_color_jitter``)."""

    def __init__(self, brightness=0.4, contrast=0.4, saturation=0.4):
        self.ranges = [(max(0.0, 1.0 - v), 1.0 + v) if v else None for v in (brightness, contrast, saturation)]

    def sample_parity(self, n):
        """records of n images from torch's CPU generator in torchvision's order (get_params: randperm(4), then b, c, s)"""
        recs = np.zeros(n, dtype=_lib.JITTER_DTYPE)
        for i in range(n):
            order = torch.randperm(4).tolist()
            f = [None if r is None else float(torch.empty(1).uniform_(r[0], r[1])) for r in self.ranges]
            recs[i]["order"] = [o if (o < 3 and f[o] is not None) else 3 for o in order]
            recs[i]["alpha"] = [np.float32(x if x is not None else 1.0) for x in f]
        return recs

    def jitter_batch(self, batch_u8, recs=None, out=None):
        """uint8 [B,H,W,3] CUDA -> uint8 [B,H,W,3] (``out`` may be the input itself)."""
        if not (isinstance(batch_u8, torch.Tensor) and batch_u8.is_cuda and batch_u8.dtype == torch.uint8 and batch_u8.is_contiguous()):
            raise ValueError("batch must be a contiguous uint8 CUDA tensor [B, H, W, 3]")
        b, h, w, _ = batch_u8.shape
        if recs is None:
            recs = self.sample_parity(b)
        d_recs = torch.from_numpy(np.ascontiguousarray(recs).view(np.uint8).reshape(-1).copy()).to(batch_u8.device)
        if out is None:
            out = torch.empty_like(batch_u8)
        with torch.cuda.device(batch_u8.device):
            import ctypes as C
            _lib.check(_lib.lib.faa_color_jitter(batch_u8.data_ptr(), out.data_ptr(), b, h, w, d_recs.data_ptr(),
                                                 C.c_void_p(torch.cuda.current_stream(batch_u8.device).cuda_stream)))
        return out


def policy_by_conf_name(aug):
    """conf['aug'] -> policy list, with the reference's error behaviour (data.py:85-109)."""
    if isinstance(aug, list):
        return aug
    if aug in archive.BY_CONF_NAME:
        return archive.BY_CONF_NAME[aug]()
    if aug in ("default",):
        return None
    raise ValueError("not found augmentations. %s" % aug)


class SubsetSampler(Sampler):
    """Reference ``SubsetSampler`` (data.py:349-362): the given indices, in order."""

    def __init__(self, indices):
        self.indices = indices

    def __iter__(self):
        return (i for i in self.indices)

    def __len__(self):
        return len(self.indices)


class DeviceDataset:
    """uint8 images [N,H,W,3] + int64 targets living on the device - the counterpart of the reference's
    torchvision dataset objects (``total_trainset`` / ``testset``, data.py:114-196); ``targets`` stays a host
    list like torchvision's so that the stratified splits see what the reference's see."""

    def __init__(self, images, targets, device="cuda"):
        self.images = torch.as_tensor(np.ascontiguousarray(images) if isinstance(images, np.ndarray) else images)
        if self.images.dtype != torch.uint8 or self.images.dim() != 4 or self.images.shape[-1] != 3:
            raise ValueError("images must be uint8 [N, H, W, 3]")
        self.images = self.images.to(device).contiguous()
        self.targets = [int(t) for t in targets]
        self.labels = torch.as_tensor(self.targets, dtype=torch.int64, device=device)

    def __len__(self):
        return self.images.shape[0]

    def subset(self, idx):
        idx = [int(i) for i in idx]
        t = torch.as_tensor(idx, dtype=torch.int64, device=self.images.device)
        d = DeviceDataset.__new__(DeviceDataset)
        d.images = self.images.index_select(0, t)
        d.targets = [self.targets[i] for i in idx]
        d.labels = self.labels.index_select(0, t)
        return d


class GpuAugmentedLoader:
    """What ``get_dataloaders`` hands to ``train.py:47`` / ``search.py:101`` instead of a torch ``DataLoader``:
    an iterable of ``(data, label)`` whose ``data`` is the augmented, normalised CUDA batch (``.cuda()`` at
    train.py:49 becomes a no-op).  Batching follows ``DataLoader(dataset, batch_size, shuffle, sampler,
    drop_last)`` (reference data.py:214-224): the index stream comes from the very sampler objects the
    reference builds; the pixels of a batch are gathered on the device and go through ONE fused launch
    (policy -> RandomCrop -> HFlip -> ToTensor -> Normalize -> CutoutDefault).

    ``parity=True`` replays the reference's per-sample draws from the global ``random`` / ``numpy.random`` /
    torch generators (a ``num_workers=0`` DataLoader); the default draws on the device with Philox.
    """

    def __init__(self, dataset: DeviceDataset, batch, policies, tail: TailSpec, sampler=None, shuffle=False,
                 drop_last=False, seed=None, parity=False):
        if not torch.cuda.is_available():
            raise _lib.FaaRuntimeError("fast_autoaugment_b200 needs a CUDA device (no CPU fallback)")
        self.dataset, self.batch_size, self.tail = dataset, int(batch), tail
        self.sampler, self.shuffle, self.drop_last, self.parity = sampler, shuffle, drop_last, parity
        self.aug = Augmentation(policies) if policies is not None else Augmentation([[("Invert", 0.0, 0.0)]])
        self.seed = int(torch.initial_seed() if seed is None else seed) & 0x7FFFFFFFFFFFFFFF
        self._drawn = 0                   # samples drawn so far: the Philox counter never repeats across epochs

    def _n(self):
        return len(self.sampler) if self.sampler is not None else len(self.dataset)

    def __len__(self):
        n = self._n()
        return n // self.batch_size if self.drop_last else math.ceil(n / self.batch_size)

    def _indices(self):
        if self.sampler is not None:
            return list(iter(self.sampler))
        n = len(self.dataset)
        return torch.randperm(n).tolist() if self.shuffle else list(range(n))

    def __iter__(self):
        idx_all = self._indices()
        dev = self.dataset.images.device
        for k in range(len(self)):
            idx = idx_all[k * self.batch_size:(k + 1) * self.batch_size]
            t = torch.as_tensor(idx, dtype=torch.int64).to(dev, non_blocking=True)
            raw = self.dataset.images.index_select(0, t)
            data = self.aug.augment_batch(raw, self.tail, seed=self.seed, first_index=self._drawn, parity=self.parity)
            self._drawn += len(idx)
            yield data, self.dataset.labels.index_select(0, t)


# ---------------------------------------------------------------------------------------------------
def _load_arrays(dataset, dataroot):
    """(train images, train targets, test images, test targets) as uint8 / int arrays.

    ``dataroot`` is the reference's dataset directory (torchvision layout, read with ``download=False`` - the
    build and GPU boxes have no network), or a directory holding ``<dataset>_train.npz`` / ``<dataset>_test.npz``
    (arrays ``data`` [N,H,W,3] uint8 and ``targets``), or - for in-memory injection (tests, synthetic benchmarks) -
    a mapping ``{"train": (images, targets), "test": (images, targets)}``."""
    base = dataset.replace("reduced_", "")
    if isinstance(dataroot, dict):
        tr, te = dataroot["train"], dataroot.get("test", dataroot["train"])
        return np.asarray(tr[0]), list(tr[1]), np.asarray(te[0]), list(te[1])
    npz = [os.path.join(str(dataroot), "%s_%s.npz" % (base, s)) for s in ("train", "test")]
    if all(os.path.exists(p) for p in npz):
        a, b = np.load(npz[0]), np.load(npz[1])
        return a["data"], list(a["targets"]), b["data"], list(b["targets"])
    import torchvision
    if base in ("cifar10", "cifar100"):
        cls = torchvision.datasets.CIFAR10 if base == "cifar10" else torchvision.datasets.CIFAR100
        tr, te = cls(root=dataroot, train=True, download=False), cls(root=dataroot, train=False, download=False)
        return tr.data, list(tr.targets), te.data, list(te.targets)
    if base == "svhn":
        def hwc(d):
            return np.ascontiguousarray(np.transpose(d.data, (0, 2, 3, 1)))
        tr = torchvision.datasets.SVHN(root=dataroot, split="train", download=False)
        te = torchvision.datasets.SVHN(root=dataroot, split="test", download=False)
        if dataset == "svhn":                                   # data.py:132-135: train + extra
            ex = torchvision.datasets.SVHN(root=dataroot, split="extra", download=False)
            return (np.concatenate([hwc(tr), hwc(ex)]), list(tr.labels) + list(ex.labels), hwc(te), list(te.labels))
        return hwc(tr), list(tr.labels), hwc(te), list(te.labels)
    raise ValueError("invalid dataset name=%s" % dataset) if "imagenet" not in dataset else NotImplementedError(
        "ImageNet needs JPEG decoding (reference imagenet.py:80), which is outside this package's hot path: "
        "pass fixed-size uint8 arrays through a mapping / .npz dataroot instead")


def get_dataloaders(dataset, batch, dataroot, split=0.15, split_idx=0, multinode=False, target_lb=-1):
    """Drop-in for reference ``get_dataloaders`` (data.py:37-225): same signature, same conf keys
    (``C.get()['aug']``, ``['cutout']``), same stratified splits and sampler objects, same return tuple
    ``(train_sampler, trainloader, validloader, testloader)`` - but the three loaders are
    ``GpuAugmentedLoader`` s over device-resident uint8 datasets, so the per-sample PIL chain of the reference's
    DataLoader workers is replaced by one fused kernel launch per batch.

    Extra conf keys (optional): ``faa_out_dtype`` ('float32' default - what the reference yields -, 'float16',
    'bfloat16'), ``faa_parity`` (replay the reference's global RNG draws per sample; tests)."""
    from sklearn.model_selection import StratifiedShuffleSplit
    import torch.distributed as dist

    conf = C.get()
    out_dtype = {"float32": torch.float32, "float16": torch.float16, "bfloat16": torch.bfloat16}[
        str(conf.get("faa_out_dtype", "float32"))]
    cutout = int(conf.get("cutout", 0) or 0)
    if "cifar" in dataset or "svhn" in dataset:                 # data.py:38-48
        tail = TailSpec((32, 32), 4, True, CIFAR_MEAN, CIFAR_STD, cutout, out_dtype)
        test_tail = TailSpec(None, 0, False, CIFAR_MEAN, CIFAR_STD, 0, out_dtype)
    elif "imagenet" in dataset:                                 # data.py:49-80 on already-sized images
        tail = TailSpec(None, 0, True, IMAGENET_MEAN, IMAGENET_STD, cutout, out_dtype)
        test_tail = TailSpec(None, 0, False, IMAGENET_MEAN, IMAGENET_STD, 0, out_dtype)
    else:
        raise ValueError("dataset=%s" % dataset)
    policies = policy_by_conf_name(conf["aug"])                 # data.py:85-109 (ValueError on an unknown name)

    tr_x, tr_y, te_x, te_y = _load_arrays(dataset, dataroot)
    if dataset in ("cifar10", "cifar100", "svhn", "imagenet"):
        total_trainset, testset = DeviceDataset(tr_x, tr_y), DeviceDataset(te_x, te_y)
    elif dataset in ("reduced_cifar10", "reduced_svhn"):        # data.py:117-126, 136-146
        test_size = 46000 if dataset == "reduced_cifar10" else 73257 - 1000
        sss = StratifiedShuffleSplit(n_splits=1, test_size=test_size, random_state=0)
        train_idx, _ = next(sss.split(list(range(len(tr_y))), tr_y))
        total_trainset, testset = DeviceDataset(tr_x, tr_y).subset(train_idx), DeviceDataset(te_x, te_y)
    else:
        raise ValueError("invalid dataset name=%s" % dataset)

    train_sampler = None
    if split > 0.0:                                             # data.py:189-203
        sss = StratifiedShuffleSplit(n_splits=5, test_size=split, random_state=0)
        sss = sss.split(list(range(len(total_trainset))), total_trainset.targets)
        for _ in range(split_idx + 1):
            train_idx, valid_idx = next(sss)
        if target_lb >= 0:
            train_idx = [i for i in train_idx if total_trainset.targets[i] == target_lb]
            valid_idx = [i for i in valid_idx if total_trainset.targets[i] == target_lb]
        train_sampler = SubsetRandomSampler(train_idx)
        valid_sampler = SubsetSampler(valid_idx)
        if multinode:
            train_sampler = torch.utils.data.distributed.DistributedSampler(
                torch.utils.data.Subset(range(len(total_trainset)), train_idx),
                num_replicas=dist.get_world_size(), rank=dist.get_rank())
    else:
        valid_sampler = SubsetSampler([])
        if multinode:
            train_sampler = torch.utils.data.distributed.DistributedSampler(
                range(len(total_trainset)), num_replicas=dist.get_world_size(), rank=dist.get_rank())

    parity = bool(conf.get("faa_parity", False))
    trainloader = GpuAugmentedLoader(total_trainset, batch, policies, tail, sampler=train_sampler,
                                     shuffle=train_sampler is None, drop_last=True, parity=parity)       # data.py:214-216
    validloader = GpuAugmentedLoader(total_trainset, batch, policies, tail, sampler=valid_sampler,
                                     shuffle=False, drop_last=False, parity=parity)                      # data.py:217-219
    testloader = GpuAugmentedLoader(testset, batch, None, test_tail, shuffle=False, drop_last=False)     # data.py:221-224
    return train_sampler, trainloader, validloader, testloader
