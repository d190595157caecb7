"""Reference ``FastAutoAugment/augmentations.py`` surface on the CUDA path.

``apply_augment(img, name, level)`` keeps the reference's signature (PIL in, PIL out,
reference ``augmentations.py:192-194``): the op runs as a one-op policy with probability 1
through the same fused kernel as the batched path; its random draws (mirror sign, Cutout
centre) come from the same global generators as the reference's.
"""
from __future__ import annotations

import numpy as np
import PIL.Image
import torch

from . import _lib
from .engine import CompiledPolicy, TailSpec, augment_batch

# (name, low, high) in the reference's registry order (augmentations.py:156-182)
_ALL = [(_lib.lib.faa_op_name(i).decode(),) for i in range(19)]


def augment_list(for_autoaug=True):
    """(op name, low, high) triples; the reference returns the functions themselves, which
    here are ``functools.partial`` views of ``apply_named`` carrying ``__name__``."""
    import ctypes as C
    out = []
    for i in range(19 if for_autoaug else 15):
        lo, hi = C.c_double(), C.c_double()
        _lib.check(_lib.lib.faa_op_range(i, C.byref(lo), C.byref(hi)))
        name = _lib.lib.faa_op_name(i).decode()

        def fn(img, v, _name=name, _lo=lo.value, _hi=hi.value):
            return _apply_magnitude(img, _name, v, _lo, _hi)
        fn.__name__ = name
        out.append((fn, lo.value, hi.value))
    return out


augment_dict = {fn.__name__: (fn, lo, hi) for fn, lo, hi in augment_list()}


def get_augment(name):
    return augment_dict[name]                      # KeyError like the reference (:189)


def _run_single(img, policy_list):
    arr = np.asarray(img.convert("RGB"))
    if not torch.cuda.is_available():
        raise _lib.FaaRuntimeError("fast_autoaugment_b200 needs a CUDA device (no CPU fallback)")
    pol = CompiledPolicy(policy_list)
    h, w = arr.shape[:2]
    samples, boxes = pol.sample_parity(1, h, w, TailSpec.raw_u8())
    x = torch.from_numpy(np.ascontiguousarray(arr)[None].copy()).cuda()
    y = augment_batch(pol, x, TailSpec.raw_u8(), samples, boxes)
    return PIL.Image.fromarray(y[0].cpu().numpy())


def _apply_magnitude(img, name, v, lo, hi):
    level = (v - lo) / (hi - lo) if hi != lo else 0.0
    return _run_single(img, [[(name, 1.0, level)]])


def apply_augment(img, name, level):
    """Reference ``apply_augment`` (augmentations.py:192-194)."""
    get_augment(name)
    return _run_single(img, [[(name, 1.0, level)]])
