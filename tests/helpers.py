"""Shared helpers of the test-suite (inputs, emulator driver, comparisons)."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")

ALL_OPS = ["ShearX", "ShearY", "TranslateX", "TranslateY", "Rotate", "AutoContrast", "Invert",
           "Equalize", "Solarize", "Posterize", "Contrast", "Color", "Brightness", "Sharpness",
           "Cutout", "CutoutAbs", "Posterize2", "TranslateXAbs", "TranslateYAbs"]


def synth(shape, kind, rng):
    """The three input families of SURVEY.md 8(d) (same generator as tests/golden/make_golden.py)."""
    h, w = shape
    if kind == 0:
        return rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    if kind == 1:
        lo = int(rng.integers(0, 200))
        hi = int(rng.integers(lo + 1, 256))
        ramp = np.linspace(lo, hi, w)[None, :, None] + rng.normal(0, 8, (h, w, 3))
        return np.clip(ramp, 0, 255).astype(np.uint8)
    return np.broadcast_to(rng.integers(0, 256, 3, dtype=np.uint8), (h, w, 3)).copy()


def synth_batch(n, shape, seed):
    rng = np.random.default_rng(seed)
    return np.stack([synth(shape, i % 3, rng) for i in range(n)])


def exact_norm_table(mean, std):
    """fp32 ToTensor+Normalize value of every byte, computed with torch itself."""
    import torch
    u = torch.arange(256, dtype=torch.uint8)
    x = u.to(torch.float32).div(255)
    m = torch.as_tensor(mean, dtype=torch.float32)[:, None]
    s = torch.as_tensor(std, dtype=torch.float32)[:, None]
    return ((x[None, :] - m) / s).numpy().astype(np.float32).copy()


def emu_augment(emu, pol, batch_u8, samples, boxes, tail=None, norm=None, partner=None, lam=1.0,
                force_generic=False, pool=None, pool_samples=None, pool_boxes=None, first=0):
    """Drive tests/emu like engine.augment_batch drives the kernels.  norm=None -> uint8 HWC."""
    from fast_autoaugment_b200.engine import TailSpec
    tail = tail or TailSpec.raw_u8()
    B, H, W, _ = batch_u8.shape
    oh, ow = tail.out_size if tail.out_size is not None else (H, W)
    table = np.ascontiguousarray(pol.compiled_table(H, W))
    src = np.ascontiguousarray(batch_u8 if pool is None else pool)
    smp = np.ascontiguousarray(samples if pool_samples is None else pool_samples)
    bxs = np.ascontiguousarray(boxes if pool_boxes is None else pool_boxes)
    cur = src
    n_op = pol.n_op
    base = 0
    while base + 2 < n_op:                      # chained windows, like the engine
        nxt = np.zeros_like(cur)
        rc = emu.faa_emu_augment(cur.ctypes.data, cur.shape[0], 0, cur.shape[0], H, W, table.ctypes.data,
                                 pol.n_sub, n_op, smp.ctypes.data, bxs.ctypes.data, base, 0, H, W, 0, None,
                                 nxt.ctypes.data, None, C.c_float(1.0), C.c_float(0.0), int(force_generic))
        assert rc == 0
        cur, base = nxt, base + 2
    if norm is None:
        out = np.zeros((B, oh, ow, 3), np.uint8)
        tab = None
    else:
        out = np.zeros((B, 3, oh, ow), np.float32)
        norm = np.ascontiguousarray(norm, dtype=np.float32)
        tab = norm.ctypes.data
    part = None
    if partner is not None:
        partner = np.ascontiguousarray(partner, dtype=np.int32)
        part = partner.ctypes.data
    rc = emu.faa_emu_augment(cur.ctypes.data, cur.shape[0], first, B, H, W, table.ctypes.data, pol.n_sub, n_op,
                             smp.ctypes.data, bxs.ctypes.data, base, 1, oh, ow, int(tail.cutout > 0), tab,
                             out.ctypes.data, part, C.c_float(np.float32(lam)), C.c_float(np.float32(1 - lam)),
                             int(force_generic))
    assert rc == 0
    return out


def set_emu_sigs(emu):
    vp = C.c_void_p
    emu.faa_emu_augment.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, C.c_int, C.c_int, vp, vp,
                                    C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, vp, C.c_float, C.c_float,
                                    C.c_int]
    emu.faa_emu_augment.restype = C.c_int
    emu.faa_emu_philox.argtypes = [vp, vp, C.c_int, C.c_int, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp]
    emu.faa_emu_philox.restype = C.c_int
    emu.faa_emu_philox_block.argtypes = [vp, vp, vp]
    emu.faa_emu_philox_block.restype = None
    return emu


def seed_all(s):
    import random
    import torch
    random.seed(s)
    np.random.seed(s)
    torch.manual_seed(s)
