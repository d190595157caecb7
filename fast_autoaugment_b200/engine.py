"""Host-side engine: policy handle, samplers and the batched launch.

``CompiledPolicy`` owns the C-ABI policy handle built from the reference's policy format
(``list[list[(op_name, prob, level)]]``, reference ``archive.py``).  Two samplers decide,
per image, what the reference's RNG draws decide:

* ``sample_parity``  - consumes the SAME global generators as the reference, in the same
  order (Python ``random``: sub-policy choice, gates, mirror signs - reference
  ``data.py:257-264``, ``augmentations.py:15,22,29,37,45,52,59``; ``numpy.random``: Cutout
  centres ``augmentations.py:131-132`` and CutoutDefault ``data.py:239-240``; torch CPU
  generator: RandomCrop / RandomHorizontalFlip), image after image, i.e. a
  ``num_workers=0`` DataLoader.  Seeding those generators like the reference reproduces
  the reference's output bit for bit.  Host cost ~2 us/image: for tests and drop-in use.
* Philox (``rng=``)  - the kernel draws the decisions itself (counter-based, keyed by
  (seed, global sample index)); same distributions, different stream; no host work.
"""
from __future__ import annotations

import ctypes as C
import random
from dataclasses import dataclass, field

import numpy as np
import torch

from . import _lib
from ._lib import lib, check

CIFAR_MEAN, CIFAR_STD = (0.4914, 0.4822, 0.4465), (0.2023, 0.1994, 0.2010)   # reference data.py:34
IMAGENET_MEAN, IMAGENET_STD = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)    # reference data.py:72

_DTYPES = {torch.float16: _lib.F16, torch.bfloat16: _lib.BF16, torch.float32: _lib.F32, torch.uint8: _lib.U8_HWC}


@dataclass
class TailSpec:
    """What follows the policy in ``transform_train`` (reference data.py:39-44,64,70-72,111-112)."""
    out_size: tuple | None = None      # RandomCrop size; None = same as the input
    crop_pad: int = 0                  # RandomCrop padding (data.py:40); 0 = no crop
    hflip: bool = False                # RandomHorizontalFlip (data.py:41,64)
    mean: tuple = CIFAR_MEAN
    std: tuple = CIFAR_STD
    cutout: int = 0                    # CutoutDefault length (data.py:111-112); 0 = off
    out_dtype: torch.dtype = torch.float16

    @staticmethod
    def cifar(cutout=16, out_dtype=torch.float16):
        return TailSpec((32, 32), 4, True, CIFAR_MEAN, CIFAR_STD, cutout, out_dtype)

    @staticmethod
    def imagenet(cutout=0, out_dtype=torch.float16):
        return TailSpec(None, 0, True, IMAGENET_MEAN, IMAGENET_STD, cutout, out_dtype)

    @staticmethod
    def raw_u8():
        return TailSpec(None, 0, False, CIFAR_MEAN, CIFAR_STD, 0, torch.uint8)

    def c_struct(self, h, w):
        oh, ow = self.out_size if self.out_size is not None else (h, w)
        t = _lib.Tail()
        t.out_h, t.out_w = int(oh), int(ow)
        t.out_dtype = _DTYPES[self.out_dtype]
        t.use_zero_box = 1 if self.cutout > 0 else 0
        t.crop_pad = int(self.crop_pad)
        for i in range(3):
            t.mean[i] = float(self.mean[i])
            t.std[i] = float(self.std[i])
        return t


class CompiledPolicy:
    """C-ABI policy handle for a reference-format policy list."""

    def __init__(self, policies):
        subs = [list(s) for s in policies]
        if not subs:
            raise IndexError("Cannot choose from an empty sequence")       # random.choice on []
        n_op = max(max(len(s) for s in subs), 1)
        # ragged sub-policies (the reference's Augmentation accepts them): short ones are padded with slots that never
        # fire (probability -1) and draw nothing - `pad[s][j]` marks them for the parity sampler
        self.pad = np.array([[j >= len(s) for j in range(n_op)] for s in subs], dtype=bool)
        subs = [s + [("Invert", -1.0, 0.0)] * (n_op - len(s)) for s in subs]
        self.policies = subs
        self.n_sub, self.n_op = len(subs), n_op
        self.names = [[str(op[0]) for op in s] for s in subs]
        ids = np.array([[_lib.op_id(op[0]) for op in s] for s in subs], dtype=np.int32)
        self.probs = np.array([[float(op[1]) for op in s] for s in subs], dtype=np.float64)
        self.levels = np.array([[float(op[2]) for op in s] for s in subs], dtype=np.float64)
        self.ids = ids
        h = C.c_void_p()
        check(lib.faa_policy_create(ids.ctypes.data, self.probs.ctypes.data, self.levels.ctypes.data,
                                    self.n_sub, self.n_op, C.byref(h)))
        self.handle = h
        self.draw = np.array([[lib.faa_policy_draw_kind(h, s, j) for j in range(n_op)]
                              for s in range(self.n_sub)], dtype=np.int8)

    def __del__(self):
        h = getattr(self, "handle", None)
        if h:
            try:
                lib.faa_policy_destroy(h)
            except Exception:
                pass
            self.handle = None

    # -- host-side compiled records (no GPU needed) -------------------------------------
    def compiled_op(self, h, w, sub, op, sign=0):
        out = np.zeros(8, dtype=np.int32)
        check(lib.faa_policy_compiled_op(self.handle, h, w, sub, op, int(sign), out.ctypes.data))
        return out

    def compiled_table(self, h, w):
        """[n_sub][n_op][2][8] int32; raises like the reference for the first bad op."""
        t = np.zeros((self.n_sub, self.n_op, 2, 8), dtype=np.int32)
        for s in range(self.n_sub):
            for j in range(self.n_op):
                for sg in range(2):
                    t[s, j, sg] = self.compiled_op(h, w, s, j, sg)
        return t

    def cutout_box(self, h, w, sub, op, ux, uy):
        b = np.zeros(1, dtype=_lib.BOX_DTYPE)
        check(lib.faa_cutout_box(self.handle, h, w, sub, op, float(ux), float(uy), b.ctypes.data))
        return b[0]

    # -- samplers -----------------------------------------------------------------------
    def sample_parity(self, batch, h, w, tail: TailSpec | None = None):
        """Replay the reference's draws from the global generators (see module doc)."""
        tail = tail or TailSpec.raw_u8()
        oh, ow = tail.out_size if tail.out_size is not None else (h, w)
        samples = np.zeros(batch, dtype=_lib.SAMPLE_DTYPE)
        boxes = np.zeros((batch, self.n_op), dtype=_lib.BOX_DTYPE)
        boxes["x1"] = -1
        boxes["y1"] = -1
        subs_range = range(self.n_sub)
        crop_rng_h = h + 2 * tail.crop_pad - oh + 1
        crop_rng_w = w + 2 * tail.crop_pad - ow + 1
        do_crop = tail.crop_pad > 0 or (oh, ow) != (h, w)
        if crop_rng_h < 1 or crop_rng_w < 1:
            raise ValueError("Required crop size %s is larger than input image size %s" %
                             ((oh, ow), (h + 2 * tail.crop_pad, w + 2 * tail.crop_pad)))
        if do_crop and (tail.crop_pad > 127 or crop_rng_h - 1 - tail.crop_pad > 127 or crop_rng_w - 1 - tail.crop_pad > 127):
            raise _lib.FaaRuntimeError("RandomCrop offsets beyond +-127 pixels are not supported (int8 records): "
                                       "crop on the host side")
        for i in range(batch):
            sub = random.choice(subs_range)                       # data.py:259
            gate = sign = 0
            for j in range(self.n_op):
                if self.pad[sub, j]:                              # padding of a ragged sub-policy: no op, no draw
                    continue
                if random.random() > self.probs[sub, j]:          # data.py:261
                    continue
                gate |= 1 << j
                d = self.draw[sub, j]
                if d < 0:
                    raise KeyError(self.names[sub][j])            # augmentations.py:189
                # validates the magnitude exactly when the reference would assert
                self.compiled_op(h, w, sub, j, 0)
                if d == _lib.DRAW_MIRROR:
                    if random.random() > 0.5:                     # augmentations.py:15 ...
                        sign |= 1 << j
                elif d == _lib.DRAW_BOX:
                    ux = np.random.random_sample()                # the u inside uniform(w), :131
                    uy = np.random.random_sample()                # :132
                    boxes[i, j] = self.cutout_box(h, w, sub, j, ux, uy)
            s = samples[i]
            s["sub"], s["gate"], s["sign"] = sub, gate, sign
            if do_crop and not (crop_rng_h == 1 and crop_rng_w == 1):     # torchvision RandomCrop.get_params
                top = int(torch.randint(0, crop_rng_h, size=(1,)).item())
                left = int(torch.randint(0, crop_rng_w, size=(1,)).item())
                s["crop_dy"], s["crop_dx"] = top - tail.crop_pad, left - tail.crop_pad
            if tail.hflip:
                s["flip"] = 1 if bool(torch.rand(1) < 0.5) else 0        # RandomHorizontalFlip
            if tail.cutout > 0:                                           # data.py:239-246
                cy = np.random.randint(oh)
                cx = np.random.randint(ow)
                half = tail.cutout // 2
                s["zero_box"] = (np.clip(cy - half, 0, oh), np.clip(cy + half, 0, oh),
                                 np.clip(cx - half, 0, ow), np.clip(cx + half, 0, ow))
        return samples, boxes

    def sample_policy_mt(self, batch, h, w, py_state=None, np_state=None):
        """C++ replay of the policy draws from explicit MT19937 states.  With the default
        arguments the live global states of ``random`` / ``numpy.random`` are read, advanced
        and written back - equivalent to the policy part of ``sample_parity``."""
        live = py_state is None and np_state is None
        if live:
            st = random.getstate()
            py_state = np.array(st[1], dtype=np.uint32)
            nst = np.random.get_state()
            np_state = np.concatenate([nst[1].astype(np.uint32), np.array([nst[2]], dtype=np.uint32)])
        py_state = np.ascontiguousarray(py_state, dtype=np.uint32)
        np_state = np.ascontiguousarray(np_state, dtype=np.uint32)
        samples = np.zeros(batch, dtype=_lib.SAMPLE_DTYPE)
        boxes = np.zeros((batch, self.n_op), dtype=_lib.BOX_DTYPE)
        check(lib.faa_sample_policy_mt(self.handle, batch, h, w, py_state.ctypes.data, np_state.ctypes.data,
                                       samples.ctypes.data, boxes.ctypes.data))
        if live:
            random.setstate((st[0], tuple(int(v) for v in py_state), st[2]))
            np.random.set_state((nst[0], np_state[:624].copy(), int(np_state[624]), nst[3], nst[4]))
        return samples, boxes, py_state, np_state


def make_rng(seed, first_index=0, tail: TailSpec | None = None):
    r = _lib.Rng()
    r.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    r.first_index = int(first_index)
    if tail is not None:
        r.crop_pad, r.hflip, r.zero_box_len = int(tail.crop_pad), int(bool(tail.hflip)), int(tail.cutout)
    return r


def _require_cuda(t, what):
    if not (isinstance(t, torch.Tensor) and t.is_cuda):
        raise _lib.FaaRuntimeError("%s must be a CUDA tensor: the augmentation path is CUDA-only" % what)


def _stream_ptr(device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def augment_batch(policy: CompiledPolicy, batch_u8: torch.Tensor, tail: TailSpec, samples=None, boxes=None,
                  rng=None, out=None, partner=None, lam=1.0, pool=None, pool_samples=None, pool_boxes=None,
                  first=0, lighting_rgb=None):
    """uint8 NHWC CUDA batch -> augmented NCHW ``tail.out_dtype`` (or uint8 NHWC).

    samples/boxes: resolved decisions (numpy structured arrays or CUDA uint8 tensors) - parity
    mode; or rng (``make_rng``) - fused Philox mode.  partner/lam: fused Mixup; ``pool`` is the
    array partner indexes into (defaults to ``batch_u8`` itself), ``first`` the position of this
    batch inside the pool (multi-GPU global pairing).
    """
    _require_cuda(batch_u8, "batch")
    if batch_u8.dtype != torch.uint8 or batch_u8.dim() != 4 or batch_u8.shape[-1] != 3:
        raise ValueError("batch must be uint8 [B, H, W, 3]")
    batch_u8 = batch_u8.contiguous()
    dev = batch_u8.device
    B, H, W, _ = batch_u8.shape
    t = tail.c_struct(H, W)
    if tail.out_dtype == torch.uint8:
        shape = (B, t.out_h, t.out_w, 3)
    else:
        shape = (B, 3, t.out_h, t.out_w)
    if out is None:
        out = torch.empty(shape, dtype=tail.out_dtype, device=dev)
    elif tuple(out.shape) != shape or out.dtype != tail.out_dtype or not out.is_contiguous():
        raise ValueError("out has the wrong shape/dtype")

    if lighting_rgb is not None:
        # Lighting (reference augmentations.py:197-215, between ToTensor and Normalize): per-image offsets [B,3] fp32
        lighting_rgb = lighting_rgb.to(device=dev, dtype=torch.float32).contiguous()
        if tuple(lighting_rgb.shape) != (B, 3):
            raise ValueError("lighting_rgb must be [B, 3]")
        policy._lighting_keep = lighting_rgb                 # stays alive until the launch has run
        check(lib.faa_policy_set_lighting(policy.handle, lighting_rgb.data_ptr(), B))
    try:
        return _augment_launch(policy, batch_u8, tail, samples, boxes, rng, out, partner, lam, pool, pool_samples, pool_boxes,
                               first, dev, B, H, W, t)
    finally:
        if lighting_rgb is not None:
            check(lib.faa_policy_set_lighting(policy.handle, None, 0))


def _augment_launch(policy, batch_u8, tail, samples, boxes, rng, out, partner, lam, pool, pool_samples, pool_boxes, first, dev,
                    B, H, W, t):
    def to_dev(a, itemsize):
        if a is None:
            return None
        if isinstance(a, torch.Tensor):
            _require_cuda(a, "records")
            return a
        flat = np.ascontiguousarray(a).view(np.uint8).reshape(-1)
        return torch.from_numpy(flat.copy()).to(dev, non_blocking=False)

    with torch.cuda.device(dev):
        stream = _stream_ptr(dev)
        rng_p = C.byref(rng) if rng is not None else None
        if partner is None:
            d_s, d_b = to_dev(samples, 16), to_dev(boxes, 8)
            n_op = policy.n_op
            if n_op <= _lib.MAX_FUSED_OPS:
                check(lib.faa_augment(policy.handle, batch_u8.data_ptr(), out.data_ptr(), B, H, W, C.byref(t),
                                      d_s.data_ptr() if d_s is not None else None,
                                      d_b.data_ptr() if d_b is not None else None, rng_p, 0, stream))
            else:
                # chained launches: every window of 2 ops but the last writes uint8 HWC
                mid = _lib.Tail()
                mid.out_h, mid.out_w, mid.out_dtype, mid.use_zero_box = H, W, _lib.U8_HWC, 0
                cur = batch_u8
                base = 0
                while base + _lib.MAX_FUSED_OPS < n_op:
                    nxt = torch.empty_like(batch_u8)
                    check(lib.faa_augment(policy.handle, cur.data_ptr(), nxt.data_ptr(), B, H, W, C.byref(mid),
                                          d_s.data_ptr() if d_s is not None else None,
                                          d_b.data_ptr() if d_b is not None else None, rng_p, base, stream))
                    cur = nxt
                    base += _lib.MAX_FUSED_OPS
                check(lib.faa_augment(policy.handle, cur.data_ptr(), out.data_ptr(), B, H, W, C.byref(t),
                                      d_s.data_ptr() if d_s is not None else None,
                                      d_b.data_ptr() if d_b is not None else None, rng_p, base, stream))
        else:
            pool_t = batch_u8 if pool is None else pool.contiguous()
            _require_cuda(pool_t, "pool")
            d_s = to_dev(samples if pool_samples is None else pool_samples, 16)
            d_b = to_dev(boxes if pool_boxes is None else pool_boxes, 8)
            if isinstance(partner, torch.Tensor):
                d_p = partner.to(device=dev, dtype=torch.int32).contiguous()
            else:
                d_p = torch.as_tensor(np.asarray(partner, dtype=np.int32), device=dev)
            check(lib.faa_augment_mixup(policy.handle, pool_t.data_ptr(), int(pool_t.shape[0]), int(first),
                                        out.data_ptr(), B, H, W, C.byref(t),
                                        d_s.data_ptr() if d_s is not None else None,
                                        d_b.data_ptr() if d_b is not None else None, rng_p,
                                        d_p.data_ptr(), float(np.float32(lam)), float(np.float32(1 - lam)),
                                        stream))
    return out


class FusedAugmenter:
    """Pre-bound production launch: fused Philox sampling + augmentation in ONE kernel launch and
    one ctypes call per batch (no per-call Python allocation), for 2-op policies.

        aug = FusedAugmenter(policy, tail, h, w, seed)
        aug(batch_u8_cuda, out_cuda, first_index)     # asynchronous on the current stream
    """

    def __init__(self, policy: CompiledPolicy, tail: TailSpec, h: int, w: int, seed: int = 0, overlap_calls: bool = False):
        """``overlap_calls=True``: consecutive ``__call__``s on one stream may overlap on the GPU (C ABI
        ``faa_policy_set_overlap``) - only if every call's input batch was complete before the PREVIOUS call was issued
        (device-resident data, a producer that runs a batch ahead).  ``run_many`` overlaps its steps regardless."""
        if policy.n_op > _lib.MAX_FUSED_OPS:
            raise ValueError("FusedAugmenter handles policies of at most 2 ops; use augment_batch")
        if overlap_calls:
            check(lib.faa_policy_set_overlap(policy.handle, 1))
        self.policy, self.tail, self.h, self.w = policy, tail, h, w
        self.t = tail.c_struct(h, w)
        self.rng = make_rng(seed, 0, tail)
        self._t_ref, self._rng_ref = C.byref(self.t), C.byref(self.rng)
        self.out_shape = ((0, self.t.out_h, self.t.out_w, 3) if tail.out_dtype == torch.uint8
                          else (0, 3, self.t.out_h, self.t.out_w))

    def empty_out(self, batch, device="cuda"):
        return torch.empty((batch,) + tuple(self.out_shape[1:]), dtype=self.tail.out_dtype, device=device)

    def __call__(self, batch_u8: torch.Tensor, out: torch.Tensor, first_index: int = 0, stream=None):
        b = batch_u8.shape[0]
        if (batch_u8.dtype != torch.uint8 or not batch_u8.is_cuda or not batch_u8.is_contiguous()
                or tuple(batch_u8.shape[1:]) != (self.h, self.w, 3)):
            raise ValueError("batch must be a contiguous uint8 CUDA tensor [B, %d, %d, 3]" % (self.h, self.w))
        if (out.dtype != self.tail.out_dtype or out.device != batch_u8.device or not out.is_contiguous()
                or tuple(out.shape) != (b,) + tuple(self.out_shape[1:])):
            raise ValueError("out must be a contiguous %s tensor %s on %s" % (self.tail.out_dtype, (b,) + tuple(self.out_shape[1:]),
                                                                             batch_u8.device))
        self.rng.first_index = first_index
        dev = batch_u8.device
        s = stream if stream is not None else torch.cuda.current_stream(dev).cuda_stream
        if dev.index == torch.cuda.current_device():
            check(lib.faa_augment(self.policy.handle, batch_u8.data_ptr(), out.data_ptr(), b, self.h,
                                  self.w, self._t_ref, None, None, self._rng_ref, 0, C.c_void_p(s)))
        else:
            with torch.cuda.device(dev):
                check(lib.faa_augment(self.policy.handle, batch_u8.data_ptr(), out.data_ptr(), b, self.h,
                                      self.w, self._t_ref, None, None, self._rng_ref, 0, C.c_void_p(s)))
        return out


    def plan_many(self, batches, outs):
        """Pointer tables for :meth:`run_many` (build once, reuse every epoch): validates the tensors like ``__call__``."""
        if len(batches) != len(outs) or not batches:
            raise ValueError("need as many outputs as batches (at least one)")
        b = batches[0].shape[0]
        for x, o in zip(batches, outs):
            if (x.dtype != torch.uint8 or not x.is_cuda or not x.is_contiguous() or tuple(x.shape) != (b, self.h, self.w, 3)):
                raise ValueError("every batch must be a contiguous uint8 CUDA tensor [%d, %d, %d, 3]" % (b, self.h, self.w))
            if (o.dtype != self.tail.out_dtype or o.device != x.device or x.device != batches[0].device or not o.is_contiguous()
                    or tuple(o.shape) != (b,) + tuple(self.out_shape[1:])):
                raise ValueError("every out must be a contiguous %s tensor %s on the batches' device"
                                 % (self.tail.out_dtype, (b,) + tuple(self.out_shape[1:])))
        n = len(batches)
        ins = (C.c_void_p * n)(*[x.data_ptr() for x in batches])
        dst = (C.c_void_p * n)(*[o.data_ptr() for o in outs])
        return (n, b, ins, dst, batches[0].device, list(batches), list(outs))      # (keeps the tensors alive)

    def run_many(self, plan, first_index: int = 0, stride=None, stream=None):
        """Augment the planned batches back to back in ONE call (C ABI ``faa_augment_many``): step k == ``self(batches[k],
        outs[k], first_index + k * stride)``; ``stride`` defaults to the batch size."""
        n, b, ins, dst, dev = plan[:5]
        self.rng.first_index = first_index
        s = stream if stream is not None else torch.cuda.current_stream(dev).cuda_stream
        with torch.cuda.device(dev):
            check(lib.faa_augment_many(self.policy.handle, n, ins, dst, b, self.h, self.w, self._t_ref, self._rng_ref,
                                       int(b if stride is None else stride), C.c_void_p(s)))
        return plan[6]


def augment_tta(policy: CompiledPolicy, batch_u8: torch.Tensor, tail: TailSpec, replicas: int, seed: int, first_index: int = 0,
                out=None):
    """Test-time-augmentation batching for the policy search (reference search.py:87-125): the reference builds
    ``num_policy`` validation loaders that each augment the SAME validation batch with their own random draws and then
    reduces the per-sample losses over the replicas.  Here ONE launch produces all replicas:

        out[r] == augment_batch(policy, batch_u8, tail, rng=make_rng(seed, first_index + r * B, tail))

    uint8 [B,H,W,3] CUDA batch -> [replicas, B, 3, out_h, out_w] (or [replicas, B, out_h, out_w, 3] uint8)."""
    _require_cuda(batch_u8, "batch")
    if batch_u8.dtype != torch.uint8 or batch_u8.dim() != 4 or batch_u8.shape[-1] != 3:
        raise ValueError("batch must be uint8 [B, H, W, 3]")
    batch_u8 = batch_u8.contiguous()
    B, H, W, _ = batch_u8.shape
    t = tail.c_struct(H, W)
    shape = (replicas, B, t.out_h, t.out_w, 3) if tail.out_dtype == torch.uint8 else (replicas, B, 3, t.out_h, t.out_w)
    if out is None:
        out = torch.empty(shape, dtype=tail.out_dtype, device=batch_u8.device)
    elif tuple(out.shape) != shape or out.dtype != tail.out_dtype or not out.is_contiguous():
        raise ValueError("out has the wrong shape/dtype")
    rng = make_rng(seed, first_index, tail)
    with torch.cuda.device(batch_u8.device):
        check(lib.faa_augment_tta(policy.handle, batch_u8.data_ptr(), out.data_ptr(), B, int(replicas), H, W, C.byref(t),
                                  C.byref(rng), _stream_ptr(batch_u8.device)))
    return out
