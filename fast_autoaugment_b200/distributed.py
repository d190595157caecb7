"""Multi-GPU use of the hot path: one process per GPU, ``torch.distributed``.

Augmentation shards trivially - every op's dependencies are inside one image - so ranks process
their own contiguous shard with no data-path collective.  The only exchange step is **Mixup with
global pairing**, i.e. a permutation over the *global* batch (reference ``aug_mixup.py:13-23`` is
rank-local under DDP, ``train.py:55``; global pairing is an extension):

* the pairing and lambda are derived from a shared seed, identically on every rank (no
  communication);
* ranks all-gather their **raw uint8** shards (3 B/px: half of the fp16 output, a quarter of the
  fp32 one) - mixing is linear, so "augment the partner here" equals "augment there, send, mix";
* the fused-Mixup kernel recomputes each partner's augmentation from the gathered raw image with
  the partner's own decisions (Philox keyed by the *global* sample index, or gathered records).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from .engine import CompiledPolicy, TailSpec, augment_batch, make_rng


def shard_bounds(n: int, rank: int, world: int):
    """Contiguous shard [lo, hi) of n samples for `rank` (DistributedSampler-style equal shards:
    n must be divisible by world, reference data.py:205-212 pads instead)."""
    if n % world:
        raise ValueError("global batch %d is not divisible by world size %d" % (n, world))
    per = n // world
    return rank * per, (rank + 1) * per


def global_pairing(global_batch: int, alpha: float, seed: int, step: int):
    """(perm, lam) of reference ``mixup`` (aug_mixup.py:14,18-19) for the GLOBAL batch, identical on
    every rank without communication: both draws come from generators seeded with (seed, step)."""
    g = torch.Generator(device="cpu")
    g.manual_seed((int(seed) * 1000003 + int(step)) & 0x7FFFFFFFFFFFFFFF)
    perm = torch.randperm(global_batch, generator=g)
    rs = np.random.RandomState((int(seed) * 7919 + int(step)) & 0xFFFFFFFF)
    lam = rs.beta(alpha, alpha)
    lam = max(lam, 1.0 - lam)
    assert 0.0 <= lam <= 1.0, lam
    return perm, float(lam)


def gather_pool(local: torch.Tensor, group=None) -> torch.Tensor:
    """All-gather equal-sized shards along dim 0 (NCCL on CUDA tensors, gloo on CPU tensors)."""
    world = dist.get_world_size(group)
    local = local.contiguous()
    out = torch.empty((world * local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    if local.is_cuda:
        dist.all_gather_into_tensor(out, local, group=group)
    else:
        dist.all_gather(list(out.chunk(world, 0)), local, group=group)
    return out


def mixup_global(policy: CompiledPolicy, local_u8: torch.Tensor, targets: torch.Tensor, tail: TailSpec, alpha: float,
                 seed: int, step: int, group=None, samples=None, boxes=None):
    """Augment this rank's shard and mix every sample with its partner from the GLOBAL batch.

    Returns ``(data, targets, partner_targets, lam)`` like reference ``mixup`` (aug_mixup.py:23).
    Decisions: fused Philox keyed by the global sample index (default), or resolved records given
    for the local shard (``samples``/``boxes`` numpy arrays, all-gathered as bytes).
    """
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    b = local_u8.shape[0]
    n = b * world
    perm, lam = global_pairing(n, alpha, seed, step)
    lo, hi = shard_bounds(n, rank, world)
    pool = gather_pool(local_u8, group)
    all_targets = gather_pool(targets, group)
    partner = perm[lo:hi]
    rng = pool_s = pool_b = None
    if samples is None:
        rng = make_rng(seed, step * n, tail)
    else:
        dev = local_u8.device
        s = torch.from_numpy(np.ascontiguousarray(samples).view(np.uint8).reshape(b, -1).copy()).to(dev)
        bx = torch.from_numpy(np.ascontiguousarray(boxes).view(np.uint8).reshape(b, -1).copy()).to(dev)
        pool_s, pool_b = gather_pool(s, group).reshape(-1), gather_pool(bx, group).reshape(-1)
    data = augment_batch(policy, local_u8, tail, rng=rng, partner=partner, lam=lam, pool=pool,
                         pool_samples=pool_s, pool_boxes=pool_b, first=lo)
    return data, targets, all_targets[partner.to(all_targets.device)], lam
