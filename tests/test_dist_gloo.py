"""World-size-2 tests of the multi-GPU host logic on CPU (gloo): shard bounds, shared-seed global
Mixup pairing, all-gather of raw shards / decision records, and - through the host emulation of the
kernel - that shards computed per rank with global pairing equal the single-process global result."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import ROOT, emu_augment, exact_norm_table, seed_all, set_emu_sigs, synth_batch


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    import ctypes
    import sys
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from fast_autoaugment_b200 import archive
    from fast_autoaugment_b200.distributed import gather_pool, global_pairing, shard_bounds
    from fast_autoaugment_b200.engine import CIFAR_MEAN, CIFAR_STD, CompiledPolicy, TailSpec

    n, b = 32, 16
    lo, hi = shard_bounds(n, rank, world)
    assert (lo, hi) == (rank * b, (rank + 1) * b)
    perm, lam = global_pairing(n, 0.2, seed=5, step=3)
    both = [None, None]
    dist.all_gather_object(both, (perm.tolist(), lam))
    assert both[0] == both[1]                                   # identical on every rank, no communication
    assert sorted(perm.tolist()) == list(range(n)) and 0.5 <= lam <= 1.0

    # every rank owns a shard of the global batch and of the decisions (parity sampler, rank 0's draw)
    batch = synth_batch(n, (32, 32), seed=9)
    policies = archive.fa_reduced_cifar10()
    pol = CompiledPolicy(policies)
    tail = TailSpec.cifar(16, torch.float32)
    seed_all(77)
    samples, boxes = pol.sample_parity(n, 32, 32, tail)          # same seed => same global decisions
    pool = gather_pool(torch.from_numpy(batch[lo:hi]))
    assert torch.equal(pool, torch.from_numpy(batch))
    s_b = torch.from_numpy(samples[lo:hi].view(np.uint8).reshape(b, -1).copy())
    x_b = torch.from_numpy(boxes[lo:hi].view(np.uint8).reshape(b, -1).copy())
    pool_s = gather_pool(s_b).numpy().reshape(-1).view(samples.dtype)
    pool_b = gather_pool(x_b).numpy().reshape(-1).view(boxes.dtype).reshape(n, -1)
    assert pool_s.tobytes() == samples.tobytes() and pool_b.tobytes() == boxes.tobytes()

    # partner-only exchange (BASELINE config 4): every rank receives exactly the raw images its samples pair with
    from fast_autoaugment_b200.distributed import exchange_partners, partner_plan
    send_idx, send_counts, recv_counts, partner_pool, recv_global = partner_plan(perm, rank, world)
    assert send_counts[rank] == 0 and recv_counts[rank] == 0 and sum(recv_counts) <= b
    recv = exchange_partners(torch.from_numpy(batch[lo:hi]), send_idx, send_counts, recv_counts)
    pool2 = torch.cat([torch.from_numpy(batch[lo:hi]), recv])
    assert torch.equal(pool2[partner_pool], torch.from_numpy(batch)[perm[lo:hi]])       # every partner is in the pool
    assert torch.equal(torch.from_numpy(batch)[recv_global], recv)                      # ... under its global index
    assert pool2.shape[0] == b + sum(recv_counts) < n                                    # less than the whole pool

    # peer route (exchange fused into the mix kernel): the pointer table addresses, for every local sample, its partner's
    # image inside the owner's buffer.  Here the "mapped buffers" are offsets into the all-gathered pool.
    from fast_autoaugment_b200.distributed import partner_pointers
    img_bytes = 32 * 32 * 3
    bases = [1000 + r * b * img_bytes for r in range(world)]              # rank r's buffer starts at bases[r]
    ptrs = partner_pointers(perm, rank, world, bases, img_bytes)
    assert ptrs.dtype == torch.int64 and ptrs.shape == (b,)
    flat = torch.from_numpy(batch).reshape(-1)
    for i in range(b):
        off = int(ptrs[i]) - 1000
        assert off % img_bytes == 0 and off // img_bytes == int(perm[lo + i])           # = the partner's global index
        assert torch.equal(flat[off:off + img_bytes], torch.from_numpy(batch[int(perm[lo + i])]).reshape(-1))

    # without a GPU the exportable pool cannot be allocated: EVERY rank must get the exception (the caller then falls
    # back to the all-to-all route) - no rank may be left waiting in a collective
    from fast_autoaugment_b200.distributed import PeerPool
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="peer pool unavailable"):
            PeerPool(b, 32, 32, torch.device("cuda", 0))

    emu = set_emu_sigs(ctypes.CDLL(os.path.join(ROOT, "tests", "emu", "libfaa_emu.so")))
    norm = exact_norm_table(CIFAR_MEAN, CIFAR_STD)
    mine = emu_augment(emu, pol, batch[lo:hi], samples[lo:hi], boxes[lo:hi], tail, norm,
                       partner=perm[lo:hi].numpy(), lam=lam, pool=pool.numpy(), pool_samples=pool_s,
                       pool_boxes=pool_b, first=lo)
    np.save(os.path.join(out_dir, "rank%d.npy" % rank), mine)
    dist.barrier()
    dist.destroy_process_group()


def test_global_mixup_sharded_equals_single_process(emu, tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    got = np.concatenate([np.load(os.path.join(tmp_path, "rank%d.npy" % r)) for r in range(world)])
    from fast_autoaugment_b200 import archive
    from fast_autoaugment_b200.distributed import global_pairing
    from fast_autoaugment_b200.engine import CIFAR_MEAN, CIFAR_STD, CompiledPolicy, TailSpec
    n = 32
    batch = synth_batch(n, (32, 32), seed=9)
    pol = CompiledPolicy(archive.fa_reduced_cifar10())
    tail = TailSpec.cifar(16, torch.float32)
    seed_all(77)
    samples, boxes = pol.sample_parity(n, 32, 32, tail)
    perm, lam = global_pairing(n, 0.2, seed=5, step=3)
    norm = exact_norm_table(CIFAR_MEAN, CIFAR_STD)
    want = emu_augment(emu, pol, batch, samples, boxes, tail, norm, partner=perm.numpy(), lam=lam)
    assert np.array_equal(got, want)
    # and equals "augment everything, then reference mixup formula"
    plain = emu_augment(emu, pol, batch, samples, boxes, tail, norm)
    ref = plain * np.float32(lam) + plain[perm.numpy()] * np.float32(1 - lam)
    assert np.array_equal(want, ref.astype(np.float32))


def test_shard_bounds_rejects_ragged():
    from fast_autoaugment_b200.distributed import shard_bounds
    with pytest.raises(ValueError):
        shard_bounds(10, 0, 4)
    assert [shard_bounds(12, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 9), (9, 12)]
