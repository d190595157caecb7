"""Multi-GPU (needs >= 2 visible GPUs): NCCL all-gather of raw uint8 shards + fused global Mixup
equals the single-GPU result on the global batch."""
import os
import socket

import numpy as np
import pytest
import torch

from helpers import ROOT, synth_batch

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, out_dir):
    import sys
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from fast_autoaugment_b200 import archive
    from fast_autoaugment_b200.distributed import mixup_global, shard_bounds
    from fast_autoaugment_b200.engine import CompiledPolicy, TailSpec
    n = 128
    batch = synth_batch(n, (64, 64), seed=3)
    lo, hi = shard_bounds(n, rank, world)
    pol = CompiledPolicy(archive.fa_resnet50_rimagenet())
    tail = TailSpec.imagenet(0, torch.float32)
    x = torch.from_numpy(batch[lo:hi]).cuda()
    y = torch.arange(lo, hi).cuda()
    data, t1, t2, lam = mixup_global(pol, x, y, tail, 0.2, seed=11, step=4)
    torch.cuda.synchronize()
    np.save(os.path.join(out_dir, "rank%d.npy" % rank), data.cpu().numpy())
    np.save(os.path.join(out_dir, "t2_%d.npy" % rank), t2.cpu().numpy())
    # the peer route (exchange fused into the mix kernel through NVLink peer memory) over several steps and both slots,
    # with CutoutDefault boxes: must equal the all-to-all route bit for bit
    from fast_autoaugment_b200.distributed import PeerPool, mixup_global_peer
    tail_c = TailSpec.imagenet(16, torch.float16)
    print("[rank %d] all-to-all route done" % rank, file=sys.stderr, flush=True)
    pool = PeerPool(hi - lo, 64, 64, x.device)
    print("[rank %d] peer pool mapped" % rank, file=sys.stderr, flush=True)
    ok = True
    for step in range(5):
        a, _, ta, lam_a = mixup_global(pol, x, y, tail_c, 0.2, seed=5, step=step)
        torch.cuda.synchronize()
        b, _, tb, lam_b = mixup_global_peer(pol, x, y, tail_c, 0.2, seed=5, step=step, pool=pool)
        torch.cuda.synchronize()
        ok = ok and torch.equal(a, b) and torch.equal(ta, tb) and lam_a == lam_b
        print("[rank %d] step %d equal=%s" % (rank, step, ok), file=sys.stderr, flush=True)
    np.save(os.path.join(out_dir, "peer_ok_%d.npy" % rank), np.array([ok]))
    pool.close()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_global_mixup_two_gpus(tmp_path):
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    got = np.concatenate([np.load(os.path.join(tmp_path, "rank%d.npy" % r)) for r in range(2)])
    t2 = np.concatenate([np.load(os.path.join(tmp_path, "t2_%d.npy" % r)) for r in range(2)])
    from fast_autoaugment_b200 import archive
    from fast_autoaugment_b200.distributed import global_pairing
    from fast_autoaugment_b200.engine import CompiledPolicy, TailSpec, augment_batch, make_rng
    n = 128
    batch = torch.from_numpy(synth_batch(n, (64, 64), seed=3)).cuda()
    pol = CompiledPolicy(archive.fa_resnet50_rimagenet())
    tail = TailSpec.imagenet(0, torch.float32)
    perm, lam = global_pairing(n, 0.2, seed=11, step=4)
    want = augment_batch(pol, batch, tail, rng=make_rng(11, 4 * n, tail), partner=perm, lam=lam).cpu().numpy()
    assert np.array_equal(got, want)
    assert np.array_equal(t2, perm.numpy())
    assert all(bool(np.load(os.path.join(tmp_path, "peer_ok_%d.npy" % r))[0]) for r in range(2))
    plain = augment_batch(pol, batch, tail, rng=make_rng(11, 4 * n, tail))
    ref = (plain * np.float32(lam) + plain[perm.cuda()] * np.float32(1 - lam)).cpu().numpy()
    assert np.array_equal(want, ref)
