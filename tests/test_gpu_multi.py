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
    plain = augment_batch(pol, batch, tail, rng=make_rng(11, 4 * n, tail))
    ref = (plain * np.float32(lam) + plain[perm.cuda()] * np.float32(1 - lam)).cpu().numpy()
    assert np.array_equal(want, ref)
