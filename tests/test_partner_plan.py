"""Host logic of the sharded Mixup (fast_autoaugment_b200.distributed), every rank simulated in one process:
partner_plan (all-to-all route) and partner_pointers (NVLink peer route) for several world sizes."""
import numpy as np
import pytest
import torch

from fast_autoaugment_b200.distributed import global_pairing, partner_plan, partner_pointers, shard_bounds


@pytest.mark.parametrize("world,b", [(1, 16), (2, 8), (3, 5), (4, 16), (8, 32)])
@pytest.mark.parametrize("step", [0, 7])
def test_all_to_all_plan_delivers_every_partner_exactly_once(world, b, step):
    n = world * b
    perm, lam = global_pairing(n, 0.2, seed=13, step=step)
    assert sorted(perm.tolist()) == list(range(n)) and 0.5 <= lam <= 1.0
    data = torch.arange(n) * 10 + 3                                   # one "image" per global sample
    plans = [partner_plan(perm, r, world) for r in range(world)]
    # what rank s sends to rank r is exactly what r expects from s, in r's order
    for r in range(world):
        lo, hi = shard_bounds(n, r, world)
        send_idx, send_counts, recv_counts, partner_pool, recv_global = plans[r]
        assert send_counts[r] == 0 and recv_counts[r] == 0
        received = []
        for s in range(world):
            s_idx, s_counts = plans[s][0], plans[s][1]
            off = sum(s_counts[:r])
            chunk = s_idx[off:off + s_counts[r]] + s * b              # global indices rank s sends to rank r
            assert len(chunk) == recv_counts[s]
            received.append(chunk)
        received = torch.cat(received) if received else torch.empty(0, dtype=torch.int64)
        assert torch.equal(received, recv_global)
        pool = torch.cat([data[lo:hi], data[received]])               # [own shard | received]
        assert torch.equal(pool[partner_pool], data[perm[lo:hi]])     # every sample finds its partner
        assert sum(recv_counts) <= b
    # every image that leaves a rank is needed by exactly one remote sample
    total_sent = sum(sum(p[1]) for p in plans)
    remote_pairs = sum(int(perm[i]) // b != i // b for i in range(n))
    assert total_sent == remote_pairs


@pytest.mark.parametrize("world,b", [(1, 4), (2, 8), (8, 16)])
def test_partner_pointers_address_the_partner_inside_its_owners_buffer(world, b):
    n = world * b
    perm, _ = global_pairing(n, 0.2, seed=2, step=1)
    img = 48
    bases = [10_000 * (r + 1) for r in range(world)]                  # distinct mappings per owner
    for rank in range(world):
        ptrs = partner_pointers(perm, rank, world, bases, img)
        lo, hi = shard_bounds(n, rank, world)
        for i in range(b):
            p = int(perm[lo + i])
            assert int(ptrs[i]) == bases[p // b] + (p % b) * img
    with pytest.raises(ValueError):
        shard_bounds(10, 0, 3)
