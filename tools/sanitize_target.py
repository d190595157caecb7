#!/usr/bin/env python
"""compute-sanitizer target: every ordered op pair through all three pixel kernels (light / mid / cluster), once with
single-band clusters (48x64) and once with multi-band clusters and DSMEM exchanges (224x224), plus the non-split
cluster-kernel-only path, fused Mixup, TTA replicas and the chained schedule.
Usage (GPU box): compute-sanitizer --tool memcheck|racecheck|synccheck python tools/sanitize_target.py"""
import os, random, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
from helpers import ALL_OPS, synth_batch
from fast_autoaugment_b200 import archive
from fast_autoaugment_b200.engine import CompiledPolicy, FusedAugmenter, TailSpec, augment_batch, augment_tta, make_rng

rng = random.Random(3)
pairs = [[(a, 1.0, rng.random()), (b, 1.0, rng.random())] for a in ALL_OPS for b in ALL_OPS]
singles = [[(a, 1.0, rng.random()), (a, 0.0, 0.5)] for a in ALL_OPS]
pol = CompiledPolicy(pairs + singles)
n = pol.n_sub
CASES = (((48, 64), "0"), ((224, 224), "0"), ((40, 56), "1000000000000"))
U8_ONLY = "u8only" in sys.argv          # diagnostic: only the uint8 launch of the 224x224 case
if U8_ONLY: CASES = CASES[1:2]
for shape, split in CASES:
    os.environ["FAA_SPLIT_MIN"] = split
    x = torch.from_numpy(synth_batch(n, shape, seed=1)).cuda()
    s = np.zeros(n, dtype=np.dtype([("sub", "<u2"), ("gate", "u1"), ("sign", "u1"), ("crop_dy", "i1"), ("crop_dx", "i1"), ("flip", "u1"),
                                    ("reserved", "u1"), ("zero_box", "<i2", (4,))]))
    s["sub"] = np.arange(n); s["gate"] = 3; s["flip"] = np.arange(n) & 1; s["sign"] = (np.arange(n) >> 1) & 3
    b = np.zeros((n, 2), dtype=np.dtype([("x0", "<i2"), ("y0", "<i2"), ("x1", "<i2"), ("y1", "<i2")]))
    b["x0"] = 3; b["y0"] = 5; b["x1"] = shape[1] // 2; b["y1"] = shape[0] // 2
    for dt in (() if U8_ONLY else (torch.float16, torch.float32)):
        out = augment_batch(pol, x, TailSpec(None, 0, True, (0.485, 0.456, 0.406), (0.229, 0.224, 0.225), 8, dt), s, b)
    out = augment_batch(pol, x, TailSpec.raw_u8(), s, b)
    torch.cuda.synchronize()
    print("ok", shape, split, float(out.float().mean()), flush=True)
if U8_ONLY: sys.exit(0)
os.environ["FAA_SPLIT_MIN"] = "0"
p2 = CompiledPolicy(archive.fa_resnet50_rimagenet())
tail = TailSpec.imagenet(0, torch.float16)
x = torch.from_numpy(synth_batch(96, (224, 224), seed=2)).cuda()
for chain in ("0", "1"):
    os.environ["FAA_CHAIN"] = chain
    f = FusedAugmenter(p2, tail, 224, 224, 5, overlap_calls=True)
    outs = [f(x, f.empty_out(96), i * 96) for i in range(4)]
    torch.cuda.synchronize()
os.environ["FAA_CHAIN"] = "0"
perm = torch.randperm(96)
augment_batch(p2, x, TailSpec.imagenet(0, torch.float32), rng=make_rng(1, 0, tail), partner=perm, lam=0.7)
augment_tta(p2, x, tail, 3, seed=2, first_index=7)
pc = CompiledPolicy(archive.fa_reduced_cifar10())
xc = torch.from_numpy(synth_batch(256, (32, 32), seed=4)).cuda()
augment_batch(pc, xc, TailSpec.cifar(16, torch.float16), rng=make_rng(3, 0, TailSpec.cifar(16)))
torch.cuda.synchronize()
print("done", flush=True)
