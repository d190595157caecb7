#!/usr/bin/env python
"""Minimal launch sequence for ncu: N steps of a workload.
   python tools/ncu_target.py <workload> <steps> [policy: real|identity|<OpName>]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import bench
from fast_autoaugment_b200 import archive
from fast_autoaugment_b200.engine import CompiledPolicy, FusedAugmenter, TailSpec

workload = sys.argv[1] if len(sys.argv) > 1 else "imagenet224_b512"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
which = sys.argv[3] if len(sys.argv) > 3 else "real"
H, W, B, pol_name, tail_kind, cutout = bench.WORKLOADS[workload]
if which == "real":
    policies = getattr(archive, pol_name)()
elif which == "identity":
    policies = [[("Invert", 0.0, 0.5), ("Invert", 0.0, 0.5)]]
else:
    policies = [[(which, 1.0, 0.7), (which, 0.0, 0.7)]]
pol = CompiledPolicy(policies)
tail = TailSpec.cifar(cutout, torch.float16) if tail_kind == "cifar" else TailSpec.imagenet(cutout, torch.float16)
f = FusedAugmenter(pol, tail, H, W, 2024)
xs = [torch.from_numpy(bench.synth_batch(B, H, W, 1234 + i)).cuda() for i in range(2)]
outs = [f.empty_out(B) for _ in range(2)]
for i in range(steps):
    f(xs[i % 2], outs[i % 2], i * B)
torch.cuda.synchronize()
print("done", workload, steps, which)
