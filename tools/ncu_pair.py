#!/usr/bin/env python
"""One two-op sub-policy applied to every image (ncu target).  Usage: ncu_pair.py OpA OpB [launches]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, bench
from fast_autoaugment_b200.engine import CompiledPolicy, FusedAugmenter, TailSpec
H = W = 224; B = 512
a, b = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 4
x = [torch.from_numpy(bench.synth_batch(B, H, W, 1 + i)).cuda() for i in range(2)]
pol = CompiledPolicy([[(a, 1.0, 0.7), (b, 1.0, 0.6)]])
f = FusedAugmenter(pol, TailSpec.imagenet(0, torch.float16), H, W, 1)
outs = [f.empty_out(B) for _ in range(2)]
for i in range(n): f(x[i % 2], outs[i % 2], i * B)
torch.cuda.synchronize()
