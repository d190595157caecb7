#!/usr/bin/env python
"""ncu target for the Mixup kernels (config 4, one GPU): mixup_global = augment to uint8 + faa_mix_u8; aug_mixup.mixup."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from fast_autoaugment_b200 import archive, aug_mixup
from fast_autoaugment_b200.distributed import mixup_global
from fast_autoaugment_b200.engine import CompiledPolicy, TailSpec
B = int(os.environ.get("MIX_B", "2048"))
pol = CompiledPolicy(archive.fa_resnet50_rimagenet())
tail = TailSpec.imagenet(0, torch.float16)
x = torch.from_numpy(bench.synth_batch(512, 224, 224, 5)).cuda().repeat(B // 512, 1, 1, 1).contiguous()
t = torch.arange(B, device="cuda") % 1000
for step in range(3):
    data, ta, tb, lam = mixup_global(pol, x, t, tail, 0.2, 11, step)
for step in range(3):
    torch.manual_seed(step)
    d2, ta, tb, lam = aug_mixup.mixup(data, t, 0.2)
torch.cuda.synchronize()
print("ok", float(data.float().mean()), float(d2.float().mean()))
