#!/usr/bin/env python
"""Is a CIFAR training step still augmentation-bound?  (BASELINE.json north-star check)

Times (a) the fused augmentation of one CIFAR batch (uint8 -> fp32 NCHW, full train chain, fused
Philox) and (b) one fp32 SGD training step of a WideResNet-40-2 (the reference's
confs/wresnet40x2_cifar.yaml model: depth 40, widen 2, batch 128 per GPU) on the same GPU.
The network below is a plain restatement of the standard WRN architecture for timing only."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.nn as nn
import torch.nn.functional as F

import bench
from fast_autoaugment_b200 import archive
from fast_autoaugment_b200.engine import CompiledPolicy, FusedAugmenter, TailSpec


class Block(nn.Module):
    def __init__(self, i, o, stride):
        super().__init__()
        self.bn1, self.conv1 = nn.BatchNorm2d(i), nn.Conv2d(i, o, 3, stride, 1, bias=False)
        self.bn2, self.conv2 = nn.BatchNorm2d(o), nn.Conv2d(o, o, 3, 1, 1, bias=False)
        self.short = None if (i == o and stride == 1) else nn.Conv2d(i, o, 1, stride, 0, bias=False)

    def forward(self, x):
        y = F.relu(self.bn1(x))
        s = x if self.short is None else self.short(y)
        y = self.conv1(y)
        y = self.conv2(F.relu(self.bn2(y)))
        return y + s


class WRN(nn.Module):
    def __init__(self, depth=40, widen=2, classes=10):
        super().__init__()
        n = (depth - 4) // 6
        w = [16, 16 * widen, 32 * widen, 64 * widen]
        layers = [nn.Conv2d(3, w[0], 3, 1, 1, bias=False)]
        for g in range(3):
            for k in range(n):
                layers.append(Block(w[g] if k == 0 else w[g + 1], w[g + 1], (1 if g == 0 else 2) if k == 0 else 1))
        self.body = nn.Sequential(*layers)
        self.bn, self.fc = nn.BatchNorm2d(w[3]), nn.Linear(w[3], classes)

    def forward(self, x):
        x = F.relu(self.bn(self.body(x)))
        return self.fc(F.adaptive_avg_pool2d(x, 1).flatten(1))


def timed(fn, n):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


B = 128
x_u8 = torch.from_numpy(bench.synth_batch(B, 32, 32, 7)).cuda()
y = torch.randint(0, 10, (B,), device="cuda")
pol = CompiledPolicy(archive.fa_reduced_cifar10())
tail = TailSpec.cifar(16, torch.float32)
aug = FusedAugmenter(pol, tail, 32, 32, 1)
out = aug.empty_out(B)
step = [0]


def do_aug():
    aug(x_u8, out, step[0] * B)
    step[0] += 1


model = WRN().cuda()
opt = torch.optim.SGD(model.parameters(), lr=0.1, momentum=0.9, nesterov=True, weight_decay=2e-4)


def do_train():
    opt.zero_grad(set_to_none=True)
    loss = F.cross_entropy(model(out), y)
    loss.backward()
    opt.step()


t_aug = timed(do_aug, 500)
t_train = timed(do_train, 30)
print("CIFAR b%d: fused augmentation %.1f us/batch ; WRN-40-2 fp32 train step %.2f ms ; augmentation = %.2f%% of the step"
      % (B, t_aug * 1e3, t_train, 100 * t_aug / t_train))
