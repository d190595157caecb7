#!/usr/bin/env python
"""Device time of two-op sub-policies (every image applies both ops), 224x224 b512."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, bench
from fast_autoaugment_b200.engine import CompiledPolicy, FusedAugmenter, TailSpec
H = W = 224; B = 512
x = [torch.from_numpy(bench.synth_batch(B, H, W, 1 + i)).cuda() for i in range(4)]
tail = TailSpec.imagenet(0, torch.float16)
pairs = [("Brightness", "Sharpness"), ("Contrast", "Sharpness"), ("Rotate", "Sharpness"), ("TranslateY", "Sharpness"),
         ("Color", "Sharpness"), ("ShearY", "Equalize"), ("Rotate", "AutoContrast"), ("Sharpness", "Rotate"),
         ("Sharpness", "Posterize"), ("Sharpness", "Sharpness"), ("Color", "Contrast"), ("Rotate", "ShearY"),
         ("AutoContrast", "Rotate"), ("Contrast", "Cutout"), ("Equalize", "AutoContrast"), ("Posterize", "Equalize")]
for a, b in pairs:
    pol = CompiledPolicy([[(a, 1.0, 0.7), (b, 1.0, 0.6)]])
    f = FusedAugmenter(pol, tail, H, W, 1, overlap_calls=True)
    outs = [f.empty_out(B) for _ in range(4)]
    for i in range(5): f(x[i % 4], outs[i % 4], i * B)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(60): f(x[i % 4], outs[i % 4], i * B)
    e1.record(); torch.cuda.synchronize()
    print("%-13s %-13s %7.1f us" % (a, b, e0.elapsed_time(e1) * 1e3 / 60), flush=True)
