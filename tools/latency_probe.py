#!/usr/bin/env python
"""Per-image latency by program class: 64 images (all clusters resident at once) of one 2-op combo."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from fast_autoaugment_b200.engine import CompiledPolicy, FusedAugmenter, TailSpec
H, W, B = 224, 224, int(sys.argv[1]) if len(sys.argv) > 1 else 64
x = torch.from_numpy(bench.synth_batch(B, H, W, 1)).cuda()
tail = TailSpec.imagenet(0, torch.float16)
combos = [("none", None), ("Invert", None), ("Color", None), ("Equalize", None), ("Rotate", None), ("Sharpness", None),
          ("Rotate", "Invert"), ("Invert", "Rotate"), ("Rotate", "ShearY"), ("Sharpness", "Rotate"), ("Rotate", "Sharpness"),
          ("Sharpness", "Sharpness"), ("Rotate", "Equalize"), ("Equalize", "AutoContrast"), ("Color", "Contrast"),
          ("Sharpness", "Equalize"), ("Equalize", "Sharpness"), ("Color", "Sharpness"), ("Equalize", "Rotate")]
for a, b in combos:
    pol = [[(a if a != "none" else "Invert", 1.0 if a != "none" else 0.0, 0.7), (b or "Invert", 1.0 if b else 0.0, 0.7)]]
    f = FusedAugmenter(CompiledPolicy(pol), tail, H, W, 1)
    out = f.empty_out(B)
    for i in range(3): f(x, out, 0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(50): f(x, out, 0)
    e1.record(); torch.cuda.synchronize()
    print("%-12s %-12s %8.1f us" % (a, b or "-", e0.elapsed_time(e1) * 1e3 / 50), flush=True)
