#!/usr/bin/env python
"""ncu target: every program family once (one warm-up call, one call inside the NVTX range `prof`).
Usage (GPU box): ncu --nvtx --nvtx-include "prof/" ... python tools/ncu_classes.py [names...]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, bench
from fast_autoaugment_b200 import archive
from fast_autoaugment_b200.engine import CompiledPolicy, FusedAugmenter, TailSpec
H = W = 224; B = 512
x = [torch.from_numpy(bench.synth_batch(B, H, W, 1 + i)).cuda() for i in range(2)]
tail = TailSpec.imagenet(0, torch.float16)
CASES = [("identity", [[("Invert", 0.0, 0.7), ("Invert", 0.0, 0.6)]])]
for nm in ("Invert", "Brightness", "Color", "Cutout", "AutoContrast", "Equalize", "Contrast", "TranslateX", "TranslateY",
           "ShearX", "ShearY", "Rotate", "Sharpness"):
    CASES.append((nm, [[(nm, 1.0, 0.7), (nm, 0.0, 0.6)]]))
for a, b in (("Rotate", "Sharpness"), ("ShearY", "Equalize"), ("Sharpness", "Rotate"), ("Color", "Contrast"),
             ("Rotate", "ShearY"), ("Equalize", "AutoContrast"), ("Brightness", "Sharpness")):
    CASES.append((a + "+" + b, [[(a, 1.0, 0.7), (b, 1.0, 0.6)]]))
CASES.append(("mix", archive.fa_resnet50_rimagenet()))
want = set(sys.argv[1:])
for name, pol_list in CASES:
    if want and name not in want:
        continue
    pol = CompiledPolicy(pol_list)
    f = FusedAugmenter(pol, tail, H, W, 1, overlap_calls=True)
    outs = [f.empty_out(B) for _ in range(2)]
    f(x[0], outs[0], 0); f(x[1], outs[1], B)
    torch.cuda.synchronize()
    torch.cuda.nvtx.range_push("prof")
    f(x[0], outs[0], 2 * B)
    torch.cuda.synchronize()
    torch.cuda.nvtx.range_pop()
    print("CASE", name, flush=True)
