#!/usr/bin/env python
"""Developer tool (GPU): the decoder's big resampling launches alone (ps_resample, two outputs with affine + ReLU).
python tools/bench_resample.py [reps]      PS_UPSAMPLE_GENERIC=1 routes bilinear through the one-thread-per-output kernel"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pixelsynth_b200 import nets  # noqa: E402
from pixelsynth_b200.conv import Out  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
N = 64
for mode, h, c in (("bilinear", 128, 128), ("bilinear", 64, 256), ("avgpool", 256, 128), ("avgpool", 128, 256)):
    x = torch.randn(N, h, h, c, device="cuda").to(torch.bfloat16)
    ho = 2 * h if mode == "bilinear" else h // 2
    o0 = torch.empty(N, ho, ho, c, dtype=torch.bfloat16, device="cuda")
    o1 = torch.empty_like(o0)
    sc, sh = torch.rand(N, c, device="cuda") + 0.5, torch.randn(N, c, device="cuda")
    run = lambda: nets.resample(x, mode, Out(o0), Out(o1, "relu", sc, sh, per_sample=True))
    for _ in range(2):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    gb = (x.numel() + 2 * o0.numel()) * 2 / 1e9
    print("%-9s %4d^2 x %3d ch -> %4d^2: %.3f ms  %.0f GB/s (%.2f GB algorithmic)" % (mode, h, c, ho, ms, gb / ms * 1e3, gb))
