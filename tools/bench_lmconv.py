"""BASELINE configs[2]: batch 32, 32x32 code grid with the right half masked (512 sampled cells per image), T = 0.7,
injected uniforms.  Prints sampler tokens/s (CUDA events around LmconvB200.sample's device work, host glue excluded
and reported) and the wavefront statistics.  Usage: python tools/bench_lmconv.py [--batch 32] [--reps 3]"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pixelsynth_b200 import _lib, lmconv, synthetic  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--reps", type=int, default=3)
args = ap.parse_args()
B = args.batch
model = lmconv.LmconvB200(synthetic.make_state("lmconv", 0))
bg = torch.zeros(B, 256, 256, dtype=torch.bool)
bg[:, :, 128:] = True
t0 = time.perf_counter()
_, order, words, smask = lmconv.glue_host(bg)
glue_ms = 1e3 * (time.perf_counter() - t0)
g = torch.Generator().manual_seed(0)
codes = torch.randint(0, 512, (B, 32, 32), generator=g)
uniforms = torch.rand(B, 1024, generator=torch.Generator().manual_seed(1))
L = _lib.lib()
best = None
for rep in range(args.reps + 1):
    _lib.kernel_time_ms(None)
    L.ps_timing_enable(1)
    try:
        out = model.sample(codes, order, words, smask, uniforms, 0.7)
        torch.cuda.synchronize()
    except Exception as e:  # noqa: BLE001
        import ctypes
        info = (ctypes.c_uint * 8)()
        print("FAULT rep %d: %s; wedge=%d info=%s" % (rep, str(e).splitlines()[0], L.ps_wedge_poll(info), [hex(x) for x in info]), flush=True)
        NW = 8 + 256 * 8 + 64 * 4
        log = (ctypes.c_uint * NW)()
        L.ps_wedge_log(log, NW)
        for k in range(256):
            w = log[8 + 8 * k: 16 + 8 * k]
            m = log[8 + 2048 + 4 * k: 12 + 2048 + 4 * k] if k < 64 else [0, 0, 0, 0]
            if m[3] >> 31:
                print("  mbar waiter: block %d thread %d (warp %d) barrier smem 0x%x (TcSmem+%d) parity %d; tile %d" %
                      (m[0], m[1], m[1] // 32, m[2], m[2] - 0x36400, m[3] & 1, (m[3] >> 4) & 0x7ffffff), flush=True)
            if w[7] >> 16 == 0xabcd:
                print("  waiter: block %d thread %d waits tiles [%d,+%d) >= %d (saw %d); own tile %d kind %d site %d" %
                      (w[0], w[1], w[2], w[3], w[4], w[5], w[6] & 0xffffff, (w[6] >> 24) & 15, w[6] >> 28), flush=True)
        import os
        os._exit(3)
    if L.ps_wedge_poll(None):
        import ctypes
        info = (ctypes.c_uint * 8)()
        L.ps_wedge_poll(info)
        print("WEDGED rep %d: info=%s" % (rep, [hex(x) for x in info]), flush=True)
    L.ps_timing_enable(0)
    ms, n = _lib.kernel_time_ms("lmconv_tc_kernel")
    if rep > 0:
        best = ms if best is None else min(best, ms)
offs = model.last_levels
tokens = int(smask.sum())
print(json.dumps({"workload": "lmconv outpaint, batch %d, 512 masked cells/image" % B, "tokens": tokens,
                  "rows_processed": int(offs[-1]), "levels": len(offs) - 1, "sampler_ms": best,
                  "tokens_per_s": tokens / (best * 1e-3), "rows_per_s": int(offs[-1]) / (best * 1e-3),
                  "tensor_tflops": 11.163e6 * int(offs[-1]) / (best * 1e-3) / 1e12, "glue_host_ms": glue_ms}))
