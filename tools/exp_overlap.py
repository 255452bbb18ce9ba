#!/usr/bin/env python
"""Experiment: two steps in flight (two model instances, two streams), optionally with the sampler confined to a CUDA
green context of S SMs and everything else on the remaining SMs.   python tools/exp_overlap.py MODE [S]
MODE: serial | two | green (sampler on S-SM partition, rest on the primary context) | split (exclusive partitions)"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from pixelsynth_b200 import _lib  # noqa: E402
from pixelsynth_b200.models.z_buffermodel import ZbufferModelPts  # noqa: E402

mode = sys.argv[1]
S = int(sys.argv[2]) if len(sys.argv) > 2 else 48
B, STEPS = 64, 12
dev = torch.device("cuda:0")
torch.cuda.init()
torch.zeros(1, device=dev)


def green_streams(s_small):
    """-> (stream on a partition of >= s_small SMs, stream on the remaining SMs), as torch external streams"""
    from cuda.bindings import driver as cu

    def ok(r):
        assert r[0] == cu.CUresult.CUDA_SUCCESS, r[0]
        return r[1:] if len(r) > 2 else r[1]
    d = ok(cu.cuDeviceGet(0))
    res = ok(cu.cuDeviceGetDevResource(d, cu.CUdevResourceType.CU_DEV_RESOURCE_TYPE_SM))
    groups, n, rem = ok(cu.cuDevSmResourceSplitByCount(1, res, 0, s_small))
    print("partition:", groups[0].sm.smCount, "+", rem.sm.smCount, "SMs", flush=True)
    out = []
    for r in (groups[0], rem):
        desc = ok(cu.cuDevResourceGenerateDesc([r], 1))
        g = ok(cu.cuGreenCtxCreate(desc, d, cu.CUgreenCtxCreate_flags.CU_GREEN_CTX_DEFAULT_STREAM))
        st = ok(cu.cuGreenCtxStreamCreate(g, cu.CUstream_flags.CU_STREAM_NON_BLOCKING, 0))
        out.append(torch.cuda.ExternalStream(int(st)))
    return out


models = [ZbufferModelPts(bench.make_opt(), device=dev) for _ in range(1 if mode == "serial" else 2)]
hb = bench.make_batch(B, [i % 8 for i in range(B)])
db = {"images": [hb["images"][0].to(dev)] * 2, "cameras": [{k: v.to(dev) for k, v in c.items()} for c in hb["cameras"]]}
g = torch.Generator().manual_seed(1)
noise = torch.randn(16, B, 20, generator=g).to(dev)
uni = torch.rand(B, 1024, generator=g)
if mode == "serial":
    main = [torch.cuda.current_stream()]
elif mode == "two":
    main = [torch.cuda.Stream(), torch.cuda.Stream()]
elif mode == "green":
    small, _ = green_streams(S)
    main = [torch.cuda.Stream(), torch.cuda.Stream()]
    for m in models:
        m.sampler_stream = small
elif mode == "split":
    small, big = green_streams(S)
    main = [big, big]      # one conv stream: steps serialise on it, only the sampler runs beside them
    for m in models:
        m.sampler_stream = small
elif mode == "split2":
    small, big = green_streams(S)
    from cuda.bindings import driver as cu
    r, gctx = cu.cuStreamGetGreenCtx(int(big.cuda_stream))
    r, st2 = cu.cuGreenCtxStreamCreate(gctx, cu.CUstream_flags.CU_STREAM_NON_BLOCKING, 0)
    main = [big, torch.cuda.ExternalStream(int(st2))]
    for m in models:
        m.sampler_stream = small
outs = [None, None]


def run(k):
    i = k % len(models)
    with torch.cuda.stream(main[i % len(main)]):
        outs[i] = models[i].forward(db, noise=noise, uniforms=uni)[1]["PredImg"]


for k in range(4):
    run(k)
torch.cuda.synchronize()
t0 = time.perf_counter()
for k in range(STEPS):
    run(k)
torch.cuda.synchronize()
ms = (time.perf_counter() - t0) * 1e3 / STEPS
_lib.check_wedge("exp_overlap")
print("%s S=%d: %.2f ms/step  %.0f views/s  checksum %.4f %.4f" % (mode, S, ms, B / ms * 1e3, float(outs[0].float().mean()),
                                                                float(outs[-1 if len(models) > 1 else 0].float().std())), flush=True)
