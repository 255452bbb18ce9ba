#!/usr/bin/env python
"""Candidate-list statistics of the splat on the bench workload: per circle view, how many 8x8 tiles exceed the
512-candidate fast path (fine_big_kernel's work).  Reads the workspace the op allocated (counts at offset 0)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from pixelsynth_b200 import ops  # noqa: E402

B = 64

captured = []
orig = ops._workspace


def grab(nbytes, device):
    t = orig(nbytes, device)
    captured.append(t)
    return t


ops._workspace = grab
from pixelsynth_b200.models.z_buffermodel import ZbufferModelPts  # noqa: E402

dev = torch.device("cuda:0")
model = ZbufferModelPts(bench.make_opt(), device=dev)
views = [i % 8 for i in range(B)]
hb = bench.make_batch(B, views)
db = {"images": [hb["images"][0].to(dev)] * 2, "cameras": [{k: v.to(dev) for k, v in c.items()} for c in hb["cameras"]]}
g = torch.Generator().manual_seed(1)
model.forward(db, noise=torch.randn(16, B, 20, generator=g).to(dev), uniforms=torch.rand(B, 1024, generator=g))
torch.cuda.synchronize()
ws = max(captured, key=lambda t: t.numel())
cnt = ws[:B * 1024 * 4].view(torch.int32).cpu().numpy().reshape(B, 1024)
novf = int(ws[B * 1024 * 4:B * 1024 * 4 + 4].view(torch.int32).item())
print("overflow tiles queued:", novf, "of", B * 1024)
d = model.last["depth"].float()
print("depth: min %.3f max %.3f mean %.3f" % (d.min().item(), d.max().item(), d.mean().item()))
for v in range(8):
    c = cnt[v]
    print("view %d: mean %.0f  p50 %d  p90 %d  max %d  >512: %d  >1024: %d" % (
        v, c.mean(), np.percentile(c, 50), np.percentile(c, 90), c.max(), (c > 512).sum(), (c > 1024).sum()))
