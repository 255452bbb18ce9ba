"""Developer tool (GPU): compares every cached activation tensor of the lmconv kernel with the fp32 oracle's
intermediates, in execution order, to localise a divergence."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from oracle import lmconv_ref, weights  # noqa: E402
import pixelsynth_b200.lmconv as lm  # noqa: E402
import make_lmconv_golden as mk  # noqa: E402

sd = weights.make_state("lmconv", 0)
model = lm.LmconvB200(sd)
bgs = mk.background_cases()
B = 1
_, order, words, smask = lm.glue_host(bgs[:B])
g = torch.Generator().manual_seed(0)
codes = torch.randint(0, 512, (B, 32, 32), generator=g)
out = model.logits(codes, order, words).cpu()
torch.cuda.synchronize()
data = torch.nn.functional.one_hot(codes, 512).permute(0, 3, 1, 2).float()
mf = [torch.cat([lmconv_ref.masks_to_float(words[b, k]) for b in range(B)]) for k in range(3)]
trace = []
with torch.no_grad():
    ref = lmconv_ref.lmconv_logits(sd, data, *mf, trace=trace)
cache = model._cache.view(torch.float16).view(B, 33, 1024, 240).float().cpu()
ids = [(0, "u_init")]
for i in range(18):
    o = model.plan.ops[i]
    if o.kind == 0:
        ids += [(o.mid, "op%d mid" % i), (o.out, "op%d out" % i)]
    else:
        ids += [(o.out, "op%d dil" % i)]
rank = np.empty(1024, int)
rank[order[0]] = np.arange(1024)
for (tid, name), t in zip(ids, trace):
    r = t[0].permute(1, 2, 0).reshape(1024, 80)
    c = cache[0, tid]
    p, n = c[:, :80], c[:, 80:160]
    x = torch.where(p > 0, p, -n)
    err = (x - r).abs()
    worst = int(err.max(1).values.argmax())
    print("%-10s tensor %2d  max err %.4f  rms %.5f  (ref std %.3f)  worst cell %d rank %d" %
          (name, tid, err.max().item(), err.pow(2).mean().sqrt().item(), r.std().item(), worst, rank[worst]))
err = (out - ref).abs()
print("logits max err %.4f rms %.5f std %.3f" % (err.max().item(), err.pow(2).mean().sqrt().item(), ref.std().item()))
