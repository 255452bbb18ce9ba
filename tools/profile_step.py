"""Developer tool (GPU): where one pipeline step's time goes -- per-kernel device time (torch.profiler / CUPTI), GPU busy
vs. step wall time, and the host phases of ZbufferModelPts.forward_image.  python tools/profile_step.py [batch]"""
import os
import sys
import time

import torch
from torch.profiler import ProfilerActivity, profile, record_function

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from bench import make_batch, make_opt  # noqa: E402
from pixelsynth_b200 import lmconv  # noqa: E402
from pixelsynth_b200.models.z_buffermodel import ZbufferModelPts  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
dev = torch.device("cuda:0")
model = ZbufferModelPts(make_opt(), device=dev)
batch = make_batch(B, [i % 8 for i in range(B)])
batch = {"images": [t.to(dev) for t in batch["images"]], "cameras": [{k: v.to(dev) for k, v in c.items()} for c in batch["cameras"]]}
g = torch.Generator().manual_seed(1)
noise = torch.randn(16, B, 20, generator=g).to(dev)
uniforms = torch.rand(B, 1024, generator=g).to(dev)

# host phases
_glue, _prep = lmconv.glue_host, model.outpaint2.prepare
host = {"glue": 0.0, "prepare": 0.0}


def glue(*a, **k):
    t = time.perf_counter()
    with record_function("host:glue"):
        r = _glue(*a, **k)
    host["glue"] += time.perf_counter() - t
    return r


def prep(*a, **k):
    t = time.perf_counter()
    with record_function("host:prepare"):
        r = _prep(*a, **k)
    host["prepare"] += time.perf_counter() - t
    return r


lmconv.glue_host = glue
model.outpaint2.prepare = prep


def step():
    model.forward(batch, noise=noise, uniforms=uniforms)


for _ in range(3):
    step()
torch.cuda.synchronize()
host["glue"] = host["prepare"] = 0.0
N = 3
t0 = time.perf_counter()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(N):
        step()
    torch.cuda.synchronize()
wall = (time.perf_counter() - t0) / N * 1e3
ev = prof.key_averages()
rows = [(e.key, e.device_time_total / N / 1e3, e.count // N) for e in ev if e.device_time_total > 0 and not e.key.startswith(("host:", "aten::", "cudaMemcpy", "Memcpy", "Memset"))]
mem = [(e.key, e.device_time_total / N / 1e3, e.count // N) for e in ev if e.device_time_total > 0 and e.key.startswith(("Memcpy", "Memset"))]
rows.sort(key=lambda r: -r[1])
busy = sum(r[1] for r in rows) + sum(r[1] for r in mem)
print("batch %d: step wall %.2f ms (under the profiler), GPU kernel+copy time %.2f ms, host glue %.2f ms, host prepare %.2f ms" %
      (B, wall, busy, host["glue"] / N * 1e3, host["prepare"] / N * 1e3))
for k, ms, n in rows[:24] + mem:
    print("  %8.3f ms  %4d x  %s" % (ms, n, k[:110]))
