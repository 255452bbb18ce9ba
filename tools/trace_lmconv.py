"""Developer tool (GPU): timeline of one lmconv_tc_kernel CTA (one mid-size level of BASELINE configs[2]) from the
clock64 trace hook: where each GEMM's time goes (gather / MMA / epilogue)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pixelsynth_b200 import _lib, lmconv, synthetic  # noqa: E402

B = 32
model = lmconv.LmconvB200(synthetic.make_state("lmconv", 0))
bg = torch.zeros(B, 256, 256, dtype=torch.bool)
bg[:, :, 128:] = True
_, order, words, smask = lmconv.glue_host(bg)
codes = torch.randint(0, 512, (B, 32, 32), generator=torch.Generator().manual_seed(0))
uniforms = torch.rand(B, 1024, generator=torch.Generator().manual_seed(1))
model.sample(codes, order, words, smask, uniforms, 0.7)
torch.cuda.synchronize()
trace = torch.zeros(8 * 1024, dtype=torch.int64, device="cuda")
_lib.lib().ps_lmconv_tc_set_trace(trace.data_ptr())
model.sample(codes, order, words, smask, uniforms, 0.7)   # the LAST level's CTA 0 is what remains in the buffer
torch.cuda.synchronize()
_lib.lib().ps_lmconv_tc_set_trace(None)
t = trace.cpu().numpy().reshape(8, 1024)
t0 = t[7, 0]
n = model.plan.n_chunks_total if t[0, model.plan.n_chunks_body] else model.plan.n_chunks_body
us = lambda c: (c - t0) / 1965.0
print("kernel %.1f us, %d chunks" % (us(t[7, 1]), n))
ef = list(model.plan.epi_first)
first = 0
print("gemm: chunks | mma_first_ready mma_last_ready | epi: acc_full acquired published | gather: first_issue last_issue")
for g in range(32):
    last = ef[g] + (3 if (ef[g] + 3 <= n and (g == 31 or ef[g] + 3 <= (ef[g + 1] if g < 31 else n))) else 2) - 1
    nxt = last + 1
    gi = [i for i in range(first, nxt) if t[2, i]]
    print("g%2d %3d-%3d | %7.1f %7.1f | %7.1f %7.1f %7.1f | %7.1f %7.1f" %
          (g, first, last, us(t[0, first]), us(t[0, last]), us(t[4, g]), us(t[6, g]) if t[6, g] else -1, us(t[5, g]),
           us(t[2, gi[0]]) if gi else -1, us(t[2, gi[-1]]) if gi else -1))
    first = nxt
d = np.diff(t[0, :n])
print("mma chunk-to-chunk gaps: median %.2f us, mean %.2f us" % (np.median(d) / 1965, d.mean() / 1965))
gw = [(t[2, i] - t[1, i]) / 1965 for i in range(n) if t[2, i]]
print("gather issue time per chunk (thread 0): median %.2f us" % np.median(gw))
print("chunk: producer_issue gather_wait_done gather_issued mma_ready  (us)")
for i in list(range(46, 72)):
    print("%4d %8.2f %8.2f %8.2f %8.2f" % (i, us(t[3, i]), us(t[1, i]) if t[1, i] else -1, us(t[2, i]) if t[2, i] else -1, us(t[0, i])))
