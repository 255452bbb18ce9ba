"""Developer tool (GPU): one decoder-sized convolution (batch 32, 256x256, 128 -> 128 channels, 3x3, per-sample
scale/shift + ReLU epilogue; optionally residual + second output) timed in isolation.  PS_CONV_DEBUG switches parts of
the kernel off (4 MMA issue, 8 TMA loads, 16 epilogue body) to see what bounds it."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pixelsynth_b200.conv import Out, PackedConv, conv_igemm  # noqa: E402

heavy = "--heavy" in sys.argv
thin = "--thin" in sys.argv      # the decoder's 128 -> 3 layer at 64 views (Cout padded to 16: grouped weight stages)
N, S, C = (64 if thin else 32), 256, 128
CO = 3 if thin else C
g = torch.Generator().manual_seed(0)
x = torch.randn(N, S, S, C, generator=g).to(device="cuda", dtype=torch.bfloat16)
w = torch.randn(CO, C, 3, 3, generator=g) * 0.03
pc = PackedConv.conv2d(w, torch.randn(CO, generator=g), padding=1)
sc = torch.rand(N, CO, device="cuda") + 0.5
sh = torch.randn(N, CO, device="cuda") * 0.1
y0 = torch.zeros(N, S, S, 8 if thin else C, device="cuda", dtype=torch.bfloat16)
y1 = torch.empty_like(y0)
res = torch.randn(N, S, S, C, generator=g).to(device="cuda", dtype=torch.bfloat16) if heavy else None
outs = [Out(y0), Out(y1, "relu", sc, sh, per_sample=True)] if heavy else [Out(y0, "relu", sc, sh, per_sample=True)]
for _ in range(3):
    conv_igemm(x, pc, outs, residual=res)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
R = 10
for _ in range(R):
    conv_igemm(x, pc, outs, residual=res)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / R
fl = 2.0 * N * S * S * C * CO * 9
print("debug=%s heavy=%d  %.3f ms  %.1f TFLOP/s" % (os.environ.get("PS_CONV_DEBUG", "0"), heavy, ms, fl / ms / 1e9))
