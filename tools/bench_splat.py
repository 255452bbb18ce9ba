"""Developer tool (GPU): the splat in its map-emitting mode (idx + z maps, 69.0 MB/view), 16 views per launch, on a
smooth synthetic depth; prints fine_kernel's time and HBM fraction.  Used for the ncu capture of the splat."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from util import synthetic_view  # noqa: E402
from pixelsynth_b200 import _lib  # noqa: E402
import pixelsynth_b200.ops  # noqa: F401,E402

B, W, K = 16, 256, 128
depth, feat, mats = synthetic_view(B, W, kind="translate", seed=0, depth_mode="smooth")
d, f, m = [torch.from_numpy(a).cuda() for a in (depth, feat, mats)]
run = lambda: torch.ops.pixelsynth_b200.splat(d, f, m, W, W, K, 4.0, 1.0, 2, 0, 13, 1e-2, True, False)
for _ in range(3):
    run()
torch.cuda.synchronize()
L = _lib.lib()
L.ps_timing_enable(1)
for _ in range(5):
    run()
torch.cuda.synchronize()
L.ps_timing_enable(0)
ms = _lib.kernel_time_ms("fine_kernel")[0] / 5
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
gbs = 69009408 * B / (ms * 1e-3) / 1e9
print(json.dumps({"fine_kernel_ms": ms, "views_per_launch": B, "GB/s": gbs, "frac_of_hbm_peak": gbs / peak}))
