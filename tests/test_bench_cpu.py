"""CPU: bench.py's contract pieces that need no GPU -- the product arm refuses to run without a CUDA device (there is
no CPU fallback to measure; it fails once, loudly: no retries, no batch downgrade), and the flag surface the driver
uses parses."""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.is_available(), reason="a GPU is present")
def test_product_arm_fails_loudly_without_a_gpu():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--gpus", "1", "--steps", "1", "--warmup", "3"],
                       capture_output=True, text=True, timeout=300)
    assert p.returncode != 0 and p.stdout.strip() == ""
    assert p.stderr.count("no CUDA device; the hot path has no CPU fallback") == 1      # one attempt: a failure is a failure
    assert "attempt" not in p.stderr


def test_flags_and_workload_description():
    sys.path.insert(0, ROOT)
    import bench

    argv, sys.argv = sys.argv, ["bench.py", "--gpus", "8", "--steps", "7", "--warmup", "4", "--impl", "reference"]
    try:
        a = bench.parse()
    finally:
        sys.argv = argv
    assert (a.gpus, a.steps, a.warmup, a.impl, a.batch) == (8, 7, 4, "reference", 64)
    cfg = bench.workload_config(a, 8)
    assert "configs[1]" in cfg["workload"] and cfg["views_per_step_per_gpu"] == 64 and cfg["global_views_per_step"] == 512
    assert "(i + r) mod 8" in cfg["view_mix"] and a.view == -1 and a.workload == "views"
    assert (a.in_flight, a.sampler_sms) == (3, 24) and cfg["in_flight"] == 3 and "24-SM green-context partition" in cfg["schedule"]
    assert bench.BYTES_PER_VIEW_SPLAT_MAPS == 69009408 and bench.BYTES_PER_VIEW_SPLAT_FUSED == 1900544
