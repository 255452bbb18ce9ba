#!/bin/bash
# Full GPU suite + bench line + step profile (one visit).
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/pytest_gpu.log)"
grep -E "^(FAILED|ERROR)" gpurun_out/pytest_gpu.log | head -20
bash tools/gpu_bench_only.sh "$@"
timeout 200 python tools/profile_step.py 128 2>/dev/null | head -12 | cut -c1-150 | tee gpurun_out/profile_step_b128.txt
