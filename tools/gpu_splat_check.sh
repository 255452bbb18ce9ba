#!/bin/bash
# Splat parity suite + map-mode timing (+ optional ncu capture with "ncu" as first argument).
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_splat_gpu.py -q -m gpu > gpurun_out/pytest_splat.log 2>&1
echo "pytest rc=$? $(tail -1 gpurun_out/pytest_splat.log)"; grep -E "^(FAILED|ERROR)" gpurun_out/pytest_splat.log | head -20
timeout 100 python tools/bench_splat.py 2>&1 | tail -1
if [ "${1:-}" = "ncu" ]; then bash tools/gpu_ncu_fine.sh; fi
