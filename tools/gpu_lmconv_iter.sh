#!/bin/bash
# lmconv iteration loop: GPU parity tests of the sampler, config-3 benchmark, CTA timeline.
mkdir -p gpurun_out
export PS_CHECK_WEDGE=1
timeout 600 python -m pytest tests/test_lmconv_gpu.py -x -q -m gpu > gpurun_out/pytest_lmconv.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest_lmconv.log
unset PS_CHECK_WEDGE
timeout 300 python tools/bench_lmconv.py --reps 3 > gpurun_out/bench_lmconv.json 2>&1; cat gpurun_out/bench_lmconv.json
timeout 300 python tools/trace_lmconv.py > gpurun_out/trace_lmconv.txt 2>&1; head -42 gpurun_out/trace_lmconv.txt
for d in 4 8 12; do echo "PS_TC_DEBUG=$d (timing only)"; PS_TC_DEBUG=$d timeout 300 python tools/bench_lmconv.py --reps 2 2>&1 | tail -1 | cut -c1-220; done
