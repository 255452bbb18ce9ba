#!/bin/bash
# lmconv sampler probes: CTA timeline + sensitivity to skipping the weight / gather copies (timing only).
mkdir -p gpurun_out
timeout 300 python tools/trace_lmconv.py > gpurun_out/trace_lmconv.txt 2>&1
for d in 0 1 2 3; do
  echo "PS_TC_DEBUG=$d" >> gpurun_out/lmconv_debug.txt
  PS_TC_DEBUG=$d timeout 300 python tools/bench_lmconv.py --reps 2 >> gpurun_out/lmconv_debug.txt 2>&1
done
cat gpurun_out/lmconv_debug.txt; head -50 gpurun_out/trace_lmconv.txt
