#!/bin/bash
mkdir -p gpurun_out
for d in 0 64 16 32 96; do echo "PS_TC_DEBUG=$d"; PS_TC_DEBUG=$d timeout 300 python tools/bench_lmconv.py --reps 2 2>&1 | tail -1 | cut -c100-200; done
