#!/bin/bash
# Final visit: the bench line (kept as the round's record) and a second run for stability.
mkdir -p gpurun_out
bash tools/gpu_bench_only.sh
cp gpurun_out/bench.json gpurun_out/bench_final.json
timeout 100 python bench.py --no-cpu-baseline --steps 3 > gpurun_out/bench_again.json 2> gpurun_out/bench_again.err; echo "again rc=$? $(cut -c1-90 gpurun_out/bench_again.json)"
