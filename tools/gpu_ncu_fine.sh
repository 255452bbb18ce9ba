#!/bin/bash
# ncu --set full capture of the splat tile kernel (maps emitted, 16 views) with source lines; PS_SPLAT_VARIANT from the caller.
mkdir -p gpurun_out
timeout 200 ncu --set full --clock-control none --import-source on -k regex:^fine_kernel -s 3 -c 1 -f \
    -o gpurun_out/prof_fine python tools/bench_splat.py > gpurun_out/ncu_fine.log 2>&1; tail -3 gpurun_out/ncu_fine.log
ls -la gpurun_out/*.ncu-rep
