#!/bin/bash
mkdir -p gpurun_out
SECT="--section SpeedOfLight --section ComputeWorkloadAnalysis --section MemoryWorkloadAnalysis --section MemoryWorkloadAnalysis_Tables --section LaunchStats --section Occupancy --section SchedulerStats --section WarpStateStats --section InstructionStats"
timeout 600 ncu $SECT --clock-control none -k regex:lmconv_tc -s 1 -c 1 -f -o gpurun_out/prof_lmconv_final \
    python tools/bench_lmconv.py --reps 1 > gpurun_out/ncu_lmconv.log 2>&1; tail -3 gpurun_out/ncu_lmconv.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_igemm -s 3 -c 1 -f -o gpurun_out/prof_conv_final \
    python tools/bench_conv.py > gpurun_out/ncu_conv.log 2>&1; tail -2 gpurun_out/ncu_conv.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_igemm -s 3 -c 1 -f -o gpurun_out/prof_conv_heavy_final \
    python tools/bench_conv.py --heavy > gpurun_out/ncu_conv2.log 2>&1; tail -2 gpurun_out/ncu_conv2.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ps::fine_kernel -s 3 -c 1 -f -o gpurun_out/prof_fine_final \
    python tools/bench_splat.py > gpurun_out/ncu_fine.log 2>&1; tail -2 gpurun_out/ncu_fine.log
timeout 200 python tools/bench_splat.py
