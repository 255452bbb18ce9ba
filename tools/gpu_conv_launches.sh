#!/bin/bash
# per-launch device time / tensor-pipe / L2 / DRAM of the 63 conv_igemm launches of one step (ncu, cold caches, serialised)
# usage: bash tools/gpu_conv_launches.sh <tag> [PS_CONV_DEBUG]
mkdir -p gpurun_out
PS_CONV_DEBUG=${2:-0} timeout 900 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,dram__throughput.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:conv_igemm -s 63 -c 63 --csv --log-file gpurun_out/conv_launches_$1.csv \
    python bench.py --steps 1 --warmup 1 --batch 64 --no-cpu-baseline > gpurun_out/ncu_conv_launch_$1.log 2>&1
tail -2 gpurun_out/ncu_conv_launch_$1.log | cut -c1-200
