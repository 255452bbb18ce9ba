#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,dram__throughput.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:conv_igemm -s 63 -c 63 --csv --log-file gpurun_out/conv_launches.csv \
    python bench.py --steps 1 --warmup 1 --batch 32 --no-cpu-baseline > gpurun_out/ncu_conv_launch.log 2>&1
tail -2 gpurun_out/ncu_conv_launch.log | cut -c1-300
