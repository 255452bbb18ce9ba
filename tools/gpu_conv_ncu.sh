#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_igemm -s 122 -c 2 -f -o gpurun_out/prof_conv2 \
    python bench.py --steps 1 --warmup 1 --batch 32 --no-cpu-baseline > gpurun_out/ncu_conv2.log 2>&1
tail -2 gpurun_out/ncu_conv2.log | cut -c1-200
