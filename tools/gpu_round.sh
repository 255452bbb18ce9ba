#!/bin/bash
# One GPU-box visit: tests, bench, per-stage times, ncu launch list + full captures. Outputs under gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
timeout 300 python tools/stage_times.py > gpurun_out/stage_times.txt 2>&1
timeout 300 python tools/bench_lmconv.py > gpurun_out/bench_lmconv.json 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 1 --batch 8 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lmconv_tc -s 1 -c 1 -f -o gpurun_out/prof_lmconv \
    python tools/bench_lmconv.py --reps 1 > gpurun_out/ncu_lmconv.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_igemm -s 60 -c 6 -f -o gpurun_out/prof_conv \
    python bench.py --steps 1 --warmup 1 --batch 8 --no-cpu-baseline > gpurun_out/ncu_conv.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/bench.json | cut -c1-600; cat gpurun_out/stage_times.txt; cat gpurun_out/bench_lmconv.json
