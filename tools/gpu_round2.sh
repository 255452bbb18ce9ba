#!/bin/bash
# Round-2 evidence visit: tests, bench, ncu launch list of the bench command, ncu --set full of the three hot kernels.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r02_gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r02_pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err; echo "bench rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02_ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lmconv_tc -s 1 -c 1 -f -o gpurun_out/r02_prof_lmconv \
    python tools/bench_lmconv.py --reps 1 > gpurun_out/r02_ncu_lmconv.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_igemm -s 3 -c 1 -f -o gpurun_out/r02_prof_conv \
    python tools/bench_conv.py > gpurun_out/r02_ncu_conv.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fine_kernel -s 3 -c 1 -f -o gpurun_out/r02_prof_fine \
    python tools/bench_splat.py > gpurun_out/r02_ncu_fine.log 2>&1
bash tools/gpu_conv_launches.sh r02 0
tail -2 gpurun_out/r02_pytest_gpu.log; cut -c1-300 gpurun_out/r02_bench.json; ls -la gpurun_out/*.ncu-rep
