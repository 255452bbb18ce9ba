#!/bin/bash
# Round-2 closing visit: GPU suite, smoke, default bench line, scene workload, ncu launch list of the bench command
# (with --in-flight 1: ncu cannot profile kernels launched into a green context, "Failed to prepare kernel for profiling").
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu_final.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r02_pytest_gpu_final.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc=$?"
timeout 600 python bench.py > gpurun_out/r02_bench_n1_final.json 2> gpurun_out/r02_bench_n1_final.err; echo "bench rc=$?"
timeout 600 python bench.py --workload scene --no-cpu-baseline > gpurun_out/r02_bench_scene.json 2> gpurun_out/r02_bench_scene.err; echo "scene rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02_launches_final.csv \
    python bench.py --steps 2 --warmup 1 --in-flight 1 --no-cpu-baseline > gpurun_out/r02_ncu_launch_final.log 2>&1
tail -2 gpurun_out/r02_pytest_gpu_final.log; tail -1 gpurun_out/r02_smoke.log; cut -c1-400 gpurun_out/r02_bench_n1_final.json; cut -c1-600 gpurun_out/r02_bench_scene.json
