#!/bin/bash
# Last GPU visit of round 1 (short budget): the splat encoding V=1 (packed ids, byte-offset slot lists, unrolled map
# writer) timed against V=0 and run through the parity suite; the op-seam + demo tests; smoke(); one ncu capture of
# fine_kernel V=1; one bench line.  Most valuable first: the call may be cut by the remaining budget.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
timeout 150 python tools/bench_splat.py > gpurun_out/splat_v0.json 2> gpurun_out/splat_v0.err; cat gpurun_out/splat_v0.json
PS_SPLAT_VARIANT=1 timeout 100 python tools/bench_splat.py > gpurun_out/splat_v1.json 2> gpurun_out/splat_v1.err; cat gpurun_out/splat_v1.json
PS_SPLAT_VARIANT=1 timeout 330 python -m pytest tests/test_splat_gpu.py tests/test_zz_ops_gpu.py -q -m gpu \
    > gpurun_out/pytest_v1.log 2>&1; echo "pytest V1 (splat + ops + demo) rc=$?"; tail -6 gpurun_out/pytest_v1.log
PS_SPLAT_VARIANT=1 timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_v1.log 2>&1; echo "smoke V1 rc=$?"; tail -3 gpurun_out/smoke_v1.log
PS_SPLAT_VARIANT=1 timeout 150 ncu --set full --clock-control none --import-source on -k regex:ps::fine_kernel -s 3 -c 1 -f \
    -o gpurun_out/prof_fine_v1 python tools/bench_splat.py > gpurun_out/ncu_fine_v1.log 2>&1; tail -2 gpurun_out/ncu_fine_v1.log
PS_SPLAT_VARIANT=1 timeout 240 python bench.py --no-cpu-baseline > gpurun_out/bench_v1.json 2> gpurun_out/bench_v1.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/bench_v1.json
