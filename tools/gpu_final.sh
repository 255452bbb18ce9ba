#!/bin/bash
# Round-end evidence: bench lines, launch list, ncu full captures of the three hot kernels.
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"
timeout 600 python bench.py --batch 128 --no-cpu-baseline > gpurun_out/bench_b128.json 2> gpurun_out/bench_b128.err; echo "bench128 rc=$?"
timeout 300 python tools/bench_lmconv.py > gpurun_out/bench_lmconv.json 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_final.csv \
    python bench.py --steps 1 --warmup 1 --batch 8 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lmconv_tc -s 1 -c 1 -f -o gpurun_out/prof_lmconv_final \
    python tools/bench_lmconv.py --reps 1 > gpurun_out/ncu_lmconv.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_igemm -s 13 -c 2 -f -o gpurun_out/prof_conv_final \
    python tools/bench_conv.py > gpurun_out/ncu_conv.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fine_kernel -s 4 -c 1 -f -o gpurun_out/prof_fine_final \
    python bench.py --steps 1 --warmup 1 --batch 16 --no-cpu-baseline > gpurun_out/ncu_fine.log 2>&1
python - <<'PY'
import json
for f in ("bench_final", "bench_b128"):
    d = json.load(open("gpurun_out/%s.json" % f))
    print(f, round(d["value"], 1), "views/s", round(d["ms_per_step"], 2), "ms/step e2e", round(d["e2e"]["value"], 1), d["clocks"])
    for k, v in d["rooflines"].items():
        print("   ", k, round(v["achieved"], 1), v["unit"], "frac", round(v["frac"], 4), "ms", round(v["ms_per_step"], 3))
    if "cpu_baseline" in d: print("    cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["sample"][:120])
PY
cat gpurun_out/bench_lmconv.json
