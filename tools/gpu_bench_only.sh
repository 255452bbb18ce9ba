#!/bin/bash
mkdir -p gpurun_out
timeout 400 python bench.py "$@" > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench.json"))
print(round(d["value"], 1), "views/s", round(d["ms_per_step"], 2), "ms/step e2e", round(d["e2e"]["value"], 1), d["clocks"], "launches", d["gpu_launches"])
for k, v in d["rooflines"].items():
    print("   ", k, round(v["achieved"], 1), v["unit"], "frac", round(v["frac"], 4), "ms", round(v["ms_per_step"], 3))
if "cpu_baseline" in d: print("    cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["sample"][:100])
PY
