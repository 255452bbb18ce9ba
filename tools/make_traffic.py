#!/usr/bin/env python
"""profiles/traffic.json from ncu --set full reports: DRAM bytes (read + write) per launch of the kernels bench.py
reports rooflines for.    python tools/make_traffic.py kernel=report.ncu-rep[:launch_index] ..."""
import csv
import io
import json
import subprocess
import sys


def dram_bytes(rep, index):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    tot = 0.0
    for name in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        i = hdr.index(name)
        v = float(data[index][i].replace(",", ""))
        tot += v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[units[i]]
    return tot


out = {"source": "ncu --set full --clock-control none, DRAM read+write bytes of one launch (profiles/*_ncu_full*.txt)"}
for arg in sys.argv[1:]:
    k, rep = arg.split("=")
    idx = 0
    if ":" in rep:
        rep, idx = rep.rsplit(":", 1)
    out[k] = dram_bytes(rep, int(idx))
json.dump(out, open("profiles/traffic.json", "w"), indent=1)
print(out)
