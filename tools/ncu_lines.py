#!/usr/bin/env python
"""Per-source-line instruction / stall-sample shares of one kernel from an ncu report.

    python tools/ncu_lines.py <report.ncu-rep> <lib.so> <kernel mangled-name substring> [min_pct]

Joins `ncu --page source --csv` (SASS view: per-address instruction counts and stall samples) with
`nvdisasm -g` line info of the cubin extracted from the .so (compiled with -lineinfo)."""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile


def sass_line_map(so, kernel):
    d = tempfile.mkdtemp()
    subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=d, stdout=subprocess.DEVNULL)
    amap = {}
    for f in os.listdir(d):
        txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(d, f)], capture_output=True, text=True).stdout
        infn, cur = False, None
        for ln in txt.splitlines():
            if ln.startswith("//---") and ".text." in ln:
                infn = kernel in ln
                cur = None
            if not infn:
                continue
            m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
            if m:
                if "inlined at" not in ln or cur is None:
                    cur = (os.path.basename(m.group(1)), int(m.group(2)))
                # keep the outermost user-code line when the inline chain mentions our file
                m2 = re.findall(r'File "([^"]+)", line (\d+)', ln)
                for fn, l in m2:
                    if fn.endswith(".cu"):
                        cur = (os.path.basename(fn), int(l))
                continue
            m = re.match(r"\s*/\*([0-9a-f]+)\*/\s+(\S.*?);", ln)
            if m and cur:
                amap[int(m.group(1), 16)] = (cur, m.group(2))
    return amap


def main():
    rep, so, kernel = sys.argv[1:4]
    minpct = float(sys.argv[4]) if len(sys.argv) > 4 else 0.7
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    h = rows[hdr]
    ci, si, ai = h.index("Instructions Executed"), h.index("# Samples"), h.index("Address")
    stall_cols = [i for i, n in enumerate(h) if n.startswith("stall_") and "Not Issued" not in n]
    amap = sass_line_map(so, kernel)
    base = None
    per = {}
    ti = ts = 0
    for r in rows[hdr + 1:]:
        try:
            a = int(r[ai], 16) if not r[ai].isdigit() else int(r[ai])
            ie, ss = int(r[ci]), int(r[si])
        except (ValueError, IndexError):
            continue
        if base is None:
            base = a
        key, _ = amap.get(a - base, (("?", 0), ""))
        e = per.setdefault(key, [0, 0, {}])
        e[0] += ie
        e[1] += ss
        for c in stall_cols:
            try:
                v = int(r[c])
            except ValueError:
                v = 0
            if v:
                e[2][h[c]] = e[2].get(h[c], 0) + v
        ti += ie
        ts += ss
    print("total warp-instructions %d, samples %d" % (ti, ts))
    src = {}
    for (fn, l), (ie, ss, st) in sorted(per.items(), key=lambda kv: kv[0]):
        if ie * 100.0 / max(ti, 1) < minpct and ss * 100.0 / max(ts, 1) < minpct:
            continue
        top = sorted(st.items(), key=lambda kv: -kv[1])[:3]
        text = ""
        for root in (".", "pixelsynth_b200/csrc"):
            p = os.path.join(root, fn)
            if os.path.exists(p):
                src.setdefault(p, open(p).read().splitlines())
                if 0 < l <= len(src[p]):
                    text = src[p][l - 1].strip()[:70]
                break
        print("%-12s:%4d inst %5.1f%% smp %5.1f%% %-44s | %s" % (fn, l, ie * 100.0 / ti, ss * 100.0 / max(ts, 1),
              " ".join("%s=%d" % (k.replace("stall_", ""), v) for k, v in top), text))


if __name__ == "__main__":
    main()
