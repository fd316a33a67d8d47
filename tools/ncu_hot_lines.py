#!/usr/bin/env python
"""tools/ncu_hot_lines.py -- per source line: stall samples and executed warp instructions, from `ncu --page source --csv --print-source cuda,sass`.
usage: ncu_hot_lines.py report.ncu-rep [top]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
cur_file = None
lines = {}
total_s = total_i = 0
hdr = None
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        i_samp = hdr.index("# Samples"); i_inst = hdr.index("Instructions Executed")
        continue
    if hdr is None or len(r) < len(hdr) or r[0] == "":
        continue
    try:
        ln = int(r[0]); s = int(r[i_samp]); i = int(r[i_inst])
    except ValueError:
        continue
    key = (cur_file, ln)
    a = lines.setdefault(key, [0, 0, r[1].strip()[:110]])
    a[0] += s; a[1] += i
    total_s += s; total_i += i
print("total samples %d, warp instructions %d" % (total_s, total_i))
for (f, ln), (s, i, src) in sorted(lines.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%5.1f%% samp %5.1f%% inst  %-16s:%-4d %s" % (100.0 * s / max(total_s, 1), 100.0 * i / max(total_i, 1), f, ln, src))
