#!/usr/bin/env python3
"""Developer tool: per-source-line hot spots of an ncu report captured with --import-source on.
usage: tools/ncu_hot.py report.ncu-rep [top]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
cur_file, hdr, lines = None, None, {}
for r in rows:
    if len(r) == 2 and r[0] in ("File Name", "File Path"):
        cur_file = r[1].split("/")[-1]
        continue
    if len(r) > 8 and r[0] == "Line No":
        hdr = r
        ie, isamp, it = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Thread Instructions Executed")
        continue
    if hdr is None or len(r) < len(hdr):
        continue
    if r[2] == "-" or r[2] == "":      # a source line row (aggregated over its SASS)
        try:
            lines[(cur_file, int(r[0]))] = (int(r[ie]), int(r[isamp]), int(r[it]), r[1].strip()[:100])
        except ValueError:
            pass
tot = sum(v[0] for v in lines.values()) or 1
tots = sum(v[1] for v in lines.values()) or 1
print("total warp-instructions %d, samples %d" % (tot, tots))
for (f, ln), v in sorted(lines.items(), key=lambda kv: -kv[1][1])[:top]:
    print("%5.1f%% smp %5.1f%% inst thr/inst %4.1f  %s:%d | %s" % (100 * v[1] / tots, 100 * v[0] / tot, v[2] / max(v[0], 1), f, ln, v[3]))
