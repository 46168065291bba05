#!/usr/bin/env python3
"""Developer tool: hottest CUDA source lines of an ncu report (--set full --import-source on): share of warp-stall samples, share
of executed warp instructions, active threads per instruction; plus per-file totals and the raw metrics that matter here.
usage: python tools/ncu_lines.py report.ncu-rep [top-n]"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
if len(rows) >= 3:
    hdr, units, vals = rows[0], rows[1], rows[2]
    want = ["gpu__time_duration.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
            "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__registers_per_thread", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps"]
    want += [h for h in hdr if "issue_stalled" in h and h.endswith("per_issue_active.ratio")]
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            print("%-95s %-12s %s" % (w, units[i], vals[i]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur = None
hdr = None
agg = collections.defaultdict(lambda: [0, 0, 0.0])
for r in csv.reader(src.splitlines()):
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        iW = hdr.index("Warp Stall Sampling (All Samples)"); iE = hdr.index("Instructions Executed"); iT = hdr.index("Thread Instructions Executed")
        continue
    if hdr and r[0].isdigit():
        try:
            w = int(r[iW]); e = int(r[iE]); t = int(r[iT])
        except ValueError:
            continue
        k = (cur, int(r[0]), r[1].strip()[:120])
        agg[k][0] += w; agg[k][1] += e; agg[k][2] += t
tw = sum(v[0] for v in agg.values()) or 1
te = sum(v[1] for v in agg.values()) or 1
print("\n# per file: share of stall samples, share of executed warp instructions")
byf = collections.defaultdict(lambda: [0, 0])
for (f, l, s), v in agg.items():
    byf[f][0] += v[0]; byf[f][1] += v[1]
for f, v in sorted(byf.items(), key=lambda kv: -kv[1][0]):
    print("%5.1f%% smp %5.1f%% inst  %s" % (100 * v[0] / tw, 100 * v[1] / te, f))
print("\n# hottest source lines (total warp-instructions %d, samples %d)" % (te, tw))
for (f, l, s), v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%5.1f%% smp %5.1f%% inst thr/inst %4.1f  %s:%d | %s" % (100 * v[0] / tw, 100 * v[1] / te, v[2] / max(v[1], 1), f, l, s))
