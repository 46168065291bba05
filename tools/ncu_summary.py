#!/usr/bin/env python3
"""Developer tool: text summary of an ncu report (key metrics + hottest source lines) for profiles/.
usage: tools/ncu_summary.py report.ncu-rep > profiles/rNN_kernel.txt"""
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
want = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__cycles_elapsed.avg.per_second"]
print("# ncu --set full --clock-control none --import-source on, report %s" % rep.split("/")[-1])
if len(rows) >= 3:
    print("kernel:", rows[2][rows[0].index("Kernel Name")] if "Kernel Name" in rows[0] else "?")
    for h, u, v in zip(*rows[:3]):
        if h in want or ("issue_stalled" in h and h.endswith("per_issue_active.ratio") and float(v or 0) >= 0.2):
            print("%-88s %-12s %s" % (h, u, v))
print()
print("# hottest source lines (share of warp-stall samples, share of executed warp instructions, active threads per instruction)")
sys.stdout.flush()
subprocess.run([sys.executable, __file__.replace("ncu_summary.py", "ncu_hot.py"), rep, "25"])
