#!/usr/bin/env python3
"""Developer tool (no GPU needed): static SASS statistics of every kernel in a built library - instruction count, registers,
stack, spill instructions, and the instruction mix (packed fp32 FFMA2 / FADD2 / FMUL2 against their scalar forms, LDS, ...).
usage: python tools/sass_stats.py [path/to/libmmgen.so] [kernel-substring ...]      (profiles/r02_sass_mix.txt is its output)"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 and sys.argv[1].endswith(".so") else os.path.join(ROOT, "mega-minecraft_b200", "libmmgen.so")
want = [a for a in sys.argv[1:] if not a.endswith(".so")]

sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
res = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True, check=True).stdout
usage = {}
for m in re.finditer(r"Function (\S+):\n\s+REG:(\d+) STACK:(\d+) SHARED:(\d+)", res):
    usage[m.group(1)] = (int(m.group(2)), int(m.group(3)), int(m.group(4)))

GROUPS = [("FFMA2", "FFMA2"), ("FADD2", "FADD2"), ("FMUL2", "FMUL2"), ("FFMA", "FFMA"), ("FADD", "FADD"), ("FMUL", "FMUL"), ("FMNMX", "FMNMX"),
          ("FSETP", "FSETP"), ("FSEL", "FSEL"), ("FRND", "FRND"), ("MUFU", "MUFU"), ("LDS", "LDS"), ("STS", "STS"), ("LDG", "LDG"), ("STG", "STG"),
          ("LDL", "LDL"), ("STL", "STL"), ("ATOM", "ATOM"), ("ATOMS", "ATOMS"), ("RED", "RED"), ("SHFL", "SHFL"), ("BAR", "BAR"), ("CALL", "CALL")]
print("%-26s %6s %4s %5s %6s | %s" % ("kernel", "instr", "reg", "stack", "smem", " ".join("%5s" % g for g, _ in GROUPS)))
for part in sass.split("Function : ")[1:]:
    name = part.split("\n", 1)[0].strip()
    short = re.sub(r"^_ZN3mmg\d+", "", name)
    short = re.split(r"(EP|ILb|Ev$|Ei)", short)[0] + ("<1>" if "ILb1" in name else "<0>" if "ILb0" in name else "")
    if want and not any(w in short for w in want):
        continue
    ops = collections.Counter()
    for m in re.finditer(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+)", part):
        ops[m.group(1)] += 1
    total = sum(ops.values()) - ops.get("NOP", 0)
    reg, stack, smem = usage.get(name, (0, 0, 0))
    print("%-26s %6d %4d %5d %6d | %s" % (short, total, reg, stack, smem, " ".join("%5d" % ops.get(k, 0) for _, k in GROUPS)))
