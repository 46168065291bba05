#!/usr/bin/env python3
"""Developer tool (GPU box): per-kernel device times of one region world for the library named by MMGEN_LIB (tuning builds).
usage: MMGEN_LIB=path/to/variant.so python tools/variant_time.py [side] [kernel ...]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mmgen_loader  # noqa: E402

mm = mmgen_loader.load()
gen = mm.ChunkGen(0)
S = int(sys.argv[1]) if len(sys.argv) > 1 else 128
want = sys.argv[2:]
w = gen.region_world(0, 0, S, S)
w.generate(mm.STAGE_ALL); w.sync()
best = {}
for rep in range(2):
    w.reset()
    gen.kernel_timing(True)
    w.generate(mm.STAGE_ALL); w.sync()
    t = gen.kernel_times(); gen.kernel_timing(False)
    for k, (ms, n) in t.items():
        best[k] = min(best.get(k, 1e30), ms)
tot = sum(best.values())
print(os.path.basename(os.environ.get("MMGEN_LIB", "libmmgen.so")), "total %.1f ms" % tot,
      " ".join("%s %.2f" % (k, v) for k, v in best.items() if (not want and v > 1.0) or k in want))
