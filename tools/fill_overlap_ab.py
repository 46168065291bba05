#!/usr/bin/env python3
"""Developer tool (GPU box): device time of a full generate of one region world under each mmgen_set_fill_overlap mode, with the
per-chunk hash sum as the parity check.  usage: python tools/fill_overlap_ab.py [side] [mode ...]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mmgen_loader  # noqa: E402

mm = mmgen_loader.load()
gen = mm.ChunkGen(0)
S = int(sys.argv[1]) if len(sys.argv) > 1 else 256
modes = [int(a) for a in sys.argv[2:]] or [0, 8, 6, 4, 24, 20]
w = gen.region_world(0, 0, S, S)
w.generate(mm.STAGE_ALL); w.sync()
for mode in modes + modes[:1]:
    gen.set_fill_overlap(mode)
    best, s6 = 1e30, 0.0
    for rep in range(3):
        w.reset()
        w.generate(mm.STAGE_ALL); w.sync()
        t = w.total_ms()
        if t < best:
            best, s6 = t, float(w.stage_ms()[6])
    print("mode %2d  total %.2f ms  S6 %.2f ms  hash %016x" % (mode, best, s6, w.chunk_hash_sum()), flush=True)
