#!/usr/bin/env python3
"""Developer tool (GPU box): prints the block hashes tests/test_region_hashes.py pins (run after a deliberate change of
results, never to paper over an accidental one)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mmgen_loader  # noqa: E402

mm = mmgen_loader.load()
gen = mm.ChunkGen(0)
for reg in ((0, 0, 48, 48), (-300, 500, 32, 32), (4000, -4000, 24, 36)):
    w = gen.region_world(*reg)
    w.generate(mm.STAGE_ALL)
    w.sync()
    print(reg, "%016x" % w.chunk_hash_sum())
    w.close()
