#!/usr/bin/env python3
"""Developer tool (GPU box): one generate of a small region world, for ncu captures (tools/gpu_profile.sh)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mmgen_loader  # noqa: E402

mm = mmgen_loader.load()
side = int(sys.argv[1]) if len(sys.argv) > 1 else 24
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
gen = mm.ChunkGen(0)
world = gen.region_world(0, 0, side, side)
for _ in range(reps):
    world.reset()
    world.generate(mm.STAGE_ALL)
world.sync()
print("stage ms", world.stage_ms(), "total", world.total_ms(), "launches", gen.launch_count())
world.close()
