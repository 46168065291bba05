#!/usr/bin/env python3
"""Developer tool (GPU box, 1 GPU): device time of each tile of a gx x gz tiling of the bench world, one after
the other on one GPU: the load imbalance a static tiling carries. usage: tile_times.py world gx,gz [gx,gz ...]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mmgen_loader  # noqa: E402

mm = mmgen_loader.load()
from mega_minecraft_b200 import tiling  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 256
grids = [tuple(int(v) for v in a.split(",")) for a in sys.argv[2:]] or [(4, 2)]
gen = mm.ChunkGen(0)
for gx, gz in grids:
    xs, zs = tiling.split_points(0, S, gx), tiling.split_points(0, S, gz)
    rows = []
    for j in range(gz):
        for i in range(gx):
            t = (xs[i], zs[j], xs[i + 1] - xs[i], zs[j + 1] - zs[j])
            w = gen.region_world(*t)
            w.generate(mm.STAGE_ALL)          # warm-up (allocations)
            w.reset()
            w.generate(mm.STAGE_ALL)
            rows.append((t, w.total_ms(), [round(float(x), 1) for x in w.stage_ms()[1:]]))
            w.close()
    tot = [r[1] for r in rows]
    print("GRID %dx%d  sum %.1f ms  max %.1f  mean %.1f  imbalance %.3f" % (gx, gz, sum(tot), max(tot), sum(tot) / len(tot), max(tot) / (sum(tot) / len(tot))))
    print("   times", [round(v, 1) for v in tot])
