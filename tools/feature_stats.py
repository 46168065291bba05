#!/usr/bin/env python3
"""Developer tool (GPU box): surface-feature census of a region: placements per type and the (column, y) pairs their boxes
put on the rasteriser (box = reach clipped to nothing, height bounds), to see which rasterisers carry the surface pass."""
import collections
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mmgen_loader  # noqa: E402

mm = mmgen_loader.load()
gen = mm.ChunkGen(0)
S = int(sys.argv[1]) if len(sys.argv) > 1 else 128
w = gen.region_world(0, 0, S, S)
w.generate(mm.STAGE_ALL & ~mm.STAGE_FILL)
F, CF = w.download_features(512)
names = "NONE SPHERE CORAL KELP ICEBERG ACACIA REDWOOD CYPRESS BIRCH PINE PINE_SHRUB RAFFLESIA LARGE_JUNGLE SMALL_JUNGLE TINY_JUNGLE MEDIUM_PURPLE_MUSHROOM PURPLE_MUSHROOM MEDIUM_CRYSTAL CRYSTAL PALM CACTUS".split()
reach = [0, 5, 8, 0, 43, 15, 20, 12, 8, 6, 6, 15, 15, 8, 1, 8, 70, 25, 25, 24, 5]
hb = [(0, 0), (-6, 6), (-3, 12), (0, 20), (0, 110), (0, 15), (-5, 75), (-3, 50), (0, 30), (0, 15), (0, 8), (0, 10), (0, 38), (0, 17), (0, 5), (0, 6), (0, 120), (-3, 32), (-6, 64), (0, 28), (0, 15)]
cnt = collections.Counter()
for f in F:
    cnt.update(f["feature"].tolist())
tot = 0
rows = []
for t, c in cnt.items():
    vox = c * (2 * reach[t] + 1) ** 2 * (hb[t][1] - hb[t][0] + 1)
    rows.append((vox, names[t], c))
    tot += vox
for vox, n, c in sorted(rows, reverse=True):
    print("%-24s placements %8d  box voxels %.3e (%.1f %%)" % (n, c, vox, 100 * vox / tot))
print("chunks with placements lists:", len(F), "cave placements", sum(len(c) for c in CF))
