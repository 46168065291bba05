#!/usr/bin/env python3
"""Developer tool (GPU box): census of k_fill_features per feature type - (column, y)
pairs offered to the filter, pairs that reached the rasteriser, hits. Needs the stats build:
  nvcc <NVCC_FLAGS> -DMMG_FEATURE_STATS -o mega-minecraft_b200/libmmgen_stats.so mega-minecraft_b200/csrc/mmgen.cu
  MMGEN_LIB=mega-minecraft_b200/libmmgen_stats.so python tools/feature_census.py 128"""
import ctypes
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
names = "NONE SPHERE CORAL KELP ICEBERG ACACIA REDWOOD CYPRESS BIRCH PINE PINE_SHRUB RAFFLESIA LARGE_JUNGLE SMALL_JUNGLE TINY_JUNGLE MEDIUM_PURPLE_MUSHROOM PURPLE_MUSHROOM MEDIUM_CRYSTAL CRYSTAL PALM CACTUS".split()
cnames = "NONE TEST_GLOWSTONE_PILLAR TEST_SHROOMLIGHT_PILLAR CAVE_VINE GLOWSTONE_CLUSTER STORMLIGHT_SPHERE CEILING_STORMLIGHT_SPHERE CRYSTAL_PILLAR WARPED_FUNGUS AMBER_FUNGUS".split()
st = np.zeros((64, 4), np.uint64)
w.generate(mm.STAGE_ALL); w.sync()
gen.L.mmgen_debug_feature_stats(st.ctypes.data_as(ctypes.c_void_p))
tot = max(float(st[:, 2].sum()), 1.0)
rows = []
for i in range(64):
    if st[i, 1] == 0:
        continue
    n = names[i] if i < 32 else "cave " + cnames[i - 32]
    rows.append((int(st[i, 0]), n, int(st[i, 1]), int(st[i, 2]), int(st[i, 3])))
print("%-30s %8s %12s %12s %12s" % ("type", "rast.%", "pairs", "rasterised", "hits"))
for c, n, a, b, h in sorted(rows, key=lambda r: -r[3]):
    print("%-30s %7.1f%% %12d %12d %12d" % (n, 100 * b / tot, a, b, h))
hs = np.zeros(9, np.uint64)
gen.L.mmgen_debug_huge_stats(hs.ctypes.data_as(ctypes.c_void_p))
print("huge-caves term: proved zero for %d of %d threshold voxels (%.1f %%), proof wrong for %d (must be 0)"
      % (hs[2], hs[1] + hs[2], 100.0 * float(hs[2]) / max(float(hs[1] + hs[2]), 1.0), hs[0]))
hgt = w.download(heightfield=True)["heightfield"]
st = w.stages().ravel()
alg = int(np.maximum(np.floor(hgt).astype(np.int64)[st >= 4], 128).sum())
print("k_caves: algorithmic voxels (0 < y <= max(h, 128)) %d, threshold evaluated %d (%.3f), warped noise + Worley evaluated %d (%.3f)"
      % (alg, hs[1] + hs[2], float(hs[1] + hs[2]) / alg, hs[3], float(hs[3]) / alg))
print("k_caves threshold bounds: %d voxels decided without fbmA, %d needed it (%.1f %%), decided wrongly %d (must be 0)"
      % (hs[4], hs[5], 100.0 * float(hs[5]) / max(float(hs[4] + hs[5]), 1.0), hs[6]))
print("k_caves jitter table: %d Worley evaluations had all 27 cells in the CTA's table, %d did not (%.2f %%)"
      % (hs[8], hs[7], 100.0 * float(hs[7]) / max(float(hs[7] + hs[8]), 1.0)))
