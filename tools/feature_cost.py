#!/usr/bin/env python3
"""Developer tool (GPU box): k_fill_features time with surface feature types switched off one at a time."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mmgen_loader
mm = mmgen_loader.load()
gen = mm.ChunkGen(0)
S = int(sys.argv[1]) if len(sys.argv) > 1 else 128
w = gen.region_world(0, 0, S, S)
names = "NONE SPHERE CORAL KELP ICEBERG ACACIA REDWOOD CYPRESS BIRCH PINE PINE_SHRUB RAFFLESIA LARGE_JUNGLE SMALL_JUNGLE TINY_JUNGLE MEDIUM_PURPLE_MUSHROOM PURPLE_MUSHROOM MEDIUM_CRYSTAL CRYSTAL PALM CACTUS".split()
def run(mask):
    gen.L.mmgen_debug_feature_mask(mask)
    w.reset(); w.generate(mm.STAGE_ALL); w.sync()
    gen.kernel_timing(True)
    w.reset(); w.generate(mm.STAGE_ALL); w.sync()
    t = gen.kernel_times(); gen.kernel_timing(False)
    return t["k_fill_features"][0]
full = run(0xffffffff)
none = run(0)
print("all types %.2f ms, no surface features %.2f ms" % (full, none))
for t in (16, 2, 4, 17):
    print("without %-24s %.2f ms  (saves %.2f)" % (names[t], run(0xffffffff & ~(1 << t)), full - run(0xffffffff & ~(1 << t))))
cnames = "NONE TEST_GLOWSTONE_PILLAR TEST_SHROOMLIGHT_PILLAR CAVE_VINE GLOWSTONE_CLUSTER STORMLIGHT_SPHERE CEILING_STORMLIGHT_SPHERE CRYSTAL_PILLAR WARPED_FUNGUS AMBER_FUNGUS".split()
nocave = run(0xffffffff & ~(0x3ff << 21))
print("no cave features %.2f ms" % nocave)
for t in range(3, 10):
    print("without cave %-26s %.2f ms  (saves %.2f)" % (cnames[t], run(0xffffffff & ~(1 << (21 + t))), full - run(0xffffffff & ~(1 << (21 + t)))))
gen.L.mmgen_debug_feature_mask(0xffffffff)
