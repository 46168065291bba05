#!/usr/bin/env python3
"""Developer tool (GPU box, 1 GPU): measures the device time of many tiles of the bench world (recompute and exchange-style
stage sets) together with their stage-1 cost features, for fitting sharding.COST_WEIGHTS offline.
usage: fit_cost_model.py [world=256] > gpurun_out/cost_samples.json"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mmgen_loader  # noqa: E402

mm = mmgen_loader.load()
from mega_minecraft_b200 import tiling  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 256
gen = mm.ChunkGen(0)
origins = np.array([[x * 16, z * 16] for z in range(-3, S + 3) for x in range(-3, S + 3)], np.int32)
feat = gen.chunk_costs(origins).reshape(S + 6, S + 6, 3)
samples = []
for gx, gz in ((4, 2), (2, 4), (3, 3)):
    xs, zs = tiling.split_points(0, S, gx), tiling.split_points(0, S, gz)
    for j in range(gz):
        for i in range(gx):
            t = (xs[i], zs[j], xs[i + 1] - xs[i], zs[j + 1] - zs[j])
            w = gen.region_world(*t)
            w.generate(mm.STAGE_ALL)
            w.reset()
            w.generate(mm.STAGE_ALL)
            ms, st = w.total_ms(), [float(v) for v in w.stage_ms()]
            w.close()
            own = feat[t[1] + 3:t[1] + 3 + t[3], t[0] + 3:t[0] + 3 + t[2]].sum(axis=(0, 1))
            ring = feat[t[1]:t[1] + t[3] + 6, t[0]:t[0] + t[2] + 6].sum(axis=(0, 1))
            samples.append({"tile": t, "ms": ms, "stage_ms": st, "own": [float(v) for v in own], "grown3": [float(v) for v in ring]})
            print(t, round(ms, 1), file=sys.stderr)
print(json.dumps(samples))
