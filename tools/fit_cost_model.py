#!/usr/bin/env python3
"""Developer tool (GPU box, 1 GPU): measures the device time of many tiles of the bench world (recompute and exchange-style
stage sets) together with their stage-1 cost features, for fitting sharding.COST_WEIGHTS offline.
usage: fit_cost_model.py [world=256] > gpurun_out/cost_samples.json        (GPU box)
       fit_cost_model.py --fit gpurun_out/cost_samples.json                  (anywhere: least squares, prints the weights and the errors)"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

if len(sys.argv) > 2 and sys.argv[1] == "--fit":
    samples = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    ms = np.array([s["ms"] for s in samples])
    n = np.array([s["tile"][2] * s["tile"][3] for s in samples], float)
    own = np.array([s["own"] for s in samples])
    # ms ~ a * cave voxels / 1e4 + b * fill voxels / 1e4 + c * land columns / 256 + d * chunks + e (per tile)
    A = np.stack([own[:, 0] / 1e4, own[:, 1] / 1e4, own[:, 2] / 256.0, n, np.ones_like(n)], axis=1)
    from scipy.optimize import nnls      # non-negative: cave and fill voxels are strongly correlated, plain least squares lets one go negative
    w, _ = nnls(A, ms)
    pred = A @ w
    area = np.linalg.lstsq(np.stack([n, np.ones_like(n)], axis=1), ms, rcond=None)[0]
    pa = np.stack([n, np.ones_like(n)], axis=1) @ area
    print("# weights (ms): cave voxels / 1e4 %.4g, fill voxels / 1e4 %.4g, land columns / 256 %.4g, per chunk %.4g, per tile %.3g" % tuple(w))
    print("# tile, measured ms, predicted ms, error %")
    for s_, m, p_ in zip(samples, ms, pred):
        print(tuple(s_["tile"]), "%.2f %.2f %+.2f" % (m, p_, 100 * (p_ - m) / m))
    err = (pred - ms) / ms
    ea = (pa - ms) / ms
    print("# rms error %.2f %%, worst %.2f %%; by area alone %.1f %% / %.1f %%" % (100 * np.sqrt((err ** 2).mean()), 100 * np.abs(err).max(),
                                                                                 100 * np.sqrt((ea ** 2).mean()), 100 * np.abs(ea).max()))
    sys.exit(0)
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
