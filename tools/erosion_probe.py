#!/usr/bin/env python3
"""Developer tool (GPU box): where do the product's eroded layers differ from the unmodified reference's in a multi-zone window,
and how does that depend on the order in which the reference erodes its zones? (tests/test_reference_tour.py)"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import mmgen_loader  # noqa: E402
from oracle import refcuda  # noqa: E402

mm = mmgen_loader.load()
from mega_minecraft_b200 import tiling  # noqa: E402

REL = 1e-5
gen = mm.ChunkGen(0)
ref = refcuda.RefCuda(0)
region = tuple(int(a) for a in sys.argv[1:5]) if len(sys.argv) >= 5 else (-300, 500, 32, 32)
x0, z0, nx, nz = tiling.apron_window(*region)
w = gen.world(x0, z0, nx, nz)
w.generate(mm.STAGE_HEIGHTFIELD | mm.STAGE_LAYERS | mm.STAGE_EROSION)
d = w.download(layers=True)
st = w.stages().ravel()
w.close()
eroded = np.nonzero(st >= 3)[0]


def deviations(a, b):
    A, B = a[eroded][:, 10:].astype(np.float64), b[eroded][:, 10:].astype(np.float64)
    bad = (np.abs(A - B) > REL * np.abs(B)).any(axis=1)
    out = []
    for k, col in np.argwhere(bad):
        c = int(eroded[k])
        cx, cz, x, z = x0 + c % nx, z0 + c // nx, int(col) % 16, int(col) // 16
        zx, zz = (cx // 12) * 12, (cz // 12) * 12
        out.append((zx, zz, (cx - zx + 6) * 16 + x, (cz - zz + 6) * 16 + z, float(np.abs(A[k, :, col] - B[k, :, col]).max())))
    return out


res = {"window": [x0, z0, nx, nz], "zones": len(eroded) // 144}
runs = []
for order in (0, 0, 1):
    ref.L.mmref_set_zone_order(order)
    runs.append(ref.generate(x0, z0, nx, nz, 3)["layers"])
ref.L.mmref_set_zone_order(0)
for name, r in zip(("asc_run1", "asc_run2", "desc"), runs):
    dev = deviations(d["layers"], r)
    by_zone = {}
    for zx, zz, gc, gr, m in dev:
        by_zone.setdefault("%d,%d" % (zx, zz), []).append((gc, gr, round(m, 4)))
    res[name] = {"columns": len(dev), "by_zone": {k: {"n": len(v), "grid_cols": sorted({a for a, _, _ in v})[:12], "grid_rows": sorted({b for _, b, _ in v})[:12],
                                                      "max_abs": max(c for _, _, c in v)} for k, v in by_zone.items()}}
res["ref_asc1_vs_asc2"] = len(deviations(runs[0], runs[1]))
res["ref_asc_vs_desc"] = len(deviations(runs[0], runs[2]))
print(json.dumps(res, indent=1))
