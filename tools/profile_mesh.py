#!/usr/bin/env python3
"""Developer tool (GPU box): one mesh pass over a filled 16x16-chunk region, for ncu captures of k_mesh_count / k_mesh_emit."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mmgen_loader  # noqa: E402

mm = mmgen_loader.load()
gen = mm.ChunkGen(0)
S = 16
world = gen.region_world(0, 0, S, S)
world.generate(mm.STAGE_ALL)
coords = np.array([[x, z] for z in range(S) for x in range(S)], np.int32)
for _ in range(2):
    counts = world.mesh(coords, download=False)
print("mesh ms", world.mesh_ms(), "verts", int(counts[:, 0].sum()))
