#!/usr/bin/env python3
"""Developer tool (GPU box): throughput of the device mesher (mmgen_world_mesh) on a filled 32x32-chunk region, next to the
reference's host createVBOs (oracle/_ref, one thread, bounded sample) on the same block volumes. One JSON line."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mmgen_loader  # noqa: E402
from oracle import refcuda  # noqa: E402

mm = mmgen_loader.load()
gen = mm.ChunkGen(0)
S = int(sys.argv[1]) if len(sys.argv) > 1 else 32
world = gen.region_world(0, 0, S, S)
world.generate(mm.STAGE_ALL)
coords = np.array([[x, z] for z in range(S) for x in range(S)], np.int32)
best = 1e9
for _ in range(5):
    counts = world.mesh(coords, download=False)
    best = min(best, world.mesh_ms())
nv = int(counts[:, 0].sum())
out = {"chunks": S * S, "vertices": nv, "device_ms": round(best, 3), "chunks_per_s": round(S * S / best * 1e3, 1),
       "bytes_written": nv * 40 + nv // 4 * 24, "bytes_read_blocks": S * S * 98304,
       "hbm_gbs": round((nv * 40 + nv // 4 * 24 + S * S * 98304) / best / 1e6, 1)}
if refcuda.available():
    blocks = world.download_region_blocks().reshape(S, S, 16, 16, 384)
    t0 = time.perf_counter()
    n = 0
    for z in range(1, 5):
        for x in range(1, 9):
            refcuda.mesh_chunk(x, z, blocks[z, x], [blocks[z + 1, x], blocks[z, x + 1], blocks[z - 1, x], blocks[z, x - 1]])
            n += 1
    dt = time.perf_counter() - t0
    out["reference_host_createVBOs"] = {"chunks": n, "seconds": round(dt, 3), "chunks_per_s": round(n / dt, 1), "threads": 1,
                                        "note": "incl. ~0.1 ms per call of ctypes marshalling of 5 block volumes"}
print(json.dumps(out))
