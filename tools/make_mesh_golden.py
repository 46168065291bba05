#!/usr/bin/env python3
"""Generates tests/golden/c2_mesh.npz: what the reference's own Chunk::createVBOs (chunk.cu:1781-2003, through
oracle/_ref/libmmref_cuda.so; a host function, no GPU needed) produces for the 36 block volumes of
tests/golden/c2_window.npz, each meshed with the neighbours that exist inside that 6x6 region (the outer ring
exercises the reference's null-neighbour rule). Stored: vertex / index counts and SHA-1 of the raw arrays for every
chunk, and the full arrays of one interior chunk. Run where /root/reference exists (after `make -C oracle ref`)."""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refcuda  # noqa: E402

g = np.load(os.path.join(ROOT, "tests", "golden", "c2_window.npz"))
x0, z0, nx, nz = (int(v) for v in g["window"])
bidx, blocks = g["block_idx"], g["blocks"]
pos = {int(i): k for k, i in enumerate(bidx)}
coords, counts, vsha, isha = [], [], [], []
full = None
for k, i in enumerate(bidx):
    cx, cz = x0 + int(i) % nx, z0 + int(i) // nx
    nbs = []
    for dx, dz in ((0, 1), (1, 0), (0, -1), (-1, 0)):        # Chunk::neighbors order: +z, +x, -z, -x
        j = (cz + dz - z0) * nx + (cx + dx - x0)
        nbs.append(blocks[pos[j]] if j in pos else None)
    v, ix = refcuda.mesh_chunk(cx, cz, blocks[k], nbs)
    coords.append((cx, cz)); counts.append((len(v), len(ix)))
    vsha.append(hashlib.sha1(v.tobytes()).hexdigest()); isha.append(hashlib.sha1(ix.tobytes()).hexdigest())
    if (cx, cz) == (5, 5):
        full = (v, ix)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "c2_mesh.npz"), coords=np.array(coords, np.int32), counts=np.array(counts, np.int32),
                    verts_sha1=np.array(vsha), idx_sha1=np.array(isha), full_coord=np.array([5, 5], np.int32), full_verts=full[0], full_idx=full[1])
print("chunks", len(coords), "verts", int(np.array(counts)[:, 0].sum()))
