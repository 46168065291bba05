#!/usr/bin/env python3
"""Developer tool (GPU box): pins for tests/test_region_hashes.py, produced by the REFERENCE.

For every region the unmodified reference CUDA pipeline (oracle/_ref/libmmref_cuda.so) generates the apron window the
region needs, the block volumes of the region's chunks are hashed on the host with the same function the product computes
on the device (mmgen_world_chunk_hash_sum: per-column 64-bit FNV-1a, then FNV-1a over (cx, cz, 256 column hashes), summed
mod 2^64), and the product's hash of the same region is printed beside it. Writes tests/golden/region_hashes.json when
called with --write (run after a deliberate change of the reference build only)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mmgen_loader  # noqa: E402
from oracle import refcuda  # noqa: E402

REGIONS = ((0, 0, 48, 48), (-300, 500, 32, 32), (4000, -4000, 24, 36))
FNV_OFF, FNV_PRIME = np.uint64(14695981039346656037), np.uint64(1099511628211)


def fnv_bytes(h, values_u64):
    """h ^= byte; h *= prime for the 8 little-endian bytes of each value of values_u64 (vectorised over the leading axes)."""
    for b in range(8):
        h = (h ^ ((values_u64 >> np.uint64(8 * b)) & np.uint64(0xff))) * FNV_PRIME
    return h


def chunk_hashes(blocks, coords):
    """blocks: uint8[n][16][16][384]; coords: (n, 2) chunk coordinates. The product's per-chunk hash (mmgen_world_chunk_hashes)."""
    n = blocks.shape[0]
    cols = blocks.reshape(n, 256, 384).astype(np.uint64)
    with np.errstate(over="ignore"):
        h = np.full((n, 256), FNV_OFF, np.uint64)
        for y in range(384):
            h = (h ^ cols[:, :, y]) * FNV_PRIME
        c = np.full(n, FNV_OFF, np.uint64)
        c = fnv_bytes(c, coords[:, 0].astype(np.int64).view(np.uint64))
        c = fnv_bytes(c, coords[:, 1].astype(np.int64).view(np.uint64))
        for k in range(256):
            c = fnv_bytes(c, h[:, k])
        return c


def chunk_hash_sum(blocks, coords):
    with np.errstate(over="ignore"):
        return int(np.add.reduce(chunk_hashes(blocks, coords), dtype=np.uint64))


def main():
    mm = mmgen_loader.load()
    from mega_minecraft_b200 import tiling
    gen = mm.ChunkGen(0)
    ref = refcuda.RefCuda(0)
    out, arrays = {}, {}
    for reg in REGIONS:
        rx0, rz0, rnx, rnz = reg
        cx0, cz0, wnx, wnz = tiling.apron_window(*reg)
        r = ref.generate(cx0, cz0, wnx, wnz, 6)
        bidx = r["block_idx"]
        coords = np.stack([cx0 + bidx % wnx, cz0 + bidx // wnx], axis=1)
        inside = (coords[:, 0] >= rx0) & (coords[:, 0] < rx0 + rnx) & (coords[:, 1] >= rz0) & (coords[:, 1] < rz0 + rnz)
        assert int(inside.sum()) == rnx * rnz, "the reference did not fill the whole region from the apron window"
        order = np.lexsort((coords[inside][:, 0], coords[inside][:, 1]))      # region raster order
        rb, rc = r["blocks"][inside][order], coords[inside][order]
        hs = chunk_hashes(rb, rc)
        with np.errstate(over="ignore"):
            href = int(np.add.reduce(hs, dtype=np.uint64))
        w = gen.region_world(*reg)
        w.generate(mm.STAGE_ALL)
        w.sync()
        hown = w.chunk_hash_sum()
        pb = w.download_region_blocks()
        wdl = w.download(layers=True)
        assert (w.cx0, w.cz0, w.nx, w.nz) == (cx0, cz0, wnx, wnz)
        w.close()
        flips = []
        diff = np.argwhere(pb != rb)
        if len(diff):
            # where a block differs: both sides' layer starts of the column (terrain thresholds) and what the oracle and its two
            # FMA-variant builds (rasteriser FMAs rounded twice / everything contracted) say - see tests/test_reference_tour.py
            from oracle import oracle as orc
            variants = {"oracle": orc.Oracle(), "oracle_unfused": orc.Oracle(variant="unfused"), "oracle_contract": orc.Oracle(variant="contract")}
            pl = wdl["layers"]
            sel = np.nonzero(inside)[0][order]                      # index into the reference's filled-chunk arrays
            pos5 = {int(c): i for i, c in enumerate(r["cave_idx"])}
            for k in sorted({int(f[0]) for f in diff}):
                kk = int(sel[k])
                c = int(bidx[kk])                                   # window raster index of the chunk
                origin = np.array([[(cx0 + c % wnx) * 16, (cz0 + c // wnx) * 16]], np.int32)
                ob = {v: o.fill(origin, r["heightfield"][c:c + 1], r["biome_weights"][c:c + 1], r["layers"][c:c + 1],
                                r["cave_layers"][pos5[c]:pos5[c] + 1], [r["gathered_features"][kk]], [r["gathered_cave_features"][kk]])[0]
                      for v, o in variants.items()}
                for _, z, x, y in diff[diff[:, 0] == k]:
                    col = int(x + 16 * z)
                    rec = {"x": int(rc[k, 0] * 16 + x), "y": int(y), "z": int(rc[k, 1] * 16 + z), "chunk": [int(rc[k, 0]), int(rc[k, 1])],
                           "product": int(pb[k, z, x, y]), "reference": int(rb[k, z, x, y]), "height": float(r["heightfield"][c, col]),
                           "layer_starts_product": [float(v) for v in pl[c, :, col]], "layer_starts_reference": [float(v) for v in r["layers"][c, :, col]]}
                    for v in variants:
                        rec[v] = int(ob[v][z, x, y])
                    # the terrain threshold the voxel straddles: the layer start closest to y on either side
                    a, b = np.array(rec["layer_starts_product"]), np.array(rec["layer_starts_reference"])
                    l = int(np.argmin(np.minimum(np.abs(a - y), np.abs(b - y))))
                    rec["nearest_layer"] = {"material": l, "product_start": float(a[l]), "reference_start": float(b[l]),
                                            "straddles_y": bool((a[l] <= y) != (b[l] <= y))}
                    flips.append(rec)
        print(reg, "reference %016x product %016x block flips %d of %d" % (href, hown, len(flips), pb.size), flush=True)
        for f in flips:
            print("   ", f, flush=True)
        key = "%d,%d,%d,%d" % reg
        out[key] = {"reference_hash_sum": "%016x" % href, "voxels": int(pb.size), "product_flips_when_pinned": flips}
        arrays[key] = hs
    if "--write" in sys.argv:
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", "region_hashes.npz"), **arrays)
        with open(os.path.join(ROOT, "tests", "golden", "region_hashes.json"), "w") as f:
            json.dump({"made_by": "tools/region_hashes.py --write: the unmodified reference chunk.cu (sm_100 build, oracle/_ref) on a B200; "
                                  "region_hashes.npz holds its per-chunk hashes in region raster order", "regions": out}, f, indent=1)
    os.makedirs(os.path.join(ROOT, "gpurun_out", "golden"), exist_ok=True)
    np.savez_compressed(os.path.join(ROOT, "gpurun_out", "golden", "region_hashes.npz"), **arrays)
    with open(os.path.join(ROOT, "gpurun_out", "golden", "region_hashes.json"), "w") as f:
        json.dump({"made_by": "tools/region_hashes.py --write: the unmodified reference chunk.cu (sm_100 build, oracle/_ref) on a B200; "
                              "region_hashes.npz holds its per-chunk hashes in region raster order", "regions": out}, f, indent=1)


if __name__ == "__main__":
    main()
