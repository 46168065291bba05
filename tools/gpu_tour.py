#!/usr/bin/env python3
"""Developer tool (GPU box): full six-stage comparison reference-CUDA vs oracle vs product on one
26x26-chunk window per surface biome (windows chosen where that biome has weight 1 over a whole
chunk; found with the oracle's stage 1). Prints one summary block per window."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mmgen_loader  # noqa: E402
from oracle import oracle as orc, refcuda  # noqa: E402

BIOME_CHUNKS = {0: (-2400, -864), 1: (-2400, -1968), 2: (-2400, -1344), 3: (-2400, 1536), 4: (-2400, -816), 5: (-2400, -912),
                6: (-2112, -1200), 7: (-2256, 1632), 8: (-2400, 672), 9: (-2400, -2352), 10: (-2352, -1776), 11: (-2400, 768),
                12: (-2400, -1776), 13: (-2400, 1008), 14: (-2400, 336), 15: (-2400, -1152), 16: (-2400, 1488), 17: (-2400, -2400),
                18: (-2400, -1536), 19: (-2400, -2256), 20: (-2352, 672), 21: (-2352, -576), 22: (-2400, -240), 23: (-2400, -2208)}
NX = NZ = 26
FIELDS = ("feature", "x", "y", "z", "canReplaceBlocks")


def bitdiff(a, b):
    return int((np.asarray(a).view(np.uint32) != np.asarray(b).view(np.uint32)).sum())


def lists_equal(A, B, cave=False):
    bad = 0
    for a, b in zip(A, B):
        if len(a) != len(b):
            bad += 1
            continue
        for f in FIELDS + (("layerHeight",) if cave else ()):
            if not np.array_equal(a[f] != 0 if f == "canReplaceBlocks" else a[f], b[f] != 0 if f == "canReplaceBlocks" else b[f]):
                bad += 1
                break
    return bad


def run_window(ref, o, gen, mm, biome, cx, cz):
    zx, zz = (cx // 12) * 12, (cz // 12) * 12
    x0, z0 = zx - 7, zz - 7
    origins = np.array([[(x0 + x) * 16, (z0 + z) * 16] for z in range(NZ) for x in range(NX)], np.int32)
    r = ref.generate(x0, z0, NX, NZ, 6)
    st = r["stage"].ravel()
    out = ["biome %2d window chunks [%d,%d)x[%d,%d) ref stage ms %s" % (biome, x0, x0 + NX, z0, z0 + NZ, np.round(r["ms"][1:], 1))]
    dom = np.bincount(r["biome_weights"].argmax(axis=1).ravel(), minlength=24)
    out.append("   dominant-biome columns: " + " ".join("%d:%d" % (b, c) for b, c in enumerate(dom) if c))
    oh, ow = o.heightfields(origins)
    world = gen.world(x0, z0, NX, NZ)
    world.generate(mm.STAGE_ALL)
    wd = world.download(heightfield=True, biome_weights=True, layers=True, cave_layers=True, blocks=True)
    out.append("   S1 height bitdiff oracle=%d product=%d | weights oracle=%d product=%d" % (
        bitdiff(oh, r["heightfield"]), bitdiff(wd["heightfield"], r["heightfield"]), bitdiff(ow, r["biome_weights"]),
        bitdiff(wd["biome_weights"], r["biome_weights"])))
    d = np.abs(oh.astype(np.float64) - r["heightfield"]) / np.abs(r["heightfield"])
    if d.max() > 0:
        out.append("      oracle height max rel %.3g, n(>1e-5)=%d" % (d.max(), int((d > 1e-5).sum())))
    # S2 on ring
    s2 = np.nonzero(st == 2)[0]
    h18 = orc.gather_h18(r["heightfield"], NX, NZ)
    inner = sorted(h18.keys())
    ol = np.full((NX * NZ, 20, 256), np.nan, np.float32)
    ol[inner] = o.layers(origins[inner], np.stack([h18[i] for i in inner]), r["biome_weights"][inner])
    written = r["layers"][s2].view(np.uint32) != refcuda.UNWRITTEN
    out.append("   S2 layers bitdiff oracle=%d product=%d (of %d written)" % (
        bitdiff(ol[s2][written], r["layers"][s2][written]), bitdiff(wd["layers"][s2][written], r["layers"][s2][written]), int(written.sum())))
    # S3
    s3 = np.nonzero(st >= 3)[0]
    lx0, lz0 = zx - 6 - x0, zz - 6 - z0
    planes = orc.gather_zone(ol, r["heightfield"], NX, lx0, lz0)
    er, sweeps = o.erode_zone(planes)
    orc.scatter_zone(er, ol, NX, lx0, lz0)
    dd = np.abs(ol[s3][:, 10:].astype(np.float64) - r["layers"][s3][:, 10:])
    out.append("   S3 eroded+backward bitdiff oracle=%d product=%d max abs %.3g (oracle sweeps %d, product %d)" % (
        bitdiff(ol[s3][:, 10:], r["layers"][s3][:, 10:]), bitdiff(wd["layers"][s3][:, 10:], r["layers"][s3][:, 10:]), dd.max(),
        sweeps, world.erosion_sweeps()))
    # S4
    cidx, rc = r["cave_idx"], r["cave_layers"]
    oc = o.caves(origins[cidx], r["heightfield"][cidx], r["biome_weights"][cidx])
    wc = wd["cave_layers"][cidx]
    out.append("   S4 cave layer fields differing oracle=%s product=%s" % (
        [int((oc[f] != rc[f]).sum()) for f in ("start", "end", "bottomBiome", "topBiome")],
        [int((wc[f] != rc[f]).sum()) for f in ("start", "end", "bottomBiome", "topBiome")]))
    # S5a (inputs = reference's own layers / caves so stages are judged independently)
    oF, oCF = o.feature_placements(origins[cidx], r["heightfield"][cidx], r["biome_weights"][cidx], r["layers"][cidx], rc)
    wF, wCF = world.download_features()
    out.append("   S5 lists differing oracle=(%d,%d) product=(%d,%d) of %d chunks; entries %d / %d" % (
        lists_equal(oF, r["features"]), lists_equal(oCF, r["cave_features"], True),
        lists_equal([wF[i] for i in cidx], r["features"]), lists_equal([wCF[i] for i in cidx], [c[:4096] for c in r["cave_features"]], True),
        len(cidx), sum(len(x) for x in r["features"]), sum(len(x) for x in r["cave_features"])))
    feats = np.bincount(np.concatenate([x["feature"] for x in r["features"]] + [np.zeros(0, np.uint8)]), minlength=21)
    out.append("      surface features by type: " + " ".join("%d:%d" % (f, c) for f, c in enumerate(feats) if c))
    # S6
    bidx, rb = r["block_idx"], r["blocks"]
    pos5 = {int(c): k for k, c in enumerate(cidx)}
    ksel = np.array([pos5[int(c)] for c in bidx])
    ob = o.fill(origins[bidx], r["heightfield"][bidx], r["biome_weights"][bidx], r["layers"][bidx], rc[ksel],
                r["gathered_features"], r["gathered_cave_features"])
    wb = wd["blocks"][bidx]
    for name, a in (("oracle", ob), ("product(world)", wb)):
        dmask = a != rb
        line = "   S6 %-14s block mismatches=%d of %d" % (name, int(dmask.sum()), dmask.size)
        if dmask.sum():
            pairs = {}
            for x, y in zip(a[dmask][:100000], rb[dmask][:100000]):
                pairs[(int(x), int(y))] = pairs.get((int(x), int(y)), 0) + 1
            line += " (got,ref): " + str(sorted(pairs.items(), key=lambda kv: -kv[1])[:8])
        out.append(line)
    world.close()
    return "\n".join(out)


def main():
    biomes = [int(a) for a in sys.argv[1:]] or sorted(BIOME_CHUNKS)
    ref = refcuda.RefCuda(0)
    o = orc.Oracle()
    mm = mmgen_loader.load()
    gen = mm.ChunkGen(0)
    for b in biomes:
        t = time.time()
        print(run_window(ref, o, gen, mm, b, *BIOME_CHUNKS[b]), flush=True)
        print("   (%.1fs)" % (time.time() - t), flush=True)


if __name__ == "__main__":
    main()
