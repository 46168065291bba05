#!/usr/bin/env python3
"""Developer tool (GPU box): runs the reference CUDA pipeline (oracle/_ref), the CPU oracle and the
product on the C2 window and reports agreement stage by stage; saves the reference outputs under
gpurun_out/ref_c2/ so golden fixtures can be cut from them."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mmgen_loader  # noqa: E402
from oracle import oracle as orc, refcuda  # noqa: E402

X0, Z0, NX, NZ = -7, -7, 26, 26


def report(name, a, b):
    a = np.asarray(a); b = np.asarray(b)
    bits = (a.view(np.uint32) != b.view(np.uint32)) if a.dtype == np.float32 else (a != b)
    rel = np.abs(a.astype(np.float64) - b) / np.maximum(np.abs(b.astype(np.float64)), 1e-30) if a.dtype == np.float32 else None
    msg = "%-28s n=%d bit-different=%d (%.3g)" % (name, a.size, int(bits.sum()), bits.mean())
    if rel is not None:
        msg += " max_rel=%.3g n(rel>1e-5)=%d" % (np.nanmax(rel), int((rel > 1e-5).sum()))
    print(msg, flush=True)
    return bits


def main():
    last = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    outdir = os.path.join(ROOT, "gpurun_out", "ref_c2")
    os.makedirs(outdir, exist_ok=True)
    origins = np.array([[(X0 + x) * 16, (Z0 + z) * 16] for z in range(NZ) for x in range(NX)], np.int32)

    t = time.time()
    ref = refcuda.RefCuda(0).generate(X0, Z0, NX, NZ, last)
    print("reference CUDA: %.2fs, per-stage wall ms %s" % (time.time() - t, np.round(ref["ms"], 1)), flush=True)
    np.save(os.path.join(outdir, "heightfield.npy"), ref["heightfield"])
    np.save(os.path.join(outdir, "biome_weights.npy"), ref["biome_weights"].astype(np.float32))

    o = orc.Oracle()
    t = time.time()
    oh, ow = o.heightfields(origins)
    print("oracle S1: %.2fs on %d threads" % (time.time() - t, o.nthreads), flush=True)

    mm = mmgen_loader.load()
    gen = mm.ChunkGen(0)
    mh, mw = gen.heightfields(origins)

    print("--- S1 ---")
    bits = report("oracle vs ref  height", oh, ref["heightfield"])
    report("oracle vs ref  weights", ow, ref["biome_weights"])
    report("product vs ref height", mh, ref["heightfield"])
    report("product vs ref weights", mw, ref["biome_weights"])
    report("product vs oracle height", mh, oh)
    report("product vs oracle weights", mw, ow)
    if last >= 2:
        print("--- S2 ---")
        st = ref["stage"].ravel()
        idx = np.nonzero(st >= 2)[0]
        rl = ref["layers"]
        np.save(os.path.join(outdir, "layers.npy"), rl)
        np.save(os.path.join(outdir, "stage.npy"), ref["stage"])
        h18 = orc.gather_h18(ref["heightfield"], NX, NZ)
        H18 = np.stack([h18[i] for i in idx])
        ol = o.layers(origins[idx], H18, ref["biome_weights"][idx])
        ml = gen.layers(origins[idx], H18, ref["biome_weights"][idx])
        world = gen.world(X0, Z0, NX, NZ)
        world.generate(mm.STAGE_HEIGHTFIELD | mm.STAGE_LAYERS | (mm.STAGE_EROSION if last >= 3 else 0) | (mm.STAGE_CAVES if last >= 4 else 0) | (mm.STAGE_FEATURES if last >= 5 else 0) | (mm.STAGE_FILL if last >= 6 else 0))
        wd = world.download(heightfield=True, layers=True, cave_layers=(last >= 4), blocks=(last >= 6))
        ws = world.stages().ravel()
        print("world stages:", np.bincount(ws), "ref stages:", np.bincount(st), "erosion sweeps", world.erosion_sweeps(), "stage ms", world.stage_ms())
        # S2 comparison on entries the reference wrote (sentinel = NaN payload) and before erosion rewrote them
        s2 = np.nonzero(st == 2)[0]
        pos = {int(c): k for k, c in enumerate(idx)}
        sel = np.array([pos[int(c)] for c in s2])
        refl = rl[s2]
        written = refl.view(np.uint32) != refcuda.UNWRITTEN
        print("written fraction of S2 entries: %.4f" % written.mean())
        report("oracle vs ref  layers(S2)", ol[sel][written], refl[written])
        report("product vs ref layers(S2)", ml[sel][written], refl[written])
        report("product vs oracle layers(S2)", ml[sel][written], ol[sel][written])
        report("world vs ref   layers(S2)", wd["layers"][s2][written], refl[written])
        for l in range(20):
            wl = written[:, l]
            if wl.sum():
                d = (ol[sel][:, l][wl].view(np.uint32) != refl[:, l][wl].view(np.uint32)).sum()
                print("   layer %2d written=%d oracle-vs-ref bitdiff=%d" % (l, wl.sum(), d))
    if last >= 3:
        print("--- S3 ---")
        s3 = np.nonzero(st >= 3)[0]
        # oracle erosion from the oracle's own S2 layers of the window around zone (0,0)
        full = np.full((NX * NZ, 20, 256), np.nan, np.float32)
        full[idx] = ol
        lx0, lz0 = 0 - 6 - X0, 0 - 6 - Z0
        planes = orc.gather_zone(full, ref["heightfield"], NX, lx0, lz0)
        t = time.time()
        er, sweeps = o.erode_zone(planes)
        print("oracle erosion: %d sweeps %.2fs" % (sweeps, time.time() - t))
        orc.scatter_zone(er, full, NX, lx0, lz0)
        mer, msweeps = gen.erode_zone(planes)
        report("product vs oracle eroded planes", mer, er[:8])
        refl = rl[s3]
        report("oracle vs ref  eroded loose layers", full[s3][:, 12:], refl[:, 12:])
        report("oracle vs ref  backward layers", full[s3][:, 10:12], refl[:, 10:12])
        report("world vs ref   eroded loose layers", wd["layers"][s3][:, 12:], refl[:, 12:])
        report("world vs oracle eroded loose", wd["layers"][s3][:, 12:], full[s3][:, 12:])
        report("world vs oracle backward", wd["layers"][s3][:, 10:12], full[s3][:, 10:12])
        d = np.abs(full[s3][:, 12:].astype(np.float64) - refl[:, 12:])
        print("   max abs diff oracle-vs-ref eroded: %.6g ; columns differing: %d of %d" % (d.max(), int((d.max(axis=1) > 0).sum()), d.shape[0] * 256))
    if last >= 4:
        print("--- S4 ---")
        cidx = ref["cave_idx"]
        rc = ref["cave_layers"]
        np.save(os.path.join(outdir, "cave_idx.npy"), cidx)
        np.save(os.path.join(outdir, "cave_layers.npy"), rc)
        t = time.time()
        oc = o.caves(origins[cidx], ref["heightfield"][cidx], ref["biome_weights"][cidx])
        print("oracle caves: %d chunks %.2fs on %d threads" % (len(cidx), time.time() - t, o.nthreads), flush=True)
        mc = gen.caves(origins[cidx], ref["heightfield"][cidx], ref["biome_weights"][cidx])
        wc = wd["cave_layers"][cidx]
        for name, a, b in (("oracle vs ref", oc, rc), ("product vs ref", mc, rc), ("product vs oracle", mc, oc), ("world vs ref", wc, rc)):
            for f in ("start", "end", "bottomBiome", "topBiome"):
                d = a[f] != b[f]
                print("  %-18s %-12s different=%d of %d ; columns affected=%d" % (name, f, int(d.sum()), d.size, int(d.any(axis=2).sum())))
        nl = (rc["start"] != 384).sum(axis=2)
        print("  ref layers/column avg %.3f max %d" % (nl.mean(), nl.max()))
    def cmp_lists(name, A, B):
        nd = sum(1 for a, b in zip(A, B) if len(a) != len(b) or (a.tobytes() != b.tobytes()))
        print("  %-34s chunks=%d lists differing=%d ; entries %d vs %d" % (name, len(A), nd, sum(len(a) for a in A), sum(len(b) for b in B)))
        if nd:
            for k, (a, b) in enumerate(zip(A, B)):
                if len(a) != len(b) or a.tobytes() != b.tobytes():
                    m = min(len(a), len(b))
                    j = next((i for i in range(m) if a[i].tobytes() != b[i].tobytes()), m)
                    print("     first diff: list %d len %d vs %d at entry %d: %s | %s" % (k, len(a), len(b), j, a[j] if j < len(a) else None, b[j] if j < len(b) else None))
                    break
    if last >= 5:
        print("--- S5a ---")
        fidx = ref["feat_idx"]
        assert (fidx == cidx).all()
        rl5 = ref["layers"][fidx]
        oF, oCF = o.feature_placements(origins[fidx], ref["heightfield"][fidx], ref["biome_weights"][fidx], rl5, rc)
        mF, mCF = gen.feature_placements(origins[fidx], ref["heightfield"][fidx], ref["biome_weights"][fidx], rl5, rc)
        wF, wCF = world.download_features()
        cmp_lists("oracle vs ref surface features", oF, ref["features"])
        cmp_lists("oracle vs ref cave features", oCF, ref["cave_features"])
        cmp_lists("product vs ref surface features", mF, ref["features"])
        cmp_lists("product vs ref cave features", mCF, ref["cave_features"])
        cmp_lists("world vs ref surface features", [wF[i] for i in fidx], ref["features"])
        cmp_lists("world vs ref cave features", [wCF[i] for i in fidx], [c[:4096] for c in ref["cave_features"]])
        import pickle
        pickle.dump({"feat_idx": fidx, "features": ref["features"], "cave_features": ref["cave_features"]}, open(os.path.join(outdir, "features.pkl"), "wb"))
    if last >= 6:
        print("--- S6 ---")
        bidx = ref["block_idx"]
        rb = ref["blocks"]
        np.save(os.path.join(outdir, "block_idx.npy"), bidx)
        np.save(os.path.join(outdir, "blocks.npy"), rb)
        pickle.dump({"gf": ref["gathered_features"], "gcf": ref["gathered_cave_features"]}, open(os.path.join(outdir, "gathered.pkl"), "wb"))
        pos5 = {int(c): k for k, c in enumerate(fidx)}
        lists = {int(c): ref["features"][k] for k, c in enumerate(fidx)}
        clists = {int(c): ref["cave_features"][k] for k, c in enumerate(fidx)}
        g = [orc.gather_features(lists, int(c) % NX, int(c) // NX, NX) for c in bidx]
        gc = [orc.gather_features(clists, int(c) % NX, int(c) // NX, NX) for c in bidx]
        cmp_lists("test-side gather vs ref gathered", g, ref["gathered_features"])
        cmp_lists("test-side gather vs ref gathered cave", gc, ref["gathered_cave_features"])
        ksel = np.array([pos5[int(c)] for c in bidx])
        t = time.time()
        ob = o.fill(origins[bidx], ref["heightfield"][bidx], ref["biome_weights"][bidx], ref["layers"][bidx], rc[ksel], g, gc)
        print("oracle fill: %d chunks %.2fs" % (len(bidx), time.time() - t), flush=True)
        t = time.time()
        mb = gen.fill(origins[bidx], ref["heightfield"][bidx], ref["biome_weights"][bidx], ref["layers"][bidx], rc[ksel], g, gc)
        print("product fill (batch op): %.3fs" % (time.time() - t), flush=True)
        wb = wd["blocks"][bidx]
        names = None
        for name, a, b in (("oracle vs ref", ob, rb), ("product vs ref", mb, rb), ("product vs oracle", mb, ob), ("world vs ref", wb, rb)):
            d = a != b
            print("  %-18s block mismatches=%d of %d (%.3g)" % (name, int(d.sum()), d.size, d.mean()))
            if d.sum():
                pairs = {}
                for x, y in zip(a[d][:200000], b[d][:200000]):
                    pairs[(int(x), int(y))] = pairs.get((int(x), int(y)), 0) + 1
                top = sorted(pairs.items(), key=lambda kv: -kv[1])[:12]
                print("     (got,ref) counts:", top)
                w_ = np.argwhere(d)[:3]
                for ci, z, x, y in w_:
                    c = int(bidx[ci])
                    print("     e.g. chunk %d (cx=%d cz=%d) local x=%d z=%d y=%d got=%d ref=%d" % (c, X0 + c % NX, Z0 + c // NX, x, z, y, a[ci, z, x, y], b[ci, z, x, y]))
    # per dominant biome breakdown of oracle-vs-ref height mismatches
    dom = ref["biome_weights"].argmax(axis=1)
    single = (ref["biome_weights"].max(axis=1) == 1.0)
    for b in range(24):
        m = (dom == b) & single
        if m.sum():
            rel = np.abs(oh[m].astype(np.float64) - ref["heightfield"][m]) / ref["heightfield"][m]
            rel2 = np.abs(mh[m].astype(np.float64) - ref["heightfield"][m]) / ref["heightfield"][m]
            print("  biome %2d single-biome cols=%6d oracle: bitdiff=%6d maxrel=%.3g | product: bitdiff=%6d maxrel=%.3g" % (
                b, m.sum(), int((oh[m].view(np.uint32) != ref["heightfield"][m].view(np.uint32)).sum()), rel.max(),
                int((mh[m].view(np.uint32) != ref["heightfield"][m].view(np.uint32)).sum()), rel2.max()))


if __name__ == "__main__":
    main()
