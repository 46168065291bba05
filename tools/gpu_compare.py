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

X0, Z0, NX, NZ = -6, -6, 24, 24


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
