#!/usr/bin/env python3
"""Cuts tests/golden/c2_window.npz from the reference-CUDA outputs that tools/gpu_compare.py saved
under gpurun_out/ref_c2/ on a B200 (the UNMODIFIED reference pipeline, oracle/_ref/libmmref_cuda.so,
window = chunks [-7,19)^2, i.e. zone (0,0) + its 6-chunk erosion pad + the 1-chunk layer border).

    gpurun -- python tools/gpu_compare.py 6     # on the GPU box; merges gpurun_out/ref_c2/*
    python tools/make_golden.py                 # here

Padding bytes of the placement structs are uninitialised in the reference and are zeroed."""
import os
import pickle

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "gpurun_out", "ref_c2")
DST = os.path.join(ROOT, "tests", "golden", "c2_window.npz")
X0, Z0, NX, NZ = -7, -7, 26, 26


def clean(a):
    a = a.copy()
    for f in ("pad0", "pad1"):
        a[f] = 0
    a["canReplaceBlocks"] = (a["canReplaceBlocks"] != 0).astype(np.uint8)
    return a


def pack(lists):
    off = np.zeros(len(lists) + 1, np.int64)
    off[1:] = np.cumsum([len(x) for x in lists])
    return np.concatenate([clean(x) for x in lists]) if off[-1] else np.zeros(0, lists[0].dtype), off


def main():
    h = np.load(os.path.join(SRC, "heightfield.npy"))
    w = np.load(os.path.join(SRC, "biome_weights.npy"))
    layers = np.load(os.path.join(SRC, "layers.npy"))
    stage = np.load(os.path.join(SRC, "stage.npy"))
    cave_idx = np.load(os.path.join(SRC, "cave_idx.npy"))
    caves = np.load(os.path.join(SRC, "cave_layers.npy"))
    caves = caves.copy()
    caves["pad"] = 0
    block_idx = np.load(os.path.join(SRC, "block_idx.npy"))
    blocks = np.load(os.path.join(SRC, "blocks.npy"))
    feats = pickle.load(open(os.path.join(SRC, "features.pkl"), "rb"))
    st = stage.ravel()
    ring = np.nonzero(st == 2)[0][::4]          # S2-only chunks (never eroded): every 4th
    zone = np.nonzero(st >= 3)[0]
    assert (zone == cave_idx).all() and (feats["feat_idx"] == cave_idx).all()
    f, foff = pack(feats["features"])
    cf, cfoff = pack(feats["cave_features"])
    np.savez_compressed(
        DST, window=np.array([X0, Z0, NX, NZ], np.int32), stage=stage, heightfield=h, biome_weights=w,
        ring_idx=ring.astype(np.int32), ring_layers=layers[ring], zone_idx=zone.astype(np.int32), zone_layers=layers[zone],
        cave_layers=caves, features=f, features_off=foff, cave_features=cf, cave_features_off=cfoff,
        block_idx=block_idx.astype(np.int32), blocks=blocks)
    print("wrote", DST, os.path.getsize(DST) / 1e6, "MB")


if __name__ == "__main__":
    main()
