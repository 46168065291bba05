"""TEST INFRASTRUCTURE. ctypes binding of oracle/_ref/libmmref_cuda.so: the UNMODIFIED reference
pipeline (/root/reference/src/terrain/chunk.cu) built for sm_100 by oracle/Makefile and driven by
oracle/refcuda_driver.cu. Needs a GPU. Used (a) on the GPU box as the binding parity oracle and to
produce tests/golden/, (b) by bench.py as the 'reference CUDA on the same B200' baseline."""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, "_ref", "libmmref_cuda.so")
# the same driver and reference state machine, but with integration/chunk_adapter.cpp linked over the
# reference's five generation entry points, i.e. the reference application running on libmmgen.so
ADAPTER_LIB = os.path.join(_HERE, "_ref", "libmmref_adapter.so")

CaveLayer = np.dtype([("start", "<i4"), ("end", "<i4"), ("bottomBiome", "u1"), ("topBiome", "u1"), ("pad", "u1", (2,))])
FeaturePlacement = np.dtype([("feature", "u1"), ("pad0", "u1", (3,)), ("x", "<i4"), ("y", "<i4"), ("z", "<i4"),
                             ("canReplaceBlocks", "u1"), ("pad1", "u1", (3,))])
CaveFeaturePlacement = np.dtype([("feature", "u1"), ("pad0", "u1", (3,)), ("x", "<i4"), ("y", "<i4"), ("z", "<i4"),
                                 ("layerHeight", "<i4"), ("canReplaceBlocks", "u1"), ("pad1", "u1", (3,))])
UNWRITTEN = 0x7FC0DEAD  # NaN payload written over dev_layers before generateLayers (chunk.cu:387-390 leaves holes)


def available():
    return os.path.exists(LIB)


def adapter_available():
    return os.path.exists(ADAPTER_LIB)


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


class RefCuda:
    def __init__(self, device=0, adapter=False):
        self.L = ctypes.CDLL(ADAPTER_LIB if adapter else LIB)
        self.L.mmref_stage_ms.restype = ctypes.c_double
        rc = self.L.mmref_init(device)
        if rc != 0:
            raise RuntimeError("mmref_init failed: %d" % rc)
        assert self.L.mmref_sizeof_feature() == 20 and self.L.mmref_sizeof_cave_feature() == 24

    def generate(self, x0, z0, nx, nz, last_stage=6):
        rc = self.L.mmref_generate(x0, z0, nx, nz, last_stage, ctypes.c_uint(UNWRITTEN))
        if rc != 0:
            raise RuntimeError("mmref_generate failed: %d" % rc)
        n = nx * nz
        L = self.L
        out = {"stage": np.array([L.mmref_stage(i) for i in range(n)], np.uint8).reshape(nz, nx),
               "ms": np.array([L.mmref_stage_ms(s) for s in range(8)])}
        st = out["stage"].ravel()
        h = np.zeros((n, 256), np.float32)
        w = np.zeros((n, 24, 256), np.float32)
        for i in range(n):
            L.mmref_get_heightfield(i, _ptr(h[i]))
            L.mmref_get_biome_weights(i, _ptr(w[i]))
        out["heightfield"], out["biome_weights"] = h, w
        if last_stage >= 2:
            l = np.zeros((n, 20, 256), np.float32)
            for i in np.nonzero(st >= 2)[0]:
                L.mmref_get_layers(int(i), _ptr(l[i]))
            out["layers"] = l
        if last_stage >= 4:
            idx = np.nonzero(st >= 4)[0]
            c = np.zeros((len(idx), 256, 32), CaveLayer)
            for k, i in enumerate(idx):
                L.mmref_get_cave_layers(int(i), _ptr(c[k]))
            out["cave_idx"], out["cave_layers"] = idx, c
        if last_stage >= 5:
            idx = np.nonzero(st >= 5)[0]
            fl, cl = [], []
            for i in idx:
                nf, nc = L.mmref_num_features(int(i)), L.mmref_num_cave_features(int(i))
                f = np.zeros(nf, FeaturePlacement)
                cf = np.zeros(nc, CaveFeaturePlacement)
                if nf:
                    L.mmref_get_features(int(i), _ptr(f))
                if nc:
                    L.mmref_get_cave_features(int(i), _ptr(cf))
                fl.append(f)
                cl.append(cf)
            out["feat_idx"], out["features"], out["cave_features"] = idx, fl, cl
        if last_stage >= 6:
            idx = np.nonzero(st >= 6)[0]
            b = np.zeros((len(idx), 16, 16, 384), np.uint8)
            gf, gc = [], []
            for k, i in enumerate(idx):
                L.mmref_get_blocks(int(i), _ptr(b[k]))
                nf, nc = L.mmref_num_gathered_features(int(i)), L.mmref_num_gathered_cave_features(int(i))
                f = np.zeros(nf, FeaturePlacement)
                cf = np.zeros(nc, CaveFeaturePlacement)
                if nf:
                    L.mmref_get_gathered_features(int(i), _ptr(f))
                if nc:
                    L.mmref_get_gathered_cave_features(int(i), _ptr(cf))
                gf.append(f)
                gc.append(cf)
            out["block_idx"], out["blocks"], out["gathered_features"], out["gathered_cave_features"] = idx, b, gf, gc
        return out


VERTEX = np.dtype([("pos", "<f4", (3,)), ("nor", "<f4", (3,)), ("uv", "<f4", (2,)), ("m", "<u8")])     # rendering/structs.hpp:25-31
_mesh_lib = None


def mesh_chunk(cx, cz, centre, neighbours):
    """The reference's own Chunk::createVBOs (chunk.cu:1781-2003) on given block volumes; needs no GPU.
    centre: uint8[16][16][384]; neighbours: 4 volumes or None in Chunk::neighbors order (+z, +x, -z, -x).
    Returns (verts as VERTEX records, idx uint32)."""
    global _mesh_lib
    if _mesh_lib is None:
        _mesh_lib = ctypes.CDLL(LIB)
        assert _mesh_lib.mmref_vertex_size() == VERTEX.itemsize
    blocks5 = np.zeros((5, 98304), np.uint8)
    blocks5[0] = np.ascontiguousarray(centre, np.uint8).ravel()
    present = np.zeros(4, np.int32)
    for i, nb in enumerate(neighbours):
        if nb is not None:
            blocks5[i + 1] = np.ascontiguousarray(nb, np.uint8).ravel()
            present[i] = 1
    cap = 1 << 20
    verts = np.zeros(cap, VERTEX)
    idx = np.zeros(cap * 3 // 2, np.uint32)
    nidx = ctypes.c_int(0)
    nv = _mesh_lib.mmref_mesh_chunk(int(cx), int(cz), _ptr(blocks5), _ptr(present), _ptr(verts), cap, _ptr(idx), len(idx), ctypes.byref(nidx))
    if nv < 0:
        raise RuntimeError("mesh larger than the oracle's buffers")
    return verts[:nv].copy(), idx[:nidx.value].copy()
