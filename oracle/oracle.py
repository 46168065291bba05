"""TEST INFRASTRUCTURE. ctypes binding of oracle/libmmoracle.so (the CPU restatement)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libmmoracle.so")


def build():
    subprocess.run(["make", "-C", _HERE, "oracle", "variants"], check=True, stdout=subprocess.DEVNULL)
    return _LIB


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


class Oracle:
    def __init__(self, nthreads=None, variant=None):
        """variant: None = the oracle; "unfused" / "contract" = diagnostic builds with other FMA contraction (see Makefile),
        only for classifying block flips - never a parity reference."""
        lib = _LIB if variant is None else os.path.join(_HERE, "libmmoracle_%s.so" % variant)
        if not os.path.exists(lib):
            build()
        self.L = ctypes.CDLL(lib)
        self.nthreads = nthreads or os.cpu_count() or 1
        for f in ("mmo_sinf", "mmo_cosf", "mmo_simplex2", "mmo_simplex3", "mmo_powf", "mmo_rng3_u01"):
            getattr(self.L, f).restype = ctypes.c_float
        self.L.mmo_sinf.argtypes = [ctypes.c_float]
        self.L.mmo_cosf.argtypes = [ctypes.c_float]
        self.L.mmo_powf.argtypes = [ctypes.c_float, ctypes.c_float]
        self.L.mmo_simplex2.argtypes = [ctypes.c_float, ctypes.c_float]
        self.L.mmo_simplex3.argtypes = [ctypes.c_float, ctypes.c_float, ctypes.c_float]
        self.L.mmo_hash.restype = ctypes.c_uint32
        self.L.mmo_hash.argtypes = [ctypes.c_uint32]

    def counters(self, reset=False):
        """Noise-primitive calls since the last reset: dict(simplex2, simplex3, sin, worley2_cells, worley3_cells)."""
        out = np.zeros(5, np.uint64)
        self.L.mmo_counters(_ptr(out), 1 if reset else 0)
        return dict(zip(("simplex2", "simplex3", "sin", "worley2_cells", "worley3_cells"), (int(v) for v in out)))

    @staticmethod
    def flops(c):
        """Canonical algorithmic FLOPs of SURVEY.md 8(d): simplex2 140, simplex3 330, Worley 45 / 60 per cell, sinf 20."""
        return 140 * c["simplex2"] + 330 * c["simplex3"] + 45 * c["worley2_cells"] + 60 * c["worley3_cells"] + 20 * c["sin"]

    def heightfields(self, origins):
        origins = np.ascontiguousarray(origins, dtype=np.int32).reshape(-1, 2)
        n = origins.shape[0]
        h = np.empty((n, 256), np.float32)
        w = np.empty((n, 24, 256), np.float32)
        self.L.mmo_heightfields(n, _ptr(origins), _ptr(h), _ptr(w), self.nthreads)
        return h, w

    def layers(self, origins, h18, weights, unwritten=np.nan):
        origins = np.ascontiguousarray(origins, dtype=np.int32).reshape(-1, 2)
        n = origins.shape[0]
        h18 = np.ascontiguousarray(h18, np.float32).reshape(n, 324)
        weights = np.ascontiguousarray(weights, np.float32).reshape(n, 24, 256)
        out = np.empty((n, 20, 256), np.float32)
        self.L.mmo_layers(n, _ptr(origins), _ptr(h18), _ptr(weights), _ptr(out), ctypes.c_float(unwritten), self.nthreads)
        return out

    def caves(self, origins, heightfield, weights):
        origins = np.ascontiguousarray(origins, dtype=np.int32).reshape(-1, 2)
        n = origins.shape[0]
        h = np.ascontiguousarray(heightfield, np.float32).reshape(n, 256)
        w = np.ascontiguousarray(weights, np.float32).reshape(n, 24, 256)
        from .refcuda import CaveLayer
        out = np.zeros((n, 256, 32), CaveLayer)
        self.L.mmo_caves(n, _ptr(origins), _ptr(h), _ptr(w), _ptr(out), self.nthreads)
        return out

    def set_cave_grid_test(self, honoured):
        """False (default): the reference as built here; True: the source-text reading (see mm_features.h)."""
        self.L.mmo_set_cave_grid_test(1 if honoured else 0)

    def feature_placements(self, origins, heightfield, weights, layers, cave_layers, max_per_chunk=4096):
        from .refcuda import FeaturePlacement, CaveFeaturePlacement
        origins = np.ascontiguousarray(origins, dtype=np.int32).reshape(-1, 2)
        n = origins.shape[0]
        F = np.zeros((n, max_per_chunk), FeaturePlacement)
        CF = np.zeros((n, max_per_chunk), CaveFeaturePlacement)
        counts = np.zeros((n, 2), np.int32)
        self.L.mmo_feature_placements(n, _ptr(origins), _ptr(np.ascontiguousarray(heightfield, np.float32)),
                                      _ptr(np.ascontiguousarray(weights, np.float32)), _ptr(np.ascontiguousarray(layers, np.float32)),
                                      _ptr(np.ascontiguousarray(cave_layers)), max_per_chunk, _ptr(F), _ptr(CF), _ptr(counts), self.nthreads)
        return [F[i, :min(counts[i, 0], max_per_chunk)].copy() for i in range(n)], \
               [CF[i, :min(counts[i, 1], max_per_chunk)].copy() for i in range(n)]

    def fill(self, origins, heightfield, weights, layers, cave_layers, gathered, gathered_cave, decorate=True):
        """gathered / gathered_cave: per-chunk lists of structured arrays (untruncated)."""
        from .refcuda import FeaturePlacement, CaveFeaturePlacement
        origins = np.ascontiguousarray(origins, dtype=np.int32).reshape(-1, 2)
        n = origins.shape[0]
        sf = max(1, max(len(g) for g in gathered))
        scf = max(1, max(len(g) for g in gathered_cave))
        F = np.zeros((n, sf), FeaturePlacement)
        CF = np.zeros((n, scf), CaveFeaturePlacement)
        counts = np.zeros((n, 2), np.int32)
        for i in range(n):
            F[i, :len(gathered[i])] = gathered[i]
            CF[i, :len(gathered_cave[i])] = gathered_cave[i]
            counts[i] = (len(gathered[i]), len(gathered_cave[i]))
        out = np.zeros((n, 16, 16, 384), np.uint8)
        self.L.mmo_fill(n, _ptr(origins), _ptr(np.ascontiguousarray(heightfield, np.float32)), _ptr(np.ascontiguousarray(weights, np.float32)),
                        _ptr(np.ascontiguousarray(layers, np.float32)), _ptr(np.ascontiguousarray(cave_layers)), _ptr(F), _ptr(CF),
                        _ptr(counts), sf, scf, _ptr(out), 1 if decorate else 0, self.nthreads)
        return out

    def erode_zone(self, planes):
        """planes: (9, 384, 384) float32 (8 loose layer starts + heightfield); returns (eroded copy, sweeps)."""
        p = np.ascontiguousarray(planes, np.float32).copy()
        sweeps = self.L.mmo_erode_zone(_ptr(p))
        return p, sweeps


# ---- host-side data movement of the reference, restated for tests (numpy) ----
def gather_h18(height, nx, nz):
    """Chunk::otherChunkGatherHeightfield (chunk.cu:237-293): height (nz*nx,256) -> dict chunk idx -> (18,18)."""
    grid = height.reshape(nz, nx, 16, 16).transpose(0, 2, 1, 3).reshape(nz * 16, nx * 16)  # [z][x] world raster
    out = {}
    for cz in range(1, nz - 1):
        for cx in range(1, nx - 1):
            out[cz * nx + cx] = grid[cz * 16 - 1:cz * 16 + 17, cx * 16 - 1:cx * 16 + 17].copy()
    return out


def gather_zone(layers, height, nx, lx0, lz0):
    """copyLayers(zone, gathered, true) (chunk.cu:603-656): -> (9, 384, 384)."""
    planes = np.empty((9, 384, 384), np.float32)
    for cz in range(24):
        for cx in range(24):
            c = (lz0 + cz) * nx + (lx0 + cx)
            for l in range(8):
                planes[l, cz * 16:(cz + 1) * 16, cx * 16:(cx + 1) * 16] = layers[c, 12 + l].reshape(16, 16)
            planes[8, cz * 16:(cz + 1) * 16, cx * 16:(cx + 1) * 16] = height[c].reshape(16, 16)
    return planes


def scatter_zone(planes, layers, nx, lx0, lz0):
    """copyLayers(zone, gathered, false) + fixBackwardStratifiedLayers (chunk.cu:603-656, 725-749), in place on layers."""
    for cz in range(6, 18):
        for cx in range(6, 18):
            c = (lz0 + cz) * nx + (lx0 + cx)
            for l in range(8):
                layers[c, 12 + l] = planes[l, cz * 16:(cz + 1) * 16, cx * 16:(cx + 1) * 16].reshape(256)
            for l in (10, 11):
                layers[c, l] = layers[c, 12] - layers[c, l]


GATHER_OFFSETS = [(0, 0), (0, 1), (1, 1), (1, 0), (1, -1), (0, -1), (-1, -1), (-1, 0), (-1, 1), (2, 0), (2, 1), (2, 2), (1, 2), (0, 2),
                  (-1, 2), (-2, 2), (-2, 1), (-2, 0), (-2, -1), (-2, -2), (-1, -2), (0, -2), (1, -2), (2, -2), (2, -1),
                  (-3, -3), (-2, -3), (-1, -3), (0, -3), (1, -3), (2, -3), (3, -3), (3, -2), (3, -1), (3, 0), (3, 1), (3, 2), (3, 3),
                  (2, 3), (1, 3), (0, 3), (-1, 3), (-2, 3), (-3, 3), (-3, 2), (-3, 1), (-3, 0), (-3, -1), (-3, -2)]


def gather_features(lists, cx, cz, nx):
    """Chunk::otherChunkGatherFeaturePlacements (chunk.cu:1158-1187): lists = dict chunk idx -> array."""
    parts = [lists[(cz + dz) * nx + (cx + dx)] for dx, dz in GATHER_OFFSETS]
    return np.concatenate(parts) if parts else None
