"""ctypes binding of libmmgen.so (include/mmgen.h).

Mirrors the reference's generation entry points (static Chunk::* functions,
/root/reference/src/terrain/chunk.hpp:100-172) with numpy arrays in the reference's wire layouts.
There is no CPU fallback: if the CUDA library is missing or no GPU is present every call raises.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.environ.get("MMGEN_LIB") or os.path.join(_HERE, "libmmgen.so")      # MMGEN_LIB: developer override for tuning builds

STAGE_HEIGHTFIELD, STAGE_LAYERS, STAGE_EROSION, STAGE_CAVES, STAGE_FEATURES, STAGE_FILL = 1, 2, 4, 8, 16, 32
STAGE_ALL = 63
FILL_OVERLAP_DEFAULT = 8      # mmgen_set_fill_overlap mode the library starts in

CaveLayer = np.dtype([("start", "<i4"), ("end", "<i4"), ("bottomBiome", "u1"), ("topBiome", "u1"), ("pad", "u1", (2,))])
FeaturePlacement = np.dtype([("feature", "u1"), ("pad0", "u1", (3,)), ("x", "<i4"), ("y", "<i4"), ("z", "<i4"),
                             ("canReplaceBlocks", "u1"), ("pad1", "u1", (3,))])
CaveFeaturePlacement = np.dtype([("feature", "u1"), ("pad0", "u1", (3,)), ("x", "<i4"), ("y", "<i4"), ("z", "<i4"),
                                 ("layerHeight", "<i4"), ("canReplaceBlocks", "u1"), ("pad1", "u1", (3,))])
assert CaveLayer.itemsize == 12 and FeaturePlacement.itemsize == 20 and CaveFeaturePlacement.itemsize == 24

NVCC_FLAGS = ["-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-fmad=false", "-lineinfo",
              "-Xcompiler", "-fPIC", "-shared"]


class MmgenError(RuntimeError):
    pass


def lib_path():
    return _LIB


def build(force=False):
    """Compile csrc/mmgen.cu for sm_100a into libmmgen.so (in-tree)."""
    src = os.path.join(_HERE, "csrc", "mmgen.cu")
    deps = [os.path.join(_HERE, "csrc", f) for f in os.listdir(os.path.join(_HERE, "csrc"))]
    deps.append(os.path.join(_HERE, "..", "include", "mmgen.h"))
    if not force and os.path.exists(_LIB) and all(os.path.getmtime(_LIB) >= os.path.getmtime(d) for d in deps):
        return _LIB
    cmd = ["nvcc"] + NVCC_FLAGS + ["-o", _LIB, src]
    subprocess.run(cmd, check=True)
    return _LIB


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


class ChunkGen:
    """Batch operators with host arrays in / host arrays out (one call per stage, like Chunk::*)."""

    _lib = None

    @classmethod
    def lib(cls):
        if cls._lib is None:
            if not os.path.exists(_LIB):
                raise MmgenError("libmmgen.so is not built (run __graft_entry__.build()); there is no CPU fallback")
            L = ctypes.CDLL(_LIB)
            L.mmgen_last_error.restype = ctypes.c_char_p
            L.mmgen_launch_count.restype = ctypes.c_uint64
            cls._lib = L
        return cls._lib

    def __init__(self, device=0):
        self.L = self.lib()
        self._check(self.L.mmgen_init(int(device)))

    def _check(self, rc):
        if rc != 0:
            raise MmgenError(self.L.mmgen_last_error().decode())

    def launch_count(self):
        return int(self.L.mmgen_launch_count())

    def kernel_timing(self, enable=True):
        """Per-kernel device timing (CUDA events around the hot kernels' launches); clears the record."""
        self._check(self.L.mmgen_kernel_timing(1 if enable else 0))

    def kernel_times(self):
        """{kernel name: (summed device ms, launches)} since the last call."""
        ms = np.zeros(32, np.float32)
        cnt = np.zeros(32, np.int32)
        n = ctypes.c_int(0)
        self._check(self.L.mmgen_kernel_times(32, _ptr(ms), _ptr(cnt), ctypes.byref(n)))
        self.L.mmgen_kernel_name.restype = ctypes.c_char_p
        return {self.L.mmgen_kernel_name(i).decode(): (float(ms[i]), int(cnt[i])) for i in range(n.value)}

    def chunk_costs(self, origins):
        """mmgen_chunk_costs: (n, 3) float32 cost features of chunks from stage 1 alone: cave-stage voxels, fill-stage voxels,
        land columns."""
        origins = np.ascontiguousarray(origins, dtype=np.int32).reshape(-1, 2)
        out = np.zeros((origins.shape[0], 3), np.float32)
        self._check(self.L.mmgen_chunk_costs(origins.shape[0], _ptr(origins), _ptr(out)))
        return out

    def selftest_packed_noise(self, n=1 << 22, seed=1):
        """mmgen_selftest_packed_noise: number of packed-noise results that differ from the scalar routines in any bit (must be 0)."""
        v = ctypes.c_uint64(0)
        self._check(self.L.mmgen_selftest_packed_noise(int(n), ctypes.c_uint32(int(seed)), ctypes.byref(v)))
        return v.value

    def set_fill_overlap(self, mode):
        """Scheduling knob (mmgen_set_fill_overlap): 0 = fill passes of a batch in sequence; g = overlap the next batch's terrain / rock
        passes with this batch's placement scan, k_fill_rock at g CTAs per SM (+16: on the high-priority stream)."""
        self._check(self.L.mmgen_set_fill_overlap(int(mode)))

    def set_serial_stages(self, serial):
        """Measurement knob (mmgen_set_serial_stages): run layers + erosion and the caves one after the other instead of overlapped."""
        self._check(self.L.mmgen_set_serial_stages(1 if serial else 0))

    def work_counters(self, reset=True):
        """mmgen_work_counters: 32 uint64 counters of the cheap stages (see include/mmgen.h)."""
        out = np.zeros(32, np.uint64)
        self._check(self.L.mmgen_work_counters(_ptr(out), 1 if reset else 0))
        return out

    def measure_fp32_peak(self):
        """Achieved FP32 FMA rate of the device in TFLOP/s (microbenchmark kernel in libmmgen)."""
        v = ctypes.c_float(0)
        self._check(self.L.mmgen_measure_fp32_peak(ctypes.byref(v)))
        return v.value

    @staticmethod
    def origins(chunk_coords):
        """(n,2) chunk coordinates -> (n,2) int32 block origins."""
        return (np.asarray(chunk_coords, dtype=np.int32).reshape(-1, 2) * 16).astype(np.int32)

    def heightfields(self, origins):
        """Chunk::generateHeightfields (chunk.cu:187-229)."""
        origins = np.ascontiguousarray(origins, dtype=np.int32).reshape(-1, 2)
        n = origins.shape[0]
        h = np.empty((n, 256), np.float32)
        w = np.empty((n, 24, 256), np.float32)
        self._check(self.L.mmgen_heightfields(n, _ptr(origins), _ptr(h), _ptr(w)))
        return h, w

    def layers(self, origins, h18, weights):
        """Chunk::generateLayers on gathered 18x18 heightfields (chunk.cu:417-469)."""
        origins = np.ascontiguousarray(origins, dtype=np.int32).reshape(-1, 2)
        n = origins.shape[0]
        h18 = np.ascontiguousarray(h18, np.float32).reshape(n, 324)
        weights = np.ascontiguousarray(weights, np.float32).reshape(n, 24, 256)
        out = np.empty((n, 20, 256), np.float32)
        self._check(self.L.mmgen_layers(n, _ptr(origins), _ptr(h18), _ptr(weights), _ptr(out)))
        return out

    def caves(self, origins, heightfield, weights):
        """Chunk::generateCaves (chunk.cu:939-993)."""
        origins = np.ascontiguousarray(origins, dtype=np.int32).reshape(-1, 2)
        n = origins.shape[0]
        h = np.ascontiguousarray(heightfield, np.float32).reshape(n, 256)
        w = np.ascontiguousarray(weights, np.float32).reshape(n, 24, 256)
        out = np.zeros((n, 256, 32), CaveLayer)
        self._check(self.L.mmgen_caves(n, _ptr(origins), _ptr(h), _ptr(w), _ptr(out)))
        return out

    def feature_placements(self, origins, heightfield, weights, layers, cave_layers, max_per_chunk=4096):
        """Chunk::generateFeaturePlacements (chunk.cu:1147-1156): per-chunk lists in column order."""
        origins = np.ascontiguousarray(origins, dtype=np.int32).reshape(-1, 2)
        n = origins.shape[0]
        F = np.zeros((n, max_per_chunk), FeaturePlacement)
        CF = np.zeros((n, max_per_chunk), CaveFeaturePlacement)
        counts = np.zeros((n, 2), np.int32)
        self._check(self.L.mmgen_feature_placements(
            n, _ptr(origins), _ptr(np.ascontiguousarray(heightfield, np.float32)), _ptr(np.ascontiguousarray(weights, np.float32)),
            _ptr(np.ascontiguousarray(layers, np.float32)), _ptr(np.ascontiguousarray(cave_layers)), max_per_chunk, _ptr(F), _ptr(CF),
            _ptr(counts)))
        return [F[i, :min(counts[i, 0], max_per_chunk)].copy() for i in range(n)], \
               [CF[i, :min(counts[i, 1], max_per_chunk)].copy() for i in range(n)]

    def gather_offsets(self):
        """The reference's 49 (dx, dz) chunk offsets in gather order (chunk.cu:1158-1167)."""
        out = np.zeros((49, 2), np.int32)
        self._check(self.L.mmgen_gather_offsets(_ptr(out)))
        return out

    def gather_features(self, neighbours, features, cave_features):
        """Chunk::gatherFeaturePlacements (chunk.cu:1158-1196). features / cave_features: own lists of a pool of chunks (as
        feature_placements returns them); neighbours: (n, 49) pool indices in the reference's order (-1 = absent).
        Returns (gathered, gathered_cave, counts): per-chunk lists cut at 2048 / 4096 and the untruncated (n, 2) counts."""
        nb = np.ascontiguousarray(neighbours, np.int32).reshape(-1, 49)
        n, m = nb.shape[0], len(features)
        sf = max(1, max(len(f) for f in features))
        sc = max(1, max(len(f) for f in cave_features))
        F = np.zeros((m, sf), FeaturePlacement)
        CF = np.zeros((m, sc), CaveFeaturePlacement)
        counts = np.zeros((m, 2), np.int32)
        for i in range(m):
            F[i, :len(features[i])] = features[i]
            CF[i, :len(cave_features[i])] = cave_features[i]
            counts[i] = (len(features[i]), len(cave_features[i]))
        oF = np.zeros((n, 2048), FeaturePlacement)
        oC = np.zeros((n, 4096), CaveFeaturePlacement)
        oc = np.zeros((n, 2), np.int32)
        self._check(self.L.mmgen_gather_features(n, _ptr(nb), m, _ptr(F), sf, _ptr(CF), sc, _ptr(counts), _ptr(oF), _ptr(oC), _ptr(oc)))
        return [oF[i, :min(oc[i, 0], 2048)].copy() for i in range(n)], [oC[i, :min(oc[i, 1], 4096)].copy() for i in range(n)], oc

    def fill(self, origins, heightfield, weights, layers, cave_layers, gathered, gathered_cave):
        """Chunk::fill incl. placeDecorators (chunk.cu:1518-1747); gathered lists per chunk, untruncated."""
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
        self._check(self.L.mmgen_fill(
            n, _ptr(origins), _ptr(np.ascontiguousarray(heightfield, np.float32)), _ptr(np.ascontiguousarray(weights, np.float32)),
            _ptr(np.ascontiguousarray(layers, np.float32)), _ptr(np.ascontiguousarray(cave_layers)), _ptr(F), _ptr(CF), _ptr(counts),
            sf, scf, _ptr(out)))
        return out

    def erode_zone(self, gathered):
        """Chunk::erodeZone's relaxation (chunk.cu:658-709): (9,384,384) -> ((8,384,384), sweeps)."""
        g = np.ascontiguousarray(gathered, np.float32).reshape(9, 384, 384)
        out = np.empty((8, 384, 384), np.float32)
        sweeps = ctypes.c_int(0)
        self._check(self.L.mmgen_erode_zone(_ptr(g), _ptr(out), ctypes.byref(sweeps)))
        return out, sweeps.value

    def world(self, cx0, cz0, nx, nz):
        return World(self, cx0, cz0, nx, nz)

    def region_world(self, rx0, rz0, rnx, rnz):
        """World sized by the apron rule to fill chunks [rx0,rx0+rnx) x [rz0,rz0+rnz)."""
        return World(self, rx0, rz0, rnx, rnz, region=True)


class World:
    """Device-resident window of chunks (mmgen_world_*)."""

    def __init__(self, gen, cx0, cz0, nx, nz, region=False):
        self.gen, self.L = gen, gen.L
        self.h = ctypes.c_void_p()
        if region:
            gen._check(self.L.mmgen_world_create_for_region(cx0, cz0, nx, nz, ctypes.byref(self.h)))
        else:
            gen._check(self.L.mmgen_world_create(cx0, cz0, nx, nz, ctypes.byref(self.h)))
        win = (ctypes.c_int * 8)()
        gen._check(self.L.mmgen_world_window(self.h, win))
        self.cx0, self.cz0, self.nx, self.nz = win[0], win[1], win[2], win[3]
        self.rx0, self.rz0, self.rnx, self.rnz = win[4], win[5], win[6], win[7]
        self.n = self.nx * self.nz

    def close(self):
        if self.h:
            self.L.mmgen_world_destroy(self.h)
            self.h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def generate(self, stage_mask=STAGE_ALL):
        self.gen._check(self.L.mmgen_world_generate(self.h, int(stage_mask)))

    def reset(self):
        self.gen._check(self.L.mmgen_world_reset(self.h))

    def set_exchange_region(self, gx0, gz0, gnx, gnz):
        """Halo exchange: placements of ring chunks inside this global region come from the neighbouring tiles' worlds
        (pack_placements / unpack_placements) instead of being recomputed. gnx <= 0 switches it off."""
        self.gen._check(self.L.mmgen_world_set_exchange_region(self.h, int(gx0), int(gz0), int(gnx), int(gnz)))

    def pack_placements(self, rect, dev_ptr, cap_bytes):
        """Packs the placement lists of the chunk rectangle rect = (cx0, cz0, nx, nz) into the DEVICE buffer at dev_ptr.
        Returns (bytes, ok): ok is False when the buffer is too small (bytes is then the size needed)."""
        n = ctypes.c_size_t(0)
        rc = self.L.mmgen_world_pack_placements(self.h, int(rect[0]), int(rect[1]), int(rect[2]), int(rect[3]), ctypes.c_void_p(int(dev_ptr)),
                                                ctypes.c_size_t(int(cap_bytes)), ctypes.byref(n))
        if rc == 2:
            return n.value, False
        self.gen._check(rc)
        return n.value, True

    def unpack_placements(self, rect, dev_ptr, nbytes):
        self.gen._check(self.L.mmgen_world_unpack_placements(self.h, int(rect[0]), int(rect[1]), int(rect[2]), int(rect[3]),
                                                             ctypes.c_void_p(int(dev_ptr)), ctypes.c_size_t(int(nbytes))))

    def rewind(self, stage):
        """Chunks beyond `stage` fall back to it; later stages can be generated again from the resident earlier products."""
        self.gen._check(self.L.mmgen_world_rewind(self.h, int(stage)))

    def generate_to_host(self, out_blocks_ptr, stage_mask=STAGE_ALL):
        """Generate and deliver the region's block volumes to host memory at address out_blocks_ptr
        (uint8[rnz][rnx][98304]; pinned memory makes the copy overlap the fill)."""
        self.gen._check(self.L.mmgen_world_generate_to_host(self.h, int(stage_mask), ctypes.c_void_p(int(out_blocks_ptr))))

    def generate_to_host_encoded(self, out_ptr, cap_bytes, index, stage_mask=STAGE_ALL):
        """Generate and deliver the region's block volumes run-length coded (format MMCH1, include/mmgen.h) into host memory
        at out_ptr; index: uint64 array (rnz * rnx, 2) that receives {offset, bytes} per chunk. Returns the payload length;
        raises MmgenError when cap_bytes is too small (the message names the size needed)."""
        n = ctypes.c_size_t(0)
        assert index.dtype == np.uint64 and index.size >= 2 * self.rnx * self.rnz
        self.gen._check(self.L.mmgen_world_generate_to_host_encoded(self.h, int(stage_mask), ctypes.c_void_p(int(out_ptr)), ctypes.c_size_t(int(cap_bytes)),
                                                                    _ptr(index), ctypes.byref(n)))
        return n.value

    def total_ms(self):
        v = ctypes.c_float(0)
        self.gen._check(self.L.mmgen_world_total_ms(self.h, ctypes.byref(v)))
        return v.value

    def download_region_blocks(self):
        b = np.empty((self.rnz * self.rnx, 16, 16, 384), np.uint8)
        self.gen._check(self.L.mmgen_world_download_region_blocks(self.h, _ptr(b)))
        return b

    def sync(self):
        self.gen._check(self.L.mmgen_world_sync(self.h))

    def stages(self):
        out = np.zeros(self.n, np.uint8)
        self.gen._check(self.L.mmgen_world_stages(self.h, _ptr(out)))
        return out.reshape(self.nz, self.nx)

    def stage_ms(self):
        out = np.zeros(7, np.float32)
        self.gen._check(self.L.mmgen_world_stage_ms(self.h, _ptr(out)))
        return out

    def erosion_sweeps(self):
        v = ctypes.c_int(0)
        self.gen._check(self.L.mmgen_world_erosion_sweeps(self.h, ctypes.byref(v)))
        return v.value

    def block_checksum(self):
        v = ctypes.c_uint64(0)
        self.gen._check(self.L.mmgen_world_block_checksum(self.h, ctypes.byref(v)))
        return v.value

    def chunk_hash_sum(self):
        v = ctypes.c_uint64(0)
        self.gen._check(self.L.mmgen_world_chunk_hash_sum(self.h, ctypes.byref(v)))
        return v.value

    def chunk_hashes(self):
        """{(cx, cz): 64-bit hash of (coordinates, block volume)} for every filled chunk; chunk_hash_sum() is their sum mod 2^64."""
        cap = self.n
        coords = np.zeros((cap, 2), np.int32)
        hs = np.zeros(cap, np.uint64)
        n = ctypes.c_int(0)
        self.gen._check(self.L.mmgen_world_chunk_hashes(self.h, cap, _ptr(coords), _ptr(hs), ctypes.byref(n)))
        return {(int(coords[k, 0]), int(coords[k, 1])): int(hs[k]) for k in range(n.value)}

    def download_features(self, max_per_chunk=4096):
        F = np.zeros((self.n, max_per_chunk), FeaturePlacement)
        CF = np.zeros((self.n, max_per_chunk), CaveFeaturePlacement)
        counts = np.zeros((self.n, 2), np.int32)
        self.gen._check(self.L.mmgen_world_download_features(self.h, max_per_chunk, _ptr(F), _ptr(CF), _ptr(counts)))
        return [F[i, :min(counts[i, 0], max_per_chunk)].copy() for i in range(self.n)], \
               [CF[i, :min(counts[i, 1], max_per_chunk)].copy() for i in range(self.n)]

    def download(self, heightfield=False, biome_weights=False, layers=False, cave_layers=False, blocks=False):
        res = {}
        h = np.empty((self.n, 256), np.float32) if heightfield else None
        w = np.empty((self.n, 24, 256), np.float32) if biome_weights else None
        l = np.empty((self.n, 20, 256), np.float32) if layers else None
        c = np.empty((self.n, 256, 32), CaveLayer) if cave_layers else None
        b = np.empty((self.n, 16, 16, 384), np.uint8) if blocks else None
        self.gen._check(self.L.mmgen_world_download(self.h, _ptr(h), _ptr(w), _ptr(l), _ptr(c), _ptr(b)))
        for k, v in (("heightfield", h), ("biome_weights", w), ("layers", l), ("cave_layers", c), ("blocks", b)):
            if v is not None:
                res[k] = v
        return res


def _codec_lib():
    L = ChunkGen.lib()
    L.mmgen_last_error.restype = ctypes.c_char_p
    return L


def decode_chunk(enc):
    """One MMCH1-encoded chunk (bytes / uint8 array) -> uint8[16][16][384] (mmgen_decode_chunk, host code, needs no GPU)."""
    enc = np.ascontiguousarray(np.frombuffer(enc, np.uint8) if isinstance(enc, (bytes, bytearray, memoryview)) else enc, np.uint8)
    out = np.empty((16, 16, 384), np.uint8)
    L = _codec_lib()
    if L.mmgen_decode_chunk(_ptr(enc), ctypes.c_size_t(enc.size), _ptr(out)) != 0:
        raise MmgenError(L.mmgen_last_error().decode())
    return out


def save_region(path, region, index, payload):
    """Writes a region file (header, index, MMCH1 payload) from what World.generate_to_host_encoded delivered."""
    L = _codec_lib()
    index = np.ascontiguousarray(index, np.uint64)
    payload = np.ascontiguousarray(payload, np.uint8)
    if L.mmgen_region_save(str(path).encode(), int(region[0]), int(region[1]), int(region[2]), int(region[3]), _ptr(index), _ptr(payload),
                           ctypes.c_size_t(payload.size)) != 0:
        raise MmgenError(L.mmgen_last_error().decode())


class RegionFile:
    """Random access to the chunks of a region file (mmgen_region_open / read_chunk / close)."""

    def __init__(self, path):
        self.L = _codec_lib()
        self.h = ctypes.c_void_p()
        rect = (ctypes.c_int32 * 4)()
        if self.L.mmgen_region_open(str(path).encode(), ctypes.byref(self.h), rect) != 0:
            raise MmgenError(self.L.mmgen_last_error().decode())
        self.region = tuple(rect)

    def read_chunk(self, cx, cz):
        out = np.empty((16, 16, 384), np.uint8)
        if self.L.mmgen_region_read_chunk(self.h, int(cx), int(cz), _ptr(out)) != 0:
            raise MmgenError(self.L.mmgen_last_error().decode())
        return out

    def close(self):
        if self.h:
            self.L.mmgen_region_close(self.h)
            self.h = ctypes.c_void_p()


class TickStats(ctypes.Structure):
    """MmgenTickStats (include/mmgen.h)."""
    _fields_ = [(k, ctypes.c_int32) for k in ("heightfields", "gatherHeightfields", "layers", "zonesEroded", "caves", "placements",
                                               "gatherPlacements", "filled", "vbos", "actionTimeLeft", "idle")] + \
               [("deviceMs", ctypes.c_float), ("meshVertices", ctypes.c_int64)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


# reference action-time costs (terrain.cpp:72-80): heightfield, gatherHeightfield, layers, erodeZone, caves, featurePlacements,
# gatherFeaturePlacements, fill, createVbos; frame cap 500, refill 60 * 500 per second (terrain.cpp:69-70)
REFERENCE_COSTS = (3, 2, 5, 500, 8, 3, 5, 8, 500 // 3)
CHUNK_FILLED, CHUNK_NEEDS_VBOS, CHUNK_DRAWABLE = 9, 10, 11


class Terrain:
    """The reference's chunk manager re-hosted on a device-resident world (mmgen_stream_*): same method names and
    meaning as `Terrain` (terrain.hpp:55-131) for the generation path - setCurrentChunkPos, tick - headless."""

    def __init__(self, gen, cx0, cz0, nx, nz):
        self.gen, self.L = gen, gen.L
        self.h = ctypes.c_void_p()
        gen._check(self.L.mmgen_stream_create(cx0, cz0, nx, nz, ctypes.byref(self.h)))
        self.cx0, self.cz0, self.nx, self.nz = cx0, cz0, nx, nz

    def close(self):
        if self.h:
            self.L.mmgen_stream_destroy(self.h)
            self.h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_radii(self, vbos_gen_radius=16, max_gen_radius=40):
        self.gen._check(self.L.mmgen_stream_set_radii(self.h, int(vbos_gen_radius), int(max_gen_radius)))

    def set_costs(self, costs=REFERENCE_COSTS, max_per_frame=500, per_second=60 * 500):
        c = (ctypes.c_int32 * 9)(*costs) if costs is not None else None
        self.gen._check(self.L.mmgen_stream_set_costs(self.h, c, int(max_per_frame), int(per_second)))

    def set_meshing(self, enable=True):
        """createVBOs on the device for the chunks that leave the VBO queue."""
        self.gen._check(self.L.mmgen_stream_set_meshing(self.h, 1 if enable else 0))

    def setCurrentChunkPos(self, cx, cz):
        self.set_player(cx * 16.0 + 8.0, cz * 16.0 + 8.0)

    def set_player(self, x, z):
        self.gen._check(self.L.mmgen_stream_set_player(self.h, ctypes.c_float(x), ctypes.c_float(z)))

    def tick(self, delta_time=1.0 / 60.0):
        st = TickStats()
        self.gen._check(self.L.mmgen_stream_tick(self.h, ctypes.c_float(delta_time), ctypes.byref(st)))
        return st

    def states(self):
        out = np.zeros(self.nx * self.nz, np.uint8)
        self.gen._check(self.L.mmgen_stream_states(self.h, _ptr(out)))
        return out.reshape(self.nz, self.nx)

    def take_filled(self, cap=4096):
        coords = np.zeros((cap, 2), np.int32)
        n = ctypes.c_int(0)
        self.gen._check(self.L.mmgen_stream_take_filled(self.h, _ptr(coords), cap, ctypes.byref(n)))
        return coords[:n.value].copy()

    def download_chunk(self, cx, cz):
        b = np.empty((16, 16, 384), np.uint8)
        self.gen._check(self.L.mmgen_stream_download_chunk(self.h, int(cx), int(cz), _ptr(b)))
        return b

    def chunk_hash_sum(self):
        """Tiling-invariant hash of every filled chunk (mmgen_world_chunk_hash_sum of the backing world)."""
        w = ctypes.c_void_p()
        self.gen._check(self.L.mmgen_stream_world(self.h, ctypes.byref(w)))
        v = ctypes.c_uint64(0)
        self.gen._check(self.L.mmgen_world_chunk_hash_sum(w, ctypes.byref(v)))
        return v.value

    def run_until_idle(self, delta_time=1.0 / 60.0, max_ticks=100000):
        """Ticks until the scheduler has nothing left to do; returns the per-tick stats."""
        log = []
        for _ in range(max_ticks):
            st = self.tick(delta_time)
            log.append(st.as_dict())
            if st.idle:
                break
        return log


Vertex = np.dtype([("pos", "<f4", (3,)), ("nor", "<f4", (3,)), ("uv", "<f4", (2,)), ("m", "<u8")])     # MmgenVertex, 40 B


def _world_mesh(self, chunk_coords, download=True):
    """Chunk::createVBOs on the device (mmgen_world_mesh) for filled chunks given as (cx, cz) pairs.
    Returns [(verts, idx)] per chunk (Vertex records, uint32 indices relative to the chunk), or the counts if not download."""
    coords = np.ascontiguousarray(chunk_coords, np.int32).reshape(-1, 2)
    n = coords.shape[0]
    counts = np.zeros((n, 2), np.int32)
    self.gen._check(self.L.mmgen_world_mesh(self.h, n, _ptr(coords), _ptr(counts)))
    if not download:
        return counts
    out = []
    for i in range(n):
        v = np.zeros(int(counts[i, 0]), Vertex)
        ix = np.zeros(int(counts[i, 1]), np.uint32)
        self.gen._check(self.L.mmgen_world_mesh_download(self.h, i, _ptr(v), _ptr(ix)))
        out.append((v, ix))
    return out


def _world_mesh_ms(self):
    v = ctypes.c_float(0)
    self.gen._check(self.L.mmgen_world_mesh_ms(self.h, ctypes.byref(v)))
    return v.value


GasInput = np.dtype([("vertexBuffer", "<u8"), ("numVertices", "<u4"), ("vertexStrideInBytes", "<u4"), ("indexBuffer", "<u8"),
                     ("numIndexTriplets", "<u4"), ("indexStrideInBytes", "<u4"), ("cx", "<i4"), ("cz", "<i4")])      # MmgenGasInput, 40 B


def _world_mesh_gas_inputs(self):
    """mmgen_world_mesh_gas_inputs: the OptiX triangle-array description (device pointers into the mesh arena) of every chunk
    of the last mesh() call - what OptixRenderer::buildChunkAccel needs instead of host vectors."""
    n = ctypes.c_int(0)
    self.gen._check(self.L.mmgen_world_mesh_gas_inputs(self.h, 0, None, ctypes.byref(n)))
    out = np.zeros(n.value, GasInput)
    self.gen._check(self.L.mmgen_world_mesh_gas_inputs(self.h, n.value, _ptr(out), ctypes.byref(n)))
    return out


World.mesh_gas_inputs = _world_mesh_gas_inputs
World.mesh = _world_mesh
World.mesh_ms = _world_mesh_ms
