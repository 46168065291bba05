"""B200-native chunk-generation path (host-side Python binding over the C ABI in include/mmgen.h).

The directory name is not a Python identifier; load it with `mmgen_loader.load()` at the repo root
(tests, bench.py and __graft_entry__.py do) - it registers this package as `mega_minecraft_b200`.
"""
from .chunkgen import (  # noqa: F401
    ChunkGen, World, Terrain, TickStats, Vertex, RegionFile, decode_chunk, save_region, REFERENCE_COSTS, MmgenError, lib_path, build, CaveLayer, FeaturePlacement, CaveFeaturePlacement,
    STAGE_HEIGHTFIELD, STAGE_LAYERS, STAGE_EROSION, STAGE_CAVES, STAGE_FEATURES, STAGE_FILL, STAGE_ALL, FILL_OVERLAP_DEFAULT,
)
