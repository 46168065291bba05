/* mmgen.h - C ABI of the B200-native chunk-generation path.
 *
 * This is the drop-in boundary for the generation entry points that the reference's chunk manager
 * (Terrain::tick, /root/reference/src/terrain/terrain.cpp:587-960) calls on `Chunk`
 * (/root/reference/src/terrain/chunk.hpp:100-172). The reference has no FFI layer: that C++ surface
 * is its operator API, so every function below names the reference entry point it replaces.
 * Plain pointers and sizes only; all functions return 0 on success, non-zero on failure, and
 * mmgen_last_error() returns a message (the reference prints and exit()s instead,
 * /root/reference/src/cuda/cuda_utils.cpp:5-17).
 *
 * Wire formats are the reference's:
 *   origins      int32[n][2]          chunk origin in blocks (x, z)            chunk.cu:199-203
 *   heightfield  float[n][256]        index x + 16*z                           chunk.hpp:59-61
 *   biomeWeights float[n][24][256]    [biome][z][x]                            chunk.hpp:70-71
 *   layers       float[n][20][256]    [material][z][x], layer START heights    chunk.hpp:64-65
 *   caveLayers   MmgenCaveLayer[n][256][32]                                    biome.hpp:108-117
 *   blocks       uint8[n][16][16][384] index y + 384*(x + 16*z)                biomeFuncs.hpp:25-30
 *   features     MmgenFeaturePlacement / MmgenCaveFeaturePlacement             biome.hpp:207-212,254-260
 *
 * Two ways to use it:
 *   (1) batch operators (mmgen_heightfields ... mmgen_fill): HOST pointers in, HOST pointers out,
 *       one call per stage like the reference's static Chunk::* functions. Host<->device copies
 *       happen inside the call; the call returns when results are in host memory.
 *   (2) device-resident world (mmgen_world_*): a rectangular window of chunks whose intermediate
 *       products never leave HBM between stages; this is what the headless bench drives and what
 *       the multi-GPU tiling uses.
 */
#ifndef MMGEN_H
#define MMGEN_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MMGEN_NUM_BIOMES 24
#define MMGEN_NUM_MATERIALS 20
#define MMGEN_MAX_CAVE_LAYERS 32
#define MMGEN_MAX_FEATURES 2048       /* MAX_GATHERED_FEATURES_PER_CHUNK, biome.hpp:7 */
#define MMGEN_MAX_CAVE_FEATURES 4096  /* MAX_GATHERED_CAVE_FEATURES_PER_CHUNK, biome.hpp:8 */
#define MMGEN_CHUNK_BLOCKS 98304
#define MMGEN_ZONE_SIZE 12            /* chunks per zone side, terrain.hpp:17 */

typedef struct { int32_t start, end; uint8_t bottomBiome, topBiome, pad[2]; } MmgenCaveLayer;          /* 12 B */
typedef struct { uint8_t feature, pad0[3]; int32_t x, y, z; uint8_t canReplaceBlocks, pad1[3]; } MmgenFeaturePlacement; /* 20 B */
typedef struct { uint8_t feature, pad0[3]; int32_t x, y, z; int32_t layerHeight; uint8_t canReplaceBlocks, pad1[3]; } MmgenCaveFeaturePlacement; /* 24 B */

/* stage bits for mmgen_world_generate */
enum {
    MMGEN_STAGE_HEIGHTFIELD = 1,   /* S1  Chunk::generateHeightfields            chunk.cu:187-229 */
    MMGEN_STAGE_LAYERS = 2,        /* S2  gatherHeightfield + generateLayers     chunk.cu:237-302, 417-469 */
    MMGEN_STAGE_EROSION = 4,       /* S3  Chunk::erodeZone                       chunk.cu:658-749 */
    MMGEN_STAGE_CAVES = 8,         /* S4  Chunk::generateCaves                   chunk.cu:939-993 */
    MMGEN_STAGE_FEATURES = 16,     /* S5  generate/gatherFeaturePlacements       chunk.cu:1147-1196 */
    MMGEN_STAGE_FILL = 32,         /* S6  Chunk::fill incl. placeDecorators      chunk.cu:1518-1747 */
    MMGEN_STAGE_ALL = 63
};

/* ---- lifetime: replaces BiomeUtils::init() (biomeFuncs.hpp:725) + Terrain::initCuda/freeCuda
 *      (terrain.cpp:154-218). device = CUDA ordinal. Fails if no CUDA device is usable: there is
 *      no CPU fallback. */
int mmgen_init(int device);
int mmgen_shutdown(void);
const char* mmgen_last_error(void);
/* number of kernel launches issued by this library since mmgen_init (bench.py's gpu_launches) */
uint64_t mmgen_launch_count(void);

/* ---- batch operators (host pointers) ----
 * Like the reference's entry points (file-static staging buffers, terrain.cpp:131-152) they share one stream and one set of
 * device scratch buffers; concurrent calls from several host threads are serialised inside the library. */

/* Chunk::generateHeightfields, chunk.hpp:100-108 / chunk.cu:187-229 */
int mmgen_heightfields(int n, const int32_t* origins, float* out_heightfield, float* out_biomeWeights);

/* Chunk::gatherHeightfield + Chunk::generateLayers, chunk.hpp:110-122 / chunk.cu:237-302, 322-469.
 * heightfield18: float[n][18*18], the chunk's heightfield with the 1-block border gathered from its
 * 8 neighbours (what otherChunkGatherHeightfield builds). Forward layers the reference leaves
 * unwritten after its early break (chunk.cu:387-390) are written as the running height here. */
int mmgen_layers(int n, const int32_t* origins, const float* heightfield18, const float* biomeWeights,
                 float* out_layers);

/* Chunk::erodeZone, chunk.hpp:124-129 / chunk.cu:477-749. gathered: float[9][384][384] = the 8 loose
 * layer-start planes (materials 12..19) followed by the heightfield plane of the 24x24-chunk window
 * around the zone, exactly what copyLayers(..., true) builds. out_eroded: float[8][384][384], the
 * relaxed loose layer starts of the whole window (the reference keeps the centre 192x192). */
int mmgen_erode_zone(const float* gathered, float* out_eroded, int* out_sweeps);

/* Chunk::generateCaves, chunk.hpp:131-141 / chunk.cu:755-993 */
int mmgen_caves(int n, const int32_t* origins, const float* heightfield, const float* biomeWeights,
                MmgenCaveLayer* out_caveLayers);

/* Chunk::generateFeaturePlacements, chunk.hpp:151 / chunk.cu:999-1156 (CPU in the reference).
 * Outputs at most maxPerChunk placements per chunk in the reference's column order (z outer, x
 * inner); out_counts[n][2] = {surface, cave} counts. */
int mmgen_feature_placements(int n, const int32_t* origins, const float* heightfield, const float* biomeWeights,
                             const float* layers, const MmgenCaveLayer* caveLayers, int maxPerChunk,
                             MmgenFeaturePlacement* out_features, MmgenCaveFeaturePlacement* out_caveFeatures,
                             int32_t* out_counts);

/* Chunk::gatherFeaturePlacements, chunk.hpp:152 / chunk.cu:1158-1196 (CPU in the reference): the own lists of a chunk's 7x7
 * neighbourhood concatenated in the reference's fixed order (mmgen_gather_offsets writes that table: 49 (dx, dz) chunk
 * offsets, chunk.cu:1158-1167). The own lists live in a pool of m chunks laid out as mmgen_feature_placements writes them
 * (features[m][featureStride], caveFeatures[m][caveFeatureStride], counts[m][2]); neighbours[n][49] holds, for each of the n
 * chunks to gather for, the pool index of the chunk at offset k (< 0: no such chunk, skipped). Outputs are the lists
 * Chunk::fill uploads: out_features[n][2048], out_caveFeatures[n][4096] cut at the reference's caps (chunk.cu:1573-1578),
 * out_counts[n][2] = the untruncated lengths. These are mmgen_fill's inputs (clamp the counts to the caps). */
int mmgen_gather_offsets(int32_t* out49x2);
int mmgen_gather_features(int n, const int32_t* neighbours, int m, const MmgenFeaturePlacement* features, int featureStride,
                          const MmgenCaveFeaturePlacement* caveFeatures, int caveFeatureStride, const int32_t* counts,
                          MmgenFeaturePlacement* out_features, MmgenCaveFeaturePlacement* out_caveFeatures, int32_t* out_counts);

/* Which reading of Chunk::tryGenerateCaveFeaturePlacement (chunk.cu:1010-1038: no return statement where its jittered-grid
 * test fails - undefined behaviour) stage 5 follows. 0 (default) = the reference as built by nvcc/g++ on Linux, which drops
 * the test - the executable oracle of the parity contract; 1 = the source text (test honoured, a failed test is `false`).
 * Applies to mmgen_feature_placements, the world and the stream from the next call on. */
int mmgen_set_cave_grid_test(int honoured);

/* Chunk::fill incl. Chunk::placeDecorators, chunk.hpp:154-172 / chunk.cu:1202-1747.
 * features / caveFeatures: the GATHERED lists per chunk (what gatherFeaturePlacements builds,
 * chunk.cu:1158-1196), numFeatures[n][2] = {surface, cave} list lengths (no terminator needed); a length above its
 * stride is an error. */
int mmgen_fill(int n, const int32_t* origins, const float* heightfield, const float* biomeWeights,
               const float* layers, const MmgenCaveLayer* caveLayers,
               const MmgenFeaturePlacement* features, const MmgenCaveFeaturePlacement* caveFeatures,
               const int32_t* numFeatures, int featureStride, int caveFeatureStride, uint8_t* out_blocks);

/* ---- device-resident world ---- */
typedef struct MmgenWorld MmgenWorld;

/* A window of nx x nz chunks whose lower corner is chunk (cx0, cz0). Every stage is run on as
 * much of the window as its halo allows, like the reference state machine would inside it:
 * S1 everywhere; S2 where the 3x3 chunk neighbourhood exists; S3 on zones (aligned to multiples of
 * 12 chunks) whose 24x24 window exists; S4/S5a on eroded zones; S5b/S6 where the 7x7 neighbourhood
 * has placements. */
int mmgen_world_create(int cx0, int cz0, int nx, int nz, MmgenWorld** out);
/* A world sized to FILL the region of rnx x rnz chunks whose lower corner is chunk (rx0, rz0), by the
 * apron rule that follows from the reference's state machine (terrain.cpp:471-522, chunk.cu:53-136):
 * placements on the region (+) 3 chunks, erosion of every zone that meets that, layers on those
 * zones (+) 6 chunks, heightfields on one more ring (the layers' 1-block slope border). Stages are restricted to what the region needs; results are
 * identical to a larger window's. This is the unit a multi-GPU tiling gives to each GPU. */
int mmgen_world_create_for_region(int rx0, int rz0, int rnx, int rnz, MmgenWorld** out);
int mmgen_world_destroy(MmgenWorld* w);
/* ---- halo exchange between tiles of one world (multi-GPU; SURVEY.md 8(e) option B). By default a region world recomputes
 * the placements (stages 4 + 5a) of the 3-chunk ring around its tile. After mmgen_world_set_exchange_region(global region in
 * chunk coordinates) it computes them only for its own tile and for ring chunks outside the global region; the rest of the
 * ring belongs to neighbouring tiles, whose worlds pack their lists (mmgen_world_pack_placements) into a DEVICE buffer that the
 * caller moves (NCCL send / recv, or a peer copy) and this world unpacks (mmgen_world_unpack_placements) before it runs the
 * fill. Message = int32 counts[n][2], then per chunk of the rectangle (raster order) its surface and cave lists packed; *bytes
 * is its length (return code 2 and *bytes set when capBytes is too small). Rectangles are in world chunk coordinates.
 * gnx <= 0 switches the exchange off again. */
int mmgen_world_set_exchange_region(MmgenWorld* w, int gx0, int gz0, int gnx, int gnz);
int mmgen_world_pack_placements(MmgenWorld* w, int cx0, int cz0, int nx, int nz, void* d_buf, size_t capBytes, size_t* bytes);
int mmgen_world_unpack_placements(MmgenWorld* w, int cx0, int cz0, int nx, int nz, const void* d_buf, size_t bytes);
/* Cost features of n chunks from stage 1 alone (runs Chunk::generateHeightfields on them and reduces on the device), for
 * cutting balanced tiles before anything expensive has run: out_costs[n][3] = {cave-stage voxels sum max(floor(h), 128),
 * fill-stage voxels sum clamp(floor(h), 1, 383), land columns sum (1 - ocean/beach weight)}. */
int mmgen_chunk_costs(int n, const int32_t* origins, float* out_costs);
/* out8 = {window cx0, cz0, nx, nz, region rx0, rz0, rnx, rnz} (region == window without a target) */
int mmgen_world_window(MmgenWorld* w, int* out8);
/* forget all progress (stages back to 0); buffers stay allocated */
int mmgen_world_reset(MmgenWorld* w);
/* forget the progress beyond `stage` (0..6): chunks further along fall back to it and the later stages can be generated again
 * from the resident products of the earlier ones (used to time one stage in isolation, BASELINE config 4) */
int mmgen_world_rewind(MmgenWorld* w, int stage);
int mmgen_world_generate(MmgenWorld* w, int stageMask);
/* mmgen_world_generate + delivery of the block volumes into HOST memory, the reference's contract for
 * Chunk::fill (results complete in host memory on return, chunk.cu:1621): out_blocks is
 * uint8[rnz][rnx][98304] in region raster order (window raster order without a target region), ideally
 * page-locked; filled batches are copied while later batches are still being filled. */
int mmgen_world_generate_to_host(MmgenWorld* w, int stageMask, uint8_t* out_blocks);
/* ---- chunk wire / on-disk format MMCH1 (the reference ships raw volumes and has no file format):
 *   encoded chunk := uint16 nRuns[256] (column order x + 16 z, little endian), then for each column nRuns x {uint8 block,
 *   uint8 length} with length 1..255 and a column's lengths adding up to 384, zero-padded to a multiple of 16 bytes; a chunk
 *   whose code would reach 98 304 bytes is stored raw and recognised by that length.
 * mmgen_world_generate_to_host_encoded: like mmgen_world_generate_to_host, but the block volumes are run-length coded on the
 * device and delivered as one payload in out_buf (ideally page-locked) plus out_index[rnz * rnx][2] = {offset, bytes} of every
 * chunk of the region in raster order; *out_bytes = payload length. Return code 2 (and *out_bytes = the size needed) when
 * capBytes is too small. mmgen_decode_chunk (host) restores uint8[16][16][384] from one encoded chunk. */
int mmgen_world_generate_to_host_encoded(MmgenWorld* w, int stageMask, uint8_t* out_buf, size_t capBytes, uint64_t* out_index,
                                         size_t* out_bytes);
int mmgen_decode_chunk(const uint8_t* enc, size_t nbytes, uint8_t* out_blocks);
/* region files: header "MMRG", version, region rectangle, the index, then the payload - exactly what
 * mmgen_world_generate_to_host_encoded delivered. read_chunk decodes one chunk (cx, cz in chunk coordinates). */
typedef struct MmgenRegionFile MmgenRegionFile;
int mmgen_region_save(const char* path, int rx0, int rz0, int rnx, int rnz, const uint64_t* index, const uint8_t* payload, size_t payloadBytes);
int mmgen_region_open(const char* path, MmgenRegionFile** out, int32_t* out_rect4);
int mmgen_region_read_chunk(MmgenRegionFile* r, int cx, int cz, uint8_t* out_blocks);
int mmgen_region_close(MmgenRegionFile* r);
/* blocks until all queued work of the world is done */
int mmgen_world_sync(MmgenWorld* w);
/* furthest completed stage per chunk (0..6), raster order i = (cz-cz0)*nx + (cx-cx0) */
int mmgen_world_stages(MmgenWorld* w, uint8_t* out);
/* device time of the last mmgen_world_generate per stage (ms, CUDA events), stage 1..6. When one call runs layers, erosion and
 * caves, stages 2-3 run on a side stream concurrently with stage 4 (mmgen_set_serial_stages): their elapsed times overlap and do
 * not add up to mmgen_world_total_ms. */
int mmgen_world_stage_ms(MmgenWorld* w, float* out7);
/* device time of the whole last mmgen_world_generate[_to_host] call (ms, CUDA events on the world's stream) */
int mmgen_world_total_ms(MmgenWorld* w, float* out);
int mmgen_world_erosion_sweeps(MmgenWorld* w, int* out);

/* downloads (host pointers), raster order; any pointer may be NULL */
int mmgen_world_download(MmgenWorld* w, float* heightfield, float* biomeWeights, float* layers,
                         MmgenCaveLayer* caveLayers, uint8_t* blocks);
/* block volumes of the target region only, uint8[rnz][rnx][98304] */
int mmgen_world_download_region_blocks(MmgenWorld* w, uint8_t* out_blocks);
/* per-chunk placement lists (own, not gathered): counts[n][2]; lists packed with stride maxPerChunk */
int mmgen_world_download_features(MmgenWorld* w, int maxPerChunk, MmgenFeaturePlacement* features,
                                  MmgenCaveFeaturePlacement* caveFeatures, int32_t* counts);
/* device pointers of the resident planes (for zero-copy consumers such as a mesher) */
int mmgen_world_device_ptrs(MmgenWorld* w, void** heightfield, void** biomeWeights, void** layers,
                            void** caveLayers, void** blocks);
/* 64-bit FNV-1a of the block volume of every filled chunk, computed on the device (cheap
 * cross-GPU / cross-run equality check for large worlds) */
int mmgen_world_block_checksum(MmgenWorld* w, uint64_t* out);
/* sum (mod 2^64) over the filled chunks of a 64-bit hash of (chunk coordinates, block volume): the same
 * number however a region is tiled over worlds / GPUs, so tiles can be checked against a single world */
int mmgen_world_chunk_hash_sum(MmgenWorld* w, uint64_t* out);
/* the terms of that sum: chunk coordinates (cx, cz) and hash of every filled chunk, raster order; *n = filled chunks (may exceed cap) */
int mmgen_world_chunk_hashes(MmgenWorld* w, int cap, int32_t* coords, uint64_t* hashes, int* n);

/* ---- streaming scheduler: Terrain::tick (terrain.cpp:587-960) re-hosted on a device-resident world.
 * The reference's ChunkState machine (chunk.hpp:18-32), spiral order (terrain.cpp:220-252), per-stage FIFO
 * queues, drain order and action-time budget (terrain.cpp:67-82) are kept, so with the default costs a tick
 * launches at most the reference's batches (166 heightfields, 100 layers, 62 caves, 62 fills, 1 zone);
 * stage products stay in HBM (no per-stage host round trip) and a zone is eroded once its whole 24x24-chunk
 * gather window has layers, which makes every chunk a pure function of its coordinates (see mm_stream.inl).
 * The session window [cx0, cx0+nx) x [cz0, cz0+nz) bounds what the stream can generate. */
typedef struct MmgenStream MmgenStream;
typedef struct {
    int32_t heightfields, gatherHeightfields, layers, zonesEroded, caves, placements, gatherPlacements, filled, vbos;
    int32_t actionTimeLeft;   /* budget left after the tick */
    int32_t idle;             /* 1 when no queue holds work and no state changed: further ticks do nothing until the player moves */
    float deviceMs;           /* device time of the tick's launches (CUDA events) */
    int64_t meshVertices;     /* vertices produced by the tick's createVBOs pass (mmgen_stream_set_meshing) */
} MmgenTickStats;
/* reference ChunkState values reported by mmgen_stream_states (chunk.hpp:18-32) */
enum {
    MMGEN_CHUNK_EMPTY = 0, MMGEN_CHUNK_HAS_HEIGHTFIELD, MMGEN_CHUNK_NEEDS_LAYERS, MMGEN_CHUNK_HAS_LAYERS, MMGEN_CHUNK_NEEDS_EROSION,
    MMGEN_CHUNK_NEEDS_CAVES, MMGEN_CHUNK_NEEDS_FEATURE_PLACEMENTS, MMGEN_CHUNK_NEEDS_GATHER_FEATURE_PLACEMENTS,
    MMGEN_CHUNK_READY_TO_FILL, MMGEN_CHUNK_FILLED, MMGEN_CHUNK_NEEDS_VBOS, MMGEN_CHUNK_DRAWABLE
};
int mmgen_stream_create(int cx0, int cz0, int nx, int nz, MmgenStream** out);
int mmgen_stream_destroy(MmgenStream* s);
/* the backing world (downloads, device pointers, checksums); owned by the stream */
int mmgen_stream_world(MmgenStream* s, MmgenWorld** out);
/* chunkVbosGenRadius / chunkMaxGenRadius (terrain.cpp:64-65; defaults 16 and 40) */
int mmgen_stream_set_radii(MmgenStream* s, int vbosGenRadius, int maxGenRadius);
/* costs9 = {heightfield, gatherHeightfield, layers, erodeZone, caves, featurePlacements, gatherFeaturePlacements, fill,
 * createVbos} (terrain.cpp:72-80; NULL keeps the current ones), frame cap and refill rate (terrain.cpp:69-70) */
int mmgen_stream_set_costs(MmgenStream* s, const int32_t* costs9, int maxActionTimePerFrame, int totalActionTimePerSecond);
/* Terrain::setCurrentChunkPos(chunkPosFromPlayerPos(player)) (terrain.cpp:254-257, 1031-1034); block coordinates */
int mmgen_stream_set_player(MmgenStream* s, float playerX, float playerZ);
/* Terrain::tick(deltaTime); returns when the tick's results are complete on the device */
int mmgen_stream_tick(MmgenStream* s, float deltaTime, MmgenTickStats* out);
/* run Chunk::createVBOs on the device (mmgen_world_mesh) for the chunks that leave the VBO queue; off by default */
int mmgen_stream_set_meshing(MmgenStream* s, int enable);
/* (cx, cz) pairs of the chunks whose meshes are in the backing world's arena (chunk i for mmgen_world_mesh_device_ptrs / _download) */
int mmgen_stream_last_meshed(MmgenStream* s, int32_t* coords, int cap, int* n);
/* ChunkState of every window chunk, raster order */
int mmgen_stream_states(MmgenStream* s, uint8_t* out);
/* chunk coordinates (cx, cz) of the chunks filled since the last call, in fill order; *n = pairs written (<= cap) */
int mmgen_stream_take_filled(MmgenStream* s, int32_t* coords, int cap, int* n);
/* block volume of one filled chunk into host memory (98 304 bytes) */
int mmgen_stream_download_chunk(MmgenStream* s, int cx, int cz, uint8_t* out_blocks);

/* ---- meshing: Chunk::createVBOs (chunk.cu:1781-2003; a host loop in the reference that re-uploads the blocks' mesh after
 * Chunk::fill downloaded them) on the device, straight from the resident block volumes. Vertex = rendering/structs.hpp:25-31. */
typedef struct { float pos[3], nor[3], uv[2]; uint64_t m; } MmgenVertex;   /* 40 B; m = Mats (structs.hpp:7-14) */
/* meshes n filled chunks given as (cx, cz) pairs, in the reference's vertex / index order, into a device arena owned by
 * the world (valid until the next call); faces towards a neighbour chunk that is not filled are skipped exactly as the
 * reference skips a null neighbour (chunk.cu:1908-1911). out_counts[n][2] = {vertices, indices} per chunk (may be NULL). */
int mmgen_world_mesh(MmgenWorld* w, int n, const int32_t* chunkCoords, int32_t* out_counts);
/* chunk i of the last mmgen_world_mesh call: copy to host memory, or the device addresses (zero-copy hand-off to a renderer) */
int mmgen_world_mesh_download(MmgenWorld* w, int i, MmgenVertex* out_verts, uint32_t* out_idx);
int mmgen_world_mesh_device_ptrs(MmgenWorld* w, int i, void** verts, void** idx, int* nVerts, int* nIdx);
/* ---- hand-off to the path tracer: OptixRenderer::buildChunkAccel (optixRenderer.cpp:223-368) copies a chunk's host vertex /
 * index vectors to the device (initFromVector, :229-230) and describes them as one OptixBuildInput triangle array. For the
 * meshes of the last mmgen_world_mesh call the description can be filled from the device arena directly: one MmgenGasInput
 * per chunk holds exactly the triangleArray fields buildChunkAccel sets (vertex format float3 at stride sizeof(Vertex) = 40,
 * index format unsigned int3 at stride 12; integration/optix_handoff.cpp shows the binding and checks the layouts against
 * rendering/structs.hpp:25-31). *n = chunks of the last mesh call (may exceed cap). */
typedef struct {
    uint64_t vertexBuffer;            /* CUdeviceptr: MmgenVertex[numVertices] */
    uint32_t numVertices, vertexStrideInBytes;
    uint64_t indexBuffer;             /* CUdeviceptr: uint32[numIndexTriplets][3], indices relative to this chunk's vertices */
    uint32_t numIndexTriplets, indexStrideInBytes;
    int32_t cx, cz;                   /* chunk coordinates */
} MmgenGasInput;
int mmgen_world_mesh_gas_inputs(MmgenWorld* w, int cap, MmgenGasInput* out, int* n);
/* device time of the last mmgen_world_mesh call (both kernels + the count read-back), ms */
int mmgen_world_mesh_ms(MmgenWorld* w, float* out);

/* ---- measurement helpers (bench.py) */
/* switch per-kernel device timing on / off (CUDA event pairs around the hot kernels' launches, on their stream); clears the record */
int mmgen_kernel_timing(int enable);
/* summed device time and launch count per kernel since the last call; kernel i is named mmgen_kernel_name(i); *n = entries written */
int mmgen_kernel_times(int cap, float* out_ms, int32_t* out_launches, int* n);
const char* mmgen_kernel_name(int slot);
/* work counters of the cheap stages since the last reset, for their roofline figures (32 values): [0..23] S1 columns in which
 * surface biome b had weight > 0 (its height function ran), [24] S1 columns, [25] S2 fbm<5> evaluations, [26] S2 columns,
 * [27] S3 32x32 tiles swept, [28] S3 tile launches that returned at the quiet-tile test, [29] S6 gathered placements of the filled
 * chunks, [30] S6 (column, y) pairs inside the placements' clipped boxes that the placement scan looked at, [31] pairs that reached
 * a rasteriser. reset != 0 clears them afterwards. */
int mmgen_work_counters(uint64_t* out32, int reset);
/* tuning knob: queue slots per chunk for the rock voxels that k_fill_terrain hands to k_fill_rock (default and maximum 49 152;
 * <= 0 restores the default). Voxels that do not fit are finished in place: results never depend on this value. */
int mmgen_set_rock_queue_per_chunk(int slots);
/* measurement knob: a generate call that runs layers, erosion and caves together overlaps layers + erosion (side stream) with the
 * caves, which only read stage-1 products; serial != 0 runs them one after the other, as the reference's state machine does. The
 * products are identical either way (tests/test_gpu_parity.py::test_stage_overlap_is_result_neutral). */
int mmgen_set_serial_stages(int serial);
/* scheduling knob of a fill that spans several batches (2048 chunks each). 0: a batch's terrain / rock / lush passes and its placement
 * scan + decorators run one after the other on the world's stream. g in 1..8: the terrain / rock / lush passes of batch b + 1 run on a
 * second stream while the placement scan of batch b runs (they write different chunks' volumes), k_fill_rock's persistent grid sized g
 * CTAs per SM (the library starts with 8); g + 16: the same on the high-priority side stream. The products are identical in every mode
 * (tests/test_gpu_parity.py::test_fill_overlap_is_result_neutral). */
int mmgen_set_fill_overlap(int mode);
/* achieved FP32 FMA rate of this device (TFLOP/s, 8 independent FFMA chains per thread on every SM): roofline denominator */
int mmgen_measure_fp32_peak(float* out_tflops);
/* self-test: the packed-fp32 noise routines (two samples per call on sm_100's FFMA2 / FADD2 / FMUL2) against the scalar routines they
 * replace, at n pseudo-random positions; *out_mismatches = results that differ in any bit (must be 0) */
int mmgen_selftest_packed_noise(int n, uint32_t seed, uint64_t* out_mismatches);

#ifdef __cplusplus
}
#endif
#endif /* MMGEN_H */
