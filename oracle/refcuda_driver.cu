// TEST INFRASTRUCTURE - NOT PRODUCT CODE.
//
// Headless driver around the UNMODIFIED reference pipeline (/root/reference/src/terrain/chunk.cu,
// compiled from where it lies by oracle/Makefile into oracle/_ref/). It plays the role of
// Terrain::tick (/root/reference/src/terrain/terrain.cpp:587-960; that file is MSVC-only and its
// scheduler is out of scope): it owns the staging buffers (terrain.cpp:111-185), creates the
// Zone/Chunk graph for a rectangular window of chunks, walks the ChunkState machine in pipeline
// order with the reference's batch caps, and copies every intermediate product out through a C ABI
// so tests/ and bench.py can use the reference's own CUDA kernels as (a) the binding parity oracle
// on the GPU box and (b) the "reference CUDA on the same B200" baseline.
//
// Nothing in mega-minecraft_b200/ may link or call this.
#define private public  // the harness needs Chunk::featurePlacements etc.; layout is unaffected
#include "terrain/chunk.hpp"
#undef private
#include "terrain/terrain.hpp"
#include "util/enums.hpp"

#include <chrono>
#include <cstdio>
#include <algorithm>
#include <cstring>
#include <map>
#include <memory>
#include <vector>

namespace BiomeUtils { void init(); }

// ---- GLEW function pointers referenced by drawable.cpp; never called headless ----
extern "C" {
PFNGLDELETEBUFFERSPROC __glewDeleteBuffers = nullptr;
PFNGLGENBUFFERSPROC __glewGenBuffers = nullptr;
PFNGLBINDBUFFERPROC __glewBindBuffer = nullptr;
PFNGLBUFFERDATAPROC __glewBufferData = nullptr;
}

namespace {

// batch caps of one tick: terrain.cpp:111-129 with the costs at terrain.cpp:71-82
constexpr int kMaxHeightfieldBatch = 166;
constexpr int kMaxLayersBatch = 100;
constexpr int kMaxCavesBatch = 62;
constexpr int kMaxFillBatch = 62;

struct Buffers
{
    Block* host_blocks = nullptr; Block* dev_blocks = nullptr;
    FeaturePlacement* dev_fp = nullptr; CaveFeaturePlacement* dev_cfp = nullptr;
    float* host_hf = nullptr; float* dev_hf = nullptr;
    float* host_bw = nullptr; float* dev_bw = nullptr;
    ivec2* host_pos = nullptr; ivec2* dev_pos = nullptr;
    float* host_layers = nullptr; float* dev_layers = nullptr;
    CaveLayer* host_cl = nullptr; CaveLayer* dev_cl = nullptr;
    float* host_gathered = nullptr; float* dev_gathered = nullptr; float* dev_accum = nullptr;
    cudaStream_t stream = nullptr;
    bool ready = false;
} B;

struct World
{
    int x0 = 0, z0 = 0, nx = 0, nz = 0;
    std::map<std::pair<int, int>, std::unique_ptr<Zone>> zones;  // key = zone origin in chunks
    std::vector<Chunk*> grid;                                    // nz * nx, raster
    std::vector<unsigned char> stage;                            // furthest completed stage per chunk
    std::vector<std::vector<FeaturePlacement>> gatheredFp;       // saved before fill() clears them
    std::vector<std::vector<CaveFeaturePlacement>> gatheredCfp;
    double ms[8] = {0};                                          // wall ms per stage
    Chunk* at(int cx, int cz) const
    {
        if (cx < x0 || cz < z0 || cx >= x0 + nx || cz >= z0 + nz) return nullptr;
        return grid[(cz - z0) * nx + (cx - x0)];
    }
    int idx(const Chunk* c) const { return (c->worldChunkPos.y - z0) * nx + (c->worldChunkPos.x - x0); }
} W;

int floorDiv(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }

double nowMs()
{
    using namespace std::chrono;
    return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

void allocBuffers()
{
    if (B.ready) return;
    cudaMallocHost((void**)&B.host_blocks, (size_t)kMaxFillBatch * devBlocksSize);
    cudaMalloc((void**)&B.dev_blocks, (size_t)kMaxFillBatch * devBlocksSize);
    cudaMalloc((void**)&B.dev_fp, (size_t)kMaxFillBatch * devFeaturePlacementsSize * sizeof(FeaturePlacement));
    cudaMalloc((void**)&B.dev_cfp, (size_t)kMaxFillBatch * devCaveFeaturePlacementsSize * sizeof(CaveFeaturePlacement));
    cudaMallocHost((void**)&B.host_hf, (size_t)kMaxHeightfieldBatch * devHeightfieldSize * sizeof(float));
    cudaMalloc((void**)&B.dev_hf, (size_t)kMaxHeightfieldBatch * devHeightfieldSize * sizeof(float));
    cudaMallocHost((void**)&B.host_bw, (size_t)kMaxHeightfieldBatch * devBiomeWeightsSize * sizeof(float));
    cudaMalloc((void**)&B.dev_bw, (size_t)kMaxHeightfieldBatch * devBiomeWeightsSize * sizeof(float));
    cudaMallocHost((void**)&B.host_pos, (size_t)kMaxHeightfieldBatch * sizeof(ivec2));
    cudaMalloc((void**)&B.dev_pos, (size_t)kMaxHeightfieldBatch * sizeof(ivec2));
    cudaMallocHost((void**)&B.host_layers, (size_t)kMaxLayersBatch * devLayersSize * sizeof(float));
    cudaMalloc((void**)&B.dev_layers, (size_t)kMaxLayersBatch * devLayersSize * sizeof(float));
    cudaMallocHost((void**)&B.host_cl, (size_t)kMaxCavesBatch * devCaveLayersSize * sizeof(CaveLayer));
    cudaMalloc((void**)&B.dev_cl, (size_t)kMaxCavesBatch * devCaveLayersSize * sizeof(CaveLayer));
    cudaMallocHost((void**)&B.host_gathered, (size_t)devGatheredLayersSize * sizeof(float));
    cudaMalloc((void**)&B.dev_gathered, (size_t)devGatheredLayersSize * sizeof(float));
    cudaMalloc((void**)&B.dev_accum, (size_t)devAccumulatedHeightsSize * sizeof(float));
    cudaStreamCreate(&B.stream);
    B.ready = true;
}

template <class F>
void forBatches(std::vector<Chunk*>& all, int cap, F f)
{
    for (size_t i = 0; i < all.size(); i += cap)
    {
        std::vector<Chunk*> batch(all.begin() + i, all.begin() + std::min(all.size(), i + cap));
        f(batch);
    }
}

}  // namespace

extern "C" {

// Stage codes reported by mmref_stage(): 0 none, 1 heightfield, 2 layers, 3 eroded, 4 caves,
// 5 feature placements, 6 gathered + filled (+decorators).
int mmref_init(int device)
{
    if (cudaSetDevice(device) != cudaSuccess) return 1;
    BiomeUtils::init();
    allocBuffers();
    return cudaGetLastError() == cudaSuccess ? 0 : 2;
}

// Order in which the window's zones are eroded: 0 = ascending (zone x, zone z), 1 = descending. The reference erodes a zone
// from whatever its 24x24-chunk gather window holds at that moment (copyLayers reads chunk->layers, chunk.cu:603-656, and
// erodeZone writes the centre chunks' layers back in place, :711-721), so a zone eroded AFTER a neighbour sees that
// neighbour's already-eroded pad: its rim depends on the order, which in the game is the player's path (terrain.cpp:471-566).
static int g_zoneOrder = 0;
void mmref_set_zone_order(int descending) { g_zoneOrder = descending ? 1 : 0; }

// Runs the reference pipeline over chunk window [x0,x0+nx) x [z0,z0+nz) up to `lastStage`
// (1..6), each chunk as far as the reference state machine allows inside that window.
// unwrittenLayerFill: bit pattern written over dev_layers before every generateLayers call so
// the forward layers the kernel never writes (chunk.cu:387-390) are recognisable.
int mmref_generate(int x0, int z0, int nx, int nz, int lastStage, unsigned int unwrittenLayerFill)
{
    allocBuffers();
    W = World();
    W.x0 = x0; W.z0 = z0; W.nx = nx; W.nz = nz;
    W.grid.assign((size_t)nx * nz, nullptr);
    W.stage.assign((size_t)nx * nz, 0);
    W.gatheredFp.assign((size_t)nx * nz, {});
    W.gatheredCfp.assign((size_t)nx * nz, {});

    // Zone / Chunk graph (terrain.cpp:254-420 restated for a fixed window)
    for (int cz = z0; cz < z0 + nz; ++cz)
        for (int cx = x0; cx < x0 + nx; ++cx)
        {
            const int zx = floorDiv(cx, ZONE_SIZE) * ZONE_SIZE, zz = floorDiv(cz, ZONE_SIZE) * ZONE_SIZE;
            auto& zoneUptr = W.zones[{zx, zz}];
            if (!zoneUptr) zoneUptr = std::make_unique<Zone>(ivec2(zx, zz));
            auto chunkUptr = std::make_unique<Chunk>(ivec2(cx, cz));
            chunkUptr->zonePtr = zoneUptr.get();
            W.grid[(cz - z0) * nx + (cx - x0)] = chunkUptr.get();
            zoneUptr->chunks[(cx - zx) + ZONE_SIZE * (cz - zz)] = std::move(chunkUptr);
        }
    for (Chunk* c : W.grid)
        for (int i = 0; i < 4; ++i)
        {
            const ivec3 d = DirectionEnums::dirVecs[i];
            c->neighbors[i] = W.at(c->worldChunkPos.x + d.x, c->worldChunkPos.y + d.z);
        }

    // S1
    double t = nowMs();
    {
        std::vector<Chunk*> all(W.grid.begin(), W.grid.end());
        forBatches(all, kMaxHeightfieldBatch, [&](std::vector<Chunk*>& batch) {
            for (Chunk* c : batch) c->setState(ChunkState::HAS_HEIGHTFIELD);
            Chunk::generateHeightfields(batch, B.host_pos, B.dev_pos, B.host_hf, B.dev_hf, B.host_bw, B.dev_bw, B.stream);
        });
        for (Chunk* c : W.grid) W.stage[W.idx(c)] = 1;
    }
    W.ms[1] = nowMs() - t;
    if (lastStage < 2) return 0;

    // S2a + S2
    t = nowMs();
    for (Chunk* c : W.grid) c->gatherHeightfield();
    {
        std::vector<Chunk*> ready;
        for (Chunk* c : W.grid) if (c->getState() == ChunkState::NEEDS_LAYERS) ready.push_back(c);
        forBatches(ready, kMaxLayersBatch, [&](std::vector<Chunk*>& batch) {
            for (Chunk* c : batch) c->setState(ChunkState::HAS_LAYERS);
            std::vector<unsigned int> fill((size_t)batch.size() * devLayersSize, unwrittenLayerFill);
            cudaMemcpy(B.dev_layers, fill.data(), fill.size() * 4, cudaMemcpyHostToDevice);
            Chunk::generateLayers(batch, B.host_hf, B.dev_hf, B.host_bw, B.dev_bw, B.host_pos, B.dev_pos, B.host_layers, B.dev_layers, B.stream);
            for (Chunk* c : batch) W.stage[W.idx(c)] = 2;
        });
    }
    W.ms[2] = nowMs() - t;
    if (lastStage < 3) return 0;

    // S3: zones whose 24x24 window lies in the window and has layers (terrain.cpp:471-522)
    t = nowMs();
    std::vector<Zone*> eroded, zoneOrder;
    for (auto& kv : W.zones) zoneOrder.push_back(kv.second.get());
    if (g_zoneOrder) std::reverse(zoneOrder.begin(), zoneOrder.end());
    for (Zone* zone : zoneOrder)
    {
        bool ok = true;
        zone->gatheredChunks.assign(ZONE_SIZE * ZONE_SIZE * 4, nullptr);
        for (int gz = 0; gz < 2 * ZONE_SIZE && ok; ++gz)
            for (int gx = 0; gx < 2 * ZONE_SIZE && ok; ++gx)
            {
                Chunk* c = W.at(zone->worldChunkPos.x - ZONE_SIZE / 2 + gx, zone->worldChunkPos.y - ZONE_SIZE / 2 + gz);
                if (c == nullptr || c->getState() < ChunkState::HAS_LAYERS) ok = false;
                else zone->gatheredChunks[gx + 2 * ZONE_SIZE * gz] = c;
            }
        if (!ok) { zone->gatheredChunks.clear(); continue; }
        Chunk::erodeZone(zone, B.host_gathered, B.dev_gathered, B.dev_accum, B.stream);
        for (const auto& c : zone->chunks) { c->setState(ChunkState::NEEDS_CAVES); W.stage[W.idx(c.get())] = 3; }
        eroded.push_back(zone);
    }
    W.ms[3] = nowMs() - t;
    if (lastStage < 4) return 0;

    // S4
    t = nowMs();
    {
        std::vector<Chunk*> ready;
        for (Chunk* c : W.grid) if (c->getState() == ChunkState::NEEDS_CAVES) ready.push_back(c);
        forBatches(ready, kMaxCavesBatch, [&](std::vector<Chunk*>& batch) {
            for (Chunk* c : batch) c->setState(ChunkState::NEEDS_FEATURE_PLACEMENTS);
            Chunk::generateCaves(batch, B.host_hf, B.dev_hf, B.host_bw, B.dev_bw, B.host_pos, B.dev_pos, B.host_cl, B.dev_cl, B.stream);
            for (Chunk* c : batch) W.stage[W.idx(c)] = 4;
        });
    }
    W.ms[4] = nowMs() - t;
    if (lastStage < 5) return 0;

    // S5a
    t = nowMs();
    for (Chunk* c : W.grid)
        if (c->getState() == ChunkState::NEEDS_FEATURE_PLACEMENTS)
        {
            c->generateFeaturePlacements();
            c->setState(ChunkState::NEEDS_GATHER_FEATURE_PLACEMENTS);
            W.stage[W.idx(c)] = 5;
        }
    W.ms[5] = nowMs() - t;
    if (lastStage < 6) return 0;

    // S5b
    t = nowMs();
    for (Chunk* c : W.grid)
        if (c->getState() == ChunkState::NEEDS_GATHER_FEATURE_PLACEMENTS) c->gatherFeaturePlacements();
    W.ms[6] = nowMs() - t;

    // S6 (+S6b inside fill)
    t = nowMs();
    {
        std::vector<Chunk*> ready;
        for (Chunk* c : W.grid)
            if (c->getState() == ChunkState::READY_TO_FILL)
            {
                ready.push_back(c);
                W.gatheredFp[W.idx(c)] = c->gatheredFeaturePlacements;
                W.gatheredCfp[W.idx(c)] = c->gatheredCaveFeaturePlacements;
            }
        forBatches(ready, kMaxFillBatch, [&](std::vector<Chunk*>& batch) {
            for (Chunk* c : batch) c->setState(ChunkState::FILLED);
            Chunk::fill(batch, B.host_hf, B.dev_hf, B.host_bw, B.dev_bw, B.host_layers, B.dev_layers, B.host_cl, B.dev_cl,
                        B.dev_fp, B.dev_cfp, B.host_blocks, B.dev_blocks, B.stream);
            for (Chunk* c : batch) W.stage[W.idx(c)] = 6;
        });
    }
    W.ms[7] = nowMs() - t;
    return cudaDeviceSynchronize() == cudaSuccess ? 0 : 3;
}

double mmref_stage_ms(int stage) { return (stage >= 0 && stage < 8) ? W.ms[stage] : -1.0; }

// All getters address chunks by window-raster index i = (cz - z0) * nx + (cx - x0).
int mmref_stage(int i) { return W.stage[i]; }
void mmref_get_heightfield(int i, float* out256) { std::memcpy(out256, W.grid[i]->heightfield.data(), 256 * 4); }
void mmref_get_biome_weights(int i, float* out) { std::memcpy(out, W.grid[i]->biomeWeights.data(), 256 * numBiomes * 4); }
void mmref_get_layers(int i, float* out) { std::memcpy(out, W.grid[i]->layers.data(), 256 * numMaterials * 4); }
void mmref_get_cave_layers(int i, void* out) { std::memcpy(out, W.grid[i]->caveLayers.data(), 256 * MAX_CAVE_LAYERS_PER_COLUMN * sizeof(CaveLayer)); }
void mmref_get_blocks(int i, unsigned char* out) { std::memcpy(out, W.grid[i]->blocks.data(), 98304); }
int mmref_num_features(int i) { return (int)W.grid[i]->featurePlacements.size(); }
int mmref_num_cave_features(int i) { return (int)W.grid[i]->caveFeaturePlacements.size(); }
void mmref_get_features(int i, void* out) { std::memcpy(out, W.grid[i]->featurePlacements.data(), W.grid[i]->featurePlacements.size() * sizeof(FeaturePlacement)); }
void mmref_get_cave_features(int i, void* out) { std::memcpy(out, W.grid[i]->caveFeaturePlacements.data(), W.grid[i]->caveFeaturePlacements.size() * sizeof(CaveFeaturePlacement)); }
int mmref_num_gathered_features(int i) { return (int)W.gatheredFp[i].size(); }
int mmref_num_gathered_cave_features(int i) { return (int)W.gatheredCfp[i].size(); }
void mmref_get_gathered_features(int i, void* out) { std::memcpy(out, W.gatheredFp[i].data(), W.gatheredFp[i].size() * sizeof(FeaturePlacement)); }
void mmref_get_gathered_cave_features(int i, void* out) { std::memcpy(out, W.gatheredCfp[i].data(), W.gatheredCfp[i].size() * sizeof(CaveFeaturePlacement)); }
int mmref_sizeof_feature() { return (int)sizeof(FeaturePlacement); }
int mmref_sizeof_cave_feature() { return (int)sizeof(CaveFeaturePlacement); }

}  // extern "C"

// ---- meshing oracle: the reference's own Chunk::createVBOs (chunk.cu:1781-2003, a host function) on block volumes
// supplied by the caller. No CUDA device is needed for this entry point (it runs in the CPU tests as well).
// blocks5: centre chunk, then the neighbours in the reference's neighbors[] order: +z, +x, -z, -x (chunk.hpp:57,
// enums.hpp:41-48), 98304 bytes each; present[i] == 0 leaves neighbors[i] null. Returns the vertex count, or -1 if the
// caps are too small; *nIdxOut = index count. vertsOut receives the reference's Vertex records verbatim (40 bytes each).
namespace BlockUtils { void init(); }
extern "C" int mmref_vertex_size() { return (int)sizeof(Vertex); }
extern "C" int mmref_mesh_chunk(int cx, int cz, const unsigned char* blocks5, const int* present, unsigned char* vertsOut, int vertCap,
                                unsigned int* idxOut, int idxCap, int* nIdxOut)
{
    static bool inited = false;
    if (!inited) { BlockUtils::init(); inited = true; }
    std::unique_ptr<Chunk> c[5];
    const ivec2 off[5] = {ivec2(0, 0), ivec2(0, 1), ivec2(1, 0), ivec2(0, -1), ivec2(-1, 0)};
    for (int i = 0; i < 5; ++i)
    {
        if (i > 0 && !present[i - 1]) continue;
        c[i] = std::make_unique<Chunk>(ivec2(cx, cz) + off[i]);
        std::memcpy(c[i]->blocks.data(), blocks5 + (size_t)i * 98304, 98304);
    }
    for (int i = 0; i < 4; ++i) c[0]->neighbors[i] = c[i + 1].get();
    c[0]->createVBOs();
    const int nv = (int)c[0]->verts.size(), ni = (int)c[0]->idx.size();
    *nIdxOut = ni;
    if (nv > vertCap || ni > idxCap) return -1;
    std::memcpy(vertsOut, c[0]->verts.data(), (size_t)nv * sizeof(Vertex));
    std::memcpy(idxOut, c[0]->idx.data(), (size_t)ni * sizeof(unsigned int));
    return nv;
}
