// TEST INFRASTRUCTURE. Headless driver of the reference's OWN chunk scheduler: the UNMODIFIED
// /root/reference/src/terrain/terrain.cpp (Terrain::tick, terrain.cpp:587-960) compiled with g++ (oracle/Makefile, target
// `terrain`: force-included ref_shim_terrain/msvc_compat.hpp for its one MSVC-ism, an OptixRenderer shim that counts calls)
// and linked with the unmodified chunk.cu object. Used by tests/test_stream.py on the GPU box to pin the re-hosted scheduler
// (mmgen_stream_*) to the real one tick by tick: batch sizes per stage, fill order, VBO builds.
//
// The generation entry points of chunk.cu are renamed in a copy of its object file (objcopy --redefine-sym, Makefile) and the
// definitions below take their names: each records its batch and forwards to the renamed original.
#include "terrain/terrain.hpp"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

// GLEW function pointers referenced by drawable.cpp; never called headless
extern "C" {
PFNGLDELETEBUFFERSPROC __glewDeleteBuffers = nullptr;
PFNGLGENBUFFERSPROC __glewGenBuffers = nullptr;
PFNGLBINDBUFFERPROC __glewBindBuffer = nullptr;
PFNGLBUFFERDATAPROC __glewBufferData = nullptr;
}

// stubs for what only Terrain::draw touches (never called here)
void ShaderProgram::setModelMat(const glm::mat4&) const {}
void ShaderProgram::draw(Drawable&) const {}
glm::vec3 Player::getForward() const { return glm::vec3(0, 0, 1); }
glm::vec3 Player::getPos() const { return glm::vec3(0); }

// the renamed originals (oracle/Makefile: REF_RENAME)
void ref_generateHeightfields(std::vector<Chunk*>&, ivec2*, ivec2*, float*, float*, float*, float*, cudaStream_t) asm("mmref_orig_generateHeightfields");
void ref_generateLayers(std::vector<Chunk*>&, float*, float*, float*, float*, ivec2*, ivec2*, float*, float*, cudaStream_t) asm("mmref_orig_generateLayers");
void ref_erodeZone(Zone*, float*, float*, float*, cudaStream_t) asm("mmref_orig_erodeZone");
void ref_generateCaves(std::vector<Chunk*>&, float*, float*, float*, float*, ivec2*, ivec2*, CaveLayer*, CaveLayer*, cudaStream_t) asm("mmref_orig_generateCaves");
void ref_fill(std::vector<Chunk*>&, float*, float*, float*, float*, float*, float*, CaveLayer*, CaveLayer*, FeaturePlacement*, CaveFeaturePlacement*, Block*,
              Block*, cudaStream_t) asm("mmref_orig_fill");
void ref_generateFeaturePlacements(Chunk*) asm("mmref_orig_generateFeaturePlacements");
void ref_gatherFeaturePlacements(Chunk*) asm("mmref_orig_gatherFeaturePlacements");
void ref_gatherHeightfield(Chunk*) asm("mmref_orig_gatherHeightfield");

#include "cuda/cudaUtils.hpp"

namespace {
// per tick: heightfields, gatherHeightfields, layers, zonesEroded, caves, placements, gatherPlacements, filled, vbos (MmgenTickStats order)
int g_counts[9];
std::vector<int> g_filled;      // (cx, cz) pairs in fill order
std::vector<int> g_eroded;      // (zone x, zone z) chunk coordinates of the zones in the order they were eroded
bool g_skipKernels = false;     // scheduler-only runs: the entry points keep their bookkeeping but launch nothing
Terrain* g_terrain = nullptr;
OptixRenderer g_optix;
}

// CudaUtils::checkCUDAError (src/cuda/cuda_utils.cpp:5-17: print and exit) is the one reference function replaced here, so that
// the scheduler can also run where there is no GPU (skipKernels: Terrain::initCuda's allocations fail and nothing uses them)
void CudaUtils::checkCUDAError(const char* msg, int line)
{
    const cudaError_t err = cudaGetLastError();
    if (err == cudaSuccess || g_skipKernels) return;
    if (line >= 0) std::fprintf(stderr, "Line %d: ", line);
    std::fprintf(stderr, "Cuda error: %s: %s.\n", msg, cudaGetErrorString(err));
    std::exit(EXIT_FAILURE);
}

void Chunk::generateHeightfields(std::vector<Chunk*>& chunks, ivec2* a, ivec2* b, float* c, float* d, float* e, float* f, cudaStream_t s)
{
    g_counts[0] += (int)chunks.size();
    if (!g_skipKernels) ref_generateHeightfields(chunks, a, b, c, d, e, f, s);
}
void Chunk::gatherHeightfield()
{
    ++g_counts[1];
    ref_gatherHeightfield(this);
}
void Chunk::generateLayers(std::vector<Chunk*>& chunks, float* a, float* b, float* c, float* d, ivec2* e, ivec2* f, float* g, float* h, cudaStream_t s)
{
    g_counts[2] += (int)chunks.size();
    if (!g_skipKernels) ref_generateLayers(chunks, a, b, c, d, e, f, g, h, s);
    else for (Chunk* ch : chunks) ch->gatheredHeightfield.clear();
}
void Chunk::erodeZone(Zone* zone, float* a, float* b, float* c, cudaStream_t s)
{
    ++g_counts[3];
    g_eroded.push_back(zone->worldChunkPos.x);
    g_eroded.push_back(zone->worldChunkPos.y);
    if (!g_skipKernels) ref_erodeZone(zone, a, b, c, s);
    else zone->gatheredChunks.clear();
}
void Chunk::generateCaves(std::vector<Chunk*>& chunks, float* a, float* b, float* c, float* d, ivec2* e, ivec2* f, CaveLayer* g, CaveLayer* h, cudaStream_t s)
{
    g_counts[4] += (int)chunks.size();
    if (!g_skipKernels) ref_generateCaves(chunks, a, b, c, d, e, f, g, h, s);
}
void Chunk::generateFeaturePlacements()
{
    ++g_counts[5];
    if (!g_skipKernels) ref_generateFeaturePlacements(this);
}
void Chunk::gatherFeaturePlacements()
{
    ++g_counts[6];
    ref_gatherFeaturePlacements(this);
}
void Chunk::fill(std::vector<Chunk*>& chunks, float* a, float* b, float* c, float* d, float* e, float* f, CaveLayer* g, CaveLayer* h, FeaturePlacement* i,
                 CaveFeaturePlacement* j, Block* k, Block* l, cudaStream_t s)
{
    g_counts[7] += (int)chunks.size();
    for (Chunk* ch : chunks) { g_filled.push_back(ch->worldChunkPos.x); g_filled.push_back(ch->worldChunkPos.y); }
    if (!g_skipKernels) ref_fill(chunks, a, b, c, d, e, f, g, h, i, j, k, l, s);
    else
        for (Chunk* ch : chunks)
        {
            // Chunk's arrays are uninitialised storage (chunk.hpp:59-72); Chunk::createVBOs will read these blocks
            std::memset(ch->blocks.data(), 0, sizeof(Block) * ch->blocks.size());
            ch->gatheredFeaturePlacements.clear();
            ch->gatheredCaveFeaturePlacements.clear();
        }
}

extern "C" {

// skipKernels != 0: the scheduler runs on empty data (no kernel launches, no CPU feature placement): its batch sizes, orders
// and budgets do not depend on the data
int mmrt_create(int device, int skipKernels)
{
    g_skipKernels = skipKernels != 0;
    if (cudaSetDevice(device) != cudaSuccess && !g_skipKernels) return 1;
    if (!g_skipKernels) BiomeUtils::init();
    delete g_terrain;
    g_terrain = new Terrain();
    g_terrain->setOptixRenderer(&g_optix);
    g_terrain->init();
    g_filled.clear();
    g_eroded.clear();
    return (cudaGetLastError() == cudaSuccess || g_skipKernels) ? 0 : 2;
}

void mmrt_set_player_chunk(int cx, int cz) { g_terrain->setCurrentChunkPos(ivec2(cx, cz)); }

// one Terrain::tick; out9 = this tick's counts (MmgenTickStats order; vbos = OptixRenderer::buildChunkAccel calls)
void mmrt_tick(float deltaTime, int* out9)
{
    std::memset(g_counts, 0, sizeof(g_counts));
    const int built0 = g_optix.built;
    g_terrain->tick(deltaTime);
    g_counts[8] = g_optix.built - built0;
    std::memcpy(out9, g_counts, sizeof(g_counts));
}

// chunk coordinates filled since the last call, in fill order; returns the number of pairs written
int mmrt_take_filled(int* coords, int cap)
{
    const int n = std::min<int>(cap, (int)g_filled.size() / 2);
    std::memcpy(coords, g_filled.data(), (size_t)n * 2 * sizeof(int));
    g_filled.erase(g_filled.begin(), g_filled.begin() + 2 * n);
    return n;
}

// zones eroded since the last call (their corner chunk coordinates), in erosion order. The reference keeps the zones it has
// to re-test in an unordered_set<Zone*> (terrain.hpp), so which of several ready zones is eroded first depends on heap addresses
int mmrt_take_eroded(int* coords, int cap)
{
    const int n = std::min<int>(cap, (int)g_eroded.size() / 2);
    std::memcpy(coords, g_eroded.data(), (size_t)n * 2 * sizeof(int));
    g_eroded.erase(g_eroded.begin(), g_eroded.begin() + 2 * n);
    return n;
}

// chunks handed to OptixRenderer::buildChunkAccel since the last call (the VBO stage), in order
int mmrt_take_built(int* coords, int cap)
{
    const int n = std::min<int>(cap, (int)g_optix.builtChunks.size());
    for (int i = 0; i < n; ++i) { coords[2 * i] = g_optix.builtChunks[i]->worldChunkPos.x; coords[2 * i + 1] = g_optix.builtChunks[i]->worldChunkPos.y; }
    g_optix.builtChunks.erase(g_optix.builtChunks.begin(), g_optix.builtChunks.begin() + n);
    return n;
}

int mmrt_num_drawable() { return (int)g_terrain->getDrawableChunks().size(); }
void mmrt_get_blocks(int cx, int cz, unsigned char* out)
{
    for (Chunk* c : g_terrain->getDrawableChunks())
        if (c->worldChunkPos.x == cx && c->worldChunkPos.y == cz) { std::memcpy(out, c->blocks.data(), 98304); return; }
}
void mmrt_destroy() { delete g_terrain; g_terrain = nullptr; }

}  // extern "C"
