// Reference-side binding of the B200-native generation path: definitions of the generation entry points
// of `Chunk` (/root/reference/src/terrain/chunk.hpp:100-172) that forward to the C ABI in include/mmgen.h.
//
// A maintainer of the reference adds this file to the build, links libmmgen.so, and removes (or
// #ifdef's out) the definitions of the same functions in src/terrain/chunk.cu: the five batch entry points (lines 187-229,
// 417-469, 658-723, 939-993, 1518-1632) and the two CPU passes of stage 5 (generateFeaturePlacements :1147-1156,
// otherChunkGatherFeaturePlacements :1169-1187, the data movement of gatherFeaturePlacements). Terrain::tick (terrain.cpp:587-960) and everything else in the
// application stay as they are: same signatures, same ownership (the staging buffers Terrain passes in
// are used as the host-side batch buffers; the dev_* pointers and the stream are unused because the
// library owns its device memory), same result contract (per-chunk host arrays complete on return).
//
// This file contains no generation logic. It is compiled against the reference's own headers by
// `make -C oracle adapter` (test infrastructure), which also links it over the reference's object file to
// run the reference state machine on the new kernels (tests/test_gpu_parity.py::test_adapter_*).
#include "terrain/chunk.hpp"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "mmgen.h"

namespace {

// the reference's convention: print and exit (src/cuda/cuda_utils.cpp:5-17)
void check(int rc, const char* what)
{
    if (rc == 0) return;
    std::fprintf(stderr, "%s failed: %s\n", what, mmgen_last_error());
    std::exit(EXIT_FAILURE);
}

int g_calls[7];      // how often each entry point below ran (mmadapter_calls: lets a test see that the overrides are the ones in use)

void ensureInit()
{
    static bool ready = false;
    if (ready) return;
    int device = 0;
    cudaGetDevice(&device);
    check(mmgen_init(device), "mmgen_init");      // replaces BiomeUtils::init()'s table upload for this path
    ready = true;
}

}  // namespace

extern "C" int mmadapter_calls(int entryPoint) { return (entryPoint >= 0 && entryPoint < 7) ? g_calls[entryPoint] : -1; }

void Chunk::generateHeightfields(std::vector<Chunk*>& chunks, ivec2* host_chunkWorldBlockPositions, ivec2*, float* host_heightfields, float*,
                                 float* host_biomeWeights, float*, cudaStream_t)
{
    ensureInit();
    ++g_calls[0];
    const int n = (int)chunks.size();
    for (int i = 0; i < n; ++i) host_chunkWorldBlockPositions[i] = ivec2(chunks[i]->worldBlockPos.x, chunks[i]->worldBlockPos.z);
    check(mmgen_heightfields(n, (const int32_t*)host_chunkWorldBlockPositions, host_heightfields, host_biomeWeights), "Chunk::generateHeightfield()");
    for (int i = 0; i < n; ++i)
    {
        std::memcpy(chunks[i]->heightfield.data(), host_heightfields + 256 * i, 256 * sizeof(float));
        std::memcpy(chunks[i]->biomeWeights.data(), host_biomeWeights + devBiomeWeightsSize * i, devBiomeWeightsSize * sizeof(float));
    }
}

void Chunk::generateLayers(std::vector<Chunk*>& chunks, float* host_heightfields, float*, float* host_biomeWeights, float*,
                           ivec2* host_chunkWorldBlockPositions, ivec2*, float* host_layers, float*, cudaStream_t)
{
    ensureInit();
    ++g_calls[1];
    const int n = (int)chunks.size();
    for (int i = 0; i < n; ++i)
    {
        Chunk* c = chunks[i];
        std::memcpy(host_heightfields + i * devHeightfieldSize, c->gatheredHeightfield.data(), devHeightfieldSize * sizeof(float));   // 18x18
        c->gatheredHeightfield.clear();
        std::memcpy(host_biomeWeights + i * devBiomeWeightsSize, c->biomeWeights.data(), devBiomeWeightsSize * sizeof(float));
        host_chunkWorldBlockPositions[i] = ivec2(c->worldBlockPos.x, c->worldBlockPos.z);
    }
    check(mmgen_layers(n, (const int32_t*)host_chunkWorldBlockPositions, host_heightfields, host_biomeWeights, host_layers), "Chunk::generateLayers()");
    for (int i = 0; i < n; ++i) std::memcpy(chunks[i]->layers.data(), host_layers + i * devLayersSize, devLayersSize * sizeof(float));
}

void Chunk::erodeZone(Zone* zonePtr, float* host_gatheredLayers, float*, float*, cudaStream_t)
{
    ensureInit();
    ++g_calls[2];
    // copyLayers(zone, gathered, true) (chunk.cu:603-656): 8 loose layer planes + the heightfield plane of the 24x24-chunk window
    constexpr int side = EROSION_GRID_SIDE_LENGTH_BLOCKS, cols = EROSION_GRID_NUM_COLS;
    for (int cz = 0; cz < ZONE_SIZE * 2; ++cz)
        for (int cx = 0; cx < ZONE_SIZE * 2; ++cx)
        {
            Chunk* c = zonePtr->gatheredChunks[cx + ZONE_SIZE * 2 * cz];
            for (int l = numStratifiedMaterials; l <= numMaterials; ++l)
                for (int bz = 0; bz < 16; ++bz)
                {
                    const float* src = (l == numMaterials) ? c->heightfield.data() + 16 * bz : c->layers.data() + 16 * bz + 256 * l;
                    std::memcpy(host_gatheredLayers + cx * 16 + side * (cz * 16 + bz) + cols * (l - numStratifiedMaterials), src, 16 * sizeof(float));
                }
        }
    zonePtr->gatheredChunks.clear();
    // relaxed planes come back in place of the first 8 planes (the 9th, the heightfield, is not needed afterwards)
    int sweeps = 0;
    check(mmgen_erode_zone(host_gatheredLayers, host_gatheredLayers, &sweeps), "Chunk::erodeZone()");
    // copyLayers(zone, gathered, false): centre 12x12 chunks
    for (int cz = 0; cz < ZONE_SIZE; ++cz)
        for (int cx = 0; cx < ZONE_SIZE; ++cx)
        {
            Chunk* c = zonePtr->chunks[cx + ZONE_SIZE * cz].get();
            for (int l = numStratifiedMaterials; l < numMaterials; ++l)
                for (int bz = 0; bz < 16; ++bz)
                    std::memcpy(c->layers.data() + 16 * bz + 256 * l,
                                host_gatheredLayers + (cx + ZONE_SIZE / 2) * 16 + side * ((cz + ZONE_SIZE / 2) * 16 + bz) + cols * (l - numStratifiedMaterials),
                                16 * sizeof(float));
        }
    for (const auto& chunkPtr : zonePtr->chunks) chunkPtr->fixBackwardStratifiedLayers();
}

void Chunk::generateCaves(std::vector<Chunk*>& chunks, float* host_heightfields, float*, float* host_biomeWeights, float*,
                          ivec2* host_chunkWorldBlockPositions, ivec2*, CaveLayer* host_caveLayers, CaveLayer*, cudaStream_t)
{
    ensureInit();
    ++g_calls[3];
    const int n = (int)chunks.size();
    for (int i = 0; i < n; ++i)
    {
        Chunk* c = chunks[i];
        std::memcpy(host_heightfields + i * 256, c->heightfield.data(), 256 * sizeof(float));
        c->gatheredHeightfield.clear();
        std::memcpy(host_biomeWeights + i * devBiomeWeightsSize, c->biomeWeights.data(), devBiomeWeightsSize * sizeof(float));
        host_chunkWorldBlockPositions[i] = ivec2(c->worldBlockPos.x, c->worldBlockPos.z);
    }
    static_assert(sizeof(CaveLayer) == sizeof(MmgenCaveLayer), "CaveLayer wire layout");
    check(mmgen_caves(n, (const int32_t*)host_chunkWorldBlockPositions, host_heightfields, host_biomeWeights, (MmgenCaveLayer*)host_caveLayers),
          "Chunk::generateCaves()");
    for (int i = 0; i < n; ++i) std::memcpy(chunks[i]->caveLayers.data(), host_caveLayers + i * devCaveLayersSize, devCaveLayersSize * sizeof(CaveLayer));
}

void Chunk::fill(std::vector<Chunk*>& chunks, float* host_heightfields, float*, float* host_biomeWeights, float*, float* host_layers, float*,
                 CaveLayer* host_caveLayers, CaveLayer*, FeaturePlacement*, CaveFeaturePlacement*, Block* host_blocks, Block*, cudaStream_t)
{
    ensureInit();
    ++g_calls[4];
    static_assert(sizeof(FeaturePlacement) == sizeof(MmgenFeaturePlacement) && sizeof(CaveFeaturePlacement) == sizeof(MmgenCaveFeaturePlacement),
                  "placement wire layouts");
    const int n = (int)chunks.size();
    // the reference uploads each chunk's gathered lists separately (chunk.cu:1576-1601); the C ABI takes them packed with one stride
    static std::vector<FeaturePlacement> fp;
    static std::vector<CaveFeaturePlacement> cfp;
    static std::vector<int32_t> counts, origins;
    size_t strideF = 1, strideC = 1;
    for (Chunk* c : chunks)
    {
        strideF = std::max(strideF, std::min<size_t>(c->gatheredFeaturePlacements.size(), MAX_GATHERED_FEATURES_PER_CHUNK));
        strideC = std::max(strideC, std::min<size_t>(c->gatheredCaveFeaturePlacements.size(), MAX_GATHERED_CAVE_FEATURES_PER_CHUNK));
    }
    fp.assign(strideF * n, FeaturePlacement{Feature::NONE});
    cfp.assign(strideC * n, CaveFeaturePlacement{CaveFeature::NONE});
    counts.resize(2 * n);
    origins.resize(2 * n);
    for (int i = 0; i < n; ++i)
    {
        Chunk* c = chunks[i];
        std::memcpy(host_heightfields + i * 256, c->heightfield.data(), 256 * sizeof(float));
        std::memcpy(host_biomeWeights + i * devBiomeWeightsSize, c->biomeWeights.data(), devBiomeWeightsSize * sizeof(float));
        std::memcpy(host_layers + i * devLayersSize, c->layers.data(), devLayersSize * sizeof(float));
        std::memcpy(host_caveLayers + i * devCaveLayersSize, c->caveLayers.data(), devCaveLayersSize * sizeof(CaveLayer));
        const size_t nf = std::min<size_t>(c->gatheredFeaturePlacements.size(), MAX_GATHERED_FEATURES_PER_CHUNK);
        const size_t nc = std::min<size_t>(c->gatheredCaveFeaturePlacements.size(), MAX_GATHERED_CAVE_FEATURES_PER_CHUNK);
        std::copy_n(c->gatheredFeaturePlacements.begin(), nf, fp.begin() + strideF * i);
        std::copy_n(c->gatheredCaveFeaturePlacements.begin(), nc, cfp.begin() + strideC * i);
        c->gatheredFeaturePlacements.clear();
        c->gatheredCaveFeaturePlacements.clear();
        counts[2 * i] = (int32_t)nf; counts[2 * i + 1] = (int32_t)nc;
        origins[2 * i] = c->worldBlockPos.x; origins[2 * i + 1] = c->worldBlockPos.z;
    }
    // includes Chunk::placeDecorators (chunk.cu:1628, 1679-1747), which the reference runs on the CPU after the download
    check(mmgen_fill(n, origins.data(), host_heightfields, host_biomeWeights, host_layers, (const MmgenCaveLayer*)host_caveLayers,
                     (const MmgenFeaturePlacement*)fp.data(), (const MmgenCaveFeaturePlacement*)cfp.data(), counts.data(), (int)strideF, (int)strideC,
                     (uint8_t*)host_blocks),
          "Chunk::fill()");
    for (int i = 0; i < n; ++i) std::memcpy(chunks[i]->blocks.data(), host_blocks + i * devBlocksSize, devBlocksSize * sizeof(Block));
}

// ---- stage 5 (CPU passes in the reference): per-chunk member functions, so one chunk per call
void Chunk::generateFeaturePlacements()
{
    ensureInit();
    ++g_calls[5];
    static std::vector<MmgenFeaturePlacement> f(256);
    static std::vector<MmgenCaveFeaturePlacement> cf(MMGEN_MAX_CAVE_FEATURES);
    const int32_t origin[2] = {worldBlockPos.x, worldBlockPos.z};
    int32_t counts[2] = {0, 0};
    // a chunk's own cave list is only ever consumed up to the gather cap (it is first or later in every concatenation)
    check(mmgen_feature_placements(1, origin, heightfield.data(), biomeWeights.data(), layers.data(), (const MmgenCaveLayer*)caveLayers.data(),
                                   MMGEN_MAX_CAVE_FEATURES, f.data(), cf.data(), counts),
          "Chunk::generateFeaturePlacements()");
    const FeaturePlacement* pf = (const FeaturePlacement*)f.data();
    const CaveFeaturePlacement* pc = (const CaveFeaturePlacement*)cf.data();
    featurePlacements.assign(pf, pf + std::min<int>(counts[0], 256));
    caveFeaturePlacements.assign(pc, pc + std::min<int>(counts[1], MMGEN_MAX_CAVE_FEATURES));
}

void Chunk::otherChunkGatherFeaturePlacements(Chunk* chunkPtr, Chunk* const (&neighborChunks)[13][13], int centerX, int centerZ)
{
    ensureInit();
    ++g_calls[6];
    static int32_t offsets[49][2];
    static bool haveOffsets = false;
    if (!haveOffsets) { check(mmgen_gather_offsets(&offsets[0][0]), "mmgen_gather_offsets"); haveOffsets = true; }
    // pool = the 49 neighbours in the reference's order (chunk.cu:1158-1167), so neighbours[k] = k
    const Chunk* nb[49];
    size_t strideF = 1, strideC = 1;
    for (int k = 0; k < 49; ++k)
    {
        nb[k] = neighborChunks[centerZ + offsets[k][1]][centerX + offsets[k][0]];
        strideF = std::max(strideF, nb[k]->featurePlacements.size());
        strideC = std::max(strideC, nb[k]->caveFeaturePlacements.size());
    }
    static std::vector<FeaturePlacement> poolF, outF(MMGEN_MAX_FEATURES);
    static std::vector<CaveFeaturePlacement> poolC, outC(MMGEN_MAX_CAVE_FEATURES);
    poolF.assign(49 * strideF, FeaturePlacement{Feature::NONE});
    poolC.assign(49 * strideC, CaveFeaturePlacement{CaveFeature::NONE});
    int32_t idx[49], counts[49][2], outCounts[2] = {0, 0};
    for (int k = 0; k < 49; ++k)
    {
        idx[k] = k;
        std::copy(nb[k]->featurePlacements.begin(), nb[k]->featurePlacements.end(), poolF.begin() + k * strideF);
        std::copy(nb[k]->caveFeaturePlacements.begin(), nb[k]->caveFeaturePlacements.end(), poolC.begin() + k * strideC);
        counts[k][0] = (int32_t)nb[k]->featurePlacements.size();
        counts[k][1] = (int32_t)nb[k]->caveFeaturePlacements.size();
    }
    check(mmgen_gather_features(1, idx, 49, (const MmgenFeaturePlacement*)poolF.data(), (int)strideF, (const MmgenCaveFeaturePlacement*)poolC.data(),
                                (int)strideC, &counts[0][0], (MmgenFeaturePlacement*)outF.data(), (MmgenCaveFeaturePlacement*)outC.data(), outCounts),
          "Chunk::gatherFeaturePlacements()");
    // the reference clears the surface list only (chunk.cu:1171); both are cleared after the fill (chunk.cu:1586, 1601)
    chunkPtr->gatheredFeaturePlacements.assign(outF.begin(), outF.begin() + std::min<int>(outCounts[0], MMGEN_MAX_FEATURES));
    chunkPtr->gatheredCaveFeaturePlacements.insert(chunkPtr->gatheredCaveFeaturePlacements.end(), outC.begin(),
                                                   outC.begin() + std::min<int>(outCounts[1], MMGEN_MAX_CAVE_FEATURES));
}
