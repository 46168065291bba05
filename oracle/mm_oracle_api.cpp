// TEST INFRASTRUCTURE (CPU oracle) - never linked into the product.
//
// C entry points of the CPU restatement, mirroring include/mmgen.h's batch operators so that
// tests/ can run the same inputs through both. Also used by bench.py's cpu_baseline / reference
// arm (the only places besides tests/ and smoke() allowed to touch oracle/).
// Threading: chunks are independent in every stage except erosion; `nthreads` std::threads split
// the chunk list.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

#include "mm_surface.h"
#include "mm_layers.h"
#include "mm_caves.h"

namespace {
template <class F>
void parallel_for(int n, int nthreads, F f)
{
    nthreads = std::max(1, std::min(nthreads, n));
    if (nthreads == 1)
    {
        for (int i = 0; i < n; ++i) f(i);
        return;
    }
    std::vector<std::thread> ts;
    for (int t = 0; t < nthreads; ++t)
        ts.emplace_back([=]() {
            for (int i = t; i < n; i += nthreads) f(i);
        });
    for (auto& t : ts) t.join();
}
}  // namespace

extern "C" {

// Chunk::generateHeightfields (chunk.cu:150-229)
void mmo_heightfields(int n, const int32_t* origins, float* out_h, float* out_w, int nthreads)
{
    parallel_for(n, nthreads, [&](int c) {
        const int ox = origins[2 * c], oz = origins[2 * c + 1];
        for (int z = 0; z < 16; ++z)
            for (int x = 0; x < 16; ++x)
            {
                const int idx = x + 16 * z;
                out_h[(size_t)c * 256 + idx] =
                    mmo::surface_column(ox + x, oz + z, out_w + (size_t)c * (mmo::NUM_BIOMES * 256) + idx, 256);
            }
    });
}

// Chunk::generateLayers on gathered 18x18 heightfields (chunk.cu:322-469). Forward layers the
// reference never writes keep the value `unwritten`.
void mmo_layers(int n, const int32_t* origins, const float* h18, const float* weights, float* out_layers, float unwritten,
                int nthreads)
{
    parallel_for(n, nthreads, [&](int c) {
        const int ox = origins[2 * c], oz = origins[2 * c + 1];
        float* L = out_layers + (size_t)c * (mmo::NUM_MATERIALS * 256);
        for (int i = 0; i < mmo::NUM_FORWARD * 256; ++i) L[i] = unwritten;
        for (int z = 0; z < 16; ++z)
            for (int x = 0; x < 16; ++x)
            {
                const int idx = x + 16 * z;
                mmo::layers_column(h18 + (size_t)c * 324, x, z, ox + x, oz + z,
                                   weights + (size_t)c * (mmo::NUM_BIOMES * 256) + idx, 256, L + idx, 256);
            }
    });
}

// Chunk::erodeZone's device part (chunk.cu:658-709) on planes[9][384*384], in place; returns sweeps
int mmo_erode_zone(float* planes) { return mmo::erode_zone(planes); }

// Chunk::generateCaves (chunk.cu:939-993); out: CaveLayer[n][256][32]
void mmo_caves(int n, const int32_t* origins, const float* heightfield, const float* weights, void* out, int nthreads)
{
    mmo::CaveLayer* cl = (mmo::CaveLayer*)out;
    parallel_for(n * 256, nthreads, [&](int i) {
        const int c = i >> 8, idx = i & 255;
        const int ox = origins[2 * c], oz = origins[2 * c + 1];
        mmo::caves_column(ox + (idx & 15), oz + (idx >> 4), heightfield[(size_t)c * 256 + idx],
                          weights + (size_t)c * (mmo::NUM_BIOMES * 256) + idx, 256, cl + ((size_t)c * 256 + idx) * mmo::MAX_CAVE_LAYERS);
    });
}
int mmo_cave_biome(int x, int y, int z, float maxHeight, int seed) { return mmo::cave_biome(x, y, z, maxHeight, seed); }

// unit probes used by tests
float mmo_sinf(float x) { return mmo::dm_sinf(x); }
float mmo_cosf(float x) { return mmo::dm_cosf(x); }
float mmo_powf(float a, float b) { return mmo::dm_powf(a, b); }
float mmo_simplex2(float x, float y) { return mmo::simplex2(x, y); }
float mmo_simplex3(float x, float y, float z) { return mmo::simplex3(x, y, z); }
uint32_t mmo_hash(uint32_t a) { return mmo::hash_u32(a); }
float mmo_rng3_u01(int x, int y, int z, int ndraw)
{
    mmo::Minstd r = mmo::make_rng3(x, y, z);
    float v = 0;
    for (int i = 0; i < ndraw; ++i) v = r.u01();
    return v;
}

}  // extern "C"
