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
#include "mm_features.h"
#include "mm_fill.h"

#include <mutex>
namespace {
mmo::OpCounters g_total = {0, 0, 0, 0, 0};
std::mutex g_totalMutex;
void flush_counters()
{
    mmo::OpCounters& c = mmo::op_counters();
    std::lock_guard<std::mutex> lock(g_totalMutex);
    g_total.simplex2 += c.simplex2; g_total.simplex3 += c.simplex3; g_total.sinCalls += c.sinCalls;
    g_total.worleyCells2 += c.worleyCells2; g_total.worleyCells3 += c.worleyCells3;
    c = mmo::OpCounters{0, 0, 0, 0, 0};
}
template <class F>
void parallel_for(int n, int nthreads, F f)
{
    nthreads = std::max(1, std::min(nthreads, n));
    if (nthreads == 1)
    {
        for (int i = 0; i < n; ++i) f(i);
        flush_counters();
        return;
    }
    std::vector<std::thread> ts;
    for (int t = 0; t < nthreads; ++t)
        ts.emplace_back([=]() {
            for (int i = t; i < n; i += nthreads) f(i);
            flush_counters();
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

// 0 (default): the reference as built here (g++ drops the grid test of tryGenerateCaveFeaturePlacement, chunk.cu:1028-1038);
// 1: the source-text reading (test honoured, failure = false). Call between stage calls only.
void mmo_set_cave_grid_test(int honoured) { mmo::cave_grid_test_honoured() = honoured ? 1 : 0; }

// Chunk::generateFeaturePlacements (chunk.cu:1147-1156). Lists are written with stride maxPerChunk;
// counts[n][2] = {surface, cave} (the true counts, even if larger than maxPerChunk).
void mmo_feature_placements(int n, const int32_t* origins, const float* heightfield, const float* weights, const float* layers,
                            const void* caveLayers, int maxPerChunk, void* outF, void* outCF, int32_t* counts, int nthreads)
{
    const mmo::CaveLayer* cl = (const mmo::CaveLayer*)caveLayers;
    mmo::FeaturePlacement* F = (mmo::FeaturePlacement*)outF;
    mmo::CaveFeaturePlacement* CF = (mmo::CaveFeaturePlacement*)outCF;
    parallel_for(n, nthreads, [&](int c) {
        std::vector<mmo::FeaturePlacement> f;
        std::vector<mmo::CaveFeaturePlacement> cf;
        const int ox = origins[2 * c], oz = origins[2 * c + 1];
        for (int z = 0; z < 16; ++z)
            for (int x = 0; x < 16; ++x)
            {
                const int idx = x + 16 * z;
                mmo::column_feature_placements(ox + x, oz + z, heightfield[(size_t)c * 256 + idx],
                                               weights + (size_t)c * (mmo::NUM_BIOMES * 256) + idx, 256,
                                               layers + (size_t)c * (mmo::NUM_MATERIALS * 256) + idx, 256,
                                               cl + ((size_t)c * 256 + idx) * mmo::MAX_CAVE_LAYERS, f, cf);
            }
        counts[2 * c] = (int)f.size();
        counts[2 * c + 1] = (int)cf.size();
        for (size_t i = 0; i < f.size() && (int)i < maxPerChunk; ++i) F[(size_t)c * maxPerChunk + i] = f[i];
        for (size_t i = 0; i < cf.size() && (int)i < maxPerChunk; ++i) CF[(size_t)c * maxPerChunk + i] = cf[i];
    });
}

// Chunk::fill incl. placeDecorators (chunk.cu:1518-1632, 1679-1747). feats / caveFeats: gathered lists
// per chunk (stride strideF / strideCF entries), numFeatures[n][2] their lengths BEFORE truncation.
void mmo_fill(int n, const int32_t* origins, const float* heightfield, const float* weights, const float* layers,
              const void* caveLayers, const void* feats, const void* caveFeats, const int32_t* numFeatures, int strideF,
              int strideCF, uint8_t* out_blocks, int decorate, int nthreads)
{
    const mmo::CaveLayer* cl = (const mmo::CaveLayer*)caveLayers;
    const mmo::FeaturePlacement* F = (const mmo::FeaturePlacement*)feats;
    const mmo::CaveFeaturePlacement* CF = (const mmo::CaveFeaturePlacement*)caveFeats;
    parallel_for(n, nthreads, [&](int c) {
        const mmo::FeaturePlacement* f = F + (size_t)c * strideF;
        const mmo::CaveFeaturePlacement* cf = CF + (size_t)c * strideCF;
        const int nf = numFeatures[2 * c], ncf = numFeatures[2 * c + 1];
        int fb[2] = {384, -1}, cfb[2] = {384, -1};
        for (int i = 0; i < nf; ++i)
        {
            fb[0] = std::min(fb[0], f[i].y + mmo::kFeatureHeightBounds[f[i].feature][0]);
            fb[1] = std::max(fb[1], f[i].y + mmo::kFeatureHeightBounds[f[i].feature][1]);
        }
        for (int i = 0; i < ncf; ++i)
        {
            cfb[0] = std::min(cfb[0], cf[i].y + mmo::kCaveFeatureHeightBounds[cf[i].feature][0]);
            cfb[1] = std::max(cfb[1], cf[i].y + cf[i].layerHeight + mmo::kCaveFeatureHeightBounds[cf[i].feature][1]);
        }
        uint8_t* blocks = out_blocks + (size_t)c * 98304;
        const float* h = heightfield + (size_t)c * 256;
        const float* w = weights + (size_t)c * (mmo::NUM_BIOMES * 256);
        const mmo::CaveLayer* ccl = cl + (size_t)c * 256 * mmo::MAX_CAVE_LAYERS;
        mmo::fill_chunk(origins[2 * c], origins[2 * c + 1], h, w, layers + (size_t)c * (mmo::NUM_MATERIALS * 256), ccl, f,
                        std::min(nf, mmo::MAX_FEATURES), cf, std::min(ncf, mmo::MAX_CAVE_FEATURES), fb, cfb, blocks);
        if (decorate) mmo::place_decorators(origins[2 * c], origins[2 * c + 1], h, w, ccl, blocks);
    });
}
// Diagnostic for block flips on RAFFLESIA petals (tests/test_reference_tour.py): the signed distances of the five petal
// cylinders at a voxel, computed exactly as place_feature's F_RAFFLESIA case does (mm_placefeature.h), plus the same
// distances with the petal rotation evaluated in double precision. out[0..4] = sd (fp32 path), out[5..9] = sd (double
// rotation), out[10] = startAngle. A voxel whose smallest |sd| is a few ulps sits on the petal's surface.
void mmo_debug_rafflesia(int px, int py, int pz, int wx, int wy, int wz, float* out)
{
    using namespace mmo;
    V3 pos = v3((float)(wx - px), (float)(wy - py), (float)(wz - pz));
    Minstd frng = make_rng4(px, py, pz, 1293012);
    const float posY = pos.y;
    pos = pos * 0.8f;
    const float u = frng.u01();
    out[10] = u * kTwoPi;
    for (int i = 0; i < 5; ++i)
    {
        const float angle = i == 0 ? u * kTwoPi : pf_fma(u, kTwoPi, ((float)i * kTwoPi) * 0.2f);
        for (int dbl = 0; dbl < 2; ++dbl)
        {
            float s, co;
            if (dbl) { s = (float)sin(-(double)angle); co = (float)cos(-(double)angle); }
            else dm_sincosf(-angle, &s, &co);
            V3 pp = v3(pf_fma(pos.x, co, pos.z * s), pf_fma(posY, 0.8f, -3.2f), pf_fma(pos.z, co, -(pos.x * s)));
            pp.y = pp.y - (float)(i % 2) * 0.53f;
            pp.y = pf_fma(fminf(fmaxf((fabsf(pp.x - 3.f) - 1.5f) / 1.5f, 0.f), 1.f), 1.3f, pp.y);
            pp.x = pp.x - 3.8f;
            pp.z = pp.z * 1.2f;
            const float dx = fabsf(len2(pp.x, pp.z)) - 2.5f, dy = fabsf(pp.y) - 0.5f;
            const float mx = fmaxf(dx, 0.f), my = fmaxf(dy, 0.f);
            out[5 * dbl + i] = fminf(fmaxf(dx, dy), 0.0f) + sqrtf(pf_fma(mx, mx, my * my));
        }
    }
}

float mmo_host_sinf(float x) { return mmo::hm_sinf(x); }

// noise-primitive call counters accumulated since the last reset: simplex2, simplex3, sin, worley cells 2-D / 3-D
void mmo_counters(unsigned long long* out5, int reset)
{
    flush_counters();
    std::lock_guard<std::mutex> lock(g_totalMutex);
    out5[0] = g_total.simplex2; out5[1] = g_total.simplex3; out5[2] = g_total.sinCalls; out5[3] = g_total.worleyCells2; out5[4] = g_total.worleyCells3;
    if (reset) g_total = mmo::OpCounters{0, 0, 0, 0, 0};
}

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


// Observed extent of placed features, by brute force over a box of voxels around each placement (tests of the
// product's culling tables): placements[n] = {feature, x, y, z, layerHeight}; for placement i every voxel with
// |dx|,|dz| <= radius and y in [ylo[i], yhi[i]] is tested; out[i] = {hits, max|dx|, max|dz|, min(y - py), max(y - py),
// min(y - py - lh), max(y - py - lh)} over the voxels the rasteriser fills (y extents are INT_MAX/INT_MIN without hits).
void mmo_feature_extent(int cave, int n, const int32_t* placements, int radius, const int32_t* ylo, const int32_t* yhi, int32_t* out, int nthreads)
{
    parallel_for(n, nthreads, [&](int i) {
        const int32_t* p = placements + 5 * i;
        int32_t* o = out + 7 * i;
        o[0] = o[1] = o[2] = 0; o[3] = o[5] = 2147483647; o[4] = o[6] = -2147483647 - 1;
        mmo::FeaturePlacement fp; mmo::CaveFeaturePlacement cp;
        std::memset(&fp, 0, sizeof(fp)); std::memset(&cp, 0, sizeof(cp));
        fp.feature = (uint8_t)p[0]; fp.x = p[1]; fp.y = p[2]; fp.z = p[3]; fp.canReplaceBlocks = 1;
        cp.feature = (uint8_t)p[0]; cp.x = p[1]; cp.y = p[2]; cp.z = p[3]; cp.layerHeight = p[4]; cp.canReplaceBlocks = 1;
        for (int dz = -radius; dz <= radius; ++dz)
            for (int dx = -radius; dx <= radius; ++dx)
                for (int y = ylo[i]; y <= yhi[i]; ++y)
                {
                    uint8_t b = 0;
                    const bool hit = cave ? mmo::place_cave_feature(cp, p[1] + dx, y, p[3] + dz, &b) : mmo::place_feature(fp, p[1] + dx, y, p[3] + dz, &b);
                    if (!hit) continue;
                    ++o[0];
                    o[1] = std::max(o[1], std::abs(dx)); o[2] = std::max(o[2], std::abs(dz));
                    o[3] = std::min(o[3], y - p[2]); o[4] = std::max(o[4], y - p[2]);
                    o[5] = std::min(o[5], y - p[2] - p[4]); o[6] = std::max(o[6], y - p[2] - p[4]);
                }
    });
}

}  // extern "C"
