// TEST INFRASTRUCTURE (CPU oracle) - never linked into the product.
//
// Stage 5 of the reference, restated: per-column feature placement (CPU code in the reference,
// /root/reference/src/terrain/chunk.cu:999-1156) and the 7x7 gather order (chunk.cu:1158-1196).
//
// Two platform facts are part of the behaviour and are frozen here as the reference build on this
// image (nvcc 12.9 host pass = g++ 13 -O3, glibc 2.39) exhibits them:
//  * isFeaturePos() runs on the host: plain IEEE ops without contraction and the C library's sinf
//    (mm_hostmath.h), not libdevice's.
//  * Chunk::tryGenerateCaveFeaturePlacement() falls off its end without a return when the grid test
//    fails (chunk.cu:1028-1038). g++ compiles that undefined behaviour by dropping the test: the
//    placement is emitted whenever the chance / ceiling / lava / min-height tests pass (seen in the
//    disassembly of the reference object and in its outputs). That is the default (cave_grid_test_honoured() == 0);
//    mmo_set_cave_grid_test(1) selects the source-text reading instead (test honoured, failure = `false`), which the
//    product offers as mmgen_set_cave_grid_test(1).
#pragma once
#include <vector>
#include "mm_hostmath.h"
#include "mm_noise.h"
#include "mm_tables.h"

namespace mmo {

inline int& cave_grid_test_honoured() { static int v = 0; return v; }      // set between calls only (mmo_set_cave_grid_test)

// chunk.cu:999-1008 (host arithmetic)
static inline bool is_feature_pos(int wx, int wz, int cell, int pad, int seed)
{
    const float fc = (float)cell;
    const int gx = (int)(floorf((float)wx / fc) * fc), gz = (int)(floorf((float)wz / fc) * fc);
    const int internal = cell - 2 * pad;
    const float vx = (float)gx, vy = (float)gz, vz = (float)seed;
    const float d1 = (vx * 238.68f + vy * 491.28f) + vz * 640.88f;
    const float d2 = (vx * 654.37f + vy * 560.45f) + vz * 151.81f;
    float r1 = hm_sinf(d1) * 39021.426f, r2 = hm_sinf(d2) * 39021.426f;
    r1 = r1 - floorf(r1);
    r2 = r2 - floorf(r2);
    const int px = gx + pad + (int)floorf(r1 * (float)internal), pz = gz + pad + (int)floorf(r2 * (float)internal);
    return wx == px && wz == pz;
}

// biomeFuncs.hpp:39-53
static inline int random_biome(const float* w, int stride, float rand)
{
    for (int i = 0; i < NUM_BIOMES; ++i)
    {
        rand -= w[stride * i];
        if (rand <= 0.f) return i;
    }
    return PLAINS;
}

// chunk.cu:1041-1145 for one column; appends to the chunk's lists
static inline void column_feature_placements(int wx, int wz, float height, const float* colWeights, int wstride,
                                             const float* colLayers, int lstride, const CaveLayer* caveLayers,
                                             std::vector<FeaturePlacement>& feats, std::vector<CaveFeaturePlacement>& caveFeats)
{
    const int groundHeight = (int)height;
    const bool honoured = cave_grid_test_honoured() != 0;
    Minstd rng = make_rng3(wx, wz, 329828101);
    bool surfaceIsCave = false;
    for (int li = 0; li < MAX_CAVE_LAYERS; ++li)
    {
        const CaveLayer& cl = caveLayers[li];
        if (cl.start == 384 || groundHeight <= cl.start) break;
        for (int pass = 0; pass < 2; ++pass)
        {
            const bool top = pass == 1;
            if (top && cl.end == 384) break;
            int n;
            const CaveFeatureGen* gens = cave_biome_feature_gens(top ? cl.topBiome : cl.bottomBiome, &n);
            for (int g = 0; g < n; ++g)
            {
                const CaveFeatureGen& gen = gens[g];
                const int seed = top ? (int)gen.feature * 58321 + li * 871503 : (int)gen.feature * 98239 + li * 191702;
                const float rand = rng.u01();
                const int layerHeight = cl.end - cl.start;
                if (rand >= gen.chance || top != gen.fromCeiling || (!gen.inLava && (top ? cl.end : cl.start + 1) <= LAVA_LEVEL) ||
                    layerHeight < gen.minLayerHeight)
                    continue;
                if (!honoured || is_feature_pos(wx, wz, gen.gridCellSize, gen.gridCellPadding, seed))
                {
                    CaveFeaturePlacement p = {};
                    p.feature = gen.feature; p.x = wx; p.y = cl.start + 1; p.z = wz; p.layerHeight = layerHeight;
                    p.canReplaceBlocks = gen.canReplace ? 1 : 0;
                    caveFeats.push_back(p);
                    break;
                }
            }
        }
        if (groundHeight > cl.start && groundHeight <= cl.end) { surfaceIsCave = true; break; }
    }
    if (surfaceIsCave) return;
    const int biome = random_biome(colWeights, wstride, rng.u01());
    int n;
    const FeatureGen* gens = biome_feature_gens(biome, &n);
    for (int g = 0; g < n; ++g)
    {
        const FeatureGen& gen = gens[g];
        if (rng.u01() >= gen.chance) continue;
        if (gen.numTop > 0)
        {
            bool canPlace = false;
            for (int t = 0; t < gen.numTop; ++t)
            {
                const int l = gen.top[t].material;
                const float ls = colLayers[lstride * l], le = colLayers[lstride * (l + 1)];
                if (ls > height || le < height || fminf(le, height) - ls < gen.top[t].minThickness) continue;
                canPlace = true;
                break;
            }
            if (!canPlace) continue;
        }
        if (is_feature_pos(wx, wz, gen.gridCellSize, gen.gridCellPadding, (int)gen.feature * 518721))
        {
            FeaturePlacement p = {};
            p.feature = gen.feature; p.x = wx; p.y = groundHeight + 1; p.z = wz; p.canReplaceBlocks = gen.canReplace ? 1 : 0;
            feats.push_back(p);
            break;
        }
    }
}

// chunk.cu:1158-1167
static const int kGatherOffsets[49][2] = {
    {0, 0}, {0, 1}, {1, 1}, {1, 0}, {1, -1}, {0, -1}, {-1, -1}, {-1, 0}, {-1, 1}, {2, 0}, {2, 1}, {2, 2}, {1, 2}, {0, 2},
    {-1, 2}, {-2, 2}, {-2, 1}, {-2, 0}, {-2, -1}, {-2, -2}, {-1, -2}, {0, -2}, {1, -2}, {2, -2}, {2, -1},
    {-3, -3}, {-2, -3}, {-1, -3}, {0, -3}, {1, -3}, {2, -3}, {3, -3}, {3, -2}, {3, -1}, {3, 0}, {3, 1}, {3, 2}, {3, 3},
    {2, 3}, {1, 3}, {0, 3}, {-1, 3}, {-2, 3}, {-3, 3}, {-3, 2}, {-3, 1}, {-3, 0}, {-3, -1}, {-3, -2}};

}  // namespace mmo
