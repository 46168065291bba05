// TEST INFRASTRUCTURE (CPU oracle) - never linked into the product.
//
// Stage 2 (terrain layers) and stage 3 (zone erosion) of the reference, restated:
//   getStratifiedMaterialThickness + kernGenerateLayers   /root/reference/src/terrain/chunk.cu:308-415
//   kernDoErosion, erodeZone loop, fixBackwardStratifiedLayers  chunk.cu:477-601, 658-749
// Rounding (fmaf vs separate ops) follows the reference's sm_100 SASS for these kernels.
#pragma once
#include <vector>
#include "mm_noise.h"
#include "mm_tables.h"

namespace mmo {

// chunk.cu:322-415 for one column. h18: the chunk's 18x18 bordered heightfield; (x,z) local.
// colWeights[b*wstride] biome weights; out layers[l*lstride]. Forward layers after the reference's
// early break (chunk.cu:387-390) are left untouched here (caller pre-fills them); `written` gets
// the number of forward layers written.
static inline int layers_column(const float* h18, int x, int z, int wx, int wz, const float* colWeights, int wstride,
                                float* layers, int lstride)
{
    const MaterialInfo* mi = material_infos();
    const float* bmw = biome_material_weights();
    float tw[NUM_MATERIALS];
    for (int m = 0; m < NUM_MATERIALS; ++m)
    {
        float acc = 0.0f;
        for (int b = 0; b < NUM_BIOMES; ++b) acc = fmaf(colWeights[b * wstride], bmw[m + NUM_MATERIALS * b], acc);
        tw[m] = acc;
    }
    const float maxHeight = h18[(x + 1) + 18 * (z + 1)];
    float slope = 0.0f;
    for (int i = 0; i < 8; ++i)
    {
        const float nh = h18[(x + 1 + kDirVecs2d[i][0]) + 18 * (z + 1 + kDirVecs2d[i][1])];
        const float d = fabsf(nh - maxHeight);
        slope = fmaxf(slope, (i & 1) ? d * 1.41421356237309504880168872420f : d);
    }
    const float fx = (float)wx, fz = (float)wz;
    auto thickness = [&](int layerIdx) -> float {
        if (!(tw[layerIdx] > 0.0f)) return 0.0f;
        const float off = (float)layerIdx * 5283.64f;
        const float f = fbm2<5>(fmaf(mi[layerIdx].v2, fx, off), fmaf(mi[layerIdx].v2, fz, off));
        return fmaxf(fmaf(f, mi[layerIdx].v1, mi[layerIdx].thickness), 0.0f) * tw[layerIdx];
    };
    float height = 0.0f;
    int written = 0;
    for (int l = 0; l < NUM_FORWARD; ++l)
    {
        layers[l * lstride] = height;
        ++written;
        if (height > maxHeight || l == NUM_FORWARD - 1) break;
        height = thickness(l) + height;
    }
    height = 0.0f;
    for (int l = NUM_STRATIFIED - 1; l >= NUM_FORWARD; --l)
    {
        height = thickness(l) + height;
        layers[l * lstride] = height;   // turned into an absolute height by fix_backward_layers()
    }
    height = maxHeight;
    for (int l = NUM_MATERIALS - 1; l >= NUM_STRATIFIED; --l)
    {
        const float lh = fmaxf(mi[l].thickness * ((mi[l].v2 - slope) / mi[l].v2), 0.0f);
        height = fmaf(-tw[l], lh, height);
        layers[l * lstride] = height;
    }
    return written;
}

// One zone: planes[9][384*384] = 8 loose layer starts + heightfield (chunk.cu:603-656 layout).
// Race-free (Jacobi) reading of kernDoErosion: every sweep reads a snapshot. Returns sweeps done.
static inline int erode_zone(float* planes)
{
    const int N = EROSION_SIDE, NC = EROSION_COLS;
    const MaterialInfo* mi = material_infos();
    std::vector<float> accum(NC, 0.0f), S(NC), E(NC);
    int sweeps = 0;
    for (int layer = NUM_ERODED - 1; layer >= 0; --layer)
    {
        const float rep = mi[NUM_STRATIFIED + layer].v1;
        const float repDiag = rep * 1.41421356237309504880168872420f;
        float* G = planes + (size_t)layer * NC;
        const float* Gup = planes + (size_t)(layer + 1) * NC;
        bool first = true, changed;
        do
        {
            changed = false;
            for (int i = 0; i < NC; ++i)
            {
                const float a = first ? accum[i] : 0.0f;
                S[i] = G[i] + a;
                E[i] = Gup[i] + a;
            }
            for (int z = 0; z < N; ++z)
                for (int x = 0; x < N; ++x)
                {
                    const int i = x + N * z;
                    float ns = S[i], maxT = E[i] - S[i];
                    for (int d = 0; d < 8; ++d)
                    {
                        int nx = x + kDirVecs2d[d][0], nz = z + kDirVecs2d[d][1];
                        nx = nx < 0 ? 0 : (nx > N - 1 ? N - 1 : nx);
                        nz = nz < 0 ? 0 : (nz > N - 1 ? N - 1 : nz);
                        const int j = nx + N * nz;
                        ns = fmaxf(ns, S[j] - ((d & 1) ? repDiag : rep));
                        maxT = fmaxf(maxT, E[j] - S[j]);
                    }
                    ns = fminf(ns, E[i]);
                    if (maxT > 0.0f)
                    {
                        G[i] = ns;
                        if (ns != S[i])
                        {
                            changed = true;
                            accum[i] = (ns - S[i]) + accum[i];
                        }
                    }
                }
            first = false;
            ++sweeps;
        } while (changed);
    }
    return sweeps;
}

// chunk.cu:725-749 for one column: layers 10, 11 become erodedStart - cumulativeThickness
static inline void fix_backward_layers(float* layers, int lstride)
{
    const float erodedStart = layers[NUM_STRATIFIED * lstride];
    for (int l = NUM_FORWARD; l < NUM_STRATIFIED; ++l) layers[l * lstride] = erodedStart - layers[l * lstride];
}

}  // namespace mmo
