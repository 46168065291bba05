// TEST INFRASTRUCTURE (CPU oracle) - never linked into the product.
//
// Stage 4 of the reference, restated: cave-biome noise and random cave biome
// (/root/reference/src/terrain/biomeFuncs.hpp:130-220), the per-voxel cave predicate
// (/root/reference/src/terrain/chunk.cu:755-810) with specialCaveNoise (rng.hpp:148-155, 282-320),
// and the per-column compaction into CaveLayers (chunk.cu:812-937).
// Rounding follows the reference's sm_100 PTX/SASS (oracle/tools/ptx_expr.py).
#pragma once
#include "mm_surface.h"

namespace mmo {

// ---------------------------------------------------------------- cave biome (biomeFuncs.hpp:135-220)
struct CaveBiomeNoise { float v[4]; };   // none, shallow, warped, rocky

static inline CaveBiomeNoise cave_biome_noise(int x, int y, int z, float maxHeight)
{
    const float px = (float)x, py = (float)y, pz = (float)z;
    const float qx = px * 0.0470f, qy = py * 0.0470f, qz = pz * 0.0470f;
    const float cx = fmaf(fbm3<3>(qx, qy, qz), 30.f, px);
    const float cy = fmaf(fbm3<3>(qx + 5923.45f, qy + 4129.42f, qz + 5790.48f), 24.f, py);
    const float cz = fmaf(fbm3<3>(qx + 1765.68f, qy + 4704.36f, qz + 5692.12f), 30.f, pz);
    const float nx = cx * 0.2000f, nz = cz * 0.2000f;
    const float top = fmaf(maxHeight + -128.f, 0.15f, 128.f);
    const float nsStart = fmaf(fbm2<3>(nx, nz), 23.f, top + -19.f);
    const float nsEnd = fmaf(fbm2<3>(nx + 3821.34f, nz + 4920.32f), 3.f, nsStart + -5.f);
    const float sdStart = fmaf(fbm2<3>(nx + -4921.34f, nz + 8402.13f), 18.f, top + -72.f);
    const float sdEnd = fmaf(fbm2<3>(nx + 9411.32f, nz + -3921.34f), 7.f, sdStart + -10.f);
    CaveBiomeNoise n;
    n.v[0] = ss_t((cy - nsEnd) / (nsStart - nsEnd));
    n.v[1] = ss_t((cy - sdEnd) / (sdStart - sdEnd));
    n.v[2] = ss_t(fmaf(simplex3_raw<true>(fmaf(cx, 0.0030f, 5821.32f), fmaf(cy, 0.0030f, 4920.12f), fmaf(cz, 0.0030f, 7931.59f)), 42.f, 0.05f) / (0.05f - -0.05f));
    n.v[3] = ss_t(fmaf(simplex3_raw<true>(fmaf(cx, 0.0022f, -9193.23f), fmaf(cy, 0.0022f, -6813.39f), fmaf(cz, 0.0022f, (float)-2171.23)), 42.f, 0.05f) / (0.05f - -0.05f));
    return n;
}

static inline int cave_biome(int x, int y, int z, float maxHeight, int seed)
{
    const CaveBiomeNoise n = cave_biome_noise(x, y, z, maxHeight);
    Minstd rng = make_rng4(x, y, z, seed);
    float rand = rng.u01();
    for (int b = 0; b < NUM_CAVE_BIOMES; ++b)
    {
        float w = 1.0f;
        for (int c = 0; c < 4; ++c)
        {
            const uint8_t t = kCaveBiomeNoiseWeights[b][c];
            if (t == 1) w *= n.v[c];
            else if (t == 2) w *= 1.0f - n.v[c];
        }
        rand -= w;
        if (rand <= 0.f) return b;
    }
    return CB_NONE;
}

// ---------------------------------------------------------------- specialCaveNoise (rng.hpp:282-320)
// hash (rng.hpp:148-155) at this call site: fma(z, Kz, fma(x, Kx, y*Ky))
static inline float special_cave_noise(float px, float py, float pz)
{
    const float fx = floorf(px), fy = floorf(py), fz = floorf(pz);
    const int ix = (int)fx, iy = (int)fy, iz = (int)fz;
    const float nfx = fx - px, nfy = fy - py, nfz = fz - pz;
    float d1 = FLT_MAX, d2 = FLT_MAX, d3 = FLT_MAX;
    for (int x = -1; x <= 1; ++x)
        for (int y = -1; y <= 1; ++y)
            for (int z = -1; z <= 1; ++z)
            {
                ++op_counters().worleyCells3;
                const float cx = (float)(ix + x), cy = (float)(iy + y), cz = (float)(iz + z);
                const float jx = hash_fract(fmaf(cz, 402.98f, fmaf(cx, 238.68f, cy * 491.28f)));
                const float jy = hash_fract(fmaf(cz, 747.42f, fmaf(cx, 654.37f, cy * 560.45f)));
                const float jz = hash_fract(fmaf(cz, 674.81f, fmaf(cx, 640.88f, cy * 151.81f)));
                const float dx = nfx + (jx + (float)x), dy = nfy + (jy + (float)y), dz = nfz + (jz + (float)z);
                const float dist = sqrtf(fmaf(dz, dz, fmaf(dx, dx, dy * dy)));
                if (dist < d1) { d3 = d2; d2 = d1; d1 = dist; }
                else if (dist < d2) { d3 = d2; d2 = dist; }
                else if (dist < d3) { d3 = dist; }
            }
    return d3 / d1 + -1.0f;
}

// y-independent part of the ravine test (chunk.cu:785-801), hoisted per column (same values)
struct Ravine { bool active; float top, depth; };

static inline Ravine ravine_column(int wx, int wz, float obw)
{
    Ravine r = {false, 0.f, 0.f};
    const float rx = (float)wx * 0.0015f, rz = (float)wz * 0.0015f;
    const float ox = fbm2<4>(rx * 10.f, rz * 10.f), oz = fbm2<4>(rx * 10.f + 5923.45f, rz * 10.f + 4129.42f);
    const Worley2 w = worley2(fmaf(ox, 0.03f, rx), fmaf(oz, 0.03f, rz));
    const float thr = (1.f - obw) * 0.12f;
    if (!(w.d1 < thr)) return r;
    const float colorX = hash_fract(fmaf(w.cpx, 238.68f, w.cpy * 491.28f));
    r.top = fmaf(colorX, 24.f, 120.f);
    const float ratio = 1.f - (w.d1 / thr);
    float depth = ss_t(ratio / 0.3f) * fmaf(fbm2<4>(fmaf(rx, 8.f, 8391.32f), fmaf(rz, 8.f, 4821.39f)), 26.f, 60.f);
    const float waveOff = fbm2<4>(fmaf(rx, 3.f, 5129.32f), fmaf(rz, 3.f, 1392.49f)) * 4.f;
    const float wave = dm_sinf(fmaf(rx + rz, 15.f, waveOff));
    depth = depth * ss_t((wave + -0.4f) / (0.6f - 0.4f));
    r.depth = depth;
    r.active = depth > 0.0001f;
    return r;
}

// chunk.cu:755-810
static inline bool cave_at_block(int wx, int y, int wz, float maxHeight, float obw, const Ravine& rav)
{
    if (y == 0) return false;
    const int hi = (int)maxHeight;
    if (y > (hi > SEA_LEVEL ? hi : SEA_LEVEL)) return true;
    const float fy = (float)y;
    const float npx = (float)wx * 0.0050f, npy = fy * 0.0050f, npz = (float)wz * 0.0050f;
    const float topRatio = ss_t((fmaf(obw, 50.f, fy) + -142.f) / (95.f - 142.f));
    const float bottomRatio = ss_t((fy + -5.f) / (20.f - 5.f));
    const float ax = npx * 0.8000f, ay = npy * 0.8000f, az = npz * 0.8000f;
    const float o1 = fbm3<5>(ax, ay, az);
    const float o2 = fbm3<5>(ax + 5923.45f, ay + 4129.42f, az + 5790.48f);
    const float o3 = fbm3<5>(ax + 1765.68f, ay + 4704.36f, az + 5692.12f);
    const float caveNoise = special_cave_noise(fmaf(o1, 1.8f, npx), fmaf(npy, 1.6f, o2 * 1.8f), fmaf(o3, 1.8f, npz));
    float thr = fmaf(fbm3<4>(npx * 4.f, npy * 4.f, npz * 4.f), 0.12f, 0.24f);
    const float huge = ss_t((fbm3<4>(npx * 0.0700f, npy * 0.0700f, npz * 0.0700f) + -0.2f) / (0.4f - 0.2f));
    thr = thr * fmaf(huge, 1.4f, 1.f);
    thr = (fmaf(bottomRatio, 0.7f, 0.3f) * topRatio) * thr;
    if (thr > 0.04f && caveNoise < thr) return true;
    if (rav.active && (rav.top - rav.depth) < fy) return true;
    return false;
}

// chunk.cu:812-937 for one column; out: 32 CaveLayers (pre-initialised here to {384,384,0,0})
static inline void caves_column(int wx, int wz, float maxHeight, const float* colWeights, int wstride, CaveLayer* out)
{
    float obw = 0.f;
    for (int b = 0; b < NUM_OCEAN_BEACH_BIOMES; ++b) obw += colWeights[b * wstride];
    const Ravine rav = ravine_column(wx, wz, obw);
    uint8_t filled[385];
    for (int y = 0; y < 384; ++y) filled[y] = cave_at_block(wx, y, wz, maxHeight, obw, rav) ? 0 : 1;
    filled[384] = 0;
    for (int l = 0; l < MAX_CAVE_LAYERS; ++l) { out[l].start = 384; out[l].end = 384; out[l].bottomBiome = 0; out[l].topBiome = 0; out[l].pad[0] = out[l].pad[1] = 0; }
    int nflips = 0;
    for (int y = 0; y < 384; ++y)
        if (filled[y] != filled[y + 1])
        {
            if (nflips < 2 * MAX_CAVE_LAYERS)
            {
                if (nflips & 1) out[nflips >> 1].end = y;
                else out[nflips >> 1].start = y;
            }
            ++nflips;
        }
    for (int l = 0; l < MAX_CAVE_LAYERS; ++l)
    {
        if (out[l].start != 384) out[l].bottomBiome = (uint8_t)cave_biome(wx, out[l].start, wz, maxHeight, 329271348);
        if (out[l].end == 384) out[l].topBiome = CB_NONE;
        else out[l].topBiome = (uint8_t)cave_biome(wx, out[l].end + 1, wz, maxHeight, 4982921);
    }
}

}  // namespace mmo
