// Per-voxel block selection of stage 6 (device functions). Behaviour:
// chunkFillPlaceBlock /root/reference/src/terrain/chunk.cu:1202-1380, biomeBlockPre/PostProcess and
// caveBiomeBlockPostProcess /root/reference/src/terrain/biomeFuncs.hpp:385-707.
#pragma once
#include "mm_stage4.cuh"
#include "mm_placefeature.cuh"

namespace mmg {

// ------------------------------------------------------------------ 3-D Worley (rng.hpp:235-278)
// hash at the LUSH_CAVES call site (biomeFuncs.hpp:661): fma(z, Kz, fma(y, Ky, x*Kx))
__device__ MMG_NOISE_INLINE float worley3_lush(float px, float py, float pz)
{
    const float fx = floorf(px), fy = floorf(py), fz = floorf(pz);
    const int ix = (int)fx, iy = (int)fy, iz = (int)fz;
    const float nfx = fx - px, nfy = fy - py, nfz = fz - pz;
    // only the smallest distance is used: the minimum is taken over the squares and rooted once (sqrtf is monotone, so
    // the smallest rounded root is the root of the smallest square)
    float q1 = FLT_MAX;
#pragma unroll 1
    for (int x = -1; x <= 1; ++x)
#pragma unroll 1
        for (int y = -1; y <= 1; ++y)
#pragma unroll 1
            for (int z = -1; z <= 1; ++z)
            {
                const float cx = (float)(ix + x), cy = (float)(iy + y), cz = (float)(iz + z);
                const float jx = hash_fract(fmaf(cz, 402.98f, fmaf(cy, 491.28f, cx * 238.68f)));
                const float jy = hash_fract(fmaf(cz, 747.42f, fmaf(cy, 560.45f, cx * 654.37f)));
                const float jz = hash_fract(fmaf(cz, 674.81f, fmaf(cy, 151.81f, cx * 640.88f)));
                const float dx = nfx + (jx + (float)x), dy = nfy + (jy + (float)y), dz = nfz + (jz + (float)z);
                q1 = fminf(q1, fmaf(dz, dz, fmaf(dx, dx, dy * dy)));
            }
    return sqrtf(q1);
}

// ------------------------------------------------------------------ biomeFuncs.hpp:385-406
__device__ __forceinline__ bool biome_pre_process(uint8_t* block, int biome, int wx, int y, int wz, float height)
{
    if (biome == CRYSTALS && height > 176.f)
    {
        const float quartzStart = fmaf(fbm2<3>((float)wx * 0.0080f, (float)wz * 0.0080f), 15.f, 140.f);
        if ((float)y > quartzStart) { *block = B_QUARTZ; return true; }
    }
    return false;
}

// ------------------------------------------------------------------ biomeFuncs.hpp:408-590
__device__ __forceinline__ bool biome_post_process(uint8_t* block, int biome, int wx, int y, int wz, float height, bool isTopBlock)
{
    const float fx = (float)wx, fz = (float)wz, fy = (float)y;
    switch (biome)
    {
    case ARCHIPELAGO:
    {
        if (y < SEA_LEVEL || *block == B_WATER) return false;
        const float dirtHeight = fmaf(fbm2<3>(fx * 0.0065f, fz * 0.0065f), 1.7f, 129.5f);
        if (fy > dirtHeight) { *block = isTopBlock ? B_GRASS_BLOCK : B_DIRT; return true; }
        return false;
    }
    case TROPICAL_BEACH:
        if (isTopBlock && *block != B_SMOOTH_SAND && *block != B_WATER) { *block = B_SMOOTH_SAND; return true; }
        return false;
    case BEACH:
        if (isTopBlock && *block != B_SAND && *block != B_WATER) { *block = B_SAND; return true; }
        return false;
    case MESA:
    {
        if (fy < 90.f || *block == B_WATER) return false;
        const float start = fmaf(fbm2<3>(fx * 0.0040f, fz * 0.0040f), 12.f, 108.f);
        if (fy < start) return false;
        if (*block == B_CLAY && fy < start + 20.f) return false;
        float sample = fmaf(simplex3<true>(fx * 0.0100f, fz * 0.0100f, fy * 0.0300f), 3.f, fy) - start;
        sample = sample - 32.f * floorf(sample / 32.f);
        uint8_t b;
        if (sample < 5.f) b = B_TERRACOTTA;
        else if (sample < 8.f) b = B_ORANGE_TERRACOTTA;
        else if (sample < 12.f) b = B_RED_TERRACOTTA;
        else if (sample < 14.f) b = B_WHITE_TERRACOTTA;
        else if (sample < 20.f) b = B_TERRACOTTA;
        else if (sample < 21.f) b = B_ORANGE_TERRACOTTA;
        else if (sample < 26.f) b = B_YELLOW_TERRACOTTA;
        else if (sample < 29.f) b = B_PURPLE_TERRACOTTA;
        else b = B_TERRACOTTA;
        *block = b;
        return true;
    }
    case FROZEN_WASTELAND:
        if (*block != B_WATER) return false;
        *block = B_PACKED_ICE;
        return true;
    case SHREKS_SWAMP:
    {
        if (fy < 100.f) return false;
        if (*block == B_DIRT || *block == B_JUNGLE_GRASS_BLOCK)
        {
            const float mudEnd = fmaf(simplex2<true>(fx * 0.0300f, fz * 0.0300f), 1.1f, 128.8f);
            if (fy < mudEnd) { *block = B_MUD; return true; }
        }
        return false;
    }
    case TIANZI_MOUNTAINS:
    {
        if (fy < 90.f || *block == B_WATER || *block == B_DIRT || *block == B_GRASS_BLOCK) return false;
        const float start = fmaf(fbm2<3>(fx * 0.0200f, fz * 0.0200f), 16.f, 112.f);
        if (fy < start) return false;
        *block = B_SMOOTH_SANDSTONE;
        return true;
    }
    case CRYSTALS:
    {
        if (!isTopBlock || *block == B_QUARTZ) return false;
        if (hash_fract(fmaf((float)(wx + 913213), 238.68f, (float)(wz + 85941) * 491.28f)) < 0.1f) { *block = B_MYCELIUM; return true; }
        return false;
    }
    case MOUNTAINS:
    {
        if (fy < 190.f) return false;
        const float snowStart = fmaf(fbm2<3>(fx * 0.0500f, fz * 0.0500f), 5.f, 202.f);
        if (fy < snowStart) return false;
        *block = B_SNOW;
        return true;
    }
    default: return false;
    }
}

// ------------------------------------------------------------------ biomeFuncs.hpp:592-707
// second half of the LUSH_CAVES case of caveBiomeBlockPostProcess (biomeFuncs.hpp:653-668)
__device__ __forceinline__ uint8_t lush_block(int wx, int y, int wz)
{
    const float nx = (float)wx * 0.025f, nz = (float)wz * 0.025f;
    float ny = (float)y * 0.025f;
    ny = ny + 192031.9821f;
    const float ax = nx * 0.4f, ay = ny * 0.4f, az = nz * 0.4f;
    const float o1 = fbm3_paired<3, false>(ax, ay, az);
    const f32x2 o23 = fbm3x2<3, true>(f2_make(ax + 5923.45f, ax + 1765.68f), f2_make(ay + 4129.42f, ay + 4704.36f), f2_make(az + 5790.48f, az + 5692.12f));
    const float o2 = f2_lo(o23), o3 = f2_hi(o23);
    const float clay = worley3_lush(fmaf(o1, 2.f, nx), fmaf(o2, 2.f, ny), fmaf(o3, 2.f, nz));
    return clay < 0.25f ? B_CLAY : B_MOSS;
}

__device__ __forceinline__ bool cave_biome_post_process(uint8_t* block, int caveBiome, int wx, int y, int wz, int caveBottomDepth, int caveTopDepth,
                                                        bool* pendingLush)
{
    if (caveBiome == CB_NONE) return false;
    const bool isTopBlock = caveBottomDepth == 0;
    switch (caveBiome)
    {
    case CB_CRYSTAL_CAVES:
    {
        if (*block != B_STONE && *block != B_DEEPSLATE && *block != B_BLACKSTONE) return false;
        const float s = (float)(wx + wz);
        const float quartz = simplex3<true>((float)(wx + y) * 0.05f, (float)(wz + 5819323) * 0.05f, (s + s) * 0.05f);
        if (quartz < -0.25f) { *block = B_QUARTZ; return true; }
        if (*block == B_BLACKSTONE) return false;
        const float chance = (*block == B_STONE) ? 0.5f : 0.4f;
        const uint8_t cobble = (*block == B_STONE) ? B_COBBLESTONE : B_COBBLED_DEEPSLATE;
        // rand1From3(worldBlockPos): fma(z, Kz, fma(x, Kx, y*Ky))
        if (hash_fract(fmaf((float)wz, 640.88f, fmaf((float)wx, 238.68f, (float)y * 491.28f))) < chance) { *block = cobble; return true; }
        return false;
    }
    case CB_LUSH_CAVES:
    {
        if (*block != B_STONE && *block != B_DEEPSLATE && *block != B_BLACKSTONE) return false;
        const float nx = (float)wx * 0.025f, nz = (float)wz * 0.025f;
        float ny = (float)y * 0.025f;
        const float threshold = fmaf(simplex3<true>(nx, ny, nz), 4.5f, 1.5f);
        const float bd = (float)caveBottomDepth, td = (float)caveTopDepth;
        if (!(bd >= 0.f && bd <= threshold) && !(td >= 0.f && td <= threshold)) return false;
        // the clay / moss decision (3 x fbm3<3> + a 27-cell Worley) is needed by a few voxels per warp only:
        // it is deferred so that the CTA can evaluate all of its pending voxels on adjacent lanes (lush_block)
        *pendingLush = true;
        return false;
    }
    case CB_WARPED_FOREST:
        if (!isTopBlock) return false;
        if (*block == B_DEEPSLATE) { *block = B_WARPED_DEEPSLATE; return true; }
        if (*block == B_BLACKSTONE) { *block = B_WARPED_BLACKSTONE; return true; }
        return false;
    case CB_AMBER_FOREST:
        if (!isTopBlock) return false;
        if (*block == B_DEEPSLATE) { *block = B_AMBER_DEEPSLATE; return true; }
        if (*block == B_BLACKSTONE) { *block = B_AMBER_BLACKSTONE; return true; }
        return false;
    }
    return false;
}

// caveBiomeBlockPostProcess (above) only ever rewrites STONE, DEEPSLATE or BLACKSTONE, whatever the cave
// biome: CRYSTAL_CAVES and LUSH_CAVES return at once for any other block, WARPED_FOREST and AMBER_FOREST
// only map DEEPSLATE / BLACKSTONE. getCaveBiome has no side effects (its RNG is its own), so for every
// other block the ~5 kFLOP cave-biome evaluation the reference performs per voxel cannot change the result.
__device__ __forceinline__ bool needs_cave_biome(uint8_t block) { return block == B_STONE || block == B_DEEPSLATE || block == B_BLACKSTONE; }

// ------------------------------------------------------------------ chunk.cu:1202-1380
// weights[24], layersAndHeight[21] (20 layer starts + height), caveLayers[32] of the column
// *pendingLush: the voxel is lush-cave rock within the moss depth; its block is lush_block(wx, y, wz)
// *pendingRock: the block is STONE / DEEPSLATE / BLACKSTONE and still has to go through getCaveBiome +
// caveBiomeBlockPostProcess with the depths *bottomDepth / *topDepth (chunk.cu:1366-1370); the caller either queues
// it for the dense kernel k_fill_rock or finishes it in place with finish_rock_block.
// Per-column constants of chunkFillPlaceBlock, built once per column by the fill kernel: the surface biomes of non-zero
// weight in index order (randomBiome subtracts the weights in index order and a zero weight changes nothing, so walking
// only these gives the same pick; the one exception, rand == 0 picking biome 0 whatever its weight, is kept) and isOcean.
struct ColumnBiomes { int n; bool isOcean; uint8_t biome[NUM_BIOMES]; float weight[NUM_BIOMES]; };
__device__ __forceinline__ int random_biome_compact(const ColumnBiomes& cb, float rand)
{
    if (rand <= 0.f) return 0;                      // biomeFuncs.hpp:39-53 with rand == 0: the first test already passes
    for (int i = 0; i < cb.n; ++i)
    {
        rand -= cb.weight[i];
        if (rand <= 0.f) return cb.biome[i];
    }
    return PLAINS;
}

__device__ __forceinline__ uint8_t fill_place_block(const ColumnBiomes& cb, const float* layersAndHeight, const CaveLayer* caveLayers, int y,
                                       float height, int wx, int wz, bool* pendingRock, int* bottomDepth, int* topDepth)
{
    if (y == 0) return B_BEDROCK;
    const float fy = (float)y;
    if (fy > height && y > SEA_LEVEL) return B_AIR;
    // the surface biome of the voxel (chunk.cu:1232-1236) is a pure function of the position: it is drawn where it is first
    // needed (the water branch, or a solid voxel below), not for the voxels that turn out to be cave air
    int randBiome = -1;
    auto voxel_biome = [&]() {
        if (randBiome < 0)
        {
            Minstd rng = make_rng3(wx, y, wz);
            randBiome = random_biome_compact(cb, rng.u01());
        }
        return randBiome;
    };
    const bool isTopBlock = fy >= height - 1.f;
    uint8_t block = B_AIR;
    if (fy > height && y <= SEA_LEVEL)
    {
        block = B_WATER;
        biome_post_process(&block, voxel_biome(), wx, y, wz, height, isTopBlock);
        if (cb.isOcean) return block;
    }
    int caveBottomDepth = -384, caveTopDepth = -384;
    for (int li = 0; li < MAX_CAVE_LAYERS; ++li)
    {
        const CaveLayer& cl = caveLayers[li];
        if (cl.start == 384) { caveBottomDepth = -384; break; }
        caveBottomDepth = cl.start - y;
        if (y <= cl.start) break;
        if (y <= cl.end)
        {
            // the reference evaluates getCaveBiome + caveBiomeBlockPostProcess here (chunk.cu:1243-1246);
            // no cave biome changes AIR or LAVA (see needs_cave_biome), so neither is evaluated
            return (y <= LAVA_LEVEL) ? B_LAVA : B_AIR;
        }
        caveTopDepth = y - (cl.end + 1);
    }
    if (fy > height) return block;
    if (biome_pre_process(&block, voxel_biome(), wx, y, wz, height))
    {
        biome_post_process(&block, randBiome, wx, y, wz, height, isTopBlock);
        return block;
    }
    const int layerStart = (fy >= layersAndHeight[NUM_FORWARD]) ? NUM_FORWARD : 0;
    int thisLayer = -1;
    for (int l = layerStart; l < NUM_MATERIALS; ++l)
        if (layersAndHeight[l] <= fy && fy < layersAndHeight[l + 1]) { thisLayer = l; break; }
    // thisLayer == -1 happens when y == height exactly or the layer stack is not monotone. The
    // reference then reads dev_materialInfos[-1].block: 16 bytes in front of that table in constant
    // memory, which in its build is dev_biomeBlocks[8].grassBlock = SAVANNA_GRASS_BLOCK (the two
    // tables are adjacent, biomeFuncs.hpp:709-710; seen in the reference's output on a B200).
    block = thisLayer < 0 ? (uint8_t)B_SAVANNA_GRASS_BLOCK : c_materialInfos[thisLayer].block;
    if (isTopBlock && block == B_DIRT) block = c_biomeGrassBlock[randBiome];
    biome_post_process(&block, randBiome, wx, y, wz, height, isTopBlock);
    if (needs_cave_biome(block))
    {
        *pendingRock = true;
        *bottomDepth = caveBottomDepth;
        *topDepth = caveTopDepth;
    }
    return block;
}

// the deferred tail of chunkFillPlaceBlock for a rock voxel (chunk.cu:1366-1370)
__device__ __forceinline__ uint8_t finish_rock_block(uint8_t block, int wx, int y, int wz, float height, int bottomDepth, int topDepth, bool* pendingLush)
{
    cave_biome_post_process(&block, cave_biome(wx, y, wz, height, 190249401), wx, y, wz, bottomDepth, topDepth, pendingLush);
    return block;
}

// A rock voxel is "bulk" when neither depth is in [0, 6]: it is not the floor block of a cave (bottom depth 0), and the
// LUSH_CAVES rule needs a depth <= threshold = fma(simplex3, 4.5, 1.5) < 6.5 (|simplex3| < 1.1: the kernel sum
// 42 * sum_i (0.6 - r_i^2)^4 r_i |grad| peaks at 1.052 over the simplex cell, |grad| <= 0.999 for all 290 table entries).
// WARPED_FOREST / AMBER_FOREST only touch floor blocks, NONE nothing: for a bulk voxel only CRYSTAL_CAVES can change the block.
__device__ __forceinline__ bool rock_is_bulk(int bottomDepth, int topDepth)
{
    return (bottomDepth < 0 || bottomDepth > 6) && (topDepth < 0 || topDepth > 6);
}
// Queue record of a rock voxel: x = chunk, y = voxel index (17 bits) | rock kind (2) | bottom depth (6) | top depth (6).
// The depths only matter as "== 0" (top block) and "in [0, threshold]" with threshold = 1.5 + 4.5 simplex3 (biomeFuncs.hpp:653-657);
// |simplex3| <= 42 * 4 * max_r((0.6 - r^2)^4 r) * |grad| < 42 * 4 * 0.0209 * 3.2 < 11.3, so threshold < 53: every depth that is
// negative or above 62 behaves like "far" and is stored as 63.
__device__ __forceinline__ uint2 pack_rock(int chunk, int voxel, uint8_t block, int bottomDepth, int topDepth)
{
    const unsigned kind = block == B_STONE ? 0u : (block == B_DEEPSLATE ? 1u : 2u);
    const unsigned bd = (bottomDepth < 0 || bottomDepth > 62) ? 63u : (unsigned)bottomDepth;
    const unsigned td = (topDepth < 0 || topDepth > 62) ? 63u : (unsigned)topDepth;
    return make_uint2((unsigned)chunk, (unsigned)voxel | kind << 17 | bd << 19 | td << 25);
}
__device__ __forceinline__ void unpack_rock(uint2 e, int* chunk, int* voxel, uint8_t* block, int* bottomDepth, int* topDepth)
{
    *chunk = (int)e.x;
    *voxel = (int)(e.y & 0x1ffffu);
    const unsigned kind = (e.y >> 17) & 3u, bd = (e.y >> 19) & 63u, td = (e.y >> 25) & 63u;
    *block = kind == 0u ? B_STONE : (kind == 1u ? B_DEEPSLATE : B_BLACKSTONE);
    *bottomDepth = bd == 63u ? -384 : (int)bd;
    *topDepth = td == 63u ? -384 : (int)td;
}

}  // namespace mmg
