// Stage 5 (feature placement + gather) and stage 6 (chunk fill + decorators) kernels.
// Replace the reference's CPU passes generateFeaturePlacements / gatherFeaturePlacements /
// placeDecorators and kernFill (/root/reference/src/terrain/chunk.cu:999-1196, 1382-1747).
//
// S5a  k_feature_placements: one CTA per chunk, one thread per column. The column routine is run
//      twice (count, then emit at the exclusive-scan offset) so that the chunk's lists come out in
//      the reference's z-outer / x-inner push_back order without atomics.
// S5b  k_gather_features: one CTA per chunk to fill; concatenates the 49 neighbour lists in the
//      reference's fixed order up to the caps (2048 / 4096) and reduces the y-bounds over the whole
//      untruncated set, as the host loop at chunk.cu:1555-1570 does.
// S6   k_fill: one CTA per column (384 threads = 384 voxels, y fastest), column inputs staged in
//      shared memory; a warp stores 32 consecutive block IDs. The reference launches one such grid
//      PER CHUNK; here one launch covers every chunk of the batch.
// S6b  k_decorators: the reference's sequential per-chunk RNG walk, one thread per chunk.
#pragma once
#include "mm_common.cuh"
#include "mm_featurefuncs.cuh"
#include "mm_fillfuncs.cuh"

namespace mmg {

constexpr int kMaxOwnFeatures = 256;        // at most one surface feature per column
constexpr int kMaxOwnCaveFeatures = 4096;   // anything beyond the gather cap can never be consumed

// chunk.cu:1041-1145 for one column. emit == nullptr: count only.
__device__ __forceinline__ void column_features(int wx, int wz, float height, const float* colWeights, const float* colLayers,
                                                const CaveLayer* caveLayers, int* nSurf, int* nCave, FeaturePlacement* outF,
                                                CaveFeaturePlacement* outCF, int maxCave)
{
    const int groundHeight = (int)height;
    Minstd rng = make_rng3(wx, wz, 329828101);
    bool surfaceIsCave = false;
    int ns = 0, nc = 0;
    for (int li = 0; li < MAX_CAVE_LAYERS; ++li)
    {
        const CaveLayer cl = caveLayers[li];
        if (cl.start == 384 || groundHeight <= cl.start) break;
        for (int pass = 0; pass < 2; ++pass)
        {
            const bool top = pass == 1;
            if (top && cl.end == 384) break;
            const int biome = top ? cl.topBiome : cl.bottomBiome;
            const int g0 = c_caveFeatureGenRange[biome][0], gn = c_caveFeatureGenRange[biome][1];
            for (int g = 0; g < gn; ++g)
            {
                const CaveFeatureGen gen = c_caveFeatureGens[g0 + g];
                const int seed = top ? (int)gen.feature * 58321 + li * 871503 : (int)gen.feature * 98239 + li * 191702;
                const float rand = rng.u01();
                const int layerHeight = cl.end - cl.start;
                if (rand >= gen.chance || top != (gen.fromCeiling != 0) || (!gen.inLava && (top ? cl.end : cl.start + 1) <= LAVA_LEVEL) ||
                    layerHeight < gen.minLayerHeight)
                    continue;
                if (kCaveGridTestIgnored || is_feature_pos(wx, wz, gen.cell, gen.pad, seed))
                {
                    if (outCF && nc < maxCave)
                    {
                        CaveFeaturePlacement p;
                        p.feature = gen.feature; p.pad0[0] = p.pad0[1] = p.pad0[2] = 0;
                        p.x = wx; p.y = cl.start + 1; p.z = wz; p.layerHeight = layerHeight;
                        p.canReplaceBlocks = gen.canReplace; p.pad1[0] = p.pad1[1] = p.pad1[2] = 0;
                        outCF[nc] = p;
                    }
                    ++nc;
                    break;
                }
            }
        }
        if (groundHeight > cl.start && groundHeight <= cl.end) { surfaceIsCave = true; break; }
    }
    if (!surfaceIsCave)
    {
        const int biome = random_biome(colWeights, 256, rng.u01());
        const int g0 = c_featureGenRange[biome][0], gn = c_featureGenRange[biome][1];
        for (int g = 0; g < gn; ++g)
        {
            const FeatureGen gen = c_featureGens[g0 + g];
            if (rng.u01() >= gen.chance) continue;
            if (gen.numTop > 0)
            {
                bool canPlace = false;
                for (int t = 0; t < gen.numTop; ++t)
                {
                    const int l = gen.topMat[t];
                    const float ls = colLayers[256 * l], le = colLayers[256 * (l + 1)];
                    if (ls > height || le < height || fminf(le, height) - ls < gen.topMin[t]) continue;
                    canPlace = true;
                    break;
                }
                if (!canPlace) continue;
            }
            if (is_feature_pos(wx, wz, gen.cell, gen.pad, (int)gen.feature * 518721))
            {
                if (outF)
                {
                    FeaturePlacement p;
                    p.feature = gen.feature; p.pad0[0] = p.pad0[1] = p.pad0[2] = 0;
                    p.x = wx; p.y = groundHeight + 1; p.z = wz;
                    p.canReplaceBlocks = gen.canReplace; p.pad1[0] = p.pad1[1] = p.pad1[2] = 0;
                    outF[0] = p;
                }
                ++ns;
                break;
            }
        }
    }
    *nSurf = ns;
    *nCave = nc;
}

// own lists: features[chunk][kMaxOwnFeatures], caveFeatures[chunk][kMaxOwnCaveFeatures], counts[chunk][2]
__global__ void __launch_bounds__(256) k_feature_placements(const int* __restrict__ chunkList, const int2* __restrict__ origins,
                                                            const float* __restrict__ heightfield, const float* __restrict__ biomeWeights,
                                                            const float* __restrict__ layers, const CaveLayer* __restrict__ caveLayers,
                                                            FeaturePlacement* __restrict__ features, CaveFeaturePlacement* __restrict__ caveFeatures,
                                                            int* __restrict__ counts)
{
    __shared__ int shS[256], shC[256];
    const int li = blockIdx.x, chunk = chunkList ? chunkList[li] : li;
    const int idx = threadIdx.x;
    const int2 o = origins[chunk];
    const int wx = o.x + (idx & 15), wz = o.y + (idx >> 4);
    const float height = heightfield[(size_t)chunk * 256 + idx];
    const float* cw = biomeWeights + (size_t)chunk * (NUM_BIOMES * 256) + idx;
    const float* cl = layers + (size_t)chunk * (NUM_MATERIALS * 256) + idx;
    const CaveLayer* ccl = caveLayers + ((size_t)chunk * 256 + idx) * MAX_CAVE_LAYERS;
    int ns, nc;
    column_features(wx, wz, height, cw, cl, ccl, &ns, &nc, nullptr, nullptr, 0);
    shS[idx] = ns;
    shC[idx] = nc;
    __syncthreads();
    // inclusive Hillis-Steele scan over 256 columns (column order = idx = x + 16 z = reference order)
    for (int d = 1; d < 256; d <<= 1)
    {
        const int a = idx >= d ? shS[idx - d] : 0, b = idx >= d ? shC[idx - d] : 0;
        __syncthreads();
        shS[idx] += a;
        shC[idx] += b;
        __syncthreads();
    }
    const int offS = shS[idx] - ns, offC = shC[idx] - nc;
    if (idx == 255)
    {
        counts[2 * chunk] = shS[255];
        counts[2 * chunk + 1] = min(shC[255], kMaxOwnCaveFeatures);
    }
    if (ns + nc > 0)
    {
        int ns2, nc2;
        const int room = max(0, kMaxOwnCaveFeatures - offC);
        column_features(wx, wz, height, cw, cl, ccl, &ns2, &nc2, features + (size_t)chunk * kMaxOwnFeatures + offS,
                        caveFeatures + (size_t)chunk * kMaxOwnCaveFeatures + offC, room);
    }
}

struct GatherInfo { int nF, nCF; int fb0, fb1, cfb0, cfb1; };

// chunk.cu:1158-1187 + 1555-1578. fillList: chunks to gather for (world raster indices).
__global__ void __launch_bounds__(256) k_gather_features(const int* __restrict__ fillList, const FeaturePlacement* __restrict__ features,
                                                         const CaveFeaturePlacement* __restrict__ caveFeatures, const int* __restrict__ counts,
                                                         int nx, FeaturePlacement* __restrict__ gF, CaveFeaturePlacement* __restrict__ gCF,
                                                         GatherInfo* __restrict__ info)
{
    __shared__ int shMin[2], shMax[2];
    const int li = blockIdx.x, chunk = fillList[li];
    const int tid = threadIdx.x;
    if (tid < 2) { shMin[tid] = 384; shMax[tid] = -1; }
    __syncthreads();
    int baseF = 0, baseC = 0;
    int mnF = 384, mxF = -1, mnC = 384, mxC = -1;
    for (int k = 0; k < 49; ++k)
    {
        const int nchunk = chunk + c_gatherOffsets[k][0] + c_gatherOffsets[k][1] * nx;
        const int nf = counts[2 * nchunk], nc = counts[2 * nchunk + 1];
        const FeaturePlacement* sf = features + (size_t)nchunk * kMaxOwnFeatures;
        const CaveFeaturePlacement* sc = caveFeatures + (size_t)nchunk * kMaxOwnCaveFeatures;
        for (int i = tid; i < nf; i += 256)
        {
            const FeaturePlacement p = sf[i];
            if (baseF + i < MAX_FEATURES) gF[(size_t)li * MAX_FEATURES + baseF + i] = p;
            mnF = min(mnF, p.y + c_featureHeightBounds[p.feature][0]);
            mxF = max(mxF, p.y + c_featureHeightBounds[p.feature][1]);
        }
        for (int i = tid; i < nc; i += 256)
        {
            const CaveFeaturePlacement p = sc[i];
            if (baseC + i < MAX_CAVE_FEATURES) gCF[(size_t)li * MAX_CAVE_FEATURES + baseC + i] = p;
            mnC = min(mnC, p.y + c_caveFeatureHeightBounds[p.feature][0]);
            mxC = max(mxC, p.y + p.layerHeight + c_caveFeatureHeightBounds[p.feature][1]);
        }
        baseF += nf;
        baseC += nc;
    }
    atomicMin(&shMin[0], mnF); atomicMax(&shMax[0], mxF);
    atomicMin(&shMin[1], mnC); atomicMax(&shMax[1], mxC);
    __syncthreads();
    if (tid == 0)
    {
        GatherInfo gi;
        gi.nF = min(baseF, MAX_FEATURES); gi.nCF = min(baseC, MAX_CAVE_FEATURES);
        gi.fb0 = shMin[0]; gi.fb1 = shMax[0]; gi.cfb0 = shMin[1]; gi.cfb1 = shMax[1];
        info[li] = gi;
    }
}

// batch-operator variant: bounds + truncated counts for lists supplied by the caller (chunk.cu:1555-1578)
__global__ void k_gather_info(const FeaturePlacement* __restrict__ gF, const CaveFeaturePlacement* __restrict__ gCF,
                              const int* __restrict__ numFeatures, int strideF, int strideCF, GatherInfo* __restrict__ info)
{
    __shared__ int shMin[2], shMax[2];
    const int li = blockIdx.x, tid = threadIdx.x;
    if (tid < 2) { shMin[tid] = 384; shMax[tid] = -1; }
    __syncthreads();
    const int nf = numFeatures[2 * li], nc = numFeatures[2 * li + 1];
    int mnF = 384, mxF = -1, mnC = 384, mxC = -1;
    for (int i = tid; i < nf; i += blockDim.x)
    {
        const FeaturePlacement p = gF[(size_t)li * strideF + i];
        mnF = min(mnF, p.y + c_featureHeightBounds[p.feature][0]);
        mxF = max(mxF, p.y + c_featureHeightBounds[p.feature][1]);
    }
    for (int i = tid; i < nc; i += blockDim.x)
    {
        const CaveFeaturePlacement p = gCF[(size_t)li * strideCF + i];
        mnC = min(mnC, p.y + c_caveFeatureHeightBounds[p.feature][0]);
        mxC = max(mxC, p.y + p.layerHeight + c_caveFeatureHeightBounds[p.feature][1]);
    }
    atomicMin(&shMin[0], mnF); atomicMax(&shMax[0], mxF);
    atomicMin(&shMin[1], mnC); atomicMax(&shMax[1], mxC);
    __syncthreads();
    if (tid == 0)
    {
        GatherInfo gi;
        gi.nF = min(nf, MAX_FEATURES); gi.nCF = min(nc, MAX_CAVE_FEATURES);
        gi.fb0 = shMin[0]; gi.fb1 = shMax[0]; gi.cfb0 = shMin[1]; gi.cfb1 = shMax[1];
        info[li] = gi;
    }
}

// kernFill (chunk.cu:1382-1510). fillList[li] = chunk index into the resident planes (or li itself);
// gathered lists are indexed by li with the given strides.
__global__ void __launch_bounds__(384) k_fill(const int* __restrict__ fillList, const int2* __restrict__ origins,
                                              const float* __restrict__ heightfield, const float* __restrict__ biomeWeights,
                                              const float* __restrict__ layers, const CaveLayer* __restrict__ caveLayers,
                                              const FeaturePlacement* __restrict__ gF, const CaveFeaturePlacement* __restrict__ gCF,
                                              const GatherInfo* __restrict__ info, int strideF, int strideCF, uint8_t* __restrict__ blocks)
{
    __shared__ float shW[NUM_BIOMES];
    __shared__ float shLH[NUM_MATERIALS + 1];
    __shared__ CaveLayer shCL[MAX_CAVE_LAYERS];
    const int li = blockIdx.x >> 8, idx = blockIdx.x & 255;
    const int chunk = fillList ? fillList[li] : li;
    const int y = threadIdx.x;
    if (y < NUM_BIOMES) shW[y] = biomeWeights[(size_t)chunk * (NUM_BIOMES * 256) + y * 256 + idx];
    else if (y < NUM_BIOMES + NUM_MATERIALS) shLH[y - NUM_BIOMES] = layers[(size_t)chunk * (NUM_MATERIALS * 256) + (y - NUM_BIOMES) * 256 + idx];
    else if (y == NUM_BIOMES + NUM_MATERIALS) shLH[NUM_MATERIALS] = heightfield[(size_t)chunk * 256 + idx];
    else if (y < NUM_BIOMES + NUM_MATERIALS + 1 + MAX_CAVE_LAYERS)
        shCL[y - (NUM_BIOMES + NUM_MATERIALS + 1)] = caveLayers[((size_t)chunk * 256 + idx) * MAX_CAVE_LAYERS + (y - (NUM_BIOMES + NUM_MATERIALS + 1))];
    __syncthreads();
    const int2 o = origins[chunk];
    const int wx = o.x + (idx & 15), wz = o.y + (idx >> 4);
    const float height = shLH[NUM_MATERIALS];
    uint8_t block = fill_place_block(shW, shLH, shCL, y, height, wx, wz);
    const GatherInfo gi = info[li];
    uint8_t fblock = 0;
    bool placed = false;
    if (y >= gi.fb0 && y <= gi.fb1)
    {
        const FeaturePlacement* f = gF + (size_t)li * strideF;
        for (int i = 0; i < gi.nF; ++i)
        {
            const FeaturePlacement fp = f[i];
            if (fp.feature == F_NONE) break;
            if (block != B_AIR && !fp.canReplaceBlocks) continue;
            if (y < fp.y + c_featureHeightBounds[fp.feature][0] || y > fp.y + c_featureHeightBounds[fp.feature][1]) continue;
            if (place_feature(fp, wx, y, wz, &fblock)) { placed = true; break; }
        }
    }
    if (!placed && y >= gi.cfb0 && y <= gi.cfb1)
    {
        const CaveFeaturePlacement* f = gCF + (size_t)li * strideCF;
        for (int i = 0; i < gi.nCF; ++i)
        {
            const CaveFeaturePlacement cp = f[i];
            if (cp.feature == CF_NONE) break;
            if (block != B_AIR && !cp.canReplaceBlocks) continue;
            if (y < cp.y + c_caveFeatureHeightBounds[cp.feature][0] || y > cp.y + cp.layerHeight + c_caveFeatureHeightBounds[cp.feature][1]) continue;
            if (place_cave_feature(cp, wx, y, wz, &fblock)) { placed = true; break; }
        }
    }
    blocks[(size_t)chunk * 98304 + (size_t)idx * 384 + y] = placed ? fblock : block;
}

// tryPlaceSingleDecorator (chunk.cu:1634-1677)
__device__ __forceinline__ void try_place_decorator(uint8_t* blocks, int x, int y, int z, const DecoratorGen& gen)
{
    if (y < 0 || y > 383) return;
    const int di = y + 384 * (x + 16 * z);
    const uint8_t cur = blocks[di];
    if (cur != gen.replace) return;
    const int under = gen.fromCeiling ? 1 : -1;
    if (y + under < 0 || y + under > 383) return;
    const uint8_t ub = blocks[di + under];
    if (ub < NUM_NON_SOLID_BLOCKS) return;
    if (gen.numUnder > 0)
    {
        bool ok = false;
        for (int i = 0; i < gen.numUnder; ++i) ok = ok || gen.under[i] == ub;
        if (!ok) return;
    }
    if (gen.second != B_AIR)
    {
        const int over = -under;
        if (y + over < 0 || y + over > 383) return;
        if (blocks[di + over] != gen.replace) return;
        blocks[di + over] = gen.second;
    }
    blocks[di] = gen.block;
}

// placeDecorators (chunk.cu:1679-1747): one sequential RNG stream per chunk -> one thread per chunk
__global__ void k_decorators(const int* __restrict__ fillList, int n, const int2* __restrict__ origins,
                             const float* __restrict__ heightfield, const float* __restrict__ biomeWeights,
                             const CaveLayer* __restrict__ caveLayers, uint8_t* __restrict__ blocks)
{
    const int li = blockIdx.x * blockDim.x + threadIdx.x;
    if (li >= n) return;
    const int chunk = fillList ? fillList[li] : li;
    const int2 o = origins[chunk];
    uint8_t* b = blocks + (size_t)chunk * 98304;
    const float* w = biomeWeights + (size_t)chunk * (NUM_BIOMES * 256);
    Minstd rng = make_rng4(o.x, 0, o.y, 7589341);
    for (int idx = 0; idx < 256; ++idx)
    {
        const int x = idx & 15, z = idx >> 4;
        const int biome = random_biome(w + idx, 256, rng.u01());
        float rand = rng.u01();
        const int g0 = c_decoratorGenRange[biome][0], gn = c_decoratorGenRange[biome][1];
        for (int g = 0; g < gn; ++g)
            if ((rand -= c_decoratorGens[g0 + g].chance) < 0.f)
            {
                try_place_decorator(b, x, (int)heightfield[(size_t)chunk * 256 + idx] + 1, z, c_decoratorGens[g0 + g]);
                break;
            }
        const CaveLayer* cl = caveLayers + ((size_t)chunk * 256 + idx) * MAX_CAVE_LAYERS;
        for (int l = 0; l < MAX_CAVE_LAYERS; ++l)
        {
            const CaveLayer c = cl[l];
            if (c.start == 384) break;
            float bottomRand = rng.u01();
            float topRand = rng.u01();
            const int c0 = c_caveDecoratorGenRange[c.bottomBiome][0], cn = c_caveDecoratorGenRange[c.bottomBiome][1];
            for (int g = 0; g < cn; ++g)
            {
                const DecoratorGen& gen = c_caveDecoratorGens[c0 + g];
                if (gen.fromCeiling)
                {
                    if ((topRand -= gen.chance) < 0.f) try_place_decorator(b, x, c.end, z, gen);
                }
                else
                {
                    if ((bottomRand -= gen.chance) < 0.f) try_place_decorator(b, x, c.start + 1, z, gen);
                }
            }
        }
    }
}


// 64-bit FNV-1a of each filled column (384 bytes), for cheap equality checks of large worlds
__global__ void k_column_hashes(const int* __restrict__ fillList, int n, const uint8_t* __restrict__ blocks, unsigned long long* __restrict__ out)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * 256) return;
    const int li = t >> 8, idx = t & 255;
    const int chunk = fillList[li];
    const uint8_t* b = blocks + (size_t)chunk * 98304 + (size_t)idx * 384;
    unsigned long long h = 14695981039346656037ull;
    for (int y = 0; y < 384; ++y) { h ^= b[y]; h *= 1099511628211ull; }
    out[t] = h;
}

}  // namespace mmg
