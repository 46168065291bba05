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
// S6   k_fill_terrain / k_fill_lush / k_fill_features (see below). The reference launches one kernFill
//      grid PER CHUNK; here one launch of each covers every chunk of the batch.
// S6b  k_decorators: the reference's sequential per-chunk RNG walk, parallel over columns by minstd skip-ahead.
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
                if (!g_caveGridTestHonoured || is_feature_pos(wx, wz, gen.cell, gen.pad, seed))
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

struct GatherInfo { int nF, nCF; int fb0, fb1, cfb0, cfb1; int needNoise, pad; };

// Extent of ONE surface placement where its own first draws fix it (the type's table values are the worst case over all
// draws). PURPLE_MUSHROOM (featurePlacement.hpp:560-640 / place_feature): pos = offset * scale (* 0.5 with probability 0.2),
// and a voxel survives the first rejection only if |pos.x|, |pos.z| <= 35 and pos.y <= height + 12 (each length in that test
// is bounded below by the component: rounding is monotone and sqrt(fl(x^2)) = |x|). The worst case is reach 70 and 120
// blocks of height; a typical mushroom needs a third of both. Same rounded operations as the rasteriser.
__device__ __forceinline__ void surface_feature_extent(const FeaturePlacement& p, int* reach, int* yHi)
{
    *reach = min(c_featureReach[p.feature], 1 << 16);
    *yHi = c_featureHeightBounds[p.feature][1];
    // rasterisers whose first statement rejects the whole placement by its own height (place_feature: F_ICEBERG, F_CORAL,
    // F_MEDIUM_CRYSTAL / F_CRYSTAL): such a placement fills nothing anywhere
    if ((p.feature == F_ICEBERG && p.y > SEA_LEVEL - 32) || (p.feature == F_CORAL && p.y > SEA_LEVEL - 6) ||
        ((p.feature == F_MEDIUM_CRYSTAL || p.feature == F_CRYSTAL) && p.y > 180))
    {
        *reach = -(1 << 20);
        return;
    }
    if (p.feature == F_PURPLE_MUSHROOM)
    {
        Minstd frng = make_rng4(p.x, p.y, p.z, 1293012);
        const float scale = fmaf(frng.u01(), 1.2f, 1.f);
        const bool half = frng.u01() < 0.2f;
        const float height = fmaf(frng.u01(), 30.f, 25.f);
        auto scaled = [&](int v) { const float s = (float)v * scale; return half ? s * 0.5f : s; };
        int t = *reach, h = *yHi;
        while (t > 0 && scaled(t) > 35.f) --t;
        while (h > 0 && scaled(h) > height + 12.f) --h;
        *reach = t;
        *yHi = h;
    }
}

// does a feature at (px, pz) with horizontal reach r touch the 16x16 footprint whose corner is (ox, oz)?
__device__ __forceinline__ bool reach_hits_chunk(int px, int pz, int r, int ox, int oz)
{
    return px + r >= ox && px - r <= ox + 15 && pz + r >= oz && pz - r <= oz + 15;
}

// chunk.cu:1158-1187 + 1555-1578. fillList: chunks to gather for (world raster indices).
// The 49 neighbour lists are concatenated in the reference's order and truncated at its caps
// (2048 / 4096) exactly as the reference does; of that list only the placements whose horizontal
// reach (c_featureReach) touches this chunk are kept, in order. The reference scans all of them per
// voxel; the dropped ones cannot contain a voxel of this chunk, so the blocks are identical.
// The 49 neighbour lists are independent except for where their kept entries go, so the warps of the CTA take them
// round-robin twice: once to count what each list keeps, and - after a prefix sum over the 49 counts - once to write the
// kept entries at their final positions (a list at a time with block-wide ordered compaction was one long dependent chain:
// 95 us per launch against 20 us).
__global__ void __launch_bounds__(256) k_gather_features(const int* __restrict__ fillList, const int2* __restrict__ origins,
                                                         const FeaturePlacement* __restrict__ features,
                                                         const CaveFeaturePlacement* __restrict__ caveFeatures, const int* __restrict__ counts,
                                                         int nx, FeaturePlacement* __restrict__ gF, CaveFeaturePlacement* __restrict__ gCF,
                                                         GatherInfo* __restrict__ info)
{
    __shared__ int shMin[2], shMax[2];
    __shared__ int shN[2][49];         // entries of list k after the reference's truncation at MAX_FEATURES / MAX_CAVE_FEATURES
    __shared__ int shKept[2][50];      // kept entries of list k, then (exclusive prefix) where they start in the output
    const int li = blockIdx.x, chunk = fillList[li];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < 2) { shMin[tid] = 384; shMax[tid] = -1; }
    const int2 o = origins[chunk];
    if (warp < 2)
    {
        // list lengths in the reference's concatenation order, cut where the concatenation reaches the cap (chunk.cu:1158-1187)
        const int cap = warp == 0 ? MAX_FEATURES : MAX_CAVE_FEATURES;
        int base = 0;
        for (int k0 = 0; k0 < 49; k0 += 32)
        {
            const int k = k0 + lane;
            const int c = k < 49 ? counts[2 * (chunk + c_gatherOffsets[k][0] + c_gatherOffsets[k][1] * nx) + warp] : 0;
            int incl = c;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1)
            {
                const int v = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += v;
            }
            const int before = base + incl - c;
            if (k < 49) shN[warp][k] = max(min(c, cap - before), 0);
            base += __shfl_sync(0xffffffffu, incl, 31);
        }
    }
    __syncthreads();
    FeaturePlacement* dstF = gF + (size_t)li * MAX_FEATURES;
    CaveFeaturePlacement* dstC = gCF + (size_t)li * MAX_CAVE_FEATURES;
    int mnF = 384, mxF = -1, mnC = 384, mxC = -1;
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass)
    {
        for (int k = warp; k < 49; k += 8)
        {
            const int nchunk = chunk + c_gatherOffsets[k][0] + c_gatherOffsets[k][1] * nx;
            const FeaturePlacement* sf = features + (size_t)nchunk * kMaxOwnFeatures;
            const CaveFeaturePlacement* sc = caveFeatures + (size_t)nchunk * kMaxOwnCaveFeatures;
            const int nf = shN[0][k], nc = shN[1][k];
            int outF = pass ? shKept[0][k] : 0, outC = pass ? shKept[1][k] : 0;
            for (int i0 = 0; i0 < nf; i0 += 32)
            {
                const int i = i0 + lane;
                FeaturePlacement p;
                bool keep = false;
                int reach = 0, yHi = 0;
                if (i < nf)
                {
                    p = sf[i];
                    surface_feature_extent(p, &reach, &yHi);
                    keep = reach_hits_chunk(p.x, p.z, reach, o.x, o.y);
                }
                const unsigned m = __ballot_sync(0xffffffffu, keep);
                if (pass && keep)
                {
                    dstF[outF + __popc(m & ((1u << lane) - 1u))] = p;
                    mnF = min(mnF, p.y + c_featureHeightBounds[p.feature][0]);
                    mxF = max(mxF, p.y + yHi);
                }
                outF += __popc(m);
            }
            for (int i0 = 0; i0 < nc; i0 += 32)
            {
                const int i = i0 + lane;
                CaveFeaturePlacement p;
                bool keep = false;
                if (i < nc)
                {
                    p = sc[i];
                    keep = reach_hits_chunk(p.x, p.z, c_caveFeatureReach[p.feature], o.x, o.y);
                }
                const unsigned m = __ballot_sync(0xffffffffu, keep);
                if (pass && keep)
                {
                    dstC[outC + __popc(m & ((1u << lane) - 1u))] = p;
                    mnC = min(mnC, p.y + c_caveFeatureHeightBounds[p.feature][0]);
                    mxC = max(mxC, p.y + p.layerHeight + c_caveFeatureHeightBounds[p.feature][1]);
                }
                outC += __popc(m);
            }
            if (!pass && lane == 0) { shKept[0][k] = outF; shKept[1][k] = outC; }
        }
        __syncthreads();
        if (!pass)
        {
            if (tid < 2)
            {
                int run = 0;
                for (int k = 0; k < 49; ++k) { const int c = shKept[tid][k]; shKept[tid][k] = run; run += c; }
                shKept[tid][49] = run;
            }
            __syncthreads();
        }
    }
    atomicMin(&shMin[0], mnF); atomicMax(&shMax[0], mxF);
    atomicMin(&shMin[1], mnC); atomicMax(&shMax[1], mxC);
    __syncthreads();
    if (tid == 0)
    {
        GatherInfo gi;
        gi.nF = shKept[0][49]; gi.nCF = shKept[1][49];
        gi.fb0 = shMin[0]; gi.fb1 = shMax[0]; gi.cfb0 = shMin[1]; gi.cfb1 = shMax[1];
        gi.needNoise = 1; gi.pad = 0;
        info[li] = gi;
    }
}

// Batch-operator form of Chunk::gatherFeaturePlacements (chunk.cu:1158-1196): one CTA per chunk to gather for. The own lists
// of its 49 neighbours (indices into a pool of chunks, in the order of the reference's offset table; < 0 = absent) are
// concatenated in that order, nothing culled, and cut at the caps the reference's fill uploads (2048 / 4096,
// chunk.cu:1573-1578); outCounts holds the untruncated lengths, which is what the reference's host vectors would hold.
__global__ void __launch_bounds__(256) k_gather_concat(const int* __restrict__ neighbours, const FeaturePlacement* __restrict__ features, int strideF,
                                                       const CaveFeaturePlacement* __restrict__ caveFeatures, int strideC,
                                                       const int* __restrict__ counts, FeaturePlacement* __restrict__ outF,
                                                       CaveFeaturePlacement* __restrict__ outC, int* __restrict__ outCounts)
{
    __shared__ int shSrc[49], shN[2][49], shOff[2][50];
    const int li = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < 49)
    {
        const int src = neighbours[li * 49 + tid];
        shSrc[tid] = src;
        shN[0][tid] = src >= 0 ? min(counts[2 * src], strideF) : 0;
        shN[1][tid] = src >= 0 ? min(counts[2 * src + 1], strideC) : 0;
    }
    __syncthreads();
    if (tid < 2)
    {
        int run = 0;
        for (int k = 0; k < 49; ++k) { shOff[tid][k] = run; run += shN[tid][k]; }
        shOff[tid][49] = run;
        outCounts[2 * li + tid] = run;
    }
    __syncthreads();
    for (int k = warp; k < 49; k += 8)
    {
        const int src = shSrc[k];
        if (src < 0) continue;
        const FeaturePlacement* sf = features + (size_t)src * strideF;
        const CaveFeaturePlacement* sc = caveFeatures + (size_t)src * strideC;
        for (int i = lane; i < shN[0][k]; i += 32)
            if (shOff[0][k] + i < MAX_FEATURES) outF[(size_t)li * MAX_FEATURES + shOff[0][k] + i] = sf[i];
        for (int i = lane; i < shN[1][k]; i += 32)
            if (shOff[1][k] + i < MAX_CAVE_FEATURES) outC[(size_t)li * MAX_CAVE_FEATURES + shOff[1][k] + i] = sc[i];
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
    const int nf = min(numFeatures[2 * li], MAX_FEATURES), nc = min(numFeatures[2 * li + 1], MAX_CAVE_FEATURES);
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
        gi.nF = nf; gi.nCF = nc;
        gi.fb0 = shMin[0]; gi.fb1 = shMax[0]; gi.cfb0 = shMin[1]; gi.cfb1 = shMax[1];
        gi.needNoise = 1; gi.pad = 0;
        info[li] = gi;
    }
}

// kernFill (chunk.cu:1382-1510) as a sequence of kernels over one batch of chunks. fillList[li] = chunk index into
// the resident planes (or li itself); gathered lists are indexed by li with the given strides.
//   k_fill_terrain   chunkFillPlaceBlock for every voxel up to the cave-biome step: one CTA per chunk row (16 columns), the
//                    voxels that are not above the terrain and the sea walked as one dense list; STONE / DEEPSLATE /
//                    BLACKSTONE voxels - the only ones getCaveBiome can change - are queued (near / bulk regions).
//   k_fill_rock      getCaveBiome + caveBiomeBlockPostProcess for the queued voxels, one per thread on dense warps;
//                    bulk voxels first decide what CRYSTAL_CAVES would do to them. Voxels that need the LUSH_CAVES
//                    clay / moss decision are queued once more.
//   k_fill_lush      decides the queued voxels, one per thread (they are rare and scattered: evaluated in
//                    place they would occupy 3-4 lanes of a warp for a 27-cell Worley + 9 simplex).
//   k_prepare_placements  per placement of the chunk's gathered lists: the RNG state after seeding and the box its own
//                    first draws allow (Prep).
//   k_fill_features  the placement scan, one CTA per 32-voxel slab of a chunk: the placements' boxes are cut into tiles of
//                    (column, y) pairs that the warps pull from a shared counter; the reference tests up to 2048 + 4096
//                    placements per voxel.
// The reference decides terrain and features in one pass per voxel; the feature test only looks at
// whether the terrain block is AIR, which the lush decision does not change, so the order
// terrain -> lush -> features gives the same blocks.
constexpr int kLushQueueCap = 1 << 22;           // queued voxels per fill batch (overflow is decided in place)
constexpr int kRockQueuePerChunk = 49152;        // rock-queue slots per chunk of a fill batch (typical need ~30 k; overflow is decided in place)

// counters of one fill batch: [0] lush queue length, [2] near-rock queue length, [3] bulk-rock queue length ([2..3] are
// advanced together by one 64-bit atomic). The rock queue has two regions: [0, nearCap) for voxels within 6 blocks of a
// cave floor / ceiling (full getCaveBiome), [nearCap, nearCap + bulkCap) for bulk voxels (only CRYSTAL_CAVES matters).
__device__ __forceinline__ int rock_near_cap(int rockQueueCap) { return (rockQueueCap / 3) & ~31; }

// what k_fill_rock / k_fill_lush would do to a rock voxel, in place - a real function so that its registers and spills stay
// out of k_fill_terrain's main path (it only runs when the rock queue of a batch overflows)
__device__ __noinline__ uint8_t finish_rock_overflow(uint8_t rb, int wx, int y, int wz, float height, int bd, int td)
{
    bool lush = false;
    uint8_t b = finish_rock_block(rb, wx, y, wz, height, bd, td, &lush);
    if (lush) b = lush_block(wx, y, wz);
    return b;
}

// ---- k_fill_terrain, one CTA per chunk ROW (16 columns x = 0 .. 15 of one z): the voxels that are not "above the terrain
// and the sea" are walked as one dense list by 384 threads. Against one 128-thread CTA per column: the row's column data
// (24 weights, 21 layer heights, 32 cave layers per column) arrives by coalesced loads once, the rock voxels of 16 columns
// take ONE queue reservation whose round trip is overlapped with the row's 6 KB of output stores, and the lanes of a warp
// are consecutive voxels that all need work (ncu of the per-column kernel, profiles/r02_k_fill_terrain_v1_src.txt: 45 % of the
// stall samples at the two barriers around the column loads and the queue atomic, 24.5 of 32 lanes active).
// Measured per 128x128-chunk region (profiles/r02_k_fill_terrain_rows.txt): per column 15.9 ms; rows with 256 threads x 3 / 4 / 5
// CTAs per SM 15.4 / 12.9 / 12.1; 320 x 5 11.7; 384 x 4 / 5 11.3 / 11.0; 512 x 4 11.0 - the kernel is latency-bound and wants
// resident warps more than registers (32 per thread at 384 x 5).
#ifndef MMG_TERRAIN_THREADS
#define MMG_TERRAIN_THREADS 384
#endif
constexpr int kRowCols = 16, kRowThreads = MMG_TERRAIN_THREADS;
__global__ void __launch_bounds__(kRowThreads, MMG_TERRAIN_MINBLOCKS) k_fill_terrain(const int* __restrict__ fillList, const int2* __restrict__ origins,
                                                              const float* __restrict__ heightfield, const float* __restrict__ biomeWeights,
                                                              const float* __restrict__ layers, const CaveLayer* __restrict__ caveLayers,
                                                              uint8_t* __restrict__ blocks, uint2* __restrict__ rockQueue, int rockQueueCap,
                                                              int* __restrict__ counters)
{
    __shared__ float shW[kRowCols][NUM_BIOMES];
    __shared__ float shLH[kRowCols][NUM_MATERIALS + 1];
    __shared__ __align__(16) CaveLayer shCL[kRowCols][MAX_CAVE_LAYERS];
    __shared__ ColumnBiomes shCB[kRowCols];
    __shared__ int shStart[kRowCols + 1];                 // dense voxel list: column c owns [shStart[c], shStart[c + 1])
    __shared__ __align__(16) uint8_t shOut[kRowCols * 384];
    __shared__ unsigned short shRock[kRowCols * 384];     // per visited voxel: 0 = not rock, else 0x8000 | bottom depth << 6 | top depth (clamped as pack_rock does)
    __shared__ int shCount[2], shLocal[2], shBase[2], shFlags;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int li = blockIdx.x >> 4, row = blockIdx.x & 15;
    const int chunk = fillList ? fillList[li] : li;
    const int idx0 = row * kRowCols;
    // ---- the row's column data, coalesced: 16 consecutive floats per biome / material plane, 6 KB of contiguous cave layers
    for (int i = t; i < NUM_BIOMES * kRowCols; i += kRowThreads)
        shW[i & 15][i >> 4] = biomeWeights[(size_t)chunk * (NUM_BIOMES * 256) + (i >> 4) * 256 + idx0 + (i & 15)];
    for (int i = t; i < NUM_MATERIALS * kRowCols; i += kRowThreads)
        shLH[i & 15][i >> 4] = layers[(size_t)chunk * (NUM_MATERIALS * 256) + (i >> 4) * 256 + idx0 + (i & 15)];
    {
        static_assert(sizeof(CaveLayer) == 12 && (kRowCols * MAX_CAVE_LAYERS * 12) % 16 == 0, "cave layers of a row are copied as uint4");
        const uint4* src = reinterpret_cast<const uint4*>(caveLayers + ((size_t)chunk * 256 + idx0) * MAX_CAVE_LAYERS);
        uint4* dst = reinterpret_cast<uint4*>(&shCL[0][0]);
        for (int i = t; i < kRowCols * MAX_CAVE_LAYERS * 12 / 16; i += kRowThreads) dst[i] = src[i];
    }
    for (int i = t; i < kRowCols * 384 / 16; i += kRowThreads) reinterpret_cast<uint4*>(shOut)[i] = make_uint4(0u, 0u, 0u, 0u);      // B_AIR == 0
    static_assert(B_AIR == 0, "the row buffer starts as air");
    if (t < kRowCols) shLH[t][NUM_MATERIALS] = heightfield[(size_t)chunk * 256 + idx0 + t];
    if (t == 0) { shCount[0] = shCount[1] = 0; shLocal[0] = shLocal[1] = 0; shFlags = 0; }
    const int2 o = origins[chunk];
    const int wz = o.y + row;
    __syncthreads();
    // ---- per column (two per warp): the biomes of non-zero weight in index order, isOcean (chunk.cu:1225-1231); does any column
    // need the simplex tables? biomePre/PostProcess evaluate noise for these biomes only (biomeFuncs.hpp:385-590); random_biome
    // can only return a biome of weight 0 when it is CORAL_REEF (rand == 0) or the PLAINS fallback, neither of which uses noise
    for (int c = warp; c < kRowCols; c += kRowThreads / 32)
    {
        const float wgt = lane < NUM_BIOMES ? shW[c][lane] : 0.f;
        const unsigned nz = __ballot_sync(0xffffffffu, wgt != 0.f);
        if (wgt != 0.f)
        {
            const int slot = __popc(nz & ((1u << lane) - 1u));
            shCB[c].biome[slot] = (uint8_t)lane;
            shCB[c].weight[slot] = wgt;
        }
        const unsigned ocean = __ballot_sync(0xffffffffu, lane < NUM_OCEAN_BIOMES && wgt > 0.f);
        const unsigned noisy = __ballot_sync(0xffffffffu, wgt > 0.f && (lane == ARCHIPELAGO || lane == MESA || lane == SHREKS_SWAMP || lane == TIANZI_MOUNTAINS ||
                                                                         lane == MOUNTAINS || lane == CRYSTALS));
        if (lane == 0)
        {
            shCB[c].n = __popc(nz); shCB[c].isOcean = ocean != 0u;
            if (noisy) atomicOr(&shFlags, 1);
        }
    }
    if (warp == 0)
    {
        // voxels of column c that are not "above the terrain and the sea" (chunkFillPlaceBlock's first exit, chunk.cu:1213-1217):
        // y <= 128 or y <= height
        int n = 0;
        if (lane < kRowCols)
        {
            const float h = shLH[lane][NUM_MATERIALS];
            int top = SEA_LEVEL;
            if (h > (float)SEA_LEVEL) top = min(383, (int)floorf(h));      // the largest y with (float)y <= h
            n = top + 1;
        }
        int incl = n;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1)
        {
            const int v = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += v;
        }
        if (lane < kRowCols) shStart[lane + 1] = incl;
        if (lane == 0) shStart[0] = 0;
    }
    __syncthreads();
    const bool needNoise = (shFlags & 1) != 0;
    if (needNoise) noise_tab_stage();
    const int total = shStart[kRowCols];
    // ---- pass 1: the block of every listed voxel; rock voxels (the only ones getCaveBiome can change) are marked
    for (int i0 = warp * 32; i0 < total; i0 += kRowThreads)
    {
        const int i = i0 + lane;
        int cls = -1;
        if (i < total)
        {
            int c = 0;
#pragma unroll
            for (int d = 8; d > 0; d >>= 1) if (shStart[c + d] <= i) c += d;
            const int y = i - shStart[c];
            const float height = shLH[c][NUM_MATERIALS];
            bool rock = false;
            int bd = -384, td = -384;
            const uint8_t blk = fill_place_block(shCB[c], shLH[c], shCL[c], y, height, o.x + c, wz, &rock, &bd, &td);
            shOut[c * 384 + y] = blk;
            unsigned short info = 0;
            if (rock)
            {
                const unsigned b6 = (bd < 0 || bd > 62) ? 63u : (unsigned)bd, t6 = (td < 0 || td > 62) ? 63u : (unsigned)td;
                info = (unsigned short)(0x8000u | b6 << 6 | t6);
                cls = rock_is_bulk(bd, td) ? 1 : 0;
            }
            shRock[c * 384 + y] = info;
        }
        const unsigned bNear = __ballot_sync(0xffffffffu, cls == 0), bBulk = __ballot_sync(0xffffffffu, cls == 1);
        if (lane == 0)
        {
            if (bNear) atomicAdd(&shCount[0], __popc(bNear));
            if (bBulk) atomicAdd(&shCount[1], __popc(bBulk));
        }
    }
    __syncthreads();
    // ---- one reservation for the row's rock voxels; its round trip overlaps the row's output stores
    const int nearCap = rock_near_cap(rockQueueCap), bulkCap = rockQueueCap - nearCap;
    unsigned long long old = 0ull;
    const unsigned nNear = (unsigned)shCount[0], nBulk = (unsigned)shCount[1];
    if (t == 0 && (nNear | nBulk)) old = atomicAdd(reinterpret_cast<unsigned long long*>(counters + 2), (unsigned long long)nBulk << 32 | nNear);
    {
        uint4* out = reinterpret_cast<uint4*>(blocks + (size_t)chunk * 98304 + (size_t)idx0 * 384);      // 16 columns = 6144 contiguous bytes
        for (int i = t; i < kRowCols * 384 / 16; i += kRowThreads) out[i] = reinterpret_cast<const uint4*>(shOut)[i];
    }
    if (t == 0)
    {
        shBase[0] = (int)(unsigned)(old & 0xffffffffull);
        shBase[1] = (int)(unsigned)(old >> 32);
        if (shBase[0] + (int)nNear > nearCap || shBase[1] + (int)nBulk > bulkCap) shFlags |= 2;      // some voxel will not fit the queue
    }
    __syncthreads();
    const bool overflow = (shFlags & 2) != 0;
    if (overflow && !needNoise) noise_tab_stage();      // the in-place path below evaluates noise
    // ---- pass 2: the marked voxels go to the queue (near / bulk regions)
    for (int i0 = warp * 32; i0 < total; i0 += kRowThreads)
    {
        const int i = i0 + lane;
        int cls = -1, c = 0, y = 0;
        unsigned info = 0;
        if (i < total)
        {
#pragma unroll
            for (int d = 8; d > 0; d >>= 1) if (shStart[c + d] <= i) c += d;
            y = i - shStart[c];
            info = shRock[c * 384 + y];
            if (info)
            {
                const unsigned b6 = (info >> 6) & 63u, t6 = info & 63u;
                cls = (b6 > 6u && t6 > 6u) ? 1 : 0;      // rock_is_bulk on the clamped depths: 63 stands for "negative or beyond 62"
            }
        }
        const unsigned bNear = __ballot_sync(0xffffffffu, cls == 0), bBulk = __ballot_sync(0xffffffffu, cls == 1);
        int baseN = 0, baseB = 0;
        if (lane == 0)
        {
            if (bNear) baseN = atomicAdd(&shLocal[0], __popc(bNear));
            if (bBulk) baseB = atomicAdd(&shLocal[1], __popc(bBulk));
        }
        baseN = __shfl_sync(0xffffffffu, baseN, 0); baseB = __shfl_sync(0xffffffffu, baseB, 0);
        if (cls >= 0)
        {
            const uint8_t blk = shOut[c * 384 + y];
            const int idx = idx0 + c;
            const unsigned kind = blk == B_STONE ? 0u : (blk == B_DEEPSLATE ? 1u : 2u);
            const uint2 rec = make_uint2((unsigned)chunk, (unsigned)(idx * 384 + y) | kind << 17 | ((info >> 6) & 63u) << 19 | (info & 63u) << 25);      // == pack_rock
            const int slot = cls ? shBase[1] + baseB + __popc(bBulk & ((1u << lane) - 1u)) : shBase[0] + baseN + __popc(bNear & ((1u << lane) - 1u));
            if (slot < (cls ? bulkCap : nearCap)) rockQueue[(cls ? nearCap : 0) + slot] = rec;
            else
            {
                // queue full (rare: the batch has more rock voxels than queue slots): same result, in place
                uint8_t rb; int c2, v2, bd, td;
                unpack_rock(rec, &c2, &v2, &rb, &bd, &td);
                blocks[(size_t)chunk * 98304 + (size_t)idx * 384 + y] = finish_rock_overflow(rb, o.x + c, y, wz, shLH[c][NUM_MATERIALS], bd, td);
            }
        }
    }
}

// getCaveBiome + caveBiomeBlockPostProcess for the queued rock voxels, one per thread on dense warps. In k_fill_terrain
// the same work ran on whatever lanes of a 32-voxel column run happened to be rock below the surface (22 of 32 on average).
// Near voxels take the full path. Bulk voxels (about 4 of 5) only ask whether the biome is CRYSTAL_CAVES, which skips one
// simplex3 always and up to twelve simplex2 (cave_biome_is_crystal) - and only where CRYSTAL_CAVES would change the block
// at all: its rule (cave_biome_post_process) turns the voxel into QUARTZ where one simplex3 is < -0.25 and otherwise STONE /
// DEEPSLATE into their cobbled form with probability 0.5 / 0.4 by a position hash; neither depends on the biome noise, so
// both are decided first (1 simplex3 + 1 sin) and the 10+ simplex3 of the biome question are spent on the voxels that would
// change (about half), compacted through a per-warp queue so that they run on full warps. The bulk region starts on a
// warp boundary.
// (Near and bulk voxels as two kernels, each with its own register allocation and occupancy target, measured in round 2: 24.3 ms
// against 24.6 per 128x128 region - not worth a second launch per batch.)
__global__ void __launch_bounds__(128, MMG_ROCK_MINBLOCKS) k_fill_rock(const int2* __restrict__ origins, const float* __restrict__ heightfield,
                                                      const uint2* __restrict__ rockQueue, int rockQueueCap, uint8_t* __restrict__ blocks,
                                                      uint2* __restrict__ lushQueue, int* __restrict__ counters)
{
    __shared__ uint2 shPending[4 * 64];      // per warp: bulk voxels CRYSTAL_CAVES would change, waiting for a full warp
    const int nearCap = rock_near_cap(rockQueueCap), bulkCap = rockQueueCap - nearCap;
    const int nNear = min(counters[2], nearCap), nBulk = min(counters[3], bulkCap);
    const int nearPad = (nNear + 31) & ~31, n = nearPad + nBulk;
    if (blockIdx.x * blockDim.x >= n) return;
    noise_tab_stage();
    const int lane = threadIdx.x & 31;
    uint2* wq = shPending + (threadIdx.x >> 5) * 64;
    int qn = 0;
#if MMG_ROCK_TWO_LEVEL
    // second level: the voxels whose `rocky` noise is not 0 (the biome question stops there for the others) wait for a full warp again
    // before the 2-D noises and the draw
    __shared__ uint2 shPend2[4 * 64];
    __shared__ float4 shWarped2[4 * 64];
    uint2* wq2 = shPend2 + (threadIdx.x >> 5) * 64;
    float4* ww2 = shWarped2 + (threadIdx.x >> 5) * 64;
    int qn2 = 0;
    auto drain2 = [&]() {
        const int cnt = min(qn2, 32);
        qn2 -= cnt;
        const int slot = qn2 + (lane < cnt ? lane : 0);
        const uint2 e = wq2[slot];
        const float4 wp = ww2[slot];
        __syncwarp();
        if (lane < cnt)
        {
            const int chunk = (int)e.x, voxel = (int)(e.y & 0x1ffffu), idx = voxel / 384, y = voxel - idx * 384;
            const int2 o = origins[chunk];
            if (cave_crystal_decide(o.x + (idx & 15), y, o.y + (idx >> 4), heightfield[(size_t)chunk * 256 + idx], 190249401, wp))
                blocks[(size_t)chunk * 98304 + voxel] = ((e.y >> 19) & 1u) ? B_QUARTZ : ((((e.y >> 17) & 3u) == 0u) ? B_COBBLESTONE : B_COBBLED_DEEPSLATE);
        }
    };
#endif
    // pops up to 32 pending bulk voxels: x = chunk, y = voxel | kind << 17 | (quartz ? 1 << 19 : 0)
    auto drain = [&]() {
        const int cnt = min(qn, 32);
        qn -= cnt;
        const uint2 e = wq[qn + (lane < cnt ? lane : 0)];
        __syncwarp();
#if MMG_ROCK_TWO_LEVEL
        bool rocky = false;
        float4 wp = make_float4(0.f, 0.f, 0.f, 0.f);
        if (lane < cnt)
        {
            const int chunk = (int)e.x, voxel = (int)(e.y & 0x1ffffu), idx = voxel / 384, y = voxel - idx * 384;
            const int2 o = origins[chunk];
            rocky = cave_crystal_rocky(o.x + (idx & 15), y, o.y + (idx >> 4), &wp);
        }
        const unsigned m = __ballot_sync(0xffffffffu, rocky);
        if (rocky)
        {
            const int slot = qn2 + __popc(m & ((1u << lane) - 1u));
            wq2[slot] = e; ww2[slot] = wp;
        }
        qn2 += __popc(m);
        __syncwarp();
        if (qn2 >= 32) drain2();
#else
        if (lane < cnt)
        {
            const int chunk = (int)e.x, voxel = (int)(e.y & 0x1ffffu), idx = voxel / 384, y = voxel - idx * 384;
            const int2 o = origins[chunk];
            if (cave_biome_is_crystal(o.x + (idx & 15), y, o.y + (idx >> 4), heightfield[(size_t)chunk * 256 + idx], 190249401))
                blocks[(size_t)chunk * 98304 + voxel] = ((e.y >> 19) & 1u) ? B_QUARTZ : ((((e.y >> 17) & 3u) == 0u) ? B_COBBLESTONE : B_COBBLED_DEEPSLATE);
        }
#endif
    };
    for (int i0 = blockIdx.x * blockDim.x; i0 < n; i0 += gridDim.x * blockDim.x)
    {
        const int i = i0 + threadIdx.x;
        const bool bulk = i >= nearPad;      // warp-uniform
        if (bulk)
        {
            bool change = false;
            uint2 e = make_uint2(0u, 0u);
            if (i < n)
            {
                e = rockQueue[nearCap + (i - nearPad)];
                const int chunk = (int)e.x, voxel = (int)(e.y & 0x1ffffu), idx = voxel / 384, y = voxel - idx * 384;
                const unsigned kind = (e.y >> 17) & 3u;      // 0 STONE, 1 DEEPSLATE, 2 BLACKSTONE (pack_rock)
                const int2 o = origins[chunk];
                const int wx = o.x + (idx & 15), wz = o.y + (idx >> 4);
                // cave_biome_post_process, CB_CRYSTAL_CAVES
                const float s = (float)(wx + wz);
                const float quartz = simplex3<true>((float)(wx + y) * 0.05f, (float)(wz + 5819323) * 0.05f, (s + s) * 0.05f);
                const bool isQuartz = quartz < -0.25f;
                change = isQuartz;
                if (!isQuartz && kind != 2u)
                    change = hash_fract(fmaf((float)wz, 640.88f, fmaf((float)wx, 238.68f, (float)y * 491.28f))) < (kind == 0u ? 0.5f : 0.4f);
                e.y = (e.y & 0x7ffffu) | (isQuartz ? 1u << 19 : 0u);
            }
            const unsigned m = __ballot_sync(0xffffffffu, change);
            if (change) wq[qn + __popc(m & ((1u << lane) - 1u))] = e;
            qn += __popc(m);
            __syncwarp();
            if (qn >= 32) drain();
            continue;
        }
        bool lush = false;
        int chunk = 0, voxel = 0, wx = 0, wz = 0, y = 0;
        if (i < nNear)
        {
            uint8_t rockBlock;
            int bd, td;
            unpack_rock(rockQueue[i], &chunk, &voxel, &rockBlock, &bd, &td);
            const int idx = voxel / 384;
            y = voxel - idx * 384;
            const int2 o = origins[chunk];
            wx = o.x + (idx & 15); wz = o.y + (idx >> 4);
            const float height = heightfield[(size_t)chunk * 256 + idx];
            const uint8_t block = finish_rock_block(rockBlock, wx, y, wz, height, bd, td, &lush);
            if (block != rockBlock) blocks[(size_t)chunk * 98304 + voxel] = block;
        }
        // warp-aggregated append of the voxels that need the lush-cave clay / moss decision
        const unsigned m = __ballot_sync(0xffffffffu, lush);
        if (m)
        {
            const int leader = __ffs(m) - 1;
            int base = 0;
            if (lane == leader) base = atomicAdd(&counters[0], __popc(m));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (lush)
            {
                const int slot = base + __popc(m & ((1u << lane) - 1u));
                if (slot < kLushQueueCap) lushQueue[slot] = make_uint2((unsigned)chunk, (unsigned)voxel);
                else blocks[(size_t)chunk * 98304 + voxel] = lush_block(wx, y, wz);
            }
        }
    }
    if (qn > 0) drain();
#if MMG_ROCK_TWO_LEVEL
    if (qn2 > 0) drain2();
#endif
}

__global__ void __launch_bounds__(128) k_fill_lush(const int2* __restrict__ origins, const uint2* __restrict__ lushQueue,
                                                   const int* __restrict__ counters, uint8_t* __restrict__ blocks)
{
    const int n = min(counters[0], kLushQueueCap);
    if (blockIdx.x * blockDim.x >= n) return;
    noise_tab_stage();
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
        const uint2 e = lushQueue[i];
        const int chunk = (int)e.x, idx = (int)e.y / 384, y = (int)e.y % 384;
        const int2 o = origins[chunk];
        blocks[(size_t)chunk * 98304 + e.y] = lush_block(o.x + (idx & 15), y, o.y + (idx >> 4));
    }
}

// What k_fill_features needs to know about a placement before it touches the rasteriser: the y band it can
// fill (reference bound intersected with the type's own band, clipped to the world), the columns of THIS
// chunk its horizontal reach covers, and whether it may overwrite terrain. lo > hi: cannot touch the chunk.
#ifdef MMG_FEATURE_STATS
__device__ unsigned g_debugFeatureMask = 0xffffffffu;      // mmgen_debug_feature_mask: developer build only
#define MMG_FEATURE_ENABLED(bit) ((g_debugFeatureMask >> (bit)) & 1u)
// developer build (tools/feature_census.py): per feature type [0..20] surface, [32..41] cave: warp clock cycles spent in the
// rasteriser loop, (column, y) pairs put on the rasteriser, pairs that reached the rasteriser call, hits
__device__ unsigned long long g_featStats[64][4];
#else
#define MMG_FEATURE_ENABLED(bit) true
#endif
struct Prep { short lo, hi; unsigned char xr, zr, canReplace, feature; uint32_t seed; };      // xr = x0 | x1 << 4 (local 0..15), zr likewise;
                                                                                              // seed = state of the placement's own RNG

// One CTA per chunk of the batch: Prep records of the chunk's (already reach-culled, ordered) lists. NONE ends
// a list scan in the reference (chunk.cu:1448-1451, 1477-1480): the list lengths are cut at the first NONE.
__global__ void __launch_bounds__(256) k_prepare_placements(const int* __restrict__ fillList, const int2* __restrict__ origins,
                                                            const FeaturePlacement* __restrict__ gF, const CaveFeaturePlacement* __restrict__ gCF,
                                                            GatherInfo* __restrict__ info, int strideF, int strideCF,
                                                            Prep* __restrict__ prepF, Prep* __restrict__ prepC)
{
    __shared__ int shFirstNone[2], shNoise;
    const int li = blockIdx.x, chunk = fillList ? fillList[li] : li, tid = threadIdx.x;
    if (tid < 2) shFirstNone[tid] = 0x7fffffff;
    if (tid == 2) shNoise = 0;
    __syncthreads();
    const int2 o = origins[chunk];
    const GatherInfo gi = info[li];
    auto columns = [&](int px, int pz, int r, Prep* k) -> bool {
        const int x0 = max(px - r, o.x) - o.x, x1 = min(px + r, o.x + 15) - o.x;
        const int z0 = max(pz - r, o.y) - o.y, z1 = min(pz + r, o.y + 15) - o.y;
        if (x0 > x1 || z0 > z1) return false;
        k->xr = (unsigned char)(x0 | x1 << 4); k->zr = (unsigned char)(z0 | z1 << 4);
        return true;
    };
    for (int i = tid; i < gi.nF; i += 256)
    {
        const FeaturePlacement p = gF[(size_t)li * strideF + i];
        Prep k;
        k.lo = 1; k.hi = 0; k.xr = k.zr = 0; k.canReplace = p.canReplaceBlocks ? 1 : 0; k.feature = p.feature;
        k.seed = make_rng4(p.x, p.y, p.z, 1293012).x;      // featurePlacement.hpp:153: seeded once per placement here, not per voxel
        int reach = 0, yHi = 0;
        if (p.feature == F_NONE) atomicMin(&shFirstNone[0], i);
        else if (!MMG_FEATURE_ENABLED(p.feature)) { /* developer build: type switched off */ }
        else if (surface_feature_extent(p, &reach, &yHi), columns(p.x, p.z, reach, &k))
        {
            k.lo = (short)max(p.y + c_featureHeightBounds[p.feature][0], 0);
            k.hi = (short)min(p.y + yHi, 383);
            shNoise = 1;      // surface rasterisers use simplex / Worley noise
        }
        prepF[(size_t)li * strideF + i] = k;
    }
    for (int i = tid; i < gi.nCF; i += 256)
    {
        const CaveFeaturePlacement p = gCF[(size_t)li * strideCF + i];
        Prep k;
        k.lo = 1; k.hi = 0; k.xr = k.zr = 0; k.canReplace = p.canReplaceBlocks ? 1 : 0; k.feature = p.feature;
        // featurePlacement.hpp:1119: the placement's own RNG is seeded once here; its first draw fixes the size of
        // most cave features, so the box is cut to what THIS placement can fill (each bound below is the rasteriser's
        // own rejection test in place_cave_feature evaluated with the same rounded operations, only tighter than the type's)
        Minstd frng = make_rng4(p.x, p.y, p.z, 398132);
        k.seed = frng.x;
        const float u0 = frng.u01();
        const int lh = p.layerHeight;
        int reach = min(c_caveFeatureReach[p.feature], 1 << 16), lo2 = -1024, hi2 = 1024;
        switch (p.feature)
        {
        case CF_CAVE_VINE:
        {
            int height = (int)fmaf(u0, 12.f, 3.f);
            height = height < lh ? height : lh;
            lo2 = p.y + lh - height; hi2 = p.y + lh;                    // ty in [-height, 0]
            break;
        }
        case CF_GLOWSTONE_CLUSTER:
        {
            // |top * scale| <= 6 with top.y scaled by 1.35 first; every component of a vector bounds its length from below
            const float scale = fmaf(u0, 0.5f, 1.f);
            int t = 6, v = 4;
            while (t > 0 && (float)t * scale > 6.f) --t;
            while (v > 0 && ((float)v * 1.35f) * scale > 6.f) --v;
            reach = t;
            lo2 = p.y + lh - v; hi2 = p.y + lh + v;
            break;
        }
        case CF_STORMLIGHT_SPHERE:
        case CF_CEILING_STORMLIGHT_SPHERE:
        {
            // dist <= radius and dist >= |component| (integer offsets: the squares and their sum are exact)
            const int R = (int)floorf(fmaf(u0, 4.f, 3.5f));
            reach = R;
            const int c = p.y + (p.feature == CF_CEILING_STORMLIGHT_SPHERE ? lh : 0);
            lo2 = c - R; hi2 = c + R;
            break;
        }
        case CF_WARPED_FUNGUS: hi2 = p.y + (int)fmaf(u0, 3.0f, 2.5f) + 3; break;      // fy <= height + 3
        case CF_AMBER_FUNGUS: hi2 = p.y + (int)fmaf(u0, 4.5f, 4.5f) + 3; break;
        default: break;
        }
        if (p.feature == CF_NONE) atomicMin(&shFirstNone[1], i);
        else if (!MMG_FEATURE_ENABLED(21 + p.feature)) { /* developer build: type switched off */ }
        else if (columns(p.x, p.z, reach, &k))
        {
            const int* band = c_caveFeatureBand[p.feature];
            k.lo = (short)max(max(max(p.y + c_caveFeatureHeightBounds[p.feature][0], p.y + band[0] + (band[1] ? lh : 0)), lo2), 0);
            k.hi = (short)min(min(min(p.y + lh + c_caveFeatureHeightBounds[p.feature][1], p.y + band[2] + (band[3] ? lh : 0)), hi2), 383);
            if (p.feature == CF_GLOWSTONE_CLUSTER || p.feature == CF_WARPED_FUNGUS || p.feature == CF_AMBER_FUNGUS) shNoise = 1;
        }
        prepC[(size_t)li * strideCF + i] = k;
    }
    __syncthreads();
    if (tid == 0)
    {
        GatherInfo g2 = gi;
        g2.nF = min(gi.nF, shFirstNone[0]);
        g2.nCF = min(gi.nCF, shFirstNone[1]);
        g2.needNoise = shNoise;
        info[li] = g2;
    }
}

constexpr int kSlab = 32;                  // voxels of a column per CTA of k_fill_features
constexpr int kSlabPitch = 257;            // shBest[yy][col], padded: lanes that differ in yy hit different banks
constexpr unsigned kNoBest = 0xffffffffu;
constexpr int kRound = 1024;               // placements scanned per round (the round's active list lives in shared memory)
#ifndef MMG_QCAP
#define MMG_QCAP 96
#endif
constexpr int kFeatQueue = MMG_QCAP;       // candidates a warp collects before it calls the rasteriser (>= 96: one round adds up to 64)
// ceil(65536 / n), n = 1..16 ([0] unused): (c * c_recip16[n]) >> 16 = c / n for c < 256
__constant__ const unsigned c_recip16[17] = {0, 65536, 32768, 21846, 16384, 13108, 10923, 9363, 8192, 7282, 6554, 5958, 5462, 5042, 4682, 4370, 4096};

// columns of a placement's box that make one unit of work (tile): about 512 (column, y) pairs, at least one warp of columns
#ifndef MMG_TILE
#define MMG_TILE 512
#endif
__device__ __forceinline__ int cols_per_tile(int ny) { return ny >= MMG_TILE / 32 ? 32 : (ny >= MMG_TILE / 64 ? 64 : (ny >= MMG_TILE / 128 ? 128 : 256)); }

// Placement scan of one 32-voxel slab of a chunk (12 slabs per chunk). Work is distributed by PLACEMENT, not by voxel: a
// placement's box clipped to this chunk and slab is a set of columns with a y band, cut into tiles of columns, and the
// warps of the CTA pull tiles from a shared counter - so the lanes of a warp rasterise the same feature on different voxels
// instead of one voxel testing hundreds of candidates (a column of a crystal-cave chunk is within reach of ~400 stormlight
// spheres), and a mushroom's 8000 pairs are shared by all warps while a vine's 15 occupy one warp for one step.
//   round (kRound placements): every thread tests the Prep records of its placements against the slab; the ones that touch
//     it enter the round's active list with their first tile number (one packed shared-memory atomic hands out both, so
//     tile numbers grow with the slot and a tile finds its placement by binary search);
//   tile: each lane takes a COLUMN of the box and builds the 32-bit mask of its candidate voxels in one step - the band,
//     ANDed with the column's air mask unless the placement may replace blocks (a cave feature's box is mostly rock: 11-15 %
//     of the pairs are candidates, and the voxel-at-a-time filter that used to find them was 57 % of the kernel's
//     instructions, ncu profiles/r01_k_fill_features_v8.txt). The set bits are handed to the warp's queue two per lane and
//     round; a voxel some placement already claimed (per-column claim mask) is dropped there if the claim is an earlier
//     placement's. The rasteriser runs on full warps popped from the queue, which is kept while the warp's next tile
//     belongs to the same placement.
// The reference's rule "the first placement in list order that contains the voxel wins, surface list before cave list"
// (chunk.cu:1444-1500) becomes an atomicMin over (list position << 8 | block) per voxel, so the order in which tiles are
// processed does not matter.
// Measured and dropped in round 2 (per 256x256 world; the kernel is bound by instruction fetch - stalled_no_instruction 4.4
// warps per issue, profiles/r02_k_fill_features_v1.txt - and by the latency of its short per-tile chains, not by the number
// of rasteriser calls): a surface pass + a cave pass so that each pass's rasterisers fit the instruction cache (214 ms
// against 162); column-level ball / disc / diamond tests that cut the rasteriser calls by 29 % (spheres: only hits reach
// the rasteriser) but lengthen the per-column chain (174 ms against 147, profiles/r02_fill_features_variants.txt).
// CTA shape, measured per 128x128-chunk region (profiles/r02_k_fill_features_shapes.txt): 256 threads x 4 CTAs per SM 37.2 ms,
// 384 x 3 34.1, 512 x 2 31.3 (the same 32 warps and 64 registers as 256 x 4, but half the shared memory - 94 KB instead of
// 220 KB per SM - which leaves the L1 its capacity for the Prep records and placement lists), 768 x 1 33.9; every shape that
// drops below 56 registers per thread is slower (384 x 4: 44.3, 512 x 3: 39.4).
#ifndef MMG_FEAT_THREADS
#define MMG_FEAT_THREADS 512
#endif
#ifndef MMG_FEAT_MINBLOCKS
#define MMG_FEAT_MINBLOCKS 2
#endif
constexpr int kFeatThreads = MMG_FEAT_THREADS;
static_assert(kFeatThreads >= 256 && kFeatThreads % 32 == 0, "thread t < 256 owns column t of the slab");
__global__ void __launch_bounds__(kFeatThreads, MMG_FEAT_MINBLOCKS) k_fill_features(const int* __restrict__ fillList, const int2* __restrict__ origins,
                                                          const FeaturePlacement* __restrict__ gF, const CaveFeaturePlacement* __restrict__ gCF,
                                                          const Prep* __restrict__ prepF, const Prep* __restrict__ prepC,
                                                          const GatherInfo* __restrict__ info, int strideF, int strideCF,
                                                          uint8_t* __restrict__ blocks)
{
    __shared__ unsigned shBest[kSlab * kSlabPitch];
    __shared__ unsigned shAir[256];                          // per column: bit yy = the terrain block is AIR
    __shared__ unsigned shClaim[256];                        // per column: bit yy = some placement has claimed the voxel (shBest != kNoBest)
    __shared__ unsigned shActBase[kRound];                   // active list of the round: first tile number ...
    __shared__ unsigned short shActE[kRound];                // ... and list position of the placement
    __shared__ unsigned shPacked;                            // tiles handed out << 11 | active placements
    __shared__ int shNextTile;
    __shared__ float shGeom[(kFeatThreads / 32) * kMushroomGeomFloats];        // per warp: purple_mushroom_geom of the placement being rasterised
    __shared__ unsigned short shQueue[(kFeatThreads / 32) * kFeatQueue];       // per warp: candidates waiting for the rasteriser
    const int slab = blockIdx.x % 12, li = blockIdx.x / 12;
    const int chunk = fillList ? fillList[li] : li;
    const int t = threadIdx.x, y0 = slab * kSlab, y1 = y0 + kSlab - 1;
    const GatherInfo gi = info[li];
    const bool segF = gi.nF > 0 && y0 <= gi.fb1 && y1 >= gi.fb0;
    const bool segC = gi.nCF > 0 && y0 <= gi.cfb1 && y1 >= gi.cfb0;
    if (!segF && !segC) return;
    const int2 o = origins[chunk];
    // thread t owns column t of the slab (32 consecutive block IDs, 16-byte aligned)
    uint8_t* colPtr = blocks + (size_t)chunk * 98304 + (size_t)(t & 255) * 384 + y0;
    if (t < 256)
    {
        const uint4 a = reinterpret_cast<const uint4*>(colPtr)[0], b = reinterpret_cast<const uint4*>(colPtr)[1];
        const unsigned w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        unsigned air = 0u;
#pragma unroll
        for (int k = 0; k < 8; ++k)
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (((w[k] >> (8 * j)) & 0xffu) == (unsigned)B_AIR) air |= 1u << (4 * k + j);
        shAir[t] = air;
        shClaim[t] = 0u;
    }
    for (int i = t; i < kSlab * kSlabPitch; i += kFeatThreads) shBest[i] = kNoBest;
    if (gi.needNoise) noise_tab_stage();
    const int lane = t & 31;
    const int nF = segF ? gi.nF : 0, nC = segC ? gi.nCF : 0, nTot = nF + nC;
    const FeaturePlacement* f = gF + (size_t)li * strideF;
    const CaveFeaturePlacement* cf = gCF + (size_t)li * strideCF;
    const Prep* pf = prepF + (size_t)li * strideF;
    const Prep* pc = prepC + (size_t)li * strideCF;
    unsigned short* wq = shQueue + (t >> 5) * kFeatQueue;
    float* wgeom = shGeom + (t >> 5) * kMushroomGeomFloats;

    if (t == 0 && slab == 0) atomicAdd(&g_work[W_S6_PLACEMENTS], (unsigned long long)(gi.nF + gi.nCF));
    unsigned nPairs = 0u, nRast = 0u;      // this warp's work counters (lane 0 adds them up at the end)
    // the warp's current placement
    int curE = -1, qn = 0;
    bool cave = false;
    Prep k = {};
    FeaturePlacement fp = {};
    CaveFeaturePlacement cp = {};
    unsigned key = 0u;
    // rasterises up to 32 queued voxels of the current placement
    auto drain = [&]() {
        const int n = min(qn, 32);
        qn -= n;
        nRast += (unsigned)n;
        const int code = wq[qn + (lane < n ? lane : 0)];
        __syncwarp();
        if (lane < n)
        {
            const int col = code >> 5, yy = code & 31, x = col & 15, z = col >> 4, y = y0 + yy;
            uint8_t fb = 0;
            const bool hit = cave ? place_cave_feature(cp, o.x + x, y, o.y + z, k.seed, &fb) : place_feature(fp, o.x + x, y, o.y + z, k.seed, wgeom, &fb);
            if (hit)
            {
                atomicMin(&shBest[yy * kSlabPitch + col], key | fb);
                atomicOr(&shClaim[col], 1u << yy);
            }
#ifdef MMG_FEATURE_STATS
            atomicAdd(&g_featStats[(cave ? 32 : 0) + k.feature][2], 1ull);
            if (hit) atomicAdd(&g_featStats[(cave ? 32 : 0) + k.feature][3], 1ull);
#endif
        }
    };

    for (int r0 = 0; r0 < nTot; r0 += kRound)
    {
        __syncthreads();                                      // first round: staging above; later rounds: the previous round's lists are free
        if (t == 0) { shPacked = 0u; shNextTile = 0; }
        __syncthreads();
        const int r1 = min(r0 + kRound, nTot);
        for (int e = r0 + t; e < r1; e += kFeatThreads)
        {
            const Prep q = e >= nF ? pc[e - nF] : pf[e];
            const int lo = max((int)q.lo, y0), hi = min((int)q.hi, y1);
            if (lo > hi) continue;
            const int ncols = ((q.xr >> 4) - (q.xr & 15) + 1) * ((q.zr >> 4) - (q.zr & 15) + 1), cpt = cols_per_tile(hi - lo + 1);
            const unsigned old = atomicAdd(&shPacked, (unsigned)((ncols + cpt - 1) / cpt) << 11 | 1u);
            shActE[old & 0x7ffu] = (unsigned short)e;
            shActBase[old & 0x7ffu] = old >> 11;
        }
        __syncthreads();
        const int nAct = (int)(shPacked & 0x7ffu), nTiles = (int)(shPacked >> 11);
        for (;;)
        {
            int tile = 0;
            if (lane == 0) tile = atomicAdd(&shNextTile, 1);
            tile = __shfl_sync(0xffffffffu, tile, 0);
            if (tile >= nTiles) break;
            int s0 = 0, s1 = nAct - 1;                        // last slot whose first tile is <= tile
            while (s0 < s1)
            {
                const int mid = (s0 + s1 + 1) >> 1;
                if ((int)shActBase[mid] <= tile) s0 = mid; else s1 = mid - 1;
            }
            const int e = shActE[s0];
            if (e != curE)
            {
                while (qn > 0) drain();                       // another placement: the queue is rasterised to the end
                curE = e;
                cave = e >= nF;
                k = cave ? pc[e - nF] : pf[e];
                key = (unsigned)e << 8;
                if (cave) cp = cf[e - nF];
                else
                {
                    fp = f[e];
                    if (k.feature == F_PURPLE_MUSHROOM)
                    {
                        __syncwarp();
                        purple_mushroom_geom(k.seed, wgeom);      // every lane writes the same values
                        __syncwarp();
                    }
                }
            }
            const int lo = max((int)k.lo, y0), hi = min((int)k.hi, y1);
            const int x0 = k.xr & 15, nxc = (k.xr >> 4) - x0 + 1, z0 = k.zr & 15, nzc = (k.zr >> 4) - z0 + 1;
            const int ny = hi - lo + 1, cpt = cols_per_tile(ny);
            const int c0 = (tile - (int)shActBase[s0]) * cpt, c1 = min(c0 + cpt, nxc * nzc);
            const unsigned band = (0xffffffffu >> (32 - ny)) << (lo - y0);      // 1 <= ny <= 32
            nPairs += (unsigned)((c1 - c0) * ny);
            const unsigned rnx = c_recip16[nxc];
#ifdef MMG_FEATURE_STATS
            if (lane == 0) atomicAdd(&g_featStats[(cave ? 32 : 0) + k.feature][1], (unsigned long long)((c1 - c0) * ny));
#endif
            for (int cb = c0; cb < c1; cb += 32)
            {
                // one column per lane: its candidate voxels as a bit mask
                const int c = cb + lane;
                int col = 0;
                unsigned cand = 0u, claimed = 0u;
                if (c < c1)
                {
                    const int dz = (int)(((unsigned)c * rnx) >> 16), dx = c - dz * nxc;
                    col = (x0 + dx) + 16 * (z0 + dz);
                    cand = band & (k.canReplace ? 0xffffffffu : shAir[col]);
                    claimed = shClaim[col];      // may miss claims made meanwhile: only costs a rasteriser call, atomicMin decides
                }
                // the set bits go to the queue, at most two per lane and round (two independent chains in flight)
                while (__any_sync(0xffffffffu, cand != 0u))
                {
                    bool ok[2];
                    int code[2];
#pragma unroll
                    for (int u = 0; u < 2; ++u)
                    {
                        ok[u] = cand != 0u;
                        const int b = __ffs((int)cand) - 1;
                        cand &= cand - 1u;
                        code[u] = col << 5 | (b & 31);
                        // claimed by an earlier placement of the list: nothing this one does there can matter
                        if (ok[u] && ((claimed >> b) & 1u)) ok[u] = shBest[b * kSlabPitch + col] > (key | 0xffu);
                    }
                    const unsigned m0 = __ballot_sync(0xffffffffu, ok[0]), m1 = __ballot_sync(0xffffffffu, ok[1]);
                    const unsigned below = (1u << lane) - 1u;
                    if (ok[0]) wq[qn + __popc(m0 & below)] = (unsigned short)code[0];
                    if (ok[1]) wq[qn + __popc(m0) + __popc(m1 & below)] = (unsigned short)code[1];
                    qn += __popc(m0) + __popc(m1);
                    __syncwarp();
                    // (deeper queues - 160, 256 entries, so that several rasteriser calls run back to back - measured no faster)
                    if (qn > kFeatQueue - 64)
                        while (qn >= 32) drain();
                }
            }
        }
        while (qn > 0) drain();
        curE = -1;
    }
    if (lane == 0 && nPairs)
    {
        atomicAdd(&g_work[W_S6_PAIRS], (unsigned long long)nPairs);
        atomicAdd(&g_work[W_S6_RASTERISED], (unsigned long long)nRast);
    }
    __syncthreads();
    // thread t writes column t back if any of its 32 voxels was claimed
    if (t < 256 && shClaim[t] != 0u)
    {
        uint4 ab[2] = {reinterpret_cast<const uint4*>(colPtr)[0], reinterpret_cast<const uint4*>(colPtr)[1]};
        uint8_t* outv = reinterpret_cast<uint8_t*>(ab);
#pragma unroll
        for (int yy = 0; yy < kSlab; ++yy)
        {
            const unsigned b = shBest[yy * kSlabPitch + t];
            if (b != kNoBest) outv[yy] = (uint8_t)(b & 0xffu);
        }
        reinterpret_cast<uint4*>(colPtr)[0] = ab[0];
        reinterpret_cast<uint4*>(colPtr)[1] = ab[1];
    }
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

// minstd skip-ahead: a^k mod (2^31 - 1), so that a column can start from the state the reference's
// sequential walk would have when it reaches that column
__device__ __forceinline__ uint32_t minstd_pow(uint32_t k)
{
    uint64_t r = 1, b = 48271u;
    while (k)
    {
        if (k & 1u) r = (r * b) % 2147483647u;
        b = (b * b) % 2147483647u;
        k >>= 1;
    }
    return (uint32_t)r;
}

// placeDecorators (chunk.cu:1679-1747). The reference walks the 256 columns of a chunk with ONE
// sequential RNG stream (z outer, x inner), drawing 2 numbers per column plus 2 per cave layer. The
// number of draws of a column does not depend on any draw, so the stream position of every column
// is a prefix sum, the state there is seed * a^position, and the columns (which only ever touch
// their own blocks) run in parallel: one CTA per chunk, one thread per column.
__global__ void __launch_bounds__(256) k_decorators(const int* __restrict__ fillList, int n, const int2* __restrict__ origins,
                                                    const float* __restrict__ heightfield, const float* __restrict__ biomeWeights,
                                                    const CaveLayer* __restrict__ caveLayers, uint8_t* __restrict__ blocks)
{
    __shared__ int shDraws[256];
    const int li = blockIdx.x;
    if (li >= n) return;
    const int chunk = fillList ? fillList[li] : li;
    const int idx = threadIdx.x, x = idx & 15, z = idx >> 4;
    const int2 o = origins[chunk];
    uint8_t* b = blocks + (size_t)chunk * 98304;
    const float* w = biomeWeights + (size_t)chunk * (NUM_BIOMES * 256);
    const CaveLayer* cl = caveLayers + ((size_t)chunk * 256 + idx) * MAX_CAVE_LAYERS;
    int nLayers = 0;
    while (nLayers < MAX_CAVE_LAYERS && cl[nLayers].start != 384) ++nLayers;
    const int draws = 2 + 2 * nLayers;
    shDraws[idx] = draws;
    __syncthreads();
    for (int d = 1; d < 256; d <<= 1)
    {
        const int a = idx >= d ? shDraws[idx - d] : 0;
        __syncthreads();
        shDraws[idx] += a;
        __syncthreads();
    }
    const int before = shDraws[idx] - draws;
    Minstd rng = make_rng4(o.x, 0, o.y, 7589341);
    rng.x = (uint32_t)(((uint64_t)rng.x * minstd_pow((uint32_t)before)) % 2147483647u);

    const int biome = random_biome(w + idx, 256, rng.u01());
    float rand = rng.u01();
    const int g0 = c_decoratorGenRange[biome][0], gn = c_decoratorGenRange[biome][1];
    for (int g = 0; g < gn; ++g)
        if ((rand -= c_decoratorGens[g0 + g].chance) < 0.f)
        {
            try_place_decorator(b, x, (int)heightfield[(size_t)chunk * 256 + idx] + 1, z, c_decoratorGens[g0 + g]);
            break;
        }
    for (int l = 0; l < nLayers; ++l)
    {
        const CaveLayer c = cl[l];
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


// ---- halo exchange of placement lists between tiles (multi-GPU, SURVEY.md 8(e) option B)
// Wire format of one message (a rectangle of n chunks, raster order): int32 counts[n][2] = {surface, cave} list lengths, then
// per chunk its surface list (20 B records) followed by its cave list (24 B records), packed. byteOff[i] = where chunk i's
// records start (a multiple of 4). One CTA per chunk, copied as 32-bit words.
__global__ void __launch_bounds__(256) k_pack_placements(const int* __restrict__ chunkIdx, const long long* __restrict__ byteOff,
                                                         const FeaturePlacement* __restrict__ features,
                                                         const CaveFeaturePlacement* __restrict__ caveFeatures, const int* __restrict__ counts,
                                                         uint8_t* __restrict__ buf)
{
    const int i = blockIdx.x, chunk = chunkIdx[i];
    const int nF = counts[2 * chunk], nC = counts[2 * chunk + 1];
    if (threadIdx.x < 2) reinterpret_cast<int*>(buf)[2 * i + threadIdx.x] = threadIdx.x ? nC : nF;
    uint32_t* dst = reinterpret_cast<uint32_t*>(buf + byteOff[i]);
    const uint32_t* sf = reinterpret_cast<const uint32_t*>(features + (size_t)chunk * kMaxOwnFeatures);
    const uint32_t* sc = reinterpret_cast<const uint32_t*>(caveFeatures + (size_t)chunk * kMaxOwnCaveFeatures);
    const int wF = nF * (int)(sizeof(FeaturePlacement) / 4), wC = nC * (int)(sizeof(CaveFeaturePlacement) / 4);
    for (int t = threadIdx.x; t < wF; t += 256) dst[t] = sf[t];
    for (int t = threadIdx.x; t < wC; t += 256) dst[wF + t] = sc[t];
}

__global__ void __launch_bounds__(256) k_unpack_placements(const int* __restrict__ chunkIdx, const long long* __restrict__ byteOff,
                                                           const uint8_t* __restrict__ buf, FeaturePlacement* __restrict__ features,
                                                           CaveFeaturePlacement* __restrict__ caveFeatures, int* __restrict__ counts)
{
    const int i = blockIdx.x, chunk = chunkIdx[i];
    const int nF = min(reinterpret_cast<const int*>(buf)[2 * i], kMaxOwnFeatures), nC = min(reinterpret_cast<const int*>(buf)[2 * i + 1], kMaxOwnCaveFeatures);
    if (threadIdx.x < 2) counts[2 * chunk + threadIdx.x] = threadIdx.x ? nC : nF;
    const uint32_t* src = reinterpret_cast<const uint32_t*>(buf + byteOff[i]);
    uint32_t* df = reinterpret_cast<uint32_t*>(features + (size_t)chunk * kMaxOwnFeatures);
    uint32_t* dc = reinterpret_cast<uint32_t*>(caveFeatures + (size_t)chunk * kMaxOwnCaveFeatures);
    const int wF = nF * (int)(sizeof(FeaturePlacement) / 4), wC = nC * (int)(sizeof(CaveFeaturePlacement) / 4);
    for (int t = threadIdx.x; t < wF; t += 256) df[t] = src[t];
    for (int t = threadIdx.x; t < wC; t += 256) dc[t] = src[wF + t];
}

// Cost features of a chunk from its stage-1 output, for cutting balanced tiles before anything expensive has run:
// out[chunk] = {sum over columns of max(floor(h), 128)  (voxels the cave stage evaluates, chunk.cu:761-764),
//               sum of clamp(floor(h), 1, 383)          (voxels the fill stage looks up a cave biome for),
//               sum of (1 - ocean/beach weight)         (land columns: where surface features grow)}
__global__ void __launch_bounds__(256) k_chunk_cost(const float* __restrict__ heightfield, const float* __restrict__ biomeWeights, float* __restrict__ out)
{
    __shared__ float sh[3][8];
    const int chunk = blockIdx.x, idx = threadIdx.x;
    const int hi = (int)floorf(heightfield[(size_t)chunk * 256 + idx]);
    float v[3] = {(float)max(hi, SEA_LEVEL), (float)min(max(hi, 1), 383), 1.f};
    for (int b = 0; b < NUM_OCEAN_BEACH_BIOMES; ++b) v[2] -= biomeWeights[(size_t)chunk * (NUM_BIOMES * 256) + b * 256 + idx];
#pragma unroll
    for (int k = 0; k < 3; ++k)
    {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], d);
        if ((idx & 31) == 0) sh[k][idx >> 5] = v[k];
    }
    __syncthreads();
    if (idx < 3)
    {
        float s = 0.f;
        for (int wq = 0; wq < 8; ++wq) s += sh[idx][wq];
        out[3 * chunk + idx] = s;
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
