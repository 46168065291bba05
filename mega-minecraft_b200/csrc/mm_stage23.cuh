// Stage 2 (terrain layers) and stage 3 (zone erosion) kernels.
// Replace kernGenerateLayers + the CPU heightfield gather (/root/reference/src/terrain/chunk.cu:237-302,
// 308-415) and kernDoErosion / copyLayers / fixBackwardStratifiedLayers (chunk.cu:477-749).
//
// S2: one CTA per chunk; the 18x18 bordered heightfield is staged in shared memory straight from the
//     resident height planes of the 3x3 chunk neighbourhood (no host gather, no 18x18 copy in HBM).
//     FP32-bound (<= 55 simplex per column); algorithmic bytes 46 360 B/chunk.
// S3: Jacobi relaxation over 384x384 zone windows, 34x34 shared tiles, 144 CTAs per zone and many
//     zones per launch (blockIdx.z), sized so the planes of a batch stay in the 126 MB L2. Convergence
//     is detected on the device (one flag per sweep, shared by the batch) and polled once per group
//     of sweeps instead of after every sweep of every zone. Pure max/min/sub: bound by L2 bandwidth and launch latency;
//     algorithmic bytes per sweep 3 planes x 589 824 B.
#pragma once
#include "mm_common.cuh"
#include "mm_arith.cuh"
#include "mm_tables.cuh"

namespace mmg {

constexpr float kSqrt2 = 1.41421356237309504880168872420f;

// ---------------------------------------------------------------- S2
// WORLD=true : heightfield is the world's plane array [chunk][256]; the border comes from the
//              neighbouring chunks (all 8 must exist: the host only lists such chunks).
// WORLD=false: h18 holds one 18x18 tile per listed chunk (the reference's gathered layout).
template <bool WORLD>
__global__ void __launch_bounds__(256) k_layers(const int* __restrict__ chunkList, const int2* __restrict__ origins,
                                                const float* __restrict__ heightOrH18, const float* __restrict__ biomeWeights,
                                                float* __restrict__ layersOut, int nx)
{
    __shared__ float sh[18 * 18];
    noise_tab_stage();
    const int li = blockIdx.x;
    const int chunk = chunkList ? chunkList[li] : li;
    const int idx = threadIdx.x, x = idx & 15, z = idx >> 4;
    if (WORLD)
    {
        for (int t = idx; t < 18 * 18; t += 256)
        {
            const int tx = t % 18 - 1, tz = t / 18 - 1;            // -1..16
            const int dcx = (tx < 0) ? -1 : (tx > 15 ? 1 : 0), dcz = (tz < 0) ? -1 : (tz > 15 ? 1 : 0);
            const int nchunk = chunk + dcx + dcz * nx;
            sh[t] = heightOrH18[(size_t)nchunk * 256 + ((tx - 16 * dcx) + 16 * (tz - 16 * dcz))];
        }
    }
    else
    {
        for (int t = idx; t < 18 * 18; t += 256) sh[t] = heightOrH18[(size_t)li * 324 + t];
    }
    __syncthreads();

    const int2 o = origins[chunk];
    const float fx = (float)(o.x + x), fz = (float)(o.y + z);
    const float* cw = biomeWeights + (size_t)chunk * (NUM_BIOMES * 256) + idx;
    float w[NUM_BIOMES];
#pragma unroll
    for (int b = 0; b < NUM_BIOMES; ++b) w[b] = cw[b * 256];

    const float maxHeight = sh[(x + 1) + 18 * (z + 1)];
    float slope = 0.0f;
#pragma unroll
    for (int i = 0; i < 8; ++i)
    {
        const float nh = sh[(x + 1 + c_dirVecs2d[i][0]) + 18 * (z + 1 + c_dirVecs2d[i][1])];
        const float d = fabsf(nh - maxHeight);
        slope = fmaxf(slope, (i & 1) ? d * kSqrt2 : d);
    }
    auto matWeight = [&](int m) -> float {
        float acc = 0.0f;
#pragma unroll
        for (int b = 0; b < NUM_BIOMES; ++b) acc = fmaf(w[b], c_biomeMaterialWeights[b][m], acc);
        return acc;
    };
    int nFbm = 0;      // work counter: fbm<5> evaluations of this column
    auto thickness = [&](int l, float tw) -> float {
        if (!(tw > 0.0f)) return 0.0f;
        ++nFbm;
        const float off = (float)l * 5283.64f;
        const MaterialInfo mi = c_materialInfos[l];
        const float f = fbm2<5>(fmaf(mi.v2, fx, off), fmaf(mi.v2, fz, off));
        return fmaxf(fmaf(f, mi.v1, mi.thickness), 0.0f) * tw;
    };
    float* out = layersOut + (size_t)chunk * (NUM_MATERIALS * 256) + idx;
    float height = 0.0f;
    bool stop = false;
    for (int l = 0; l < NUM_FORWARD; ++l)
    {
        out[l * 256] = height;     // after the reference's break point this repeats the last height
        if (stop || l == NUM_FORWARD - 1) continue;
        if (height > maxHeight) { stop = true; continue; }
        height = thickness(l, matWeight(l)) + height;
    }
    height = 0.0f;
    for (int l = NUM_STRATIFIED - 1; l >= NUM_FORWARD; --l)
    {
        height = thickness(l, matWeight(l)) + height;
        out[l * 256] = height;
    }
    height = maxHeight;
#pragma unroll
    for (int l = NUM_MATERIALS - 1; l >= NUM_STRATIFIED; --l)
    {
        const MaterialInfo mi = c_materialInfos[l];
        const float lh = fmaxf(mi.thickness * ((mi.v2 - slope) / mi.v2), 0.0f);
        height = fmaf(-matWeight(l), lh, height);
        out[l * 256] = height;
    }
    {
        int v = nFbm;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
        if ((idx & 31) == 0) atomicAdd(&g_work[W_S2_FBM5], (unsigned long long)v);
        if (idx == 32) atomicAdd(&g_work[W_S2_COLUMNS], 256ull);
    }
}

// ---------------------------------------------------------------- S3
constexpr int kErosionSide = 384, kErosionCols = kErosionSide * kErosionSide;

constexpr int kZonePlanes = 12;   // per zone: [0..8] gathered planes, [9] ping-pong scratch, [10..11] accumulated heights

// gather: zone planes[9][384][384] from chunk-major layers/height of the 24x24-chunk window whose
// lower corner is chunk zoneCorner[blockIdx.z] in window raster coordinates (copyLayers(..., true), chunk.cu:603-656)
__global__ void k_zone_gather(const float* __restrict__ layers, const float* __restrict__ height, float* __restrict__ zones,
                              const int2* __restrict__ zoneCorner, int nx)
{
    const int2 zc = zoneCorner[blockIdx.z];
    float* planes = zones + (size_t)blockIdx.z * kZonePlanes * kErosionCols;
    const int gx = blockIdx.x * 32 + threadIdx.x, gz = blockIdx.y * 32 + threadIdx.y;
    const int chunk = (zc.x + (gx >> 4)) + (zc.y + (gz >> 4)) * nx;
    const int idx = (gx & 15) + 16 * (gz & 15);
    const int col = gx + kErosionSide * gz;
#pragma unroll
    for (int l = 0; l < NUM_ERODED; ++l)
        planes[(size_t)l * kErosionCols + col] = layers[(size_t)chunk * (NUM_MATERIALS * 256) + (NUM_STRATIFIED + l) * 256 + idx];
    planes[(size_t)NUM_ERODED * kErosionCols + col] = height[(size_t)chunk * 256 + idx];
    planes[(size_t)10 * kErosionCols + col] = 0.0f;    // accumulated heights start at 0 (chunk.cu:680)
}

// one Jacobi sweep of one loose layer over every zone of the batch (blockIdx.z): reads the start plane
// pIn (+ accumulated heights on the layer's first sweep), writes pOut. Plane roles are indices into the
// zone's 12 planes; they swap in lockstep for all zones. A zone that has converged is a fixed point of
// the sweep, so sweeping it again (while other zones of the batch still move) changes nothing.
//
// tileChanged[3][kMaxZoneBatch][144]: did sweep n change a cell of 32x32 tile t of zone z? Sweep n reads row (n-1)%3, sets
// row n%3 and clears row (n+1)%3. A tile's result depends on the 34x34 cells around it, i.e. on its 3x3 tile neighbourhood.
// If the previous sweep changed none of those nine tiles (because it found nothing to change there, or because it was
// itself skipped), their input and output planes are equal (outS = sIn, accumOut = accumIn for every cell), this sweep
// would recompute what the previous one found and write back what pOut already holds: the CTA returns at once. Erosion
// activity is local (steep slopes), so most tiles of a zone go quiet long before its last one does. `force` disables the
// shortcut for the first two sweeps of a layer (the first sweep folds the carried heights into its comparison, so
// "nothing flagged" does not imply "planes equal" there).
constexpr int kMaxZoneBatch = 32;
constexpr int kZoneTiles = 144;      // 12 x 12 tiles of 32 x 32 cells
// (A persistent variant - two 1024-thread CTAs per SM walking the tiles - was measured and is slower: 54.7 vs 39.4 ms per
// 256x256 world; with one tile per CTA the block scheduler overlaps the staging of one tile with the arithmetic of others.)
// One CTA of 32 x 4 threads per tile, eight rows per thread: a quiet tile costs the launch of 4 warps, not 32, and a tile CTA
// (4 096 registers) fits next to the five resident CTAs of k_caves, with which erosion shares the SMs in a full generate
// (profiles/r02_erode_rows.txt: 256x256 world, overlapped: 2 / 4 / 8 / 16 rows of threads 457.9 / 457.0 / 460.5 / 467.4 ms).
// (2 / 3 / 4 / 6 tiles per CTA, so that the mostly quiet late sweeps launch fewer blocks, measured in round 2: 7.3 / 7.4 / 8.1 /
// 8.4 ms against 6.1 per 128x128 region - the live tiles of a CTA run one after the other.)
#ifndef MMG_ERODE_ROWS
#define MMG_ERODE_ROWS 4
#endif
constexpr int kErodeRows = MMG_ERODE_ROWS;
__global__ void __launch_bounds__(32 * kErodeRows) k_erode_sweep(float* __restrict__ zones, int pIn, int pOut, int pUp, int pAccIn, int pAccOut,
                                                      float rep, int isFirst, int* __restrict__ changedFlag, int* __restrict__ tileChanged,
                                                      int sweepNo, int force)
{
    __shared__ float shS[34 * 34];
    __shared__ float shE[34 * 34];
    const int rowSize = kMaxZoneBatch * kZoneTiles;
    const int* rowPrev = tileChanged + ((sweepNo + 2) % 3) * rowSize + blockIdx.z * kZoneTiles;
    int* rowCur = tileChanged + (sweepNo % 3) * rowSize + blockIdx.z * kZoneTiles;
    int* rowNext = tileChanged + ((sweepNo + 1) % 3) * rowSize + blockIdx.z * kZoneTiles;
    const int tile = blockIdx.x + 12 * blockIdx.y;
    const int lx = threadIdx.x, lid = lx + 32 * threadIdx.y;
    {
        if (lid == 9) rowNext[tile] = 0;
        int moved = force;
        if (!force && lid < 9)
        {
            const int tx = (int)blockIdx.x + lid % 3 - 1, tz = (int)blockIdx.y + lid / 3 - 1;
            if (tx >= 0 && tx < 12 && tz >= 0 && tz < 12) moved = rowPrev[tx + 12 * tz];
        }
        const int live = __syncthreads_or(moved);
        if (lid == 10) atomicAdd(&g_work[live ? W_S3_TILES_SWEPT : W_S3_TILES_QUIET], 1ull);
        if (!live) return;
    }
    float* zone = zones + (size_t)blockIdx.z * kZonePlanes * kErosionCols;
    const float* sIn = zone + (size_t)pIn * kErosionCols;
    float* sOut = zone + (size_t)pOut * kErosionCols;
    const float* eUp = zone + (size_t)pUp * kErosionCols;
    const float* accumIn = zone + (size_t)pAccIn * kErosionCols;
    float* accumOut = zone + (size_t)pAccOut * kErosionCols;
    const int bx0 = blockIdx.x * 32, bz0 = blockIdx.y * 32;
    for (int t = lid; t < 34 * 34; t += 32 * kErodeRows)
    {
        int gx = bx0 - 1 + (t % 34), gz = bz0 - 1 + (t / 34);
        gx = min(max(gx, 0), kErosionSide - 1);       // clamp-to-edge halo, chunk.cu:545
        gz = min(max(gz, 0), kErosionSide - 1);
        const int j = gx + kErosionSide * gz;
        const float a = isFirst ? accumIn[j] : 0.0f;
        shS[t] = sIn[j] + a;
        shE[t] = eUp[j] + a;
    }
    __syncthreads();
    const float repDiag = rep * kSqrt2;
    bool changed = false;
#pragma unroll
    for (int r = 0; r < 32 / kErodeRows; ++r)
    {
        const int lz = threadIdx.y + kErodeRows * r;
        const int gx = bx0 + lx, gz = bz0 + lz, i = gx + kErosionSide * gz;
        const int c = (lx + 1) + 34 * (lz + 1);
        const float s0 = shS[c], e0 = shE[c];
        float ns = s0, maxT = e0 - s0;
#pragma unroll
        for (int d = 0; d < 8; ++d)
        {
            const int j = c + c_dirVecs2d[d][0] + 34 * c_dirVecs2d[d][1];
            const float sj = shS[j];
            ns = fmaxf(ns, sj - ((d & 1) ? repDiag : rep));
            maxT = fmaxf(maxT, shE[j] - sj);
        }
        ns = fminf(ns, e0);
        float outS = sIn[i];
        float acc = accumIn[i];
        if (maxT > 0.0f)
        {
            outS = ns;
            if (ns != s0)
            {
                acc = (ns - s0) + acc;
                changed = true;
            }
        }
        sOut[i] = outS;
        accumOut[i] = acc;
    }
    if (changed)
    {
        *changedFlag = 1;
        rowCur[tile] = 1;
    }
}

// scatter the centre 12x12 chunks back (copyLayers(..., false)) into the eroded layer set and apply
// fixBackwardStratifiedLayers (chunk.cu:725-749); the stratified layers are copied through.
__global__ void k_zone_scatter(const float* __restrict__ zones, const float* __restrict__ layersIn, float* __restrict__ layersOut,
                               const int2* __restrict__ zoneCorner, int nx)
{
    const int2 zc = zoneCorner[blockIdx.z];
    const float* planes = zones + (size_t)blockIdx.z * kZonePlanes * kErosionCols;
    const int gx = 96 + blockIdx.x * 32 + threadIdx.x, gz = 96 + blockIdx.y * 32 + threadIdx.y;   // centre 192x192
    const int chunk = (zc.x + (gx >> 4)) + (zc.y + (gz >> 4)) * nx;
    const int idx = (gx & 15) + 16 * (gz & 15);
    const int col = gx + kErosionSide * gz;
    const float* in = layersIn + (size_t)chunk * (NUM_MATERIALS * 256) + idx;
    float* out = layersOut + (size_t)chunk * (NUM_MATERIALS * 256) + idx;
#pragma unroll
    for (int l = 0; l < NUM_FORWARD; ++l) out[l * 256] = in[l * 256];
    const float erodedStart = planes[col];   // loose layer 0 = material 12
#pragma unroll
    for (int l = NUM_FORWARD; l < NUM_STRATIFIED; ++l) out[l * 256] = erodedStart - in[l * 256];
#pragma unroll
    for (int l = 0; l < NUM_ERODED; ++l) out[(NUM_STRATIFIED + l) * 256] = planes[(size_t)l * kErosionCols + col];
}

}  // namespace mmg
