// Stage 4: cave-biome noise, the per-voxel cave predicate with specialCaveNoise, and the per-column
// compaction into CaveLayers. Replaces kernGenerateCaves / shouldGenerateCaveAtBlock / getCaveBiome
// (/root/reference/src/terrain/chunk.cu:755-937, biomeFuncs.hpp:130-220, rng.hpp:282-320).
//
// Mapping: k_cave_columns computes the y-independent terms once per column (ocean+beach weight and
// the whole ravine test: 24 simplex + a 9-cell Worley that the reference re-evaluates for every one
// of 384 voxels); k_caves then runs one 288-thread CTA per pair of columns: the voxels that are not decided by y alone
// (at most 141 per column) are listed and evaluated one per thread on dense warps, the solid/air columns are kept as
// 12 words of bits each, flips come from bit scans and the <= 64 cave-biome lookups per column go to a queue.
// FP32-pipe bound: ~23 simplex3 + 27 hashed cells per evaluated voxel; algorithmic bytes 107 528 B/chunk.
#pragma once
#include "mm_surface.cuh"

namespace mmg {

#ifdef MMG_FEATURE_STATS
// developer build: voxels whose "huge caves" term was proved 0 but is not (must stay 0); voxels evaluated without / with proof
__device__ unsigned long long g_hugeMismatch, g_hugeVoxels[2], g_caveWarped, g_cavePending[3], g_caveTable[2];      // g_caveTable: survivors whose 27 Worley cells were not / were all in the jitter table;      // g_cavePending: decided by bounds / pending / decided wrongly (must stay 0);      // g_caveWarped: voxels that evaluated the warped specialCaveNoise
#endif

// ---------------------------------------------------------------- cave biome (biomeFuncs.hpp:135-220)
struct CaveBiomeNoise { float v[4]; };   // none, shallow, warped, rocky

__device__ MMG_NOISE_INLINE CaveBiomeNoise cave_biome_noise(int x, int y, int z, float maxHeight)
{
    const float px = (float)x, py = (float)y, pz = (float)z;
    const float qx = px * 0.0470f, qy = py * 0.0470f, qz = pz * 0.0470f;
    float o1, o2, o3;
    fbm3_from3<3>(qx, qy, qz, &o1, &o2, &o3);      // noise evaluations in pairs (packed fp32, mm_arith.cuh): same values
    const float cx = fmaf(o1, 30.f, px);
    const float cy = fmaf(o2, 24.f, py);
    const float cz = fmaf(o3, 30.f, pz);
    const float nx = cx * 0.2000f, nz = cz * 0.2000f;
    const float top = fmaf(maxHeight + -128.f, 0.15f, 128.f);
    const f32x2 ns = fbm2x2<3>(f2_make(nx, nx + 3821.34f), f2_make(nz, nz + 4920.32f));
    const f32x2 sd = fbm2x2<3>(f2_make(nx + -4921.34f, nx + 9411.32f), f2_make(nz + 8402.13f, nz + -3921.34f));
    const float nsStart = fmaf(f2_lo(ns), 23.f, top + -19.f);
    const float nsEnd = fmaf(f2_hi(ns), 3.f, nsStart + -5.f);
    const float sdStart = fmaf(f2_lo(sd), 18.f, top + -72.f);
    const float sdEnd = fmaf(f2_hi(sd), 7.f, sdStart + -10.f);
    CaveBiomeNoise n;
    n.v[0] = ss_t((cy - nsEnd) / (nsStart - nsEnd));
    n.v[1] = ss_t((cy - sdEnd) / (sdStart - sdEnd));
    const f32x2 wr = simplex3x2_raw<true>(f2_make(fmaf(cx, 0.0030f, 5821.32f), fmaf(cx, 0.0022f, -9193.23f)),
                                          f2_make(fmaf(cy, 0.0030f, 4920.12f), fmaf(cy, 0.0022f, -6813.39f)),
                                          f2_make(fmaf(cz, 0.0030f, 7931.59f), fmaf(cz, 0.0022f, (float)-2171.23)));
    n.v[2] = ss_t(fmaf(f2_lo(wr), 42.f, 0.05f) / (0.05f - -0.05f));
    n.v[3] = ss_t(fmaf(f2_hi(wr), 42.f, 0.05f) / (0.05f - -0.05f));
    return n;
}

__device__ __forceinline__ int cave_biome(int x, int y, int z, float maxHeight, int seed)
{
    const CaveBiomeNoise n = cave_biome_noise(x, y, z, maxHeight);
    Minstd rng = make_rng4(x, y, z, seed);
    float rand = rng.u01();
    for (int b = 0; b < NUM_CAVE_BIOMES; ++b)
    {
        float w = 1.0f;
        for (int c = 0; c < 4; ++c)
        {
            const uint8_t t = c_caveBiomeNoiseWeights[b][c];
            if (t == 1) w *= n.v[c];
            else if (t == 2) w *= 1.0f - n.v[c];
        }
        rand -= w;
        if (rand <= 0.f) return b;
    }
    return CB_NONE;
}

// "Is the cave biome at (x, y, z) CRYSTAL_CAVES?" - the same arithmetic as cave_biome() with the terms that cannot
// matter left out. CRYSTAL_CAVES is biome 1 with weight ((1 - none) * shallow) * rocky, drawn after NONE (weight none):
// it is chosen iff rand - none > 0 and (rand - none) - weight <= 0. A factor that is exactly 0 makes the weight 0 and the
// second test false whatever the other noises are, and the `warped` noise does not enter at all.
__device__ MMG_NOISE_INLINE bool cave_biome_is_crystal(int x, int y, int z, float maxHeight, int seed)
{
    const float px = (float)x, py = (float)y, pz = (float)z;
    const float qx = px * 0.0470f, qy = py * 0.0470f, qz = pz * 0.0470f;
    float o1, o2, o3;
    fbm3_from3<3>(qx, qy, qz, &o1, &o2, &o3);
    const float cx = fmaf(o1, 30.f, px);
    const float cy = fmaf(o2, 24.f, py);
    const float cz = fmaf(o3, 30.f, pz);
    const float rocky = ss_t(fmaf(simplex3_raw<true>(fmaf(cx, 0.0022f, -9193.23f), fmaf(cy, 0.0022f, -6813.39f), fmaf(cz, 0.0022f, (float)-2171.23)), 42.f, 0.05f) / (0.05f - -0.05f));
    if (rocky == 0.f) return false;
    const float nx = cx * 0.2000f, nz = cz * 0.2000f;
    const float top = fmaf(maxHeight + -128.f, 0.15f, 128.f);
    const f32x2 sd = fbm2x2<3>(f2_make(nx + -4921.34f, nx + 9411.32f), f2_make(nz + 8402.13f, nz + -3921.34f));
    const float sdStart = fmaf(f2_lo(sd), 18.f, top + -72.f);
    const float sdEnd = fmaf(f2_hi(sd), 7.f, sdStart + -10.f);
    const float shallow = ss_t((cy - sdEnd) / (sdStart - sdEnd));
    if (shallow == 0.f) return false;
    const f32x2 ns = fbm2x2<3>(f2_make(nx, nx + 3821.34f), f2_make(nz, nz + 4920.32f));
    const float nsStart = fmaf(f2_lo(ns), 23.f, top + -19.f);
    const float nsEnd = fmaf(f2_hi(ns), 3.f, nsStart + -5.f);
    const float none = ss_t((cy - nsEnd) / (nsStart - nsEnd));
    Minstd rng = make_rng4(x, y, z, seed);
    float rand = rng.u01();
    rand -= none;                                   // biome 0: NONE
    if (rand <= 0.f) return false;
    float w = 1.0f;                                 // biome 1: c_caveBiomeNoiseWeights[1] = {2, 1, 0, 1}
    w *= 1.0f - none;
    w *= shallow;
    w *= rocky;
    rand -= w;
    return rand <= 0.f;
}

// cave_biome_is_crystal in two steps, for callers that regroup their voxels in between (k_fill_rock): the first step stops where
// `rocky` is exactly 0 (nearly half of the voxels), the second carries on from the warped position with the 2-D noises and the draw.
// The same operations in the same order as above: rocky(...) && decide(...) == cave_biome_is_crystal(...).
__device__ MMG_NOISE_INLINE bool cave_crystal_rocky(int x, int y, int z, float4* warped)      // *warped = (cx, cy, cz, rocky)
{
    const float px = (float)x, py = (float)y, pz = (float)z;
    const float qx = px * 0.0470f, qy = py * 0.0470f, qz = pz * 0.0470f;
    float o1, o2, o3;
    fbm3_from3<3>(qx, qy, qz, &o1, &o2, &o3);
    const float cx = fmaf(o1, 30.f, px);
    const float cy = fmaf(o2, 24.f, py);
    const float cz = fmaf(o3, 30.f, pz);
    const float rocky = ss_t(fmaf(simplex3_raw<true>(fmaf(cx, 0.0022f, -9193.23f), fmaf(cy, 0.0022f, -6813.39f), fmaf(cz, 0.0022f, (float)-2171.23)), 42.f, 0.05f) / (0.05f - -0.05f));
    *warped = make_float4(cx, cy, cz, rocky);
    return !(rocky == 0.f);
}
__device__ MMG_NOISE_INLINE bool cave_crystal_decide(int x, int y, int z, float maxHeight, int seed, float4 warped)
{
    const float cx = warped.x, cy = warped.y, cz = warped.z, rocky = warped.w;
    const float nx = cx * 0.2000f, nz = cz * 0.2000f;
    const float top = fmaf(maxHeight + -128.f, 0.15f, 128.f);
    const f32x2 sd = fbm2x2<3>(f2_make(nx + -4921.34f, nx + 9411.32f), f2_make(nz + 8402.13f, nz + -3921.34f));
    const float sdStart = fmaf(f2_lo(sd), 18.f, top + -72.f);
    const float sdEnd = fmaf(f2_hi(sd), 7.f, sdStart + -10.f);
    const float shallow = ss_t((cy - sdEnd) / (sdStart - sdEnd));
    if (shallow == 0.f) return false;
    const f32x2 ns = fbm2x2<3>(f2_make(nx, nx + 3821.34f), f2_make(nz, nz + 4920.32f));
    const float nsStart = fmaf(f2_lo(ns), 23.f, top + -19.f);
    const float nsEnd = fmaf(f2_hi(ns), 3.f, nsStart + -5.f);
    const float none = ss_t((cy - nsEnd) / (nsStart - nsEnd));
    Minstd rng = make_rng4(x, y, z, seed);
    float rand = rng.u01();
    rand -= none;                                   // biome 0: NONE
    if (rand <= 0.f) return false;
    float w = 1.0f;                                 // biome 1: c_caveBiomeNoiseWeights[1] = {2, 1, 0, 1}
    w *= 1.0f - none;
    w *= shallow;
    w *= rocky;
    rand -= w;
    return rand <= 0.f;
}

// ---------------------------------------------------------------- specialCaveNoise (rng.hpp:282-320)
// hash (rng.hpp:148-155) at this call site: fma(z, Kz, fma(x, Kx, y*Ky))
__device__ MMG_NOISE_INLINE float special_cave_noise(float px, float py, float pz)
{
    const float fx = floorf(px), fy = floorf(py), fz = floorf(pz);
    const int ix = (int)fx, iy = (int)fy, iz = (int)fz;
    const float nfx = fx - px, nfy = fy - py, nfz = fz - pz;
    float d1 = FLT_MAX, d2 = FLT_MAX, d3 = FLT_MAX;
#pragma unroll 1
    for (int x = -1; x <= 1; ++x)
#pragma unroll 1
        for (int y = -1; y <= 1; ++y)
#pragma unroll 1
            for (int z = -1; z <= 1; ++z)
            {
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

__device__ __forceinline__ Ravine ravine_column(int wx, int wz, float obw)
{
    Ravine r = {false, 0.f, 0.f};
    const float rx = (float)wx * 0.0015f, rz = (float)wz * 0.0015f;
    const f32x2 oxz = fbm2x2<4>(f2_make(rx * 10.f, rx * 10.f + 5923.45f), f2_make(rz * 10.f, rz * 10.f + 4129.42f));      // noise in pairs (mm_arith.cuh)
    const float ox = f2_lo(oxz), oz = f2_hi(oxz);
    const Worley2 w = worley2(fmaf(ox, 0.03f, rx), fmaf(oz, 0.03f, rz));
    const float thr = (1.f - obw) * 0.12f;
    if (!(w.d1 < thr)) return r;
    const float colorX = hash_fract(fmaf(w.cpx, 238.68f, w.cpy * 491.28f));
    r.top = fmaf(colorX, 24.f, 120.f);
    const float ratio = 1.f - (w.d1 / thr);
    const f32x2 dw = fbm2x2<4>(f2_make(fmaf(rx, 8.f, 8391.32f), fmaf(rx, 3.f, 5129.32f)), f2_make(fmaf(rz, 8.f, 4821.39f), fmaf(rz, 3.f, 1392.49f)));
    float depth = ss_t(ratio / 0.3f) * fmaf(f2_lo(dw), 26.f, 60.f);
    const float waveOff = f2_hi(dw) * 4.f;
    const float wave = sinf(fmaf(rx + rz, 15.f, waveOff));
    depth = depth * ss_t((wave + -0.4f) / (0.6f - 0.4f));
    r.depth = depth;
    r.active = depth > 0.0001f;
    return r;
}

// jitter of one Worley cell (rand3From3 of the integer cell corner, rng.hpp:148-155, 296-300)
__device__ __forceinline__ void cave_cell_jitter(int icx, int icy, int icz, float* jx, float* jy, float* jz)
{
    const float cx = (float)icx, cy = (float)icy, cz = (float)icz;
    *jx = hash_fract(fmaf(cz, 402.98f, fmaf(cx, 238.68f, cy * 491.28f)));
    *jy = hash_fract(fmaf(cz, 747.42f, fmaf(cx, 654.37f, cy * 560.45f)));
    *jz = hash_fract(fmaf(cz, 674.81f, fmaf(cx, 640.88f, cy * 151.81f)));
}

// specialCaveNoise with the cell jitters taken from a table covering cells [box, box + kCaveBox)^3 that the
// CTA filled once (a jitter depends only on the integer cell, and the voxels of one column share a few
// dozen cells; the reference recomputes 27 x 3 sin() hashes per voxel). Cells outside the table are
// computed in place. Same cells, same order, same comparisons as special_cave_noise.
constexpr int kCaveBox = 6;
// The three smallest distances are selected on the SQUARED distances and only d1 and d3 are square-rooted: sqrtf is
// monotone (correctly rounded), so the k-th smallest of the rounded roots is the root of the k-th smallest square - the
// same two numbers the reference's insertion on sqrt'ed distances ends with (equal roots are interchangeable), for 2
// instead of 27 IEEE square roots per voxel.
__device__ __forceinline__ float special_cave_noise_cached(float px, float py, float pz, int bx, int by, int bz, int ex, int ey, int ez, const float* shJit)
{
    constexpr int N3 = kCaveBox * kCaveBox * kCaveBox;
    const float fx = floorf(px), fy = floorf(py), fz = floorf(pz);
    const int ix = (int)fx, iy = (int)fy, iz = (int)fz;
    const float nfx = fx - px, nfy = fy - py, nfz = fz - pz;
    float q1 = FLT_MAX, q2 = FLT_MAX, q3 = FLT_MAX;
    // keeps the three smallest squares (a 5-instruction min / max network; the values, not their order of arrival, matter)
    auto insert = [&](float q) {
        const float a = fminf(q1, q); q = fmaxf(q1, q); q1 = a;
        const float b = fminf(q2, q); q = fmaxf(q2, q); q2 = b;
        q3 = fminf(q3, q);
    };
    const int ux0 = ix - 1 - bx, uy0 = iy - 1 - by, uz0 = iz - 1 - bz;
    // the table holds cells [box, box + (ex, ey, ez)) (k_caves: a fixed kCaveBox^3 block placed around the expected sample positions)
    if ((unsigned)ux0 <= (unsigned)(ex - 3) && (unsigned)uy0 <= (unsigned)(ey - 3) && (unsigned)uz0 <= (unsigned)(ez - 3))
    {
        // the whole 3x3x3 neighbourhood is in the table (nearly always): no per-cell bounds tests
        const float* J = shJit + (ux0 * kCaveBox + uy0) * kCaveBox + uz0;
        // cells two at a time with packed fp32 (mm_arith.cuh): per x the rows y = 0 / y = 1 pair up for each z, the row y = 2 pairs
        // z = 0 / z = 1 and leaves z = 2 - same sums, same squares; the order of arrival does not matter to insert()
        const f32x2 nfx2 = f2_dup(nfx), nfy2 = f2_dup(nfy), nfz2 = f2_dup(nfz);
        float ox = -1.f;
#pragma unroll 1
        for (int x = 0; x < 3; ++x, ox += 1.f)
        {
            const float* Jc = J + x * kCaveBox * kCaveBox;
            const f32x2 ox2 = f2_dup(ox), oy01 = f2_make(-1.f, 0.f);
#pragma unroll
            for (int z = 0; z < 3; ++z)
            {
                const f32x2 dx = f2_add(nfx2, f2_add(f2_make(Jc[z], Jc[kCaveBox + z]), ox2));
                const f32x2 dy = f2_add(nfy2, f2_add(f2_make(Jc[z + N3], Jc[kCaveBox + z + N3]), oy01));
                const f32x2 dz = f2_add(nfz2, f2_add(f2_make(Jc[z + 2 * N3], Jc[kCaveBox + z + 2 * N3]), f2_dup((float)(z - 1))));
                const f32x2 q = f2_fma(dz, dz, f2_fma(dx, dx, f2_mul(dy, dy)));
                insert(f2_lo(q));
                insert(f2_hi(q));
            }
            const float* Jr = Jc + 2 * kCaveBox;
            {
                const f32x2 dx = f2_add(nfx2, f2_add(f2_make(Jr[0], Jr[1]), ox2));
                const f32x2 dy = f2_add(nfy2, f2_add(f2_make(Jr[N3], Jr[1 + N3]), f2_dup(1.f)));
                const f32x2 dz = f2_add(nfz2, f2_add(f2_make(Jr[2 * N3], Jr[1 + 2 * N3]), f2_make(-1.f, 0.f)));
                const f32x2 q = f2_fma(dz, dz, f2_fma(dx, dx, f2_mul(dy, dy)));
                insert(f2_lo(q));
                insert(f2_hi(q));
            }
            {
                const float dx = nfx + (Jr[2] + ox), dy = nfy + (Jr[2 + N3] + 1.f), dz = nfz + (Jr[2 + 2 * N3] + 1.f);
                insert(fmaf(dz, dz, fmaf(dx, dx, dy * dy)));      // dist = sqrtf(this) in the reference
            }
        }
    }
    else
    {
#pragma unroll 1
        for (int x = -1; x <= 1; ++x)
#pragma unroll 1
            for (int y = -1; y <= 1; ++y)
#pragma unroll 1
                for (int z = -1; z <= 1; ++z)
                {
                    const int ux = ix + x - bx, uy = iy + y - by, uz = iz + z - bz;
                    float jx, jy, jz;
                    if ((unsigned)ux < (unsigned)ex && (unsigned)uy < (unsigned)ey && (unsigned)uz < (unsigned)ez)
                    {
                        const int c = (ux * kCaveBox + uy) * kCaveBox + uz;
                        jx = shJit[c]; jy = shJit[c + N3]; jz = shJit[c + 2 * N3];
                    }
                    else
                        cave_cell_jitter(ix + x, iy + y, iz + z, &jx, &jy, &jz);
                    const float dx = nfx + (jx + (float)x), dy = nfy + (jy + (float)y), dz = nfz + (jz + (float)z);
                    insert(fmaf(dz, dz, fmaf(dx, dx, dy * dy)));
                }
    }
    return sqrtf(q3) / sqrtf(q1) + -1.0f;
}

// chunk.cu:755-810, first half: everything up to the cave-noise threshold
//     thr = (fma(bottomRatio, 0.7, 0.3) * topRatio) * (fma(fbmA, 0.12, 0.24) * fma(huge, 1.4, 1)),   air <=> thr > 0.04 && noise < thr,
// split so that fbmA = fbm3<4>(pos * 0.02) - four simplex3 of the 23 per voxel - is only evaluated where it can matter.
// Every operation of thr is monotone in fbmA (rounding is monotone, the other factors are >= 0), and |fbmA| < 1
// (|42 * simplex3_raw| <= 1.052, see mm_fillfuncs.cuh; amplitudes sum to 0.9375), so cave_thr(-1) <= thr <= cave_thr(+1)
// exactly as computed. A voxel whose upper bound fails `> 0.04` is solid without any noise; one whose warped cave noise is
// >= the upper bound is solid, one whose noise is < the lower bound (and that bound > 0.04) is air; only the rest
// (k_caves queues them and evaluates fbmA on dense warps) need the exact threshold.
struct CaveThr { float ratio, hugeFactor; };      // ratio = fma(bottomRatio, 0.7, 0.3) * topRatio
__device__ __forceinline__ float cave_thr(const CaveThr& c, float fbmA)
{
    float thr = fmaf(fbmA, 0.12f, 0.24f);
    thr = thr * c.hugeFactor;                     // huge == 0: fma(0, 1.4, 1) = 1 and thr * 1 = thr, the same bits
    return c.ratio * thr;
}
__device__ __forceinline__ float cave_fbm_a(int wx, int y, int wz)
{
    const float npx = (float)wx * 0.0050f, npy = (float)y * 0.0050f, npz = (float)wz * 0.0050f;
    return fbm3_paired<4>(npx * 4.f, npy * 4.f, npz * 4.f);
}
__device__ __forceinline__ float cave_huge_factor(int wx, int y, int wz)
{
    const float npx = (float)wx * 0.0050f, npy = (float)y * 0.0050f, npz = (float)wz * 0.0050f;
    const float huge = ss_t((fbm3_paired<4>(npx * 0.0700f, npy * 0.0700f, npz * 0.0700f) + -0.2f) / (0.4f - 0.2f));
    return fmaf(huge, 1.4f, 1.f);
}
// The predicate in two steps. cave_threshold_cheap (no noise): 0 = solid, 1 = air, 2 = "survivor": the noise has to be asked.
// hugeZero: the caller has proved that the "huge caves" term is exactly 0 at this voxel (huge_zero_mask below); the upper-bound
// test then needs no noise either.
__device__ __forceinline__ int cave_threshold_cheap(int y, float maxHeight, float obw, bool hugeZero, CaveThr* c)
{
    if (y == 0) return 0;
    const int hi = (int)maxHeight;
    if (y > (hi > SEA_LEVEL ? hi : SEA_LEVEL)) return 1;
    const float fy = (float)y;
    const float topRatio = ss_t((fmaf(obw, 50.f, fy) + -142.f) / (95.f - 142.f));
    const float bottomRatio = ss_t((fy + -5.f) / (20.f - 5.f));
    // topRatio is exactly 0 from y = 142 - 50 obw upwards; the threshold is (finite * 0) * finite = 0 there and
    // the test `thr > 0.04` fails whatever the noises say
    if (topRatio == 0.f) return 0;
    c->ratio = fmaf(bottomRatio, 0.7f, 0.3f) * topRatio;
    c->hugeFactor = 1.f;
    // the reference always evaluates the warped specialCaveNoise (15 simplex + 27 hashed cells) and then
    // tests `thr > 0.04 && caveNoise < thr` (chunk.cu:776-783); where even the largest possible threshold fails the first
    // test (topRatio -> 0 towards y = 142 - 50 obw) no noise can matter
    if (hugeZero && !(cave_thr(*c, 1.f) > 0.04f)) return 0;
    return 2;
}
// cave_threshold_noise, for survivors: 0 = solid after all, 2 = the warped specialCaveNoise at (*px, *py, *pz) has to be
// compared with the threshold *c.
__device__ __forceinline__ int cave_threshold_noise(int wx, int y, int wz, bool hugeZero, CaveThr* c, float* px, float* py, float* pz)
{
    const float npx = (float)wx * 0.0050f, npy = (float)y * 0.0050f, npz = (float)wz * 0.0050f;
#ifdef MMG_FEATURE_STATS
    const bool hugeCheck = hugeZero;      // developer build: evaluate the term anyway and count voxels where the proof was wrong
    hugeZero = false;
    atomicAdd(&g_hugeVoxels[hugeCheck ? 1 : 0], 1ull);
#endif
    if (!hugeZero)
    {
        const float huge = ss_t((fbm3_paired<4>(npx * 0.0700f, npy * 0.0700f, npz * 0.0700f) + -0.2f) / (0.4f - 0.2f));
#ifdef MMG_FEATURE_STATS
        if (hugeCheck && huge != 0.f) atomicAdd(&g_hugeMismatch, 1ull);
#endif
        c->hugeFactor = fmaf(huge, 1.4f, 1.f);
        if (!(cave_thr(*c, 1.f) > 0.04f)) return 0;
    }
    const float ax = npx * 0.8000f, ay = npy * 0.8000f, az = npz * 0.8000f;
    float o1, o2, o3;
    fbm3_from3<5>(ax, ay, az, &o1, &o2, &o3);
    *px = fmaf(o1, 1.8f, npx); *py = fmaf(npy, 1.6f, o2 * 1.8f); *pz = fmaf(o3, 1.8f, npz);
    return 2;
}

// The "huge caves" term of the threshold, smoothstep(0.2, 0.4, fbm3<4>(pos * 0.00035)), is exactly 0 wherever the fbm is
// <= 0.2 - most of the world - yet the reference pays its four simplex3 at every voxel. The fbm is Lipschitz: one octave is
// 42 * simplex3_raw, whose gradient norm stays below 7.6 (20 M finite-difference samples of the oracle's restatement;
// kSimplex3Lipschitz = 10 is used), so the fbm moves by at most sum_i 2^-(i+1) * 2^i * 0.00035 * 10 = 0.007 per voxel of
// distance, in any direction. One sample at the middle of each 16-voxel run (y = 16 s + 8, s = 0..8: every voxel the threshold
// is evaluated for has y < 142) that is <= 0.2 - (distance to the farthest voxel it speaks for) * 0.007 - 0.004 (rounding) proves
// huge == 0 there. Validated by the census build (-DMMG_FEATURE_STATS evaluates the term anyway and counts disagreements: none),
// by tests/test_exact_shortcuts.py on the oracle's fbm and by the unchanged world hash.
constexpr float kSimplex3Lipschitz = 10.f;
constexpr int kHugeRun = 16, kHugeSamples = 9;
// The bound holds in every direction, so one sample serves a 4 x 4 block of columns: taken at the block's centre (x0 + 1.5, z0 + 1.5)
// and the middle of the run, it is at most sqrt(1.5^2 + 1.5^2 + 8^2) = 8.28 voxels from any voxel of the 4 x 4 x 16 box - the margin
// grows from 8 to 8.3 voxels' worth and the chunk needs 144 evaluations of the fbm instead of 2 304.
// Bit s of the result = proved for y in [16 s, 16 s + 16) of every column x0 .. x0 + 3, z0 .. z0 + 3.
__device__ __forceinline__ bool huge_zero_sample(int wx0, int wz0, int s)
{
    constexpr float perVoxel = 4.f * 0.5f * (0.0050f * 0.0700f) * kSimplex3Lipschitz;
    constexpr float limit = 0.2f - 8.3f * perVoxel - 0.004f;
    const float npx = ((float)wx0 + 1.5f) * 0.0050f, npz = ((float)wz0 + 1.5f) * 0.0050f;
    const float npy = (float)(kHugeRun * s + kHugeRun / 2) * 0.0050f;
    return fbm3_paired<4>(npx * 0.0700f, npy * 0.0700f, npz * 0.0700f) <= limit;
}

struct CaveColumn { float obw, ravTop, ravDepth; int ravActive; unsigned hugeZeroMask; };

__global__ void __launch_bounds__(256) k_cave_columns(const int* __restrict__ chunkList, const int2* __restrict__ origins,
                                                      const float* __restrict__ biomeWeights, CaveColumn* __restrict__ cols)
{
    __shared__ unsigned shHuge[16];      // per 4 x 4 block of columns: huge-caves term proved 0 for run s
    const int idx = threadIdx.x;
    if (idx < 16) shHuge[idx] = 0u;
    noise_tab_stage();
    const int li = blockIdx.x, chunk = chunkList ? chunkList[li] : li;
    const int2 o = origins[chunk];
    if (idx < 16 * kHugeSamples)
    {
        const int b = idx / kHugeSamples, sRun = idx % kHugeSamples;
        if (huge_zero_sample(o.x + 4 * (b & 3), o.y + 4 * (b >> 2), sRun)) atomicOr(&shHuge[b], 1u << sRun);
    }
    __syncthreads();
    const float* cw = biomeWeights + (size_t)chunk * (NUM_BIOMES * 256) + idx;
    float obw = 0.f;
#pragma unroll
    for (int b = 0; b < NUM_OCEAN_BEACH_BIOMES; ++b) obw += cw[b * 256];
    const Ravine r = ravine_column(o.x + (idx & 15), o.y + (idx >> 4), obw);
    CaveColumn c;
    c.obw = obw; c.ravTop = r.top; c.ravDepth = r.depth; c.ravActive = r.active ? 1 : 0;
    c.hugeZeroMask = shHuge[((idx & 15) >> 2) + 4 * (idx >> 6)];
    cols[(size_t)li * 256 + idx] = c;
}

// kCaveCols consecutive columns (one chunk row, x = 0 .. 15) per CTA. Only voxels 1 <= y <= 141 can need the noise (topRatio is 0
// from y = 142 - 50 obw upwards), so the survivors of all the columns - at most 141 each - are listed and the threads walk the
// list: the noise runs on full warps (one partly filled warp per CTA) and the CTA synchronises four times in all, where one
// column per 128-thread CTA in 128-voxel slabs spent a whole warp on the few voxels from 128 up and synchronised four times per
// slab (ncu, profiles/r02_k_caves_v2_src.txt: 26 of 32 lanes active in the noise code, 14 % of the stall samples at one slab
// barrier). The columns are 0.005 Worley cells apart and share one jitter table.
// Measured per 128x128-chunk region (profiles/r02_k_caves_columns.txt): 1 column / 128 threads 60.8 ms, 2 / 288 58.9, 4 / 128 52.5,
// 8 / 256 49.9, 16 / 256 48.6 (5 CTAs per SM; 4: 49.1; 384 threads x 3: 49.4; 512 x 2: 51.7).
#ifndef MMG_CAVE_COLS
#define MMG_CAVE_COLS 16
#endif
#ifndef MMG_CAVE_THREADS
#define MMG_CAVE_THREADS 256
#endif
constexpr int kCaveCols = MMG_CAVE_COLS, kCaveThreads = MMG_CAVE_THREADS, kCaveSurvivorTop = 141, kCaveListCap = kCaveCols * kCaveSurvivorTop + 32;
static_assert(256 % kCaveCols == 0 && kCaveCols <= 64, "whole groups of columns per chunk; 6 bits of column in a list entry");
__global__ void __launch_bounds__(kCaveThreads, MMG_CAVES_MINBLOCKS) k_caves(const int* __restrict__ chunkList, const int2* __restrict__ origins,
                                               const float* __restrict__ heightfield, const CaveColumn* __restrict__ cols,
                                               CaveLayer* __restrict__ caveLayers, uint2* __restrict__ biomeQueue, int* __restrict__ biomeCount,
                                               int biomeQueueCap)
{
    __shared__ unsigned int shFilled[kCaveCols][13];  // bit y = 1 if solid; word 12 = 0 (y = 384 is "not filled")
    __shared__ int shFlips[kCaveThreads / 32][2 * MAX_CAVE_LAYERS];      // per warp
    __shared__ float shJit[3 * kCaveBox * kCaveBox * kCaveBox];
    __shared__ int shNumList, shNumPending;
    __shared__ unsigned short shList[kCaveListCap];      // survivors: column << 9 | y
    __shared__ unsigned short shPendCY[kCaveListCap];    // voxels whose exact threshold is still needed (cave_fbm_a)
    __shared__ float shPendNoise[kCaveListCap];          // (their threshold terms are recomputed: a dozen instructions, and 6 bytes per entry let 16 columns fit)
    __shared__ float shHeight[kCaveCols];
    __shared__ CaveColumn shCol[kCaveCols];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int li = blockIdx.x / (256 / kCaveCols);
    const int chunk = chunkList ? chunkList[li] : li;
    const int idx0 = (blockIdx.x % (256 / kCaveCols)) * kCaveCols;      // columns idx0 .. idx0 + kCaveCols - 1: same z, consecutive x
    if (tid < kCaveCols)
    {
        shHeight[tid] = heightfield[(size_t)chunk * 256 + idx0 + tid];
        shCol[tid] = cols[(size_t)li * 256 + idx0 + tid];
    }
    if (tid == 0) { shNumList = 0; shNumPending = 0; }
    noise_tab_stage();      // its barrier also publishes the column data above
    const int2 o = origins[chunk];
    const int wx0 = o.x + (idx0 & 15), wz0 = o.y + (idx0 >> 4);      // column c of the CTA: x = (idx0 + c) & 15, z = (idx0 + c) >> 4
    // ---- the Worley jitter table: kCaveBox^3 cells around where the survivors' warped sample positions are expected -
    // p = pos * 0.005 * (1, 1.6, 1) + 1.8 * fbm3From3, y = 1 .. ~132, |fbm| mostly < 0.5 - chosen BEFORE the noise is known,
    // so that no barrier separates the noise from the Worley evaluation; a voxel whose 3x3x3 cells are not all in the table
    // computes the missing jitters in place (special_cave_noise_cached), the result is the same either way.
    const int bx = (int)floorf((float)wx0 * 0.0050f - 2.5f), by = (int)floorf(0.53f - 2.5f), bz = (int)floorf((float)wz0 * 0.0050f - 2.5f);
    {
        constexpr int N3 = kCaveBox * kCaveBox * kCaveBox;
        for (int i = tid; i < N3; i += kCaveThreads)
        {
            const int uz = i % kCaveBox, r = i / kCaveBox, uy = r % kCaveBox, ux = r / kCaveBox;
            float jx, jy, jz;
            cave_cell_jitter(bx + ux, by + uy, bz + uz, &jx, &jy, &jz);
            shJit[i] = jx; shJit[i + N3] = jy; shJit[i + 2 * N3] = jz;
        }
    }
    // ---- every voxel's state without noise: one warp per 32-voxel word of a column
    for (int wd = warp; wd < 12 * kCaveCols; wd += kCaveThreads / 32)
    {
        const int c = wd / 12, y = (wd % 12) * 32 + lane;
        const CaveColumn& cc = shCol[c];
        const bool hugeZero = y < kHugeRun * kHugeSamples && ((cc.hugeZeroMask >> (y / kHugeRun)) & 1u);
        CaveThr ct;
        int st;
        if (y >= 160)      // warp-uniform: topRatio is 0 from y = 142 up, only "above the terrain and the sea" is left to ask
        {
            const int hi = (int)shHeight[c];
            st = y > (hi > SEA_LEVEL ? hi : SEA_LEVEL) ? 1 : 0;
        }
        else
            st = cave_threshold_cheap(y, shHeight[c], cc.obw, hugeZero, &ct);
        // solid voxels fall to the ravine rule (chunk.cu:785-808); survivors count as solid until their noise says otherwise
        bool solid = st != 1;
        if (st == 0) solid = !(cc.ravActive != 0 && (cc.ravTop - cc.ravDepth) < (float)y && y != 0);
        const unsigned int bits = __ballot_sync(0xffffffffu, solid);
        const unsigned int sv = __ballot_sync(0xffffffffu, st == 2);
        if (lane == 0) shFilled[c][wd % 12] = bits;
        if (sv)
        {
            int base = 0;
            if (lane == 0) base = atomicAdd(&shNumList, __popc(sv));
            base = __shfl_sync(0xffffffffu, base, 0);
            const int slot = base + __popc(sv & ((1u << lane) - 1u));
            if (st == 2 && slot < kCaveListCap) shList[slot] = (unsigned short)(c << 9 | y);
        }
    }
    if (tid < kCaveCols) shFilled[tid][12] = 0u;
    __syncthreads();
    // ---- the survivors: threshold bounds, the warped sample position, the Worley noise there
    const int nList = min(shNumList, kCaveListCap);      // only 1 <= y <= 141 survive
    for (int i = tid; i < nList; i += kCaveThreads)
    {
        const int e = shList[i];
        const int c = e >> 9, y = e & 511;
        const CaveColumn& cc = shCol[c];
        const bool hugeZero = y < kHugeRun * kHugeSamples && ((cc.hugeZeroMask >> (y / kHugeRun)) & 1u);
        CaveThr ct = {0.f, 1.f};
        float px = 0.f, py = 0.f, pz = 0.f;
        cave_threshold_cheap(y, shHeight[c], cc.obw, hugeZero, &ct);      // recomputes ratio (a dozen instructions) instead of storing it
        const int wx = o.x + ((idx0 + c) & 15), wz = o.y + ((idx0 + c) >> 4);
        const int st = cave_threshold_noise(wx, y, wz, hugeZero, &ct, &px, &py, &pz);
        bool air = false, pending = false;
        float noise = 0.f;
        if (st == 2)
        {
            noise = special_cave_noise_cached(px, py, pz, bx, by, bz, kCaveBox, kCaveBox, kCaveBox, shJit);
            const float thrLo = cave_thr(ct, -1.f), thrHi = cave_thr(ct, 1.f);      // thrLo <= thr <= thrHi, thrHi > 0.04 here
            if (noise < thrHi)      // else solid: noise < thr is impossible
            {
                if (thrLo > 0.04f && noise < thrLo) air = true;      // thr > 0.04 && noise < thr whatever fbmA is
                else pending = true;
            }
#ifdef MMG_FEATURE_STATS
            atomicAdd(&g_caveWarped, 1ull);
            atomicAdd(&g_cavePending[pending ? 1 : 0], 1ull);
            if (!pending)
            {
                const float thr = cave_thr(ct, cave_fbm_a(wx, y, wz));
                if ((thr > 0.04f && noise < thr) != air) atomicAdd(&g_cavePending[2], 1ull);
            }
            {
                const int ux0 = (int)floorf(px) - 1 - bx, uy0 = (int)floorf(py) - 1 - by, uz0 = (int)floorf(pz) - 1 - bz;
                const bool inTable = (unsigned)ux0 <= (unsigned)(kCaveBox - 3) && (unsigned)uy0 <= (unsigned)(kCaveBox - 3) && (unsigned)uz0 <= (unsigned)(kCaveBox - 3);
                atomicAdd(&g_caveTable[inTable ? 1 : 0], 1ull);
            }
#endif
        }
        if (!air && !pending) air = cc.ravActive != 0 && (cc.ravTop - cc.ravDepth) < (float)y;      // chunk.cu:785-808 (y != 0 here)
        if (air) atomicAnd(&shFilled[c][y >> 5], ~(1u << (y & 31)));
        // pending voxels (a few per warp) are listed and get their exact threshold on adjacent lanes below; until then solid
        if (pending)
        {
            const int slot = atomicAdd(&shNumPending, 1);
            shPendCY[slot] = (unsigned short)(c << 9 | y); shPendNoise[slot] = noise;
        }
    }
    __syncthreads();
    // the pending voxels of all columns (~25 each) get their exact threshold on adjacent lanes
    for (int i = tid; i < shNumPending; i += kCaveThreads)
    {
        const int e = shPendCY[i], c2 = e >> 9, y2 = e & 511;
        const CaveColumn& cc = shCol[c2];
        const int wx = o.x + ((idx0 + c2) & 15), wz = o.y + ((idx0 + c2) >> 4);
        const bool hugeZero = y2 < kHugeRun * kHugeSamples && ((cc.hugeZeroMask >> (y2 / kHugeRun)) & 1u);
        CaveThr ct;
        cave_threshold_cheap(y2, shHeight[c2], cc.obw, hugeZero, &ct);
        if (!hugeZero) ct.hugeFactor = cave_huge_factor(wx, y2, wz);      // 0.5 % of the voxels
        const float thr = cave_thr(ct, cave_fbm_a(wx, y2, wz));
        bool air2 = thr > 0.04f && shPendNoise[i] < thr;
        if (!air2) air2 = cc.ravActive != 0 && (cc.ravTop - cc.ravDepth) < (float)y2;      // y2 != 0: y == 0 never gets here
        if (air2) atomicAnd(&shFilled[c2][y2 >> 5], ~(1u << (y2 & 31)));
    }
    __syncthreads();
    // ---- a warp per column turns the bit column into cave layers
    for (int c = warp; c < kCaveCols; c += kCaveThreads / 32)
    {
    const int idx = idx0 + c, wx = o.x + (idx & 15), wz = o.y + (idx >> 4);
    const float maxHeight = shHeight[c];
    CaveLayer* out = caveLayers + ((size_t)chunk * 256 + idx) * MAX_CAVE_LAYERS;
    int nflips;
    {
        // flips: filled[y] != filled[y+1]
        unsigned int f = 0u;
        if (lane < 12)
        {
            const unsigned int w0 = shFilled[c][lane], w1 = shFilled[c][lane + 1];
            f = w0 ^ ((w0 >> 1) | (w1 << 31));
        }
        const int cnt = __popc(f);
        int base = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1)
        {
            const int v = __shfl_up_sync(0xffffffffu, base, d);
            if (lane >= d) base += v;
        }
        nflips = min(__shfl_sync(0xffffffffu, base, 11), 2 * MAX_CAVE_LAYERS);
        base -= cnt;
        while (f)
        {
            const int b = __ffs(f) - 1;
            f &= f - 1;
            if (base < 2 * MAX_CAVE_LAYERS) shFlips[warp][base] = 32 * lane + b;
            ++base;
        }
    }
    __syncwarp();
    {
        // layer l: (start, end] from consecutive flips; the two cave biomes are looked up by k_cave_biomes
        const int l = lane;
        const int start = (2 * l < nflips) ? shFlips[warp][2 * l] : 384;
        const int end = (2 * l + 1 < nflips) ? shFlips[warp][2 * l + 1] : 384;
        CaveLayer cl;
        cl.start = start; cl.end = end; cl.bottomBiome = CB_NONE; cl.topBiome = CB_NONE; cl.pad[0] = cl.pad[1] = 0;
        const int need = (start != 384 ? 1 : 0) + (end != 384 ? 1 : 0);
        // warp-aggregated append of this column's lookups (the warp holds all 32 layers)
        int incl = need;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1)
        {
            const int v = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += v;
        }
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        int base = 0;
        if (lane == 0 && total > 0) base = atomicAdd(biomeCount, total);
        base = __shfl_sync(0xffffffffu, base, 0) + incl - need;
        const unsigned col = (unsigned)(chunk * 256 + idx);
        if (start != 384)
        {
            if (base < biomeQueueCap) biomeQueue[base] = make_uint2(col, (unsigned)(l << 10 | start));
            else cl.bottomBiome = (uint8_t)cave_biome(wx, start, wz, maxHeight, 329271348);
            ++base;
        }
        if (end != 384)
        {
            if (base < biomeQueueCap) biomeQueue[base] = make_uint2(col, (unsigned)(l << 10 | 1 << 9 | (end + 1)));
            else cl.topBiome = (uint8_t)cave_biome(wx, end + 1, wz, maxHeight, 4982921);
        }
        out[l] = cl;
    }
    __syncwarp();
    }      // columns of this warp
}

// getCaveBiome for the bottom / top of every cave layer (chunk.cu:915-935), one queued lookup per thread.
// k_caves finds ~5 lookups per column; run there they would occupy 5 lanes of a warp for ~5 kFLOP each.
__global__ void __launch_bounds__(128) k_cave_biomes(const int2* __restrict__ origins, const float* __restrict__ heightfield,
                                                     const uint2* __restrict__ biomeQueue, const int* __restrict__ biomeCount, int biomeQueueCap,
                                                     CaveLayer* __restrict__ caveLayers)
{
    const int n = min(*biomeCount, biomeQueueCap);
    if (blockIdx.x * blockDim.x >= n) return;
    noise_tab_stage();
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
        const uint2 e = biomeQueue[i];
        const int chunk = (int)(e.x >> 8), idx = (int)(e.x & 255u);
        const int l = (int)(e.y >> 10), top = (int)((e.y >> 9) & 1u), y = (int)(e.y & 511u);
        const int2 o = origins[chunk];
        const int b = cave_biome(o.x + (idx & 15), y, o.y + (idx >> 4), heightfield[e.x], top ? 4982921 : 329271348);
        CaveLayer* cl = caveLayers + (size_t)e.x * MAX_CAVE_LAYERS + l;
        if (top) cl->topBiome = (uint8_t)b;
        else cl->bottomBiome = (uint8_t)b;
    }
}

}  // namespace mmg
