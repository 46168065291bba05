// Stage 1 kernel: per-column surface-biome noise, 24 biome weights and the blended height.
// Replaces kernGenerateHeightfield (/root/reference/src/terrain/chunk.cu:150-185).
//
// Layout: one thread per column, 256-thread CTAs = one chunk each; outputs are the reference's
// wire layouts (height[chunk][z][x], weights[chunk][biome][z][x]) so a warp writes 128 contiguous
// bytes per plane. The work is FP32-pipe bound (11 simplex for the biome noise + 5..16 per active
// biome); algorithmic bytes are 25 608 B per chunk.
#pragma once
#include "mm_common.cuh"
#include "mm_surface.cuh"

namespace mmg {

__global__ void __launch_bounds__(256) k_heightfield(const int* __restrict__ chunkList, const int2* __restrict__ origins,
                                                     float* __restrict__ heightfield, float* __restrict__ biomeWeights)
{
    noise_tab_stage();
    const int chunk = chunkList ? chunkList[blockIdx.x] : blockIdx.x;
    const int idx = threadIdx.x;            // x + 16*z
    const int2 o = origins[chunk];
    const int wx = o.x + (idx & 15), wz = o.y + (idx >> 4);
    float* w = biomeWeights + (size_t)chunk * (NUM_BIOMES * 256) + idx;
    unsigned active = 0u;
    const float h = surface_column(wx, wz, w, 256, &active);
    heightfield[(size_t)chunk * 256 + idx] = h;
    // work counters (mm_common.cuh): per biome, the columns of this CTA that evaluated its height function
    __shared__ unsigned shCount[NUM_BIOMES];
    if (idx < NUM_BIOMES) shCount[idx] = 0u;
    __syncthreads();
    const int lane = idx & 31;
#pragma unroll
    for (int b = 0; b < NUM_BIOMES; ++b)
    {
        const unsigned m = __ballot_sync(0xffffffffu, (active >> b) & 1u);
        if (lane == 0 && m) atomicAdd(&shCount[b], (unsigned)__popc(m));
    }
    __syncthreads();
    if (idx < NUM_BIOMES && shCount[idx]) atomicAdd(&g_work[W_S1_BIOME0 + idx], (unsigned long long)shCount[idx]);
    if (idx == NUM_BIOMES) atomicAdd(&g_work[W_S1_COLUMNS], 256ull);
}

}  // namespace mmg
