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
    const float h = surface_column(wx, wz, w, 256);
    heightfield[(size_t)chunk * 256 + idx] = h;
}

}  // namespace mmg
