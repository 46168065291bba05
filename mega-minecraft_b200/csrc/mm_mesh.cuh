// Chunk meshing on the device: Chunk::createVBOs (/root/reference/src/terrain/chunk.cu:1753-2003, a host loop in the
// reference that re-uploads what Chunk::fill just downloaded) as two kernels over the resident block volumes.
//
// The reference walks z, x, y and appends 4 vertices + 6 indices per visible cube face (8 + 12 for an X-shaped plant),
// so a chunk's vertex order is the voxel memory order (y fastest) and, inside a voxel, the direction order
// +z, +x, -z, -x, +y, -y (enums.hpp:41-48). k_mesh_count counts the vertices of every column (one thread per column),
// scans the 256 counts of a chunk and leaves each column's offset; k_mesh_emit walks the columns again and writes the
// records at those offsets: same records, same order as the reference's vectors. HBM-bound: 96 KB of block IDs read
// (plus the 4 neighbour faces), 40 B per vertex + 6 B per vertex of indices written.
#pragma once
#include "mm_blockdata.cuh"
#include "mm_common.cuh"
#include "mm_arith.cuh"
#include "mm_hostmath.cuh"
#include "mm_tables.cuh"

namespace mmg {

// rendering/structs.hpp:25-31: vec3 pos, vec3 nor, vec2 uv, Mats m (enum class : size_t) - 40 bytes
struct MeshVertex { float px, py, pz, nx, ny, nz, u, v; unsigned long long m; };
static_assert(sizeof(MeshVertex) == 40, "Vertex layout");

struct MeshChunk { int chunk; int nb[4]; int2 origin; };   // nb: +z, +x, -z, -x neighbour chunk index or -1 (Chunk::neighbors)

enum { MAT_DIFFUSE, MAT_WATER, MAT_CRYSTAL, MAT_SMOOTH_MICRO, MAT_MICRO, MAT_ROUGH_MICRO };   // structs.hpp:7-14

// chunk.cu:1799-1828
__device__ __forceinline__ int mesh_material(uint8_t b)
{
    switch (b)
    {
    case B_WATER: return MAT_WATER;
    case B_CYAN_CRYSTAL: case B_GREEN_CRYSTAL: case B_MAGENTA_CRYSTAL: return MAT_CRYSTAL;
    case B_MARBLE: case B_QUARTZ: case B_ICE: case B_PACKED_ICE: case B_BLUE_ICE: return MAT_SMOOTH_MICRO;
    case B_SNOW: case B_SNOWY_GRASS_BLOCK: return MAT_MICRO;
    case B_SAND: case B_GRAVEL: return MAT_ROUGH_MICRO;
    default: return MAT_DIFFUSE;
    }
}

__constant__ const int c_meshDir[6][3] = {{0, 0, 1}, {1, 0, 0}, {0, 0, -1}, {-1, 0, 0}, {0, 1, 0}, {0, -1, 0}};       // enums.hpp:41-48
__constant__ const int c_meshFaceVerts[24][3] = {                                                                     // chunk.cu:1768-1775
    {0, 0, 1}, {1, 0, 1}, {1, 1, 1}, {0, 1, 1}, {1, 0, 1}, {1, 0, 0}, {1, 1, 0}, {1, 1, 1}, {1, 0, 0}, {0, 0, 0}, {0, 1, 0}, {1, 1, 0},
    {0, 0, 0}, {0, 0, 1}, {0, 1, 1}, {0, 1, 0}, {0, 1, 1}, {1, 1, 1}, {1, 1, 0}, {0, 1, 0}, {0, 0, 0}, {1, 0, 0}, {1, 0, 1}, {0, 0, 1}};
__constant__ const int c_meshUvOff[4][2] = {{0, 0}, {1, 0}, {1, 1}, {0, 1}};                                          // chunk.cu:1777-1779
// 0.5f * sinf(radians(45.f)) and normalize(vec3(1, 0, -+1)) as the host computes them (chunk.cu:1753, 1765-1766)
constexpr float kXOff = 0.35355338f, kXNor = 0.70710677f;
__constant__ const float c_meshXVerts[8][3] = {{kXOff, 0.f, kXOff}, {-kXOff, 0.f, -kXOff}, {-kXOff, 1.f, -kXOff}, {kXOff, 1.f, kXOff},
                                               {-kXOff, 0.f, kXOff}, {kXOff, 0.f, -kXOff}, {kXOff, 1.f, -kXOff}, {-kXOff, 1.f, kXOff}};

// Visible faces of voxel (x, y, z) of the chunk as a 6-bit mask (chunk.cu:1879-1936); self = its block (not AIR, not X-shaped).
__device__ __forceinline__ unsigned mesh_face_mask(const uint8_t* __restrict__ blocks, const MeshChunk& mc, int x, int y, int z, uint32_t selfData)
{
    const unsigned selfTrans = selfData & 3u;
    unsigned mask = 0u;
#pragma unroll
    for (int d = 0; d < 6; ++d)
    {
        int nx = x + c_meshDir[d][0], ny = y + c_meshDir[d][1], nz = z + c_meshDir[d][2];
        if (ny >= 0 && ny < 384)
        {
            int nchunk = mc.chunk;
            if (nx < 0) { nchunk = mc.nb[3]; nx += 16; }
            else if (nx >= 16) { nchunk = mc.nb[1]; nx -= 16; }
            else if (nz < 0) { nchunk = mc.nb[2]; nz += 16; }
            else if (nz >= 16) { nchunk = mc.nb[0]; nz -= 16; }
            if (nchunk < 0) continue;                                  // no neighbour chunk: the face is not emitted (chunk.cu:1908-1911)
            const uint8_t nb = blocks[(size_t)nchunk * 98304 + ny + 384 * (nx + 16 * nz)];
            const unsigned nTrans = c_blockData[nb] & 3u;
            const bool show = (selfTrans == 2u) ? (nb == B_AIR || nTrans == 1u) : (nTrans != 0u);
            if (!show) continue;
        }
        mask |= 1u << d;                                               // faces on the world's floor and ceiling are always emitted
    }
    return mask;
}

// per chunk: colOff[256] = exclusive scan of the columns' vertex counts, totals[li] = vertices of the chunk
__global__ void __launch_bounds__(256) k_mesh_count(const MeshChunk* __restrict__ list, const uint8_t* __restrict__ blocks,
                                                    int* __restrict__ colOff, int* __restrict__ totals)
{
    __shared__ int sh[256];
    const int li = blockIdx.x, t = threadIdx.x, x = t & 15, z = t >> 4;
    const MeshChunk mc = list[li];
    const uint8_t* col = blocks + (size_t)mc.chunk * 98304 + (size_t)t * 384;
    int n = 0;
    for (int y0 = 0; y0 < 384; y0 += 16)
    {
        const uint4 v = *reinterpret_cast<const uint4*>(col + y0);
        if ((v.x | v.y | v.z | v.w) == 0u) continue;                  // 16 AIR voxels
        const unsigned w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 16; ++k)
        {
            const uint8_t b = (uint8_t)(w[k >> 2] >> (8 * (k & 3)));
            if (b == B_AIR) continue;
            const uint32_t data = c_blockData[b];
            n += ((data & 3u) == 3u) ? 8 : 4 * __popc(mesh_face_mask(blocks, mc, x, y0 + k, z, data));
        }
    }
    sh[t] = n;
    __syncthreads();
    for (int d = 1; d < 256; d <<= 1)
    {
        const int a = t >= d ? sh[t - d] : 0;
        __syncthreads();
        sh[t] += a;
        __syncthreads();
    }
    colOff[li * 256 + t] = sh[t] - n;
    if (t == 255) totals[li] = sh[255];
}

__device__ __forceinline__ void mesh_store(MeshVertex* out, float px, float py, float pz, float nx, float ny, float nz, float u, float v, int m)
{
    float2* o = reinterpret_cast<float2*>(out);
    o[0] = make_float2(px, py); o[1] = make_float2(pz, nx); o[2] = make_float2(ny, nz); o[3] = make_float2(u, v);
    reinterpret_cast<unsigned long long*>(out)[4] = (unsigned long long)m;
}

// vertBase[li]: first vertex of the chunk in the arena (indices are relative to the chunk, like the reference's idx vector)
__global__ void __launch_bounds__(256) k_mesh_emit(const MeshChunk* __restrict__ list, const uint8_t* __restrict__ blocks,
                                                   const int* __restrict__ colOff, const long long* __restrict__ vertBase,
                                                   MeshVertex* __restrict__ verts, uint32_t* __restrict__ idx)
{
    const int li = blockIdx.x, t = threadIdx.x, x = t & 15, z = t >> 4;
    const MeshChunk mc = list[li];
    const uint8_t* col = blocks + (size_t)mc.chunk * 98304 + (size_t)t * 384;
    int local = colOff[li * 256 + t];                                  // vertex index inside the chunk
    MeshVertex* vout = verts + vertBase[li];
    uint32_t* iout = idx + (vertBase[li] / 4) * 6;                     // 6 indices per 4 vertices throughout
    auto quad = [&](int first) {
        uint32_t* q = iout + (first / 4) * 6;
        q[0] = first; q[1] = first + 1; q[2] = first + 2; q[3] = first; q[4] = first + 2; q[5] = first + 3;
    };
    for (int y0 = 0; y0 < 384; y0 += 16)
    {
        const uint4 v = *reinterpret_cast<const uint4*>(col + y0);
        if ((v.x | v.y | v.z | v.w) == 0u) continue;
        const unsigned w[4] = {v.x, v.y, v.z, v.w};
        for (int k = 0; k < 16; ++k)
        {
            const uint8_t b = (uint8_t)(w[k >> 2] >> (8 * (k & 3)));
            if (b == B_AIR) continue;
            const int y = y0 + k;
            const uint32_t data = c_blockData[b];
            const int mat = mesh_material(b);
            if ((data & 3u) == 3u)
            {
                // X-shaped plant (chunk.cu:1833-1875): jittered by rand2From2 of the world column, evaluated on the HOST in the
                // reference (glibc sinf, no FMA contraction): hm_sinf and separately rounded products
                const float wxf = (float)(mc.origin.x + x), wzf = (float)(mc.origin.y + z);
                const float d1 = wxf * 238.68f + wzf * 491.28f, d2 = wxf * 654.37f + wzf * 560.45f;
                float r1 = hm_sinf(d1) * 39021.426f, r2 = hm_sinf(d2) * 39021.426f;
                r1 = r1 - floorf(r1); r2 = r2 - floorf(r2);
                const float bx = ((float)x + 0.5f) + 0.4f * (r1 - 0.5f), by = (float)y, bz = ((float)z + 0.5f) + 0.4f * (r2 - 0.5f);
                const int su = (data >> 2) & 15, sv = (data >> 6) & 15;
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    mesh_store(vout + local + i, bx + c_meshXVerts[i][0], by + c_meshXVerts[i][1], bz + c_meshXVerts[i][2],
                               kXNor, 0.f, i < 4 ? -kXNor : kXNor,
                               (float)(su + c_meshUvOff[i & 3][0]) * 0.0625f, (float)(sv + c_meshUvOff[i & 3][1]) * 0.0625f, mat);
                quad(local);
                quad(local + 4);
                local += 8;
                continue;
            }
            unsigned mask = mesh_face_mask(blocks, mc, x, y, z, data);
            while (mask)
            {
                const int d = __ffs(mask) - 1;
                mask &= mask - 1;
                const int side = c_meshDir[d][1] == 1 ? 1 : (c_meshDir[d][1] == -1 ? 2 : 0);      // 0 side, 1 top, 2 bottom
                const int cu = (data >> (2 + 8 * side)) & 15, cv = (data >> (6 + 8 * side)) & 15;
                int uvStart = 0, uvFlip = -1;
                const bool rot = (data >> (26 + side)) & 1u, flip = (data >> (29 + side)) & 1u;
                if (rot || flip)
                {
                    // makeSeededRandomEngine(worldPos, dirIdx) + u04 (chunk.cu:1954-1967)
                    Minstd rng = make_rng4(mc.origin.x + x, y, mc.origin.y + z, d);
                    if (rot) uvStart = (int)(rng.u01() * 4.f + 0.f);
                    if (flip) uvFlip = (int)(rng.u01() * 4.f + 0.f);
                }
#pragma unroll
                for (int j = 0; j < 4; ++j)
                {
                    int ou = c_meshUvOff[(uvStart + j) & 3][0], ov = c_meshUvOff[(uvStart + j) & 3][1];
                    if (uvFlip != -1)
                    {
                        if (uvFlip & 1) ou = 1 - ou;
                        if (uvFlip & 2) ov = 1 - ov;
                    }
                    mesh_store(vout + local + j, (float)(x + c_meshFaceVerts[d * 4 + j][0]), (float)(y + c_meshFaceVerts[d * 4 + j][1]),
                               (float)(z + c_meshFaceVerts[d * 4 + j][2]), (float)c_meshDir[d][0], (float)c_meshDir[d][1], (float)c_meshDir[d][2],
                               (float)(cu + ou) * 0.0625f, (float)(cv + ov) * 0.0625f, mat);
                }
                quad(local);
                local += 4;
            }
        }
    }
}

}  // namespace mmg
