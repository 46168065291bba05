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

// 0.5f * sinf(radians(45.f)) and normalize(vec3(1, 0, -+1)) as the host computes them (chunk.cu:1753, 1765-1766)
constexpr float kXOff = 0.35355338f, kXNor = 0.70710677f;
// xShapedVertPositions (chunk.cu:1754-1764): corner i of the two crossed quads is (+-kXOff, 0 | 1, +-kXOff)
__device__ __forceinline__ float mesh_x_vx(int i) { return ((0x69u >> i) & 1u) ? kXOff : -kXOff; }      // + for i in {0, 3, 5, 6}
__device__ __forceinline__ float mesh_x_vy(int i) { return ((0xCCu >> i) & 1u) ? 1.f : 0.f; }            // 1 for i in {2, 3, 6, 7}
__device__ __forceinline__ float mesh_x_vz(int i) { return ((0x99u >> i) & 1u) ? kXOff : -kXOff; }      // + for i in {0, 3, 4, 7}
// uvOffsets (chunk.cu:1777-1779): (0,0) (1,0) (1,1) (0,1)
__device__ __forceinline__ int mesh_uv_u(int j) { return (j == 1 || j == 2) ? 1 : 0; }
__device__ __forceinline__ int mesh_uv_v(int j) { return j >= 2 ? 1 : 0; }

// Lane-varying lookups go through shared memory or arithmetic: a __constant__ read with 32 different indices is serialised.
__device__ __forceinline__ int mesh_dx(int d) { return (d == 1) - (d == 3); }     // enums.hpp:41-48: +z, +x, -z, -x, +y, -y
__device__ __forceinline__ int mesh_dy(int d) { return (d == 4) - (d == 5); }
__device__ __forceinline__ int mesh_dz(int d) { return (d == 0) - (d == 2); }
// corner j of face d (chunk.cu:1768-1775) as bits x | y << 1 | z << 2, eight faces-corners per 32-bit word
__device__ __forceinline__ int mesh_face_corner(int d, int j)
{
    // d:      0 (+z)        1 (+x)        2 (-z)        3 (-x)        4 (+y)        5 (-y)
    // j = 0:  (0,0,1)=4     (1,0,1)=5     (1,0,0)=1     (0,0,0)=0     (0,1,1)=6     (0,0,0)=0
    // j = 1:  (1,0,1)=5     (1,0,0)=1     (0,0,0)=0     (0,0,1)=4     (1,1,1)=7     (1,0,0)=1
    // j = 2:  (1,1,1)=7     (1,1,0)=3     (0,1,0)=2     (0,1,1)=6     (1,1,0)=3     (1,0,1)=5
    // j = 3:  (0,1,1)=6     (1,1,1)=7     (1,1,0)=3     (0,1,0)=2     (0,1,0)=2     (0,0,1)=4
    const unsigned long long tab = 06754ull << 0 | 07315ull << 12 | 03201ull << 24 | 02640ull << 36 | 02376ull << 48;   // octal, j = 0 lowest
    const unsigned last = 04510u;                                                                                       // d = 5
    const unsigned row = d < 5 ? (unsigned)(tab >> (12 * d)) & 07777u : last;
    return (int)(row >> (3 * j)) & 7;
}
__device__ __forceinline__ void mesh_stage_block_data(uint32_t* shBD)
{
    for (int i = threadIdx.x; i < NUM_BLOCKS; i += blockDim.x) shBD[i] = c_blockData[i];
    __syncthreads();
}

// Visible faces of a voxel as a 6-bit mask (chunk.cu:1879-1936): mesh_face_mask_strip below; self = its block (not AIR, not X-shaped).
// The column a warp works on and its four horizontal neighbour columns (+z, +x, -z, -x; from the neighbour chunk at the rim), 384 bytes
// each, staged in shared memory by coalesced word loads (15 independent loads per lane) before the column is walked: the face tests then
// read shared memory instead of chasing six dependent global byte loads per voxel (ncu, profiles/r02_k_mesh_{count,emit}.txt: 34 % / 21 %
// of the stall samples sat on those loads). Returns the mask of horizontal directions whose neighbour chunk does not exist.
struct MeshStrip { uint32_t w[5][96]; };
__device__ __forceinline__ unsigned mesh_stage_strip(MeshStrip& S, const uint8_t* __restrict__ blocks, const MeshChunk& mc, int x, int z, int lane)
{
    unsigned missing = 0u;
#pragma unroll
    for (int k = 0; k < 5; ++k)
    {
        int nchunk = mc.chunk, nx = x, nz = z;
        if (k > 0)
        {
            const int d = k - 1;
            nx += mesh_dx(d); nz += mesh_dz(d);
            if (nx < 0) { nchunk = mc.nb[3]; nx += 16; }
            else if (nx >= 16) { nchunk = mc.nb[1]; nx -= 16; }
            else if (nz < 0) { nchunk = mc.nb[2]; nz += 16; }
            else if (nz >= 16) { nchunk = mc.nb[0]; nz -= 16; }
            if (nchunk < 0) { missing |= 1u << d; continue; }      // warp-uniform
        }
        const uint32_t* src = reinterpret_cast<const uint32_t*>(blocks + (size_t)nchunk * 98304 + 384 * (nx + 16 * nz));
        S.w[k][lane] = src[lane]; S.w[k][lane + 32] = src[lane + 32]; S.w[k][lane + 64] = src[lane + 64];
    }
    __syncwarp();
    return missing;
}
// mesh_face_mask on a staged strip: same tests, same mask
__device__ __forceinline__ unsigned mesh_face_mask_strip(const MeshStrip& S, unsigned missing, int y, uint32_t selfData, const uint32_t* shBD)
{
    const unsigned selfTrans = selfData & 3u;
    unsigned mask = 0u;
#pragma unroll
    for (int d = 0; d < 6; ++d)
    {
        const int ny = y + mesh_dy(d);
        if (ny >= 0 && ny < 384)
        {
            if (d < 4 && ((missing >> d) & 1u)) continue;             // no neighbour chunk: the face is not emitted (chunk.cu:1908-1911)
            const uint8_t nb = reinterpret_cast<const uint8_t*>(S.w[d < 4 ? d + 1 : 0])[ny];
            const unsigned nTrans = shBD[nb] & 3u;
            const bool show = (selfTrans == 2u) ? (nb == B_AIR || nTrans == 1u) : (nTrans != 0u);
            if (!show) continue;
        }
        mask |= 1u << d;                                               // faces on the world's floor and ceiling are always emitted
    }
    return mask;
}

// Both kernels: one CTA (12 warps) per chunk; a warp takes columns warp, warp + 12, ..., stages the column and its four neighbours, and
// its lanes take 32 consecutive voxels of the column at a time, so the faces of 32 voxels are known together.
constexpr int kMeshWarps = 12;
constexpr int kMeshStripBytes = kMeshWarps * (int)sizeof(MeshStrip);

// per chunk: colOff[256] = exclusive scan of the columns' vertex counts, totals[li] = vertices of the chunk
__global__ void __launch_bounds__(32 * kMeshWarps) k_mesh_count(const MeshChunk* __restrict__ list, const uint8_t* __restrict__ blocks,
                                                                int* __restrict__ colOff, int* __restrict__ totals)
{
    __shared__ int sh[256];
    __shared__ uint32_t shBD[NUM_BLOCKS];
    __shared__ MeshStrip shStrip[kMeshWarps];
    const int li = blockIdx.x, t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const MeshChunk mc = list[li];
    mesh_stage_block_data(shBD);
    MeshStrip& S = shStrip[warp];
    for (int c = warp; c < 256; c += kMeshWarps)
    {
        const int x = c & 15, z = c >> 4;
        __syncwarp();                                                      // the previous column's strip has been read
        const unsigned missing = mesh_stage_strip(S, blocks, mc, x, z, lane);
        const uint8_t* col = reinterpret_cast<const uint8_t*>(S.w[0]);
        int n = 0;
        for (int y0 = 0; y0 < 384; y0 += 32)
        {
            const uint8_t b = col[y0 + lane];
            if (b == B_AIR) continue;
            const uint32_t data = shBD[b];
            n += ((data & 3u) == 3u) ? 8 : 4 * __popc(mesh_face_mask_strip(S, missing, y0 + lane, data, shBD));
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) n += __shfl_xor_sync(0xffffffffu, n, d);
        if (lane == 0) sh[c] = n;
    }
    __syncthreads();
    const int n = t < 256 ? sh[t] : 0;
    for (int d = 1; d < 256; d <<= 1)
    {
        const int a = (t < 256 && t >= d) ? sh[t - d] : 0;
        __syncthreads();
        if (t < 256) sh[t] += a;
        __syncthreads();
    }
    if (t < 256) colOff[li * 256 + t] = sh[t] - n;
    if (t == 255) totals[li] = sh[255];
}

// One quad of the output, as the emitting warp keeps it in shared memory between finding it and writing it.
// code: y (9 bits) | kind << 9 (0-5 = cube face direction, 6 / 7 = first / second quad of an X-shaped plant) |
//       uvStart << 12 | (uvFlip + 1) << 14 | texture cell u << 17 | v << 21 | material << 25; bx, bz: base position of a plant.
struct MeshQuad { uint32_t code; float bx, bz; };

// vertBase[li]: first vertex of the chunk in the arena (indices are relative to the chunk, like the reference's idx vector).
// The lanes of a warp first list the quads of their 32 voxels in order (warp prefix sum). Then every lane builds one vertex
// (8 quads per round) into a staging strip in shared memory and the warp copies the strip out 8-byte word by word: a warp
// store covers 256 contiguous bytes of the vertex array.
constexpr int kMeshWarpQuads = 32 * 6;
__global__ void __launch_bounds__(32 * kMeshWarps) k_mesh_emit(const MeshChunk* __restrict__ list, const uint8_t* __restrict__ blocks,
                                                               const int* __restrict__ colOff, const long long* __restrict__ vertBase,
                                                               MeshVertex* __restrict__ verts, uint32_t* __restrict__ idx)
{
    __shared__ MeshQuad shQ[kMeshWarps][kMeshWarpQuads];
    __shared__ __align__(16) uint2 shStage[kMeshWarps][32 * 5];
    __shared__ uint32_t shBD[NUM_BLOCKS];
    MeshStrip* shStrip = reinterpret_cast<MeshStrip*>(mmg_dyn_smem);      // kMeshStripBytes of dynamic shared memory (the static arrays take 44 KB)
    const int li = blockIdx.x, t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const MeshChunk mc = list[li];
    mesh_stage_block_data(shBD);
    uint2* vout = reinterpret_cast<uint2*>(verts + vertBase[li]);
    uint2* iout = reinterpret_cast<uint2*>(idx + (vertBase[li] / 4) * 6);     // 6 indices per 4 vertices throughout; 8-byte aligned
    MeshQuad* Q = shQ[warp];
    uint2* stage = shStage[warp];
    MeshStrip& S = shStrip[warp];
    for (int c = warp; c < 256; c += kMeshWarps)
    {
        const int x = c & 15, z = c >> 4;
        __syncwarp();                                                      // the previous column's strip has been read
        const unsigned missing = mesh_stage_strip(S, blocks, mc, x, z, lane);
        const uint8_t* col = reinterpret_cast<const uint8_t*>(S.w[0]);
        int quadBase = colOff[li * 256 + c] / 4;                           // quad index inside the chunk
        for (int y0 = 0; y0 < 384; y0 += 32)
        {
            const int y = y0 + lane;
            const uint8_t b = col[y];
            unsigned mask = 0u;
            int nq = 0;
            uint32_t data = 0u;
            if (b != B_AIR)
            {
                data = shBD[b];
                if ((data & 3u) == 3u) nq = 2;
                else { mask = mesh_face_mask_strip(S, missing, y, data, shBD); nq = __popc(mask); }
            }
            if (!__ballot_sync(0xffffffffu, nq != 0)) continue;            // air above the terrain, rock below it: most words hold no quad
            int incl = nq;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1)
            {
                const int v = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += v;
            }
            const int total = __shfl_sync(0xffffffffu, incl, 31);
            if (total == 0) continue;
            int slot = incl - nq;
            if (nq)
            {
                const uint32_t mat = (uint32_t)mesh_material(b) << 25;
                if ((data & 3u) == 3u)
                {
                    // X-shaped plant (chunk.cu:1833-1875): jittered by rand2From2 of the world column, evaluated on the HOST in the
                    // reference (glibc sinf, no FMA contraction): hm_sinf and separately rounded products
                    const float wxf = (float)(mc.origin.x + x), wzf = (float)(mc.origin.y + z);
                    const float d1 = wxf * 238.68f + wzf * 491.28f, d2 = wxf * 654.37f + wzf * 560.45f;
                    float r1 = hm_sinf(d1) * 39021.426f, r2 = hm_sinf(d2) * 39021.426f;
                    r1 = r1 - floorf(r1); r2 = r2 - floorf(r2);
                    MeshQuad q;
                    q.bx = ((float)x + 0.5f) + 0.4f * (r1 - 0.5f); q.bz = ((float)z + 0.5f) + 0.4f * (r2 - 0.5f);
                    const uint32_t common = (uint32_t)y | mat | ((data >> 2) & 15u) << 17 | ((data >> 6) & 15u) << 21;     // side cell
                    q.code = common | 6u << 9; Q[slot] = q;
                    q.code = common | 7u << 9; Q[slot + 1] = q;
                }
                else
                    while (mask)
                    {
                        const int d = __ffs(mask) - 1;
                        mask &= mask - 1;
                        const int side = d == 4 ? 1 : (d == 5 ? 2 : 0);                         // 0 side, 1 top, 2 bottom (chunk.cu:1940-1952)
                        int uvStart = 0, uvFlip = -1;
                        const bool rot = (data >> (26 + side)) & 1u, flip = (data >> (29 + side)) & 1u;
                        if (rot || flip)
                        {
                            // makeSeededRandomEngine(worldPos, dirIdx) + u04 (chunk.cu:1954-1967)
                            Minstd rng = make_rng4(mc.origin.x + x, y, mc.origin.y + z, d);
                            if (rot) uvStart = (int)(rng.u01() * 4.f + 0.f);
                            if (flip) uvFlip = (int)(rng.u01() * 4.f + 0.f);
                        }
                        MeshQuad q;
                        q.bx = q.bz = 0.f;
                        q.code = (uint32_t)y | (uint32_t)d << 9 | (uint32_t)uvStart << 12 | (uint32_t)(uvFlip + 1) << 14 |
                                 ((data >> (2 + 8 * side)) & 15u) << 17 | ((data >> (6 + 8 * side)) & 15u) << 21 | mat;
                        Q[slot++] = q;
                    }
            }
            __syncwarp();
            // vertices: lane = (quad, corner); 8 quads per round
            for (int f0 = 0; f0 < total; f0 += 8)
            {
                const int f = f0 + (lane >> 2), j = lane & 3;
                if (f < total)
                {
                    const MeshQuad q = Q[f];
                    const int qy = q.code & 511, kind = (q.code >> 9) & 7;
                    float px, py, pz, nx, ny, nz;
                    int ou, ov;
                    if (kind >= 6)
                    {
                        const int i = (kind - 6) * 4 + j;
                        px = q.bx + mesh_x_vx(i); py = (float)qy + mesh_x_vy(i); pz = q.bz + mesh_x_vz(i);
                        nx = kXNor; ny = 0.f; nz = kind == 6 ? -kXNor : kXNor;
                        ou = mesh_uv_u(j); ov = mesh_uv_v(j);
                    }
                    else
                    {
                        const int corner = mesh_face_corner(kind, j);
                        px = (float)(x + (corner & 1)); py = (float)(qy + ((corner >> 1) & 1)); pz = (float)(z + (corner >> 2));
                        nx = (float)mesh_dx(kind); ny = (float)mesh_dy(kind); nz = (float)mesh_dz(kind);
                        const int uvStart = (q.code >> 12) & 3, uvFlip = (int)((q.code >> 14) & 7) - 1;
                        ou = mesh_uv_u((uvStart + j) & 3); ov = mesh_uv_v((uvStart + j) & 3);
                        if (uvFlip != -1)
                        {
                            if (uvFlip & 1) ou = 1 - ou;
                            if (uvFlip & 2) ov = 1 - ov;
                        }
                    }
                    const int cu = (q.code >> 17) & 15, cv = (q.code >> 21) & 15;
                    uint2* o = stage + lane * 5;                                   // Vertex: pos, nor, uv, m (40 bytes)
                    o[0] = make_uint2(__float_as_uint(px), __float_as_uint(py));
                    o[1] = make_uint2(__float_as_uint(pz), __float_as_uint(nx));
                    o[2] = make_uint2(__float_as_uint(ny), __float_as_uint(nz));
                    o[3] = make_uint2(__float_as_uint((float)(cu + ou) * 0.0625f), __float_as_uint((float)(cv + ov) * 0.0625f));
                    o[4] = make_uint2(q.code >> 25, 0u);
                }
                __syncwarp();
                // a quad is 160 bytes and a chunk's vertices start on a quad boundary of the arena: 16-byte words throughout
                const int words = min(8, total - f0) * 10;
                uint4* vo = reinterpret_cast<uint4*>(vout + (size_t)(quadBase + f0) * 20);
                const uint4* st4 = reinterpret_cast<const uint4*>(stage);
#pragma unroll
                for (int r = 0; r < 3; ++r)
                    if (r * 32 + lane < words) vo[r * 32 + lane] = st4[r * 32 + lane];
                __syncwarp();
            }
            // indices: total quads x 3 words: (f, f+1) (f+2, f) (f+2, f+3) with f = first vertex of the quad
            uint2* io = iout + (size_t)quadBase * 3;
            for (int wi = lane; wi < total * 3; wi += 32)
            {
                const int f = wi / 3, k = wi - f * 3;
                const uint32_t first = (uint32_t)(quadBase + f) * 4u;
                io[wi] = make_uint2(first + (k == 0 ? 0u : 2u), first + (k == 0 ? 1u : (k == 1 ? 0u : 3u)));
            }
            quadBase += total;
        }
    }
}

}  // namespace mmg
