// Chunk wire / on-disk format "MMCH1" and its device-side encoder.
//
// The reference keeps a chunk's blocks as a raw uint8[16][16][384] array, index y + 384 * (x + 16 * z)
// (/root/reference/src/terrain/biomeFuncs.hpp:25-30, chunk.hpp:59-72), downloads it raw after Chunk::fill (chunk.cu:1604-1621)
// and has no on-disk format at all. A generated column is a few long vertical runs - bedrock, rock
// strata, cave air, soil, air or water above - so the volume is run-length coded per column:
//
//   encoded chunk := u16 nRuns[256]                          column order x + 16 z, little endian        (512 bytes)
//                    for each column: nRuns x { u8 block, u8 length }   length 1..255, a column's lengths sum to 384
//
// padded with zero bytes to a multiple of 16. A chunk whose code would not be shorter than the raw volume is stored raw
// (98 304 bytes; recognised by its length).
// Typical chunks take 5-9 KB (11-19x smaller), which is what makes the device->host delivery of large worlds cheap
// (mmgen_world_generate_to_host_encoded) and region files small (mmgen_region_*).
#pragma once
#include "mm_common.cuh"

namespace mmg {

constexpr int kChunkBytes = 98304;
constexpr int kCodecHeaderBytes = 512;

// runs of one column: calls emit(block, length) for every run, lengths capped at 255; returns the number of runs
template <typename Emit>
__device__ __forceinline__ int column_runs(const uint8_t* __restrict__ col, Emit emit)
{
    int n = 0, len = 0;
    unsigned cur = 0;
#pragma unroll 1
    for (int q = 0; q < 24; ++q)
    {
        const uint4 v = reinterpret_cast<const uint4*>(col)[q];
        const unsigned w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
            for (int j = 0; j < 4; ++j)
            {
                const unsigned b = (w[k] >> (8 * j)) & 0xffu;
                if (len > 0 && (b != cur || len == 255))
                {
                    emit(cur, len, n);
                    ++n;
                    len = 0;
                }
                cur = b;
                ++len;
            }
    }
    emit(cur, len, n);
    return n + 1;
}

// pass 1: run count of every column and the encoded size of every chunk of the batch
__global__ void __launch_bounds__(256) k_encode_count(const int* __restrict__ list, const uint8_t* __restrict__ blocks,
                                                      unsigned short* __restrict__ nRuns, unsigned* __restrict__ sizes)
{
    __shared__ int shSum[8];
    const int li = blockIdx.x, chunk = list[li], t = threadIdx.x;
    const int n = column_runs(blocks + (size_t)chunk * kChunkBytes + (size_t)t * 384, [](unsigned, int, int) {});
    nRuns[(size_t)li * 256 + t] = (unsigned short)n;
    int v = n;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    if ((t & 31) == 0) shSum[t >> 5] = v;
    __syncthreads();
    if (t == 0)
    {
        int s = 0;
        for (int k = 0; k < 8; ++k) s += shSum[k];
        const unsigned bytes = (unsigned)(kCodecHeaderBytes + 2 * s + 15) & ~15u;       // padded: every chunk starts 16-byte aligned
        sizes[li] = bytes < (unsigned)kChunkBytes ? bytes : (unsigned)kChunkBytes;      // raw escape
    }
}

// pass 2 (one CTA): where every chunk of the batch goes in the arena, appended after what earlier batches wrote.
// index[slot] = {offset, bytes}; batchInfo = {arena offset of the batch, bytes of the batch}
constexpr int kEncodePlaceThreads = 1024;
__global__ void __launch_bounds__(kEncodePlaceThreads) k_encode_place(int m, const int* __restrict__ slots, const unsigned* __restrict__ sizes,
                                                      unsigned long long* __restrict__ arenaUsed, unsigned long long* __restrict__ offsets,
                                                      unsigned long long* __restrict__ index, unsigned long long* __restrict__ batchInfo)
{
    __shared__ unsigned long long sh[kEncodePlaceThreads];
    const int t = threadIdx.x;
    const unsigned long long base = *arenaUsed;
    unsigned long long carry = 0ull;      // bytes of the batch's chunks before this group of kEncodePlaceThreads
    for (int c0 = 0; c0 < m; c0 += kEncodePlaceThreads)
    {
        const int i = c0 + t;
        const unsigned long long mine = i < m ? sizes[i] : 0ull;
        __syncthreads();      // the previous group's totals have been read
        sh[t] = mine;
        __syncthreads();
        for (int d = 1; d < kEncodePlaceThreads; d <<= 1)
        {
            const unsigned long long a = t >= d ? sh[t - d] : 0ull;
            __syncthreads();
            sh[t] += a;
            __syncthreads();
        }
        if (i < m)
        {
            const unsigned long long off = base + carry + sh[t] - mine;
            offsets[i] = off;
            index[2 * (size_t)slots[i]] = off;
            index[2 * (size_t)slots[i] + 1] = mine;
        }
        carry += sh[kEncodePlaceThreads - 1];
    }
    if (t == 0)
    {
        batchInfo[0] = base;
        batchInfo[1] = carry;
        *arenaUsed = base + carry;
    }
}

// pass 3: header + runs of every chunk at its place in the arena (or the raw volume)
__global__ void __launch_bounds__(256) k_encode_emit(const int* __restrict__ list, const uint8_t* __restrict__ blocks,
                                                     const unsigned short* __restrict__ nRuns, const unsigned* __restrict__ sizes,
                                                     const unsigned long long* __restrict__ offsets, uint8_t* __restrict__ arena)
{
    __shared__ int shScan[256];
    const int li = blockIdx.x, chunk = list[li], t = threadIdx.x;
    const uint8_t* src = blocks + (size_t)chunk * kChunkBytes;
    uint8_t* dst = arena + offsets[li];
    if (sizes[li] == (unsigned)kChunkBytes)
    {
        for (int i = t; i < kChunkBytes / 16; i += 256) reinterpret_cast<uint4*>(dst)[i] = reinterpret_cast<const uint4*>(src)[i];
        return;
    }
    const int n = nRuns[(size_t)li * 256 + t];
    shScan[t] = n;
    __syncthreads();
    for (int d = 1; d < 256; d <<= 1)
    {
        const int a = t >= d ? shScan[t - d] : 0;
        __syncthreads();
        shScan[t] += a;
        __syncthreads();
    }
    reinterpret_cast<unsigned short*>(dst)[t] = (unsigned short)n;
    unsigned short* runs = reinterpret_cast<unsigned short*>(dst + kCodecHeaderBytes) + (shScan[t] - n);
    column_runs(src + (size_t)t * 384, [&](unsigned b, int len, int k) { runs[k] = (unsigned short)(b | (unsigned)len << 8); });
    if (t == 255)      // zero padding up to the chunk's 16-byte aligned size
        for (unsigned b = (unsigned)(kCodecHeaderBytes + 2 * shScan[255]); b < sizes[li]; b += 2) *reinterpret_cast<unsigned short*>(dst + b) = 0;
}

}  // namespace mmg
