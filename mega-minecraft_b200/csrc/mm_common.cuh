// Launch/error utilities of the generation path. Replaces the reference's CUDA_CHECK /
// CudaUtils::checkCUDAError (/root/reference/src/util/common.h:9-16,
// /root/reference/src/cuda/cuda_utils.cpp:5-17), which print and exit(); a library must not exit,
// so errors are recorded and returned through the C ABI instead.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <string>

namespace mmg {

extern thread_local std::string g_lastError;
extern uint64_t g_launchCount;

inline int fail(const char* what, cudaError_t err, const char* file, int line)
{
    char buf[512];
    snprintf(buf, sizeof(buf), "%s: %s (%s) at %s:%d", what, cudaGetErrorName(err), cudaGetErrorString(err), file, line);
    g_lastError = buf;
    return 1;
}

#define MMG_CUDA(call)                                                        \
    do {                                                                      \
        cudaError_t mmg_err_ = (call);                                        \
        if (mmg_err_ != cudaSuccess) return ::mmg::fail(#call, mmg_err_, __FILE__, __LINE__); \
    } while (0)

// every kernel launch of this library goes through here so launches can be counted and checked
#define MMG_LAUNCH(kernel, grid, block, smem, stream, ...)                    \
    do {                                                                      \
        kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);           \
        ++::mmg::g_launchCount;                                               \
        MMG_CUDA(cudaGetLastError());                                         \
    } while (0)

extern int g_numSMs;            // multiprocessor count of the bound device (148 on a B200), set by mmgen_init
#define kNumSMs (::mmg::g_numSMs)

// Work counters for the roofline figures of the cheap stages (mmgen_work_counters; bench.py turns them into algorithmic
// FLOPs / bytes with SURVEY.md 8(d)'s canonical costs). Always on: a few warp-aggregated atomics per CTA of S1 / S2 / S3.
//   [0..23] S1: columns in which surface biome b has weight > 0 (getHeight is evaluated for exactly those, chunk.cu:171-179)
//   [24]    S1: columns;  [25] S2: fbm<5> evaluations (one per stratified layer with weight > 0 that is reached, chunk.cu:308-320)
//   [26]    S2: columns;  [27] S3: 32x32 tiles actually swept (k_erode_sweep CTAs that did not return at the quiet-tile test)
//   [28]    S3: tile CTAs that returned at the quiet-tile test (swept + quiet = launched)
//   [29]    S6: gathered placements of the filled chunks (the reference tests each of them at each of the chunk's 98 304 voxels, chunk.cu:1444-1500)
//   [30]    S6: (column, y) pairs inside the placements' clipped boxes that k_fill_features looked at;  [31] pairs that reached a rasteriser
enum { W_S1_BIOME0 = 0, W_S1_COLUMNS = 24, W_S2_FBM5 = 25, W_S2_COLUMNS = 26, W_S3_TILES_SWEPT = 27, W_S3_TILES_QUIET = 28,
       W_S6_PLACEMENTS = 29, W_S6_PAIRS = 30, W_S6_RASTERISED = 31, W_NUM = 32 };
__device__ unsigned long long g_work[W_NUM];

// resident CTAs per SM the register allocator is asked to allow (tuned on a B200)
#ifndef MMG_CAVES_MINBLOCKS
#define MMG_CAVES_MINBLOCKS 5
#endif
#ifndef MMG_TERRAIN_MINBLOCKS
#define MMG_TERRAIN_MINBLOCKS 5
#endif
#ifndef MMG_ROCK_TWO_LEVEL
#define MMG_ROCK_TWO_LEVEL 1
#endif
#ifndef MMG_ROCK_MINBLOCKS
#define MMG_ROCK_MINBLOCKS 8
#endif

}  // namespace mmg
