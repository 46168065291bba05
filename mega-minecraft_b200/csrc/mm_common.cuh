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

// resident CTAs per SM the register allocator is asked to allow (tuned on a B200)
#ifndef MMG_CAVES_MINBLOCKS
#define MMG_CAVES_MINBLOCKS 10
#endif
#ifndef MMG_ROCK_MINBLOCKS
#define MMG_ROCK_MINBLOCKS 8
#endif

}  // namespace mmg
