// Shim: stands in for the OptiX SDK header the reference's common headers include
// (/root/reference/src/util/common.h:4, src/cuda/cudaUtils.hpp:3). Nothing on the
// chunk-generation path calls OptiX; only the names below must exist to compile.
#pragma once
typedef unsigned long long OptixTraversableHandle;
typedef int OptixResult;
#define OPTIX_SUCCESS 0
static inline const char* optixGetErrorName(OptixResult) { return "optix-shim"; }
static inline const char* optixGetErrorString(OptixResult) { return "optix-shim"; }
