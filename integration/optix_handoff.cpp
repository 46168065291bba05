// Reference-side binding of the mesh hand-off: how OptixRenderer::buildChunkAccel
// (/root/reference/src/rendering/optixRenderer.cpp:223-368) takes a chunk mesh that ALREADY lives in device memory
// (mmgen_world_mesh + mmgen_world_mesh_gas_inputs, include/mmgen.h) instead of uploading Chunk::verts / Chunk::idx
// (initFromVector, optixRenderer.cpp:229-230).
//
// A maintainer replaces the first lines of buildChunkAccel by fillTriangleArray(gas, triangleInput.triangleArray) and drops
// the two CUBuffers; everything after it (SBT record, optixAccelComputeMemoryUsage, optixAccelBuild, compaction) is unchanged,
// and the whole tick's chunks can go into ONE optixAccelBuild call as an array of build inputs.
//
// The OptiX SDK is not part of this repository's environment, so the triangle-array descriptor is mirrored here with the
// fields buildChunkAccel sets (optixRenderer.cpp:267-285, OptiX 7 names); with the SDK present the template parameter is
// OptixBuildInputTriangleArray itself. What IS checked at compile time (make -C oracle adapter) is the part that can go wrong
// silently: the vertex / index layouts of the reference (rendering/structs.hpp:25-31) against the library's.
#include <cstddef>
#include <cstdint>

#include "rendering/structs.hpp"
#include "mmgen.h"

static_assert(sizeof(Vertex) == sizeof(MmgenVertex) && sizeof(Vertex) == 40, "Vertex layout (structs.hpp:25-31)");
static_assert(offsetof(Vertex, pos) == offsetof(MmgenVertex, pos) && offsetof(Vertex, nor) == offsetof(MmgenVertex, nor) &&
              offsetof(Vertex, uv) == offsetof(MmgenVertex, uv) && offsetof(Vertex, m) == offsetof(MmgenVertex, m), "Vertex field offsets");
static_assert(sizeof(glm::uvec3) == 12, "index triplets are unsigned int3 (optixRenderer.cpp:276-277)");
static_assert(sizeof(Mats) == sizeof(uint64_t), "Mats is a size_t enum (structs.hpp:7)");

// the fields of OptixBuildInputTriangleArray that buildChunkAccel fills
struct TriangleArrayFields
{
    const uint64_t* vertexBuffers;      // CUdeviceptr*
    unsigned int numVertices;
    int vertexFormat;                   // OPTIX_VERTEX_FORMAT_FLOAT3
    unsigned int vertexStrideInBytes;
    uint64_t indexBuffer;               // CUdeviceptr
    unsigned int numIndexTriplets;
    int indexFormat;                    // OPTIX_INDICES_FORMAT_UNSIGNED_INT3
    unsigned int indexStrideInBytes;
    const unsigned int* flags;
    unsigned int numSbtRecords;
};
constexpr int kVertexFormatFloat3 = 0x2121, kIndicesFormatUnsignedInt3 = 0x2103;      // optix_types.h (OptiX 7)

template <typename TriangleArray>
void fillTriangleArray(const MmgenGasInput& gas, TriangleArray& t, const unsigned int* flags)
{
    t.vertexFormat = kVertexFormatFloat3;
    t.vertexStrideInBytes = gas.vertexStrideInBytes;      // sizeof(Vertex)
    t.numVertices = gas.numVertices;
    t.vertexBuffers = &gas.vertexBuffer;                   // device memory of the mesh arena: no initFromVector
    t.indexFormat = kIndicesFormatUnsignedInt3;
    t.indexStrideInBytes = gas.indexStrideInBytes;        // sizeof(glm::uvec3)
    t.numIndexTriplets = gas.numIndexTriplets;
    t.indexBuffer = gas.indexBuffer;
    t.flags = flags;
    t.numSbtRecords = 1;
}

// keeps the template instantiated (and therefore compiled) in the adapter build
void mmgenFillTriangleArrayForTest(const MmgenGasInput& gas, TriangleArrayFields& t, const unsigned int* flags) { fillTriangleArray(gas, t, flags); }
