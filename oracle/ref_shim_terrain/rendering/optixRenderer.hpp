// TEST INFRASTRUCTURE. Shim that replaces /root/reference/src/rendering/optixRenderer.hpp (included by terrain.hpp:15) when
// the UNMODIFIED terrain.cpp is compiled headless for oracle/_ref/libmmref_terrain.so: the three calls Terrain::tick makes
// (terrain.cpp:602, 654, 667) are counted instead of building OptiX acceleration structures.
#pragma once
#include <vector>
class Chunk;
class OptixRenderer
{
public:
    int destroyed = 0, built = 0, rootBuilds = 0;
    void destroyChunk(const Chunk*) { ++destroyed; }
    std::vector<const Chunk*> builtChunks;
    void buildChunkAccel(const Chunk* c) { ++built; builtChunks.push_back(c); }
    void buildRootAccel() { ++rootBuilds; }
};
