// Shim: replaces /root/reference/src/rendering/optixRenderer.hpp (pulled in by
// src/terrain/terrain.hpp:15) so the chunk-generation translation unit compiles headless.
// terrain.hpp only needs the class name.
#pragma once
class OptixRenderer;
