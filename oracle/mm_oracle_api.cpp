// TEST INFRASTRUCTURE (CPU oracle) - never linked into the product.
//
// C entry points of the CPU restatement, mirroring include/mmgen.h's batch operators so that
// tests/ can run the same inputs through both. Also used by bench.py's cpu_baseline / reference
// arm (the only places besides tests/ and smoke() allowed to touch oracle/).
// Threading: chunks are independent in every stage except erosion; `nthreads` std::threads split
// the chunk list.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

#include "mm_surface.h"

namespace {
template <class F>
void parallel_for(int n, int nthreads, F f)
{
    nthreads = std::max(1, std::min(nthreads, n));
    if (nthreads == 1)
    {
        for (int i = 0; i < n; ++i) f(i);
        return;
    }
    std::vector<std::thread> ts;
    for (int t = 0; t < nthreads; ++t)
        ts.emplace_back([=]() {
            for (int i = t; i < n; i += nthreads) f(i);
        });
    for (auto& t : ts) t.join();
}
}  // namespace

extern "C" {

// Chunk::generateHeightfields (chunk.cu:150-229)
void mmo_heightfields(int n, const int32_t* origins, float* out_h, float* out_w, int nthreads)
{
    parallel_for(n, nthreads, [&](int c) {
        const int ox = origins[2 * c], oz = origins[2 * c + 1];
        for (int z = 0; z < 16; ++z)
            for (int x = 0; x < 16; ++x)
            {
                const int idx = x + 16 * z;
                out_h[(size_t)c * 256 + idx] =
                    mmo::surface_column(ox + x, oz + z, out_w + (size_t)c * (mmo::NUM_BIOMES * 256) + idx, 256);
            }
    });
}

// unit probes used by tests
float mmo_sinf(float x) { return mmo::dm_sinf(x); }
float mmo_cosf(float x) { return mmo::dm_cosf(x); }
float mmo_powf(float a, float b) { return mmo::dm_powf(a, b); }
float mmo_simplex2(float x, float y) { return mmo::simplex2(x, y); }
float mmo_simplex3(float x, float y, float z) { return mmo::simplex3(x, y, z); }
uint32_t mmo_hash(uint32_t a) { return mmo::hash_u32(a); }
float mmo_rng3_u01(int x, int y, int z, int ndraw)
{
    mmo::Minstd r = mmo::make_rng3(x, y, z);
    float v = 0;
    for (int i = 0; i < ndraw; ++i) v = r.u01();
    return v;
}

}  // extern "C"
