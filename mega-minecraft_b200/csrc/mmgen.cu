// C ABI (include/mmgen.h) and the device-resident world of the B200-native generation path.
// Host side of the boundary that Terrain::tick drives in the reference
// (/root/reference/src/terrain/terrain.cpp:587-960 calling /root/reference/src/terrain/chunk.hpp:100-172).
#include "../../include/mmgen.h"

#include <algorithm>
#include <cstring>
#include <mutex>
#include <vector>

#include "mm_common.cuh"
#include "mm_stage1.cuh"

namespace mmg {

thread_local std::string g_lastError;
uint64_t g_launchCount = 0;
static bool g_ready = false;
static int g_device = -1;

// grow-only device scratch for the batch operators (the reference's Terrain owns fixed staging
// buffers sized for one tick, terrain.cpp:111-185; here the callee owns them)
struct Scratch
{
    void* ptr = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes)
    {
        if (bytes <= cap) return 0;
        if (ptr) cudaFree(ptr);
        ptr = nullptr;
        cap = 0;
        MMG_CUDA(cudaMalloc(&ptr, bytes));
        cap = bytes;
        return 0;
    }
};
static Scratch g_scratch[8];
static cudaStream_t g_stream = nullptr;

static int requireReady()
{
    if (!g_ready)
    {
        g_lastError = "mmgen_init() has not succeeded: no CUDA device bound (there is no CPU fallback)";
        return 1;
    }
    return 0;
}

}  // namespace mmg

using namespace mmg;

struct MmgenWorld
{
    int cx0 = 0, cz0 = 0, nx = 0, nz = 0, n = 0;
    int2* d_origins = nullptr;
    float* d_height = nullptr;
    float* d_weights = nullptr;
    std::vector<uint8_t> stage;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[8] = {};
    float stageMs[7] = {0};
};

extern "C" {

const char* mmgen_last_error(void) { return g_lastError.c_str(); }
uint64_t mmgen_launch_count(void) { return g_launchCount; }

int mmgen_init(int device)
{
    int count = 0;
    cudaError_t err = cudaGetDeviceCount(&count);
    if (err != cudaSuccess || count == 0)
    {
        g_lastError = std::string("mmgen_init: no CUDA device available (") + cudaGetErrorString(err) +
                      "); this library has no CPU fallback";
        return 1;
    }
    MMG_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    MMG_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
    {
        g_lastError = "mmgen_init: kernels are built for sm_100a only";
        return 1;
    }
    if (!g_stream) MMG_CUDA(cudaStreamCreateWithFlags(&g_stream, cudaStreamNonBlocking));
    g_device = device;
    g_ready = true;
    return 0;
}

int mmgen_shutdown(void)
{
    if (!g_ready) return 0;
    for (auto& s : g_scratch)
    {
        if (s.ptr) cudaFree(s.ptr);
        s.ptr = nullptr;
        s.cap = 0;
    }
    if (g_stream) cudaStreamDestroy(g_stream);
    g_stream = nullptr;
    g_ready = false;
    return 0;
}

int mmgen_heightfields(int n, const int32_t* origins, float* out_heightfield, float* out_biomeWeights)
{
    if (requireReady()) return 1;
    if (n <= 0) return 0;
    if (g_scratch[0].ensure((size_t)n * sizeof(int2))) return 1;
    if (g_scratch[1].ensure((size_t)n * 256 * sizeof(float))) return 1;
    if (g_scratch[2].ensure((size_t)n * NUM_BIOMES * 256 * sizeof(float))) return 1;
    int2* d_o = (int2*)g_scratch[0].ptr;
    float* d_h = (float*)g_scratch[1].ptr;
    float* d_w = (float*)g_scratch[2].ptr;
    MMG_CUDA(cudaMemcpyAsync(d_o, origins, (size_t)n * sizeof(int2), cudaMemcpyHostToDevice, g_stream));
    MMG_LAUNCH(k_heightfield, n, 256, 0, g_stream, d_o, d_h, d_w);
    if (out_heightfield) MMG_CUDA(cudaMemcpyAsync(out_heightfield, d_h, (size_t)n * 256 * sizeof(float), cudaMemcpyDeviceToHost, g_stream));
    if (out_biomeWeights) MMG_CUDA(cudaMemcpyAsync(out_biomeWeights, d_w, (size_t)n * NUM_BIOMES * 256 * sizeof(float), cudaMemcpyDeviceToHost, g_stream));
    MMG_CUDA(cudaStreamSynchronize(g_stream));
    return 0;
}

// ------------------------------------------------------------------ world
int mmgen_world_create(int cx0, int cz0, int nx, int nz, MmgenWorld** out)
{
    if (requireReady()) return 1;
    if (nx <= 0 || nz <= 0 || !out)
    {
        g_lastError = "mmgen_world_create: bad arguments";
        return 1;
    }
    MmgenWorld* w = new MmgenWorld();
    w->cx0 = cx0; w->cz0 = cz0; w->nx = nx; w->nz = nz; w->n = nx * nz;
    w->stage.assign(w->n, 0);
    MMG_CUDA(cudaStreamCreateWithFlags(&w->stream, cudaStreamNonBlocking));
    for (auto& e : w->ev) MMG_CUDA(cudaEventCreate(&e));
    std::vector<int2> origins(w->n);
    for (int z = 0; z < nz; ++z)
        for (int x = 0; x < nx; ++x) origins[z * nx + x] = make_int2((cx0 + x) * 16, (cz0 + z) * 16);
    MMG_CUDA(cudaMalloc(&w->d_origins, (size_t)w->n * sizeof(int2)));
    MMG_CUDA(cudaMemcpy(w->d_origins, origins.data(), (size_t)w->n * sizeof(int2), cudaMemcpyHostToDevice));
    *out = w;
    return 0;
}

int mmgen_world_destroy(MmgenWorld* w)
{
    if (!w) return 0;
    cudaFree(w->d_origins);
    cudaFree(w->d_height);
    cudaFree(w->d_weights);
    for (auto& e : w->ev) if (e) cudaEventDestroy(e);
    if (w->stream) cudaStreamDestroy(w->stream);
    delete w;
    return 0;
}

int mmgen_world_generate(MmgenWorld* w, int stageMask)
{
    if (requireReady()) return 1;
    if (stageMask & MMGEN_STAGE_HEIGHTFIELD)
    {
        if (!w->d_height) MMG_CUDA(cudaMalloc(&w->d_height, (size_t)w->n * 256 * sizeof(float)));
        if (!w->d_weights) MMG_CUDA(cudaMalloc(&w->d_weights, (size_t)w->n * NUM_BIOMES * 256 * sizeof(float)));
        MMG_CUDA(cudaEventRecord(w->ev[0], w->stream));
        MMG_LAUNCH(k_heightfield, w->n, 256, 0, w->stream, w->d_origins, w->d_height, w->d_weights);
        MMG_CUDA(cudaEventRecord(w->ev[1], w->stream));
        for (auto& s : w->stage) s = std::max<uint8_t>(s, 1);
    }
    return 0;
}

int mmgen_world_sync(MmgenWorld* w)
{
    MMG_CUDA(cudaStreamSynchronize(w->stream));
    return 0;
}

int mmgen_world_stages(MmgenWorld* w, uint8_t* out)
{
    std::memcpy(out, w->stage.data(), w->n);
    return 0;
}

int mmgen_world_stage_ms(MmgenWorld* w, float* out7)
{
    MMG_CUDA(cudaStreamSynchronize(w->stream));
    for (int s = 0; s < 7; ++s) out7[s] = 0.f;
    MMG_CUDA(cudaEventElapsedTime(&out7[1], w->ev[0], w->ev[1]));
    return 0;
}

int mmgen_world_download(MmgenWorld* w, float* heightfield, float* biomeWeights, float* layers,
                         MmgenCaveLayer* caveLayers, uint8_t* blocks)
{
    MMG_CUDA(cudaStreamSynchronize(w->stream));
    if (heightfield && w->d_height) MMG_CUDA(cudaMemcpy(heightfield, w->d_height, (size_t)w->n * 256 * sizeof(float), cudaMemcpyDeviceToHost));
    if (biomeWeights && w->d_weights) MMG_CUDA(cudaMemcpy(biomeWeights, w->d_weights, (size_t)w->n * NUM_BIOMES * 256 * sizeof(float), cudaMemcpyDeviceToHost));
    (void)layers; (void)caveLayers; (void)blocks;
    return 0;
}

}  // extern "C"
