// C ABI (include/mmgen.h) and the device-resident world of the B200-native generation path.
// Host side of the boundary that Terrain::tick drives in the reference
// (/root/reference/src/terrain/terrain.cpp:587-960 calling /root/reference/src/terrain/chunk.hpp:100-172).
#include "../../include/mmgen.h"

#include <algorithm>
#include <cstring>
#include <deque>
#include <queue>
#include <mutex>
#include <vector>

#include "mm_common.cuh"
#include "mm_stage1.cuh"
#include "mm_stage23.cuh"
#include "mm_stage4.cuh"
#include "mm_stage56.cuh"
#include "mm_mesh.cuh"
#include "mm_codec.cuh"

namespace mmg {

thread_local std::string g_lastError;
uint64_t g_launchCount = 0;
int g_numSMs = 148;
static bool g_ready = false;
// The batch operators share one stream, one set of scratch buffers and the kernel timer (like the reference, whose staging
// buffers are file-static, terrain.cpp:131-152): calls into them are serialised by this mutex. World / stream objects own
// their buffers; one world must not be driven from two threads at once.
static std::mutex g_batchMutex;
#define MMG_BATCH_LOCK() std::lock_guard<std::mutex> mmg_batch_lock_(g_batchMutex)

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
static Scratch g_scratch[17];   // [0..7] stage inputs/outputs, [8..11] fill extras, [12..13] cave-biome queue, [14..15] placement Prep records, [16] rock-voxel queue
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


// ---- per-kernel device timing (mmgen_kernel_timing): CUDA event pairs around the launches of the hot kernels,
// on the stream they are launched on. Off by default; bench.py switches it on for its timed region.
enum KernelSlot { K_HEIGHTFIELD, K_LAYERS, K_ERODE_SWEEPS, K_CAVE_COLUMNS, K_CAVES, K_CAVE_BIOMES, K_PLACEMENTS, K_GATHER, K_FILL_TERRAIN,
                  K_FILL_ROCK, K_FILL_LUSH, K_PREPARE, K_FILL_FEATURES, K_DECORATORS, K_NUM };
static const char* const kKernelNames[K_NUM] = {"k_heightfield", "k_layers", "k_erode_sweep", "k_cave_columns", "k_caves", "k_cave_biomes",
                                                "k_feature_placements", "k_gather_features", "k_fill_terrain", "k_fill_rock", "k_fill_lush",
                                                "k_prepare_placements", "k_fill_features", "k_decorators"};
struct KernelTimer
{
    bool on = false;
    std::vector<cudaEvent_t> pool;      // pairs: [2i] before, [2i+1] after
    std::vector<int> slotOf, countOf;
    size_t used = 0;
    void begin(int slot, cudaStream_t st, int launches)
    {
        if (!on) return;
        if (2 * used + 2 > pool.size())
        {
            cudaEvent_t a, b;
            if (cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess) { on = false; return; }
            pool.push_back(a); pool.push_back(b);
        }
        slotOf.resize(used + 1); countOf.resize(used + 1);
        slotOf[used] = slot; countOf[used] = launches;
        cudaEventRecord(pool[2 * used], st);
    }
    void end(cudaStream_t st)
    {
        if (!on) return;
        cudaEventRecord(pool[2 * used + 1], st);
        ++used;
    }
};
static KernelTimer g_kt;
#define MMG_TIMED(slot, stream, launches, stmt) do { g_kt.begin(slot, stream, launches); stmt; g_kt.end(stream); } while (0)

// dependent-FMA microbenchmark for the FP32 roofline denominator: 8 independent chains per thread, all lanes busy
__global__ void __launch_bounds__(256) k_fp32_peak(float* out, int iters)
{
    float a0 = threadIdx.x * 1e-3f, a1 = a0 + 1.f, a2 = a0 + 2.f, a3 = a0 + 3.f, a4 = a0 + 4.f, a5 = a0 + 5.f, a6 = a0 + 6.f, a7 = a0 + 7.f;
    const float m = 0.9999f, c = 1e-4f;
#pragma unroll 1
    for (int i = 0; i < iters; ++i)
    {
#pragma unroll
        for (int k = 0; k < 16; ++k)
        {
            a0 = fmaf(a0, m, c); a1 = fmaf(a1, m, c); a2 = fmaf(a2, m, c); a3 = fmaf(a3, m, c);
            a4 = fmaf(a4, m, c); a5 = fmaf(a5, m, c); a6 = fmaf(a6, m, c); a7 = fmaf(a7, m, c);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
}

}  // namespace mmg

using namespace mmg;

struct CodecState;
static void codecFree(CodecState* c);
struct EncodedOut { uint8_t* buf; size_t cap; uint64_t* index; size_t bytes; };      // mmgen_world_generate_to_host_encoded

struct MmgenWorld
{
    CodecState* codec = nullptr;         // encoder scratch + arena (mm_codec.inl), created on first use
    int cx0 = 0, cz0 = 0, nx = 0, nz = 0, n = 0;
    int2* d_origins = nullptr;
    std::vector<int2> h_origins;
    float* d_height = nullptr;
    float* d_weights = nullptr;
    float* d_layers = nullptr;        // S2 output (never modified afterwards: erosion pads read it)
    float* d_eroded = nullptr;        // S3 output: full 20-layer set of eroded chunks
    float* d_zone = nullptr;          // kZoneBatch zones x (9 planes + 1 scratch plane + 2 accum planes)
    int2* d_zoneCorners = nullptr;
    int* d_flags = nullptr;           // one "changed" flag per sweep of a batch
    int* d_list = nullptr;            // chunk index list of the stage call being queued: a slice of d_listRing
    int* d_listRing = nullptr;        // kListSlots slices of n ints; a stage call takes the next one, so that the calls of one generate /
    int listSlot = 0;                 // one tick can be queued back to back without waiting for the previous stage's kernels
    CaveLayer* d_caves = nullptr;     // [chunk][256][32]
    CaveColumn* d_caveCols = nullptr; // per-column hoisted cave terms of one cave batch
    uint2* d_caveQueue = nullptr;     // cave-biome lookups of one cave batch
    int* d_caveCount = nullptr;
    FeaturePlacement* d_features = nullptr;          // own lists [chunk][kMaxOwnFeatures]
    CaveFeaturePlacement* d_caveFeatures = nullptr;  // own lists [chunk][kMaxOwnCaveFeatures]
    int* d_counts = nullptr;                         // [chunk][2]
    FeaturePlacement* d_gF = nullptr;                // gathered lists of one fill batch
    CaveFeaturePlacement* d_gCF = nullptr;
    GatherInfo* d_info = nullptr;
    Prep* d_prepF = nullptr;                         // per-placement culling records of one fill batch
    Prep* d_prepC = nullptr;
    uint2* d_lushQueue = nullptr;                    // voxels of one fill batch waiting for the lush-cave decision
    int* d_lushCount = nullptr;                      // [0] lush queue length, [2] near-rock, [3] bulk-rock queue lengths
    uint2* d_rockQueue = nullptr;                    // rock voxels of one fill batch waiting for getCaveBiome (k_fill_rock)
    size_t fillCap = 0;                              // chunks per fill batch the scratch above (d_gF .. d_rockQueue) is sized for
    // second set of gathered lists + Prep records: a call that fills several batches gathers and prepares batch b + 1 on the side
    // stream while batch b's terrain / rock passes run (allocated by the first such call)
    FeaturePlacement* d_gF2 = nullptr; CaveFeaturePlacement* d_gCF2 = nullptr; GatherInfo* d_info2 = nullptr; Prep* d_prepF2 = nullptr; Prep* d_prepC2 = nullptr;
    cudaEvent_t evGather[2] = {}, evScan[2] = {};    // per set: gathered + prepared (side stream); the placement scan is done with it (main stream)
    size_t fillHint = 0;                             // streaming sessions: the most chunks a tick can fill (0 = a batch world)
    bool reserving = false;                          // worldReserve: the stage runners allocate their buffers and return
    uint8_t* d_blocks = nullptr;                     // [chunk][16][16][384]
    // meshing (mmgen_world_mesh): arena of the last call
    MeshChunk* d_meshList = nullptr;
    int* d_meshColOff = nullptr;
    int* d_meshTotals = nullptr;
    long long* d_meshBase = nullptr;
    MeshVertex* d_meshVerts = nullptr;
    uint32_t* d_meshIdx = nullptr;
    size_t meshListCap = 0, meshVertCap = 0;
    std::vector<long long> h_meshBase;
    std::vector<int> h_meshTotals;
    std::vector<int32_t> h_meshCoords;   // (cx, cz) of the chunks of the last mmgen_world_mesh call
    float meshMs = 0.f;
    int erosionSweeps = 0;
    std::vector<uint8_t> stage;
    cudaStream_t stream = nullptr;
    cudaStream_t copyStream = nullptr;   // device->host block copies overlapped with the fill (generate_to_host)
    cudaStream_t sideStream = nullptr;   // layers + erosion of a full generate, while the caves (which only need stage 1) run on `stream`
    cudaStream_t fillStream = nullptr;   // terrain / rock / lush passes of the next fill batch under mmgen_set_fill_overlap (default priority)
    cudaEvent_t evSide[2] = {};          // fork / join of the side stream
    cudaEvent_t ev[14] = {};             // [2s-2, 2s-1] bracket stage s; [12, 13] bracket the whole generate
    cudaEvent_t evMesh[2] = {};          // bracket mmgen_world_mesh (its own pair: total_ms keeps the last generate's time)
    cudaEvent_t evBatch[2] = {};
    // target region (mmgen_world_create_for_region): only what filling it needs is computed
    bool hasTarget = false;
    int tx0 = 0, tz0 = 0, tnx = 0, tnz = 0;   // window-local chunk coordinates
    bool inTarget(int x, int z, int grow) const
    {
        return !hasTarget || (x >= tx0 - grow && x < tx0 + tnx + grow && z >= tz0 - grow && z < tz0 + tnz + grow);
    }
    // halo exchange (mmgen_world_set_exchange_region): chunks of the global region [gx0, gx0+gnx) x [gz0, gz0+gnz) (window-local)
    // outside the own target tile belong to another world, which computes their placements and sends them
    bool hasGlobal = false;
    int gx0 = 0, gz0 = 0, gnx = 0, gnz = 0;
    // does THIS world have to compute the placements (S4 + S5a) of window chunk (x, z)?
    bool ownsPlacements(int x, int z) const
    {
        if (!inTarget(x, z, 3)) return false;
        if (!hasGlobal || inTarget(x, z, 0)) return true;
        return !(x >= gx0 && x < gx0 + gnx && z >= gz0 && z < gz0 + gnz);
    }
    // scratch of the exchange: chunk indices and byte offsets of one message
    int* d_xIdx = nullptr;
    long long* d_xOff = nullptr;
    size_t xCap = 0;
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
    cudaFuncAttributes fattr;
    MMG_CUDA(cudaGetDeviceProperties(&prop, device));
    // the library carries sm_100a SASS only (arch-specific: no PTX fallback for other Blackwell parts)
    if (prop.major != 10 || prop.minor != 0 || cudaFuncGetAttributes(&fattr, k_init_noise_tables) != cudaSuccess)
    {
        cudaGetLastError();
        g_lastError = "mmgen_init: device " + std::to_string(device) + " is sm_" + std::to_string(prop.major) + std::to_string(prop.minor) +
                      "; the kernels of this library are built for sm_100a (B200) only and there is no fallback";
        return 1;
    }
    g_numSMs = prop.multiProcessorCount;
    if (!g_stream) MMG_CUDA(cudaStreamCreateWithFlags(&g_stream, cudaStreamNonBlocking));
    // k_fill_features: 40 KB static + the 10 KB noise tables exceed the 48 KB default
    MMG_CUDA(cudaFuncSetAttribute(k_fill_features, cudaFuncAttributeMaxDynamicSharedMemorySize, kNoiseSmemBytes));
    MMG_CUDA(cudaFuncSetAttribute(k_mesh_emit, cudaFuncAttributeMaxDynamicSharedMemorySize, kMeshStripBytes));
    MMG_LAUNCH(k_init_noise_tables, 3, 256, 0, g_stream);     // simplex lattice tables (mm_arith.cuh)
    MMG_CUDA(cudaStreamSynchronize(g_stream));
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
    MMG_BATCH_LOCK();
    if (n <= 0) return 0;
    if (g_scratch[0].ensure((size_t)n * sizeof(int2))) return 1;
    if (g_scratch[1].ensure((size_t)n * 256 * sizeof(float))) return 1;
    if (g_scratch[2].ensure((size_t)n * NUM_BIOMES * 256 * sizeof(float))) return 1;
    int2* d_o = (int2*)g_scratch[0].ptr;
    float* d_h = (float*)g_scratch[1].ptr;
    float* d_w = (float*)g_scratch[2].ptr;
    MMG_CUDA(cudaMemcpyAsync(d_o, origins, (size_t)n * sizeof(int2), cudaMemcpyHostToDevice, g_stream));
    MMG_LAUNCH(k_heightfield, n, 256, kNoiseSmemBytes, g_stream, (const int*)nullptr, (const int2*)d_o, d_h, d_w);
    if (out_heightfield) MMG_CUDA(cudaMemcpyAsync(out_heightfield, d_h, (size_t)n * 256 * sizeof(float), cudaMemcpyDeviceToHost, g_stream));
    if (out_biomeWeights) MMG_CUDA(cudaMemcpyAsync(out_biomeWeights, d_w, (size_t)n * NUM_BIOMES * 256 * sizeof(float), cudaMemcpyDeviceToHost, g_stream));
    MMG_CUDA(cudaStreamSynchronize(g_stream));
    return 0;
}

// ------------------------------------------------------------------ erosion driver
// zones: nZones x 12 planes of 384x384 floats (kZonePlanes). Runs Chunk::erodeZone's loop
// (chunk.cu:682-705) for all zones of the batch in lockstep, with the convergence test on the device.
static const float kTanRepose[NUM_ERODED] = {1.42814791f, 0.839099586f, 1.0f, 0.839099586f,
                                             0.577350318f, 0.700207531f, 2.14450693f, 1.0f};
#ifndef MMG_SWEEP_GROUP
#define MMG_SWEEP_GROUP 8
#endif
constexpr int kSweepGroup = MMG_SWEEP_GROUP;     // sweeps between two convergence polls (even)
constexpr int kZoneBatch = kMaxZoneBatch;     // zones per launch: 32 x 4 live planes x 590 KB = 75 MB, L2-resident

static int erodeZonesDevice(float* d_zones, int nZones, int* d_flags, cudaStream_t stream, int* sweepsOut)
{
    int accIn = 10, accOut = 11;     // plane 10 was zeroed by the gather (or by the caller)
    int sweeps = 0;
    int h_flags[kSweepGroup];
    int* d_zoneChanged = d_flags + kSweepGroup;      // [3][kMaxZoneBatch][kZoneTiles], see k_erode_sweep
    MMG_CUDA(cudaMemsetAsync(d_zoneChanged, 0, 3 * kMaxZoneBatch * kZoneTiles * sizeof(int), stream));
    for (int layer = NUM_ERODED - 1; layer >= 0; --layer)
    {
        int pIn = layer, pOut = 9;
        int layerSweeps = 0;
        bool first = true, converged = false;
        while (!converged)
        {
            MMG_CUDA(cudaMemsetAsync(d_flags, 0, kSweepGroup * sizeof(int), stream));
            g_kt.begin(K_ERODE_SWEEPS, stream, kSweepGroup);
            for (int b = 0; b < kSweepGroup; ++b)
            {
                MMG_LAUNCH(k_erode_sweep, dim3(12, 12, nZones), dim3(32, kErodeRows), 0, stream, d_zones, pIn, pOut, layer + 1, accIn, accOut,
                           kTanRepose[layer], first ? 1 : 0, d_flags + b, d_zoneChanged, sweeps, layerSweeps < 2 ? 1 : 0);
                ++layerSweeps;
                std::swap(pIn, pOut);
                std::swap(accIn, accOut);
                first = false;
                ++sweeps;
            }
            g_kt.end(stream);
            MMG_CUDA(cudaMemcpyAsync(h_flags, d_flags, sizeof(h_flags), cudaMemcpyDeviceToHost, stream));
            MMG_CUDA(cudaStreamSynchronize(stream));
            converged = (h_flags[kSweepGroup - 1] == 0);
        }
        // an even number of sweeps per group: the result is back in the layer's own plane
        static_assert(kSweepGroup % 2 == 0, "ping-pong must end in the layer plane");
    }
    if (sweepsOut) *sweepsOut = sweeps;
    return 0;
}

extern "C" int mmgen_layers(int n, const int32_t* origins, const float* heightfield18, const float* biomeWeights, float* out_layers)
{
    if (requireReady()) return 1;
    MMG_BATCH_LOCK();
    if (n <= 0) return 0;
    if (g_scratch[0].ensure((size_t)n * sizeof(int2))) return 1;
    if (g_scratch[1].ensure((size_t)n * 324 * sizeof(float))) return 1;
    if (g_scratch[2].ensure((size_t)n * NUM_BIOMES * 256 * sizeof(float))) return 1;
    if (g_scratch[3].ensure((size_t)n * NUM_MATERIALS * 256 * sizeof(float))) return 1;
    MMG_CUDA(cudaMemcpyAsync(g_scratch[0].ptr, origins, (size_t)n * sizeof(int2), cudaMemcpyHostToDevice, g_stream));
    MMG_CUDA(cudaMemcpyAsync(g_scratch[1].ptr, heightfield18, (size_t)n * 324 * sizeof(float), cudaMemcpyHostToDevice, g_stream));
    MMG_CUDA(cudaMemcpyAsync(g_scratch[2].ptr, biomeWeights, (size_t)n * NUM_BIOMES * 256 * sizeof(float), cudaMemcpyHostToDevice, g_stream));
    MMG_LAUNCH(k_layers<false>, n, 256, kNoiseSmemBytes, g_stream, (const int*)nullptr, (const int2*)g_scratch[0].ptr,
               (const float*)g_scratch[1].ptr, (const float*)g_scratch[2].ptr, (float*)g_scratch[3].ptr, 0);
    MMG_CUDA(cudaMemcpyAsync(out_layers, g_scratch[3].ptr, (size_t)n * NUM_MATERIALS * 256 * sizeof(float), cudaMemcpyDeviceToHost, g_stream));
    MMG_CUDA(cudaStreamSynchronize(g_stream));
    return 0;
}

extern "C" int mmgen_erode_zone(const float* gathered, float* out_eroded, int* out_sweeps)
{
    if (requireReady()) return 1;
    MMG_BATCH_LOCK();
    const size_t P = kErosionCols;
    if (g_scratch[4].ensure(kZonePlanes * P * sizeof(float))) return 1;
    if (g_scratch[5].ensure((kSweepGroup + 3 * kMaxZoneBatch * kZoneTiles) * sizeof(int))) return 1;
    float* d_zone = (float*)g_scratch[4].ptr;
    MMG_CUDA(cudaMemcpyAsync(d_zone, gathered, 9 * P * sizeof(float), cudaMemcpyHostToDevice, g_stream));
    MMG_CUDA(cudaMemsetAsync(d_zone + 10 * P, 0, P * sizeof(float), g_stream));
    if (erodeZonesDevice(d_zone, 1, (int*)g_scratch[5].ptr, g_stream, out_sweeps)) return 1;
    MMG_CUDA(cudaMemcpyAsync(out_eroded, d_zone, 8 * P * sizeof(float), cudaMemcpyDeviceToHost, g_stream));
    MMG_CUDA(cudaStreamSynchronize(g_stream));
    return 0;
}

// (8192 / 16384 chunks per group measured: 453.3 / 452.4 ms per 256x256 world against 454.0 - not worth 2x / 4x the queue memory)
#ifndef MMG_CAVE_BATCH
#define MMG_CAVE_BATCH 4096
#endif
constexpr int kCaveBatch = MMG_CAVE_BATCH;             // chunks per cave launch group
constexpr int kCaveBiomeQueueCap = kCaveBatch * 256 * 8;  // cave-biome lookups queued per group (avg ~5.4 per column)

// the kernel sequence of Chunk::generateCaves for m chunks (d_list: chunk indices or null; column terms indexed by batch position)
static int launchCaves(int m, const int* d_list, const int2* d_origins, const float* d_height, const float* d_weights, CaveColumn* d_cols,
                       CaveLayer* d_caves, uint2* d_queue, int* d_count, cudaStream_t stream)
{
    for (int c0 = 0; c0 < m; c0 += kCaveBatch)
    {
        const int mb = std::min(kCaveBatch, m - c0);
        const int* dl = d_list ? d_list + c0 : nullptr;
        // without a list, chunk index == batch position: offset every per-chunk pointer instead
        const size_t off = d_list ? 0 : (size_t)c0;
        MMG_CUDA(cudaMemsetAsync(d_count, 0, sizeof(int), stream));
        MMG_TIMED(K_CAVE_COLUMNS, stream, 1, MMG_LAUNCH(k_cave_columns, mb, 256, kNoiseSmemBytes, stream, dl, d_origins + off,
                                                        d_weights + off * NUM_BIOMES * 256, d_cols));
        MMG_TIMED(K_CAVES, stream, 1, MMG_LAUNCH(k_caves, mb * (256 / kCaveCols), kCaveThreads, kNoiseSmemBytes, stream, dl, d_origins + off, d_height + off * 256,
                                                 (const CaveColumn*)d_cols, d_caves + off * 256 * MAX_CAVE_LAYERS, d_queue, d_count, kCaveBiomeQueueCap));
        MMG_TIMED(K_CAVE_BIOMES, stream, 1, MMG_LAUNCH(k_cave_biomes, kNumSMs * 16, 128, kNoiseSmemBytes, stream, d_origins + off, d_height + off * 256,
                                                       (const uint2*)d_queue, (const int*)d_count, kCaveBiomeQueueCap, d_caves + off * 256 * MAX_CAVE_LAYERS));
    }
    return 0;
}

extern "C" int mmgen_caves(int n, const int32_t* origins, const float* heightfield, const float* biomeWeights,
                           MmgenCaveLayer* out_caveLayers)
{
    if (requireReady()) return 1;
    MMG_BATCH_LOCK();
    if (n <= 0) return 0;
    const size_t clBytes = (size_t)n * 256 * MAX_CAVE_LAYERS * sizeof(CaveLayer);
    if (g_scratch[0].ensure((size_t)n * sizeof(int2))) return 1;
    if (g_scratch[1].ensure((size_t)n * 256 * sizeof(float))) return 1;
    if (g_scratch[2].ensure((size_t)n * NUM_BIOMES * 256 * sizeof(float))) return 1;
    if (g_scratch[3].ensure(clBytes)) return 1;
    if (g_scratch[6].ensure((size_t)std::min(n, kCaveBatch) * 256 * sizeof(CaveColumn))) return 1;
    MMG_CUDA(cudaMemcpyAsync(g_scratch[0].ptr, origins, (size_t)n * sizeof(int2), cudaMemcpyHostToDevice, g_stream));
    MMG_CUDA(cudaMemcpyAsync(g_scratch[1].ptr, heightfield, (size_t)n * 256 * sizeof(float), cudaMemcpyHostToDevice, g_stream));
    MMG_CUDA(cudaMemcpyAsync(g_scratch[2].ptr, biomeWeights, (size_t)n * NUM_BIOMES * 256 * sizeof(float), cudaMemcpyHostToDevice, g_stream));
    Scratch* Q = g_scratch + 12;
    if (Q[0].ensure((size_t)kCaveBiomeQueueCap * sizeof(uint2)) || Q[1].ensure(sizeof(int))) return 1;
    if (launchCaves(n, nullptr, (const int2*)g_scratch[0].ptr, (const float*)g_scratch[1].ptr, (const float*)g_scratch[2].ptr,
                    (CaveColumn*)g_scratch[6].ptr, (CaveLayer*)g_scratch[3].ptr, (uint2*)Q[0].ptr, (int*)Q[1].ptr, g_stream))
        return 1;
    MMG_CUDA(cudaMemcpyAsync(out_caveLayers, g_scratch[3].ptr, clBytes, cudaMemcpyDeviceToHost, g_stream));
    MMG_CUDA(cudaStreamSynchronize(g_stream));
    return 0;
}

// chunks per fill batch: every fill kernel ends in a tail of half-empty SMs, and k_fill_features (12 CTAs per chunk, two per SM,
// uneven durations) has the longest one. Per 128x128 region (profiles/r02_fill_batch.txt): 256 chunks 146.9 ms, 512 133.3, 1024
// 126.8, 2048 123.5, 4096 121.9; 2048 keeps the last batch's host copy (the only one not overlapped) at 200 MB.
#ifndef MMG_FILL_BATCH
#define MMG_FILL_BATCH 2048
#endif
constexpr int kFillBatch = MMG_FILL_BATCH;
static bool g_serialStages = false;                    // mmgen_set_serial_stages: no overlap of layers + erosion with the caves (measurement knob)
static int g_rockQueuePerChunk = kRockQueuePerChunk;   // mmgen_set_rock_queue_per_chunk (tuning / test knob, <= kRockQueuePerChunk)
// mmgen_set_fill_overlap: 0 = a batch's terrain / rock / lush passes and its placement scan run one after the other on the world's
// stream; g > 0 = the terrain / rock / lush passes of batch b + 1 run on the side stream while the placement scan of batch b runs on
// the main stream (they touch different chunks' volumes), with k_fill_rock's persistent grid at g CTAs per SM
static int g_fillOverlap = 8;                          // 256x256 world on a B200: 454.2 ms off, 446.7 ms at 8 (profiles/r02_fill_overlap.txt)

// the kernel sequence of Chunk::fill for one batch of m chunks (lists indexed by batch position)
// the fill of one batch in three parts, so that the world path can run the placement preparation of the NEXT batch on its side
// stream while this batch's terrain / rock passes run: (1) terrain, rock, lush; (2) Prep records of the gathered lists;
// (3) the placement scan and the decorators
static int launchFillTerrain(int m, const int* d_list, const int2* d_origins, const float* d_height, const float* d_weights, const float* d_layers,
                             const CaveLayer* d_caves, uint8_t* d_blocks, uint2* d_rockQueue, uint2* d_lushQueue, int* d_counters, cudaStream_t stream,
                             int rockCtasPerSM = MMG_ROCK_MINBLOCKS)
{
    const int rockCap = (int)std::min<size_t>((size_t)m * g_rockQueuePerChunk, (size_t)kFillBatch * kRockQueuePerChunk);
    MMG_CUDA(cudaMemsetAsync(d_counters, 0, 4 * sizeof(int), stream));
    MMG_TIMED(K_FILL_TERRAIN, stream, 1, MMG_LAUNCH(k_fill_terrain, m * 16, kRowThreads, kNoiseSmemBytes, stream, d_list, d_origins, d_height,
                                                    d_weights, d_layers, d_caves, d_blocks, d_rockQueue, rockCap, d_counters));
    MMG_TIMED(K_FILL_ROCK, stream, 1, MMG_LAUNCH(k_fill_rock, kNumSMs * rockCtasPerSM, 128, kNoiseSmemBytes, stream, d_origins, d_height,
                                                 (const uint2*)d_rockQueue, rockCap, d_blocks, d_lushQueue, d_counters));
    MMG_TIMED(K_FILL_LUSH, stream, 1, MMG_LAUNCH(k_fill_lush, kNumSMs * 8, 128, kNoiseSmemBytes, stream, d_origins, (const uint2*)d_lushQueue,
                                                 (const int*)d_counters, d_blocks));
    return 0;
}
static int launchFillPrepare(int m, const int* d_list, const int2* d_origins, const FeaturePlacement* d_gF, const CaveFeaturePlacement* d_gCF,
                             GatherInfo* d_info, Prep* d_prepF, Prep* d_prepC, int strideF, int strideCF, cudaStream_t stream)
{
    MMG_TIMED(K_PREPARE, stream, 1, MMG_LAUNCH(k_prepare_placements, m, 256, 0, stream, d_list, d_origins, d_gF, d_gCF, d_info, strideF, strideCF,
                                               d_prepF, d_prepC));
    return 0;
}
static int launchFillFeatures(int m, const int* d_list, const int2* d_origins, const float* d_height, const float* d_weights, const CaveLayer* d_caves,
                              const FeaturePlacement* d_gF, const CaveFeaturePlacement* d_gCF, const GatherInfo* d_info, const Prep* d_prepF,
                              const Prep* d_prepC, int strideF, int strideCF, uint8_t* d_blocks, cudaStream_t stream)
{
    MMG_TIMED(K_FILL_FEATURES, stream, 1, MMG_LAUNCH(k_fill_features, m * 12, kFeatThreads, kNoiseSmemBytes, stream, d_list, d_origins, d_gF, d_gCF,
                                                     d_prepF, d_prepC, d_info, strideF, strideCF, d_blocks));
    MMG_TIMED(K_DECORATORS, stream, 1, MMG_LAUNCH(k_decorators, m, 256, 0, stream, d_list, m, d_origins, d_height, d_weights, d_caves, d_blocks));
    return 0;
}
static int launchFill(int m, const int* d_list, const int2* d_origins, const float* d_height, const float* d_weights, const float* d_layers,
                      const CaveLayer* d_caves, const FeaturePlacement* d_gF, const CaveFeaturePlacement* d_gCF, GatherInfo* d_info,
                      Prep* d_prepF, Prep* d_prepC, int strideF, int strideCF, uint8_t* d_blocks, uint2* d_rockQueue, uint2* d_lushQueue,
                      int* d_counters, cudaStream_t stream)
{
    return launchFillTerrain(m, d_list, d_origins, d_height, d_weights, d_layers, d_caves, d_blocks, d_rockQueue, d_lushQueue, d_counters, stream) ||
           launchFillPrepare(m, d_list, d_origins, d_gF, d_gCF, d_info, d_prepF, d_prepC, strideF, strideCF, stream) ||
           launchFillFeatures(m, d_list, d_origins, d_height, d_weights, d_caves, d_gF, d_gCF, d_info, d_prepF, d_prepC, strideF, strideCF, d_blocks, stream);
}   // chunks gathered + filled per launch group (bounds the gathered-list buffers)

extern "C" int mmgen_feature_placements(int n, const int32_t* origins, const float* heightfield, const float* biomeWeights,
                                        const float* layers, const MmgenCaveLayer* caveLayers, int maxPerChunk,
                                        MmgenFeaturePlacement* out_features, MmgenCaveFeaturePlacement* out_caveFeatures,
                                        int32_t* out_counts)
{
    if (requireReady()) return 1;
    MMG_BATCH_LOCK();
    if (n <= 0) return 0;
    const size_t clBytes = (size_t)n * 256 * MAX_CAVE_LAYERS * sizeof(CaveLayer);
    Scratch* S = g_scratch;
    if (S[0].ensure((size_t)n * sizeof(int2)) || S[1].ensure((size_t)n * 256 * 4) || S[2].ensure((size_t)n * NUM_BIOMES * 256 * 4) ||
        S[3].ensure(clBytes) || S[4].ensure((size_t)n * NUM_MATERIALS * 256 * 4) ||
        S[5].ensure((size_t)n * kMaxOwnFeatures * sizeof(FeaturePlacement)) ||
        S[6].ensure((size_t)n * kMaxOwnCaveFeatures * sizeof(CaveFeaturePlacement)) || S[7].ensure((size_t)n * 2 * sizeof(int)))
        return 1;
    MMG_CUDA(cudaMemcpyAsync(S[0].ptr, origins, (size_t)n * sizeof(int2), cudaMemcpyHostToDevice, g_stream));
    MMG_CUDA(cudaMemcpyAsync(S[1].ptr, heightfield, (size_t)n * 256 * 4, cudaMemcpyHostToDevice, g_stream));
    MMG_CUDA(cudaMemcpyAsync(S[2].ptr, biomeWeights, (size_t)n * NUM_BIOMES * 256 * 4, cudaMemcpyHostToDevice, g_stream));
    MMG_CUDA(cudaMemcpyAsync(S[3].ptr, caveLayers, clBytes, cudaMemcpyHostToDevice, g_stream));
    MMG_CUDA(cudaMemcpyAsync(S[4].ptr, layers, (size_t)n * NUM_MATERIALS * 256 * 4, cudaMemcpyHostToDevice, g_stream));
    MMG_LAUNCH(k_feature_placements, n, 256, 0, g_stream, (const int*)nullptr, (const int2*)S[0].ptr, (const float*)S[1].ptr,
               (const float*)S[2].ptr, (const float*)S[4].ptr, (const CaveLayer*)S[3].ptr, (FeaturePlacement*)S[5].ptr,
               (CaveFeaturePlacement*)S[6].ptr, (int*)S[7].ptr);
    std::vector<int> counts((size_t)n * 2);
    MMG_CUDA(cudaMemcpyAsync(counts.data(), S[7].ptr, counts.size() * sizeof(int), cudaMemcpyDeviceToHost, g_stream));
    MMG_CUDA(cudaStreamSynchronize(g_stream));
    for (int c = 0; c < n; ++c)
    {
        out_counts[2 * c] = counts[2 * c];
        out_counts[2 * c + 1] = counts[2 * c + 1];
        const int nf = std::min(counts[2 * c], maxPerChunk), nc = std::min(counts[2 * c + 1], maxPerChunk);
        if (nf) MMG_CUDA(cudaMemcpyAsync(out_features + (size_t)c * maxPerChunk, (FeaturePlacement*)S[5].ptr + (size_t)c * kMaxOwnFeatures,
                                         (size_t)nf * sizeof(FeaturePlacement), cudaMemcpyDeviceToHost, g_stream));
        if (nc) MMG_CUDA(cudaMemcpyAsync(out_caveFeatures + (size_t)c * maxPerChunk, (CaveFeaturePlacement*)S[6].ptr + (size_t)c * kMaxOwnCaveFeatures,
                                         (size_t)nc * sizeof(CaveFeaturePlacement), cudaMemcpyDeviceToHost, g_stream));
    }
    MMG_CUDA(cudaStreamSynchronize(g_stream));
    return 0;
}

extern "C" int mmgen_gather_offsets(int32_t* out49x2)
{
    static const int off[49][2] = {
        {0, 0}, {0, 1}, {1, 1}, {1, 0}, {1, -1}, {0, -1}, {-1, -1}, {-1, 0}, {-1, 1}, {2, 0}, {2, 1}, {2, 2}, {1, 2}, {0, 2},
        {-1, 2}, {-2, 2}, {-2, 1}, {-2, 0}, {-2, -1}, {-2, -2}, {-1, -2}, {0, -2}, {1, -2}, {2, -2}, {2, -1},
        {-3, -3}, {-2, -3}, {-1, -3}, {0, -3}, {1, -3}, {2, -3}, {3, -3}, {3, -2}, {3, -1}, {3, 0}, {3, 1}, {3, 2}, {3, 3},
        {2, 3}, {1, 3}, {0, 3}, {-1, 3}, {-2, 3}, {-3, 3}, {-3, 2}, {-3, 1}, {-3, 0}, {-3, -1}, {-3, -2}};      // == c_gatherOffsets
    std::memcpy(out49x2, off, sizeof(off));
    return 0;
}

extern "C" int mmgen_gather_features(int n, const int32_t* neighbours, int m, const MmgenFeaturePlacement* features, int featureStride,
                                     const MmgenCaveFeaturePlacement* caveFeatures, int caveFeatureStride, const int32_t* counts,
                                     MmgenFeaturePlacement* out_features, MmgenCaveFeaturePlacement* out_caveFeatures, int32_t* out_counts)
{
    if (requireReady()) return 1;
    MMG_BATCH_LOCK();
    if (n <= 0) return 0;
    if (m <= 0 || featureStride <= 0 || caveFeatureStride <= 0 || !neighbours || !features || !caveFeatures || !counts || !out_counts)
    {
        g_lastError = "mmgen_gather_features: bad arguments";
        return 1;
    }
    for (int i = 0; i < n * 49; ++i)
        if (neighbours[i] >= m)
        {
            g_lastError = "mmgen_gather_features: neighbour index " + std::to_string(neighbours[i]) + " outside the pool of " + std::to_string(m) + " chunks";
            return 1;
        }
    Scratch* S = g_scratch;
    if (S[0].ensure((size_t)n * 49 * sizeof(int)) || S[5].ensure((size_t)m * featureStride * sizeof(FeaturePlacement)) ||
        S[6].ensure((size_t)m * caveFeatureStride * sizeof(CaveFeaturePlacement)) || S[7].ensure((size_t)(m + n) * 2 * sizeof(int)) ||
        g_scratch[8].ensure((size_t)n * MAX_FEATURES * sizeof(FeaturePlacement)) || g_scratch[9].ensure((size_t)n * MAX_CAVE_FEATURES * sizeof(CaveFeaturePlacement)))
        return 1;
    int* d_counts = (int*)S[7].ptr;
    int* d_outCounts = d_counts + (size_t)m * 2;
    MMG_CUDA(cudaMemcpyAsync(S[0].ptr, neighbours, (size_t)n * 49 * sizeof(int), cudaMemcpyHostToDevice, g_stream));
    MMG_CUDA(cudaMemcpyAsync(S[5].ptr, features, (size_t)m * featureStride * sizeof(FeaturePlacement), cudaMemcpyHostToDevice, g_stream));
    MMG_CUDA(cudaMemcpyAsync(S[6].ptr, caveFeatures, (size_t)m * caveFeatureStride * sizeof(CaveFeaturePlacement), cudaMemcpyHostToDevice, g_stream));
    MMG_CUDA(cudaMemcpyAsync(d_counts, counts, (size_t)m * 2 * sizeof(int), cudaMemcpyHostToDevice, g_stream));
    MMG_LAUNCH(k_gather_concat, n, 256, 0, g_stream, (const int*)S[0].ptr, (const FeaturePlacement*)S[5].ptr, featureStride,
               (const CaveFeaturePlacement*)S[6].ptr, caveFeatureStride, (const int*)d_counts, (FeaturePlacement*)g_scratch[8].ptr,
               (CaveFeaturePlacement*)g_scratch[9].ptr, d_outCounts);
    MMG_CUDA(cudaMemcpyAsync(out_counts, d_outCounts, (size_t)n * 2 * sizeof(int), cudaMemcpyDeviceToHost, g_stream));
    MMG_CUDA(cudaStreamSynchronize(g_stream));
    for (int i = 0; i < n; ++i)
    {
        const int nf = std::min(out_counts[2 * i], MAX_FEATURES), nc = std::min(out_counts[2 * i + 1], MAX_CAVE_FEATURES);
        if (nf && out_features) MMG_CUDA(cudaMemcpyAsync(out_features + (size_t)i * MAX_FEATURES, (FeaturePlacement*)g_scratch[8].ptr + (size_t)i * MAX_FEATURES,
                                                         (size_t)nf * sizeof(FeaturePlacement), cudaMemcpyDeviceToHost, g_stream));
        if (nc && out_caveFeatures) MMG_CUDA(cudaMemcpyAsync(out_caveFeatures + (size_t)i * MAX_CAVE_FEATURES, (CaveFeaturePlacement*)g_scratch[9].ptr + (size_t)i * MAX_CAVE_FEATURES,
                                                             (size_t)nc * sizeof(CaveFeaturePlacement), cudaMemcpyDeviceToHost, g_stream));
    }
    MMG_CUDA(cudaStreamSynchronize(g_stream));
    return 0;
}

extern "C" int mmgen_fill(int n, const int32_t* origins, const float* heightfield, const float* biomeWeights, const float* layers,
                          const MmgenCaveLayer* caveLayers, const MmgenFeaturePlacement* features,
                          const MmgenCaveFeaturePlacement* caveFeatures, const int32_t* numFeatures, int featureStride,
                          int caveFeatureStride, uint8_t* out_blocks)
{
    if (requireReady()) return 1;
    MMG_BATCH_LOCK();
    if (n <= 0) return 0;
    // the lists are read with the caller's strides: a count beyond its stride (or beyond the reference's caps, which the
    // gather never exceeds, chunk.cu:1573-1578) would read past the caller's arrays
    if (featureStride < 0 || caveFeatureStride < 0 || !numFeatures)
    {
        g_lastError = "mmgen_fill: bad strides / numFeatures";
        return 1;
    }
    for (int i = 0; i < n; ++i)
        if (numFeatures[2 * i] < 0 || numFeatures[2 * i + 1] < 0 || numFeatures[2 * i] > featureStride || numFeatures[2 * i + 1] > caveFeatureStride ||
            (numFeatures[2 * i] > 0 && !features) || (numFeatures[2 * i + 1] > 0 && !caveFeatures))
        {
            g_lastError = "mmgen_fill: chunk " + std::to_string(i) + " has {" + std::to_string(numFeatures[2 * i]) + ", " +
                          std::to_string(numFeatures[2 * i + 1]) + "} placements but the strides are {" + std::to_string(featureStride) + ", " +
                          std::to_string(caveFeatureStride) + "}";
            return 1;
        }
    Scratch* X = g_scratch + 8;
    const size_t clBytes = (size_t)n * 256 * MAX_CAVE_LAYERS * sizeof(CaveLayer);
    Scratch* S = g_scratch;
    if (S[0].ensure((size_t)n * sizeof(int2)) || S[1].ensure((size_t)n * 256 * 4) || S[2].ensure((size_t)n * NUM_BIOMES * 256 * 4) ||
        S[3].ensure(clBytes) || S[4].ensure((size_t)n * NUM_MATERIALS * 256 * 4) ||
        S[5].ensure((size_t)n * featureStride * sizeof(FeaturePlacement) + 16) ||
        S[6].ensure((size_t)n * caveFeatureStride * sizeof(CaveFeaturePlacement) + 16) || S[7].ensure((size_t)n * 2 * sizeof(int)) ||
        X[0].ensure((size_t)n * sizeof(GatherInfo)) || X[1].ensure((size_t)n * 98304) ||
        X[2].ensure((size_t)kLushQueueCap * sizeof(uint2)) || X[3].ensure(4 * sizeof(int)) ||
        g_scratch[16].ensure((size_t)std::min(n, kFillBatch) * kRockQueuePerChunk * sizeof(uint2)) ||
        g_scratch[14].ensure((size_t)n * featureStride * sizeof(Prep) + 16) || g_scratch[15].ensure((size_t)n * caveFeatureStride * sizeof(Prep) + 16))
        return 1;
    MMG_CUDA(cudaMemcpyAsync(S[0].ptr, origins, (size_t)n * sizeof(int2), cudaMemcpyHostToDevice, g_stream));
    MMG_CUDA(cudaMemcpyAsync(S[1].ptr, heightfield, (size_t)n * 256 * 4, cudaMemcpyHostToDevice, g_stream));
    MMG_CUDA(cudaMemcpyAsync(S[2].ptr, biomeWeights, (size_t)n * NUM_BIOMES * 256 * 4, cudaMemcpyHostToDevice, g_stream));
    MMG_CUDA(cudaMemcpyAsync(S[3].ptr, caveLayers, clBytes, cudaMemcpyHostToDevice, g_stream));
    MMG_CUDA(cudaMemcpyAsync(S[4].ptr, layers, (size_t)n * NUM_MATERIALS * 256 * 4, cudaMemcpyHostToDevice, g_stream));
    if (featureStride) MMG_CUDA(cudaMemcpyAsync(S[5].ptr, features, (size_t)n * featureStride * sizeof(FeaturePlacement), cudaMemcpyHostToDevice, g_stream));
    if (caveFeatureStride) MMG_CUDA(cudaMemcpyAsync(S[6].ptr, caveFeatures, (size_t)n * caveFeatureStride * sizeof(CaveFeaturePlacement), cudaMemcpyHostToDevice, g_stream));
    MMG_CUDA(cudaMemcpyAsync(S[7].ptr, numFeatures, (size_t)n * 2 * sizeof(int), cudaMemcpyHostToDevice, g_stream));
    MMG_LAUNCH(k_gather_info, n, 256, 0, g_stream, (const FeaturePlacement*)S[5].ptr, (const CaveFeaturePlacement*)S[6].ptr,
               (const int*)S[7].ptr, featureStride, caveFeatureStride, (GatherInfo*)X[0].ptr);
    if (launchFill(n, nullptr, (const int2*)S[0].ptr, (const float*)S[1].ptr, (const float*)S[2].ptr, (const float*)S[4].ptr,
                   (const CaveLayer*)S[3].ptr, (const FeaturePlacement*)S[5].ptr, (const CaveFeaturePlacement*)S[6].ptr,
                   (GatherInfo*)X[0].ptr, (Prep*)g_scratch[14].ptr, (Prep*)g_scratch[15].ptr, featureStride, caveFeatureStride, (uint8_t*)X[1].ptr,
                   (uint2*)g_scratch[16].ptr, (uint2*)X[2].ptr, (int*)X[3].ptr, g_stream))
        return 1;
    MMG_CUDA(cudaMemcpyAsync(out_blocks, X[1].ptr, (size_t)n * 98304, cudaMemcpyDeviceToHost, g_stream));
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
    MMG_CUDA(cudaStreamCreateWithFlags(&w->copyStream, cudaStreamNonBlocking));
    MMG_CUDA(cudaStreamCreateWithFlags(&w->fillStream, cudaStreamNonBlocking));
    {
        // erosion is hundreds of short dependent launches: its blocks go first whenever the cave kernel frees an SM slot
        int prLow = 0, prHigh = 0;
        MMG_CUDA(cudaDeviceGetStreamPriorityRange(&prLow, &prHigh));
        MMG_CUDA(cudaStreamCreateWithPriority(&w->sideStream, cudaStreamNonBlocking, prHigh));
    }
    for (auto& e : w->evSide) MMG_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    for (auto& e : w->evGather) MMG_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    for (auto& e : w->evScan) MMG_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    for (auto& e : w->ev) MMG_CUDA(cudaEventCreate(&e));
    for (auto& e : w->evMesh) MMG_CUDA(cudaEventCreate(&e));
    for (auto& e : w->evBatch) MMG_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    std::vector<int2> origins(w->n);
    for (int z = 0; z < nz; ++z)
        for (int x = 0; x < nx; ++x) origins[z * nx + x] = make_int2((cx0 + x) * 16, (cz0 + z) * 16);
    MMG_CUDA(cudaMalloc(&w->d_origins, (size_t)w->n * sizeof(int2)));
    MMG_CUDA(cudaMemcpy(w->d_origins, origins.data(), (size_t)w->n * sizeof(int2), cudaMemcpyHostToDevice));
    w->h_origins.swap(origins);
    *out = w;
    return 0;
}

int mmgen_world_destroy(MmgenWorld* w)
{
    if (!w) return 0;
    cudaFree(w->d_origins);
    cudaFree(w->d_height);
    cudaFree(w->d_weights);
    cudaFree(w->d_layers);
    cudaFree(w->d_eroded);
    cudaFree(w->d_zone);
    cudaFree(w->d_zoneCorners);
    cudaFree(w->d_flags);
    cudaFree(w->d_listRing);
    cudaFree(w->d_caves);
    cudaFree(w->d_caveCols);
    cudaFree(w->d_caveQueue);
    cudaFree(w->d_caveCount);
    cudaFree(w->d_features);
    cudaFree(w->d_caveFeatures);
    cudaFree(w->d_counts);
    cudaFree(w->d_gF);
    cudaFree(w->d_gCF);
    cudaFree(w->d_info);
    cudaFree(w->d_prepF);
    cudaFree(w->d_prepC);
    cudaFree(w->d_lushQueue);
    cudaFree(w->d_lushCount);
    cudaFree(w->d_rockQueue);
    cudaFree(w->d_blocks);
    codecFree(w->codec);
    cudaFree(w->d_xIdx);
    cudaFree(w->d_xOff);
    cudaFree(w->d_meshList);
    cudaFree(w->d_meshColOff);
    cudaFree(w->d_meshTotals);
    cudaFree(w->d_meshBase);
    cudaFree(w->d_meshVerts);
    cudaFree(w->d_meshIdx);
    for (auto& e : w->ev) if (e) cudaEventDestroy(e);
    for (auto& e : w->evMesh) if (e) cudaEventDestroy(e);
    for (auto& e : w->evBatch) if (e) cudaEventDestroy(e);
    if (w->stream) cudaStreamDestroy(w->stream);
    if (w->copyStream) cudaStreamDestroy(w->copyStream);
    if (w->sideStream) cudaStreamDestroy(w->sideStream);
    if (w->fillStream) cudaStreamDestroy(w->fillStream);
    for (auto& e : w->evSide) if (e) cudaEventDestroy(e);
    for (auto& e : w->evGather) if (e) cudaEventDestroy(e);
    for (auto& e : w->evScan) if (e) cudaEventDestroy(e);
    cudaFree(w->d_gF2); cudaFree(w->d_gCF2); cudaFree(w->d_info2); cudaFree(w->d_prepF2); cudaFree(w->d_prepC2);
    delete w;
    return 0;
}

}  // extern "C" (everything below is declared with C linkage by mmgen.h; templates cannot sit inside the block)

// ------------------------------------------------------------------ stage runners of the resident world
// Each runs one stage over an explicit list of window chunks (raster indices) and keeps w->stage up to date.
// worldGenerate (batch mode) and the streaming scheduler (mm_stream.cuh, Terrain::tick re-hosted) both
// drive the world through these. The runners only queue work (erosion excepted: its convergence is polled); callers that
// read results synchronise (downloads, checksums, mmgen_world_sync, the end of a stream tick).
constexpr int kListSlots = 8;
static int worldUploadList(MmgenWorld* w, const std::vector<int>& list)
{
    if (!w->d_listRing) MMG_CUDA(cudaMalloc(&w->d_listRing, (size_t)kListSlots * w->n * sizeof(int)));
    w->d_list = w->d_listRing + (size_t)w->listSlot * w->n;
    w->listSlot = (w->listSlot + 1) % kListSlots;
    // pageable source: staged before the call returns, the device copy runs in stream order (after the kernels that read this
    // slice kListSlots stage calls ago)
    MMG_CUDA(cudaMemcpyAsync(w->d_list, list.data(), list.size() * sizeof(int), cudaMemcpyHostToDevice, w->stream));
    return 0;
}

// S1 (Chunk::generateHeightfields); list == nullptr: every chunk of the window
static int worldHeightfields(MmgenWorld* w, const std::vector<int>* list)
{
    if (list && list->empty() && !w->reserving) return 0;
    if (!w->d_height) MMG_CUDA(cudaMalloc(&w->d_height, (size_t)w->n * 256 * sizeof(float)));
    if (!w->d_weights) MMG_CUDA(cudaMalloc(&w->d_weights, (size_t)w->n * NUM_BIOMES * 256 * sizeof(float)));
    if (w->reserving) return worldUploadList(w, std::vector<int>());
    if (list)
    {
        if (worldUploadList(w, *list)) return 1;
        MMG_TIMED(K_HEIGHTFIELD, w->stream, 1, MMG_LAUNCH(k_heightfield, (int)list->size(), 256, kNoiseSmemBytes, w->stream, (const int*)w->d_list,
                                                          (const int2*)w->d_origins, w->d_height, w->d_weights));
        for (int i : *list) w->stage[i] = std::max<uint8_t>(w->stage[i], 1);
    }
    else
    {
        MMG_TIMED(K_HEIGHTFIELD, w->stream, 1, MMG_LAUNCH(k_heightfield, w->n, 256, kNoiseSmemBytes, w->stream, (const int*)nullptr,
                                                          (const int2*)w->d_origins, w->d_height, w->d_weights));
        for (auto& s : w->stage) s = std::max<uint8_t>(s, 1);
    }
    return 0;
}

// S2 (gatherHeightfield + Chunk::generateLayers): every listed chunk has its 3x3 neighbourhood at stage >= 1
static int worldLayers(MmgenWorld* w, const std::vector<int>& list)
{
    if (list.empty() && !w->reserving) return 0;
    if (!w->d_layers) MMG_CUDA(cudaMalloc(&w->d_layers, (size_t)w->n * NUM_MATERIALS * 256 * sizeof(float)));
    if (w->reserving) return 0;
    if (worldUploadList(w, list)) return 1;
    MMG_TIMED(K_LAYERS, w->stream, 1, MMG_LAUNCH(k_layers<true>, (int)list.size(), 256, kNoiseSmemBytes, w->stream, (const int*)w->d_list,
                                                 (const int2*)w->d_origins, (const float*)w->d_height, (const float*)w->d_weights, w->d_layers, w->nx));
    for (int i : list) w->stage[i] = std::max<uint8_t>(w->stage[i], 2);
    return 0;
}

// S3 (Chunk::erodeZone) for zones given by the window-local corner of their 24x24-chunk gather window
static int worldErode(MmgenWorld* w, const std::vector<int2>& corners)
{
    if (corners.empty() && !w->reserving) return 0;
    const int nx = w->nx;
    if (!w->d_zone) MMG_CUDA(cudaMalloc(&w->d_zone, (size_t)kZoneBatch * kZonePlanes * kErosionCols * sizeof(float)));
    if (!w->d_zoneCorners) MMG_CUDA(cudaMalloc(&w->d_zoneCorners, (size_t)kZoneBatch * sizeof(int2)));
    if (!w->d_flags) MMG_CUDA(cudaMalloc(&w->d_flags, (kSweepGroup + 3 * kMaxZoneBatch * kZoneTiles) * sizeof(int)));
    if (!w->d_eroded) MMG_CUDA(cudaMalloc(&w->d_eroded, (size_t)w->n * NUM_MATERIALS * 256 * sizeof(float)));
    if (w->reserving) return 0;
    for (size_t z0 = 0; z0 < corners.size(); z0 += kZoneBatch)
    {
        const int m = (int)std::min<size_t>(kZoneBatch, corners.size() - z0);
        MMG_CUDA(cudaMemcpyAsync(w->d_zoneCorners, corners.data() + z0, (size_t)m * sizeof(int2), cudaMemcpyHostToDevice, w->stream));
        MMG_LAUNCH(k_zone_gather, dim3(12, 12, m), dim3(32, 32), 0, w->stream, (const float*)w->d_layers,
                   (const float*)w->d_height, w->d_zone, (const int2*)w->d_zoneCorners, nx);
        int sweeps = 0;
        if (erodeZonesDevice(w->d_zone, m, w->d_flags, w->stream, &sweeps)) return 1;
        w->erosionSweeps += sweeps;
        MMG_LAUNCH(k_zone_scatter, dim3(6, 6, m), dim3(32, 32), 0, w->stream, (const float*)w->d_zone,
                   (const float*)w->d_layers, w->d_eroded, (const int2*)w->d_zoneCorners, nx);
        for (int k = 0; k < m; ++k)
            for (int z = 6; z < 18; ++z)
                for (int x = 6; x < 18; ++x) w->stage[(corners[z0 + k].y + z) * nx + corners[z0 + k].x + x] = 3;
    }
    MMG_CUDA(cudaStreamSynchronize(w->stream));   // the corner buffer is reused
    return 0;
}

// S4 (Chunk::generateCaves)
static int worldCaves(MmgenWorld* w, const std::vector<int>& list)
{
    if (list.empty() && !w->reserving) return 0;
    const int m = (int)list.size();
    if (!w->d_caves) MMG_CUDA(cudaMalloc(&w->d_caves, (size_t)w->n * 256 * MAX_CAVE_LAYERS * sizeof(CaveLayer)));
    if (!w->d_caveCols) MMG_CUDA(cudaMalloc(&w->d_caveCols, (size_t)kCaveBatch * 256 * sizeof(CaveColumn)));
    if (!w->d_caveQueue) MMG_CUDA(cudaMalloc(&w->d_caveQueue, (size_t)kCaveBiomeQueueCap * sizeof(uint2)));
    if (!w->d_caveCount) MMG_CUDA(cudaMalloc(&w->d_caveCount, sizeof(int)));
    if (w->reserving) return 0;
    if (worldUploadList(w, list)) return 1;
    if (launchCaves(m, (const int*)w->d_list, (const int2*)w->d_origins, (const float*)w->d_height, (const float*)w->d_weights,
                    w->d_caveCols, w->d_caves, w->d_caveQueue, w->d_caveCount, w->stream))
        return 1;
    for (int i : list) w->stage[i] = 4;
    return 0;
}

// S5a (Chunk::generateFeaturePlacements, a CPU pass in the reference)
static int worldPlacements(MmgenWorld* w, const std::vector<int>& list)
{
    if (list.empty() && !w->reserving) return 0;
    const int m = (int)list.size();
    if (!w->d_features) MMG_CUDA(cudaMalloc(&w->d_features, (size_t)w->n * kMaxOwnFeatures * sizeof(FeaturePlacement)));
    if (!w->d_caveFeatures) MMG_CUDA(cudaMalloc(&w->d_caveFeatures, (size_t)w->n * kMaxOwnCaveFeatures * sizeof(CaveFeaturePlacement)));
    if (!w->d_counts)
    {
        MMG_CUDA(cudaMalloc(&w->d_counts, (size_t)w->n * 2 * sizeof(int)));
        MMG_CUDA(cudaMemsetAsync(w->d_counts, 0, (size_t)w->n * 2 * sizeof(int), w->stream));
    }
    if (w->reserving) return 0;
    if (worldUploadList(w, list)) return 1;
    MMG_TIMED(K_PLACEMENTS, w->stream, 1, MMG_LAUNCH(k_feature_placements, m, 256, 0, w->stream, (const int*)w->d_list, (const int2*)w->d_origins,
                                                     (const float*)w->d_height, (const float*)w->d_weights, (const float*)w->d_eroded,
                                                     (const CaveLayer*)w->d_caves, w->d_features, w->d_caveFeatures, w->d_counts));
    for (int i : list) w->stage[i] = 5;
    return 0;
}

// S5b + S6 (gatherFeaturePlacements + Chunk::fill + placeDecorators): every listed chunk has its 7x7 neighbourhood at
// stage >= 5. hostBlocks != nullptr: each finished batch is copied to host memory while the next one is being filled,
// chunk i of the list to slot hostSlot(i).
static int codecEnsure(MmgenWorld* w, size_t targets, size_t batches);
static int codecEncodeBatch(MmgenWorld* w, int b, int m, const int* dl, const int* h_slots);
static int codecDeliver(MmgenWorld* w, int batches, size_t targets, EncodedOut* enc);

template <typename SlotFn>
static int worldFill(MmgenWorld* w, const std::vector<int>& list, uint8_t* hostBlocks, SlotFn hostSlot, EncodedOut* enc = nullptr)
{
    if (list.empty() && !w->reserving) return 0;
    const int nBatches = (int)((list.size() + kFillBatch - 1) / kFillBatch);
    std::vector<int> slots;
    if (enc)
    {
        size_t targets = 0;
        for (int c : list) targets = std::max(targets, hostSlot(c) + 1);
        if (codecEnsure(w, targets, (size_t)nBatches)) return 1;
        enc->bytes = targets;      // number of index entries, replaced by the byte count in codecDeliver
    }
    const int nx = w->nx;
    if (!w->d_blocks) MMG_CUDA(cudaMalloc(&w->d_blocks, (size_t)w->n * 98304));
    // scratch of one fill batch (0.66 MB per chunk: 1.3 GB for a full batch). A batch world gets it for min(batch, its chunks); a
    // streaming session for the most chunks one tick can fill under its action-time budget (mmgen_stream_set_costs) - cudaMalloc /
    // cudaFree of these sizes take tens of milliseconds (measured on BASELINE config 3), so it is sized once, not grown step by step
    const size_t hint = w->fillHint ? w->fillHint : std::min<size_t>((size_t)kFillBatch, (size_t)w->n);
    const size_t need = w->reserving ? hint : std::min<size_t>((size_t)kFillBatch, list.size());
    if (need > w->fillCap)
    {
        const size_t cap = std::min<size_t>((size_t)kFillBatch, std::max(need, hint));
        MMG_CUDA(cudaStreamSynchronize(w->stream));
        cudaFree(w->d_gF); cudaFree(w->d_gCF); cudaFree(w->d_info); cudaFree(w->d_prepF); cudaFree(w->d_prepC); cudaFree(w->d_rockQueue);
        cudaFree(w->d_gF2); cudaFree(w->d_gCF2); cudaFree(w->d_info2); cudaFree(w->d_prepF2); cudaFree(w->d_prepC2);
        w->d_gF = nullptr; w->d_gCF = nullptr; w->d_info = nullptr; w->d_prepF = nullptr; w->d_prepC = nullptr; w->d_rockQueue = nullptr;
        w->d_gF2 = nullptr; w->d_gCF2 = nullptr; w->d_info2 = nullptr; w->d_prepF2 = nullptr; w->d_prepC2 = nullptr;
        w->fillCap = 0;
        MMG_CUDA(cudaMalloc(&w->d_gF, cap * MAX_FEATURES * sizeof(FeaturePlacement)));
        MMG_CUDA(cudaMalloc(&w->d_gCF, cap * MAX_CAVE_FEATURES * sizeof(CaveFeaturePlacement)));
        MMG_CUDA(cudaMalloc(&w->d_info, cap * sizeof(GatherInfo)));
        MMG_CUDA(cudaMalloc(&w->d_prepF, cap * MAX_FEATURES * sizeof(Prep)));
        MMG_CUDA(cudaMalloc(&w->d_prepC, cap * MAX_CAVE_FEATURES * sizeof(Prep)));
        MMG_CUDA(cudaMalloc(&w->d_rockQueue, cap * kRockQueuePerChunk * sizeof(uint2)));
        w->fillCap = cap;
    }
    if (!w->d_lushQueue) MMG_CUDA(cudaMalloc(&w->d_lushQueue, (size_t)kLushQueueCap * sizeof(uint2)));
    if (!w->d_lushCount) MMG_CUDA(cudaMalloc(&w->d_lushCount, 4 * sizeof(int)));
    if (w->reserving) return 0;
    if (worldUploadList(w, list)) return 1;
    const bool pipelined = nBatches > 1;
    if (pipelined && !w->d_gF2)
    {
        MMG_CUDA(cudaMalloc(&w->d_gF2, w->fillCap * MAX_FEATURES * sizeof(FeaturePlacement)));
        MMG_CUDA(cudaMalloc(&w->d_gCF2, w->fillCap * MAX_CAVE_FEATURES * sizeof(CaveFeaturePlacement)));
        MMG_CUDA(cudaMalloc(&w->d_info2, w->fillCap * sizeof(GatherInfo)));
        MMG_CUDA(cudaMalloc(&w->d_prepF2, w->fillCap * MAX_FEATURES * sizeof(Prep)));
        MMG_CUDA(cudaMalloc(&w->d_prepC2, w->fillCap * MAX_CAVE_FEATURES * sizeof(Prep)));
    }
    FeaturePlacement* const gF[2] = {w->d_gF, w->d_gF2};
    CaveFeaturePlacement* const gCF[2] = {w->d_gCF, w->d_gCF2};
    GatherInfo* const info[2] = {w->d_info, w->d_info2};
    Prep* const prepF[2] = {w->d_prepF, w->d_prepF2};
    Prep* const prepC[2] = {w->d_prepC, w->d_prepC2};
    // gathered lists + Prep records of batch b into set b & 1, on stream st
    auto gather = [&](size_t b, cudaStream_t st) -> int {
        const size_t b0 = b * (size_t)kFillBatch;
        const int m = (int)std::min<size_t>(kFillBatch, list.size() - b0), set = pipelined ? (int)(b & 1) : 0;
        const int* dl = w->d_list + b0;
        MMG_TIMED(K_GATHER, st, 1, MMG_LAUNCH(k_gather_features, m, 256, 0, st, dl, (const int2*)w->d_origins,
                                              (const FeaturePlacement*)w->d_features, (const CaveFeaturePlacement*)w->d_caveFeatures,
                                              (const int*)w->d_counts, nx, gF[set], gCF[set], info[set]));
        return launchFillPrepare(m, dl, (const int2*)w->d_origins, gF[set], gCF[set], info[set], prepF[set], prepC[set], MAX_FEATURES, MAX_CAVE_FEATURES, st);
    };
    const bool overlapFill = pipelined && (g_fillOverlap & 15) > 0 && !g_serialStages;
    const int rockCtasPerSM = overlapFill ? std::min(g_fillOverlap & 15, (int)MMG_ROCK_MINBLOCKS) : (int)MMG_ROCK_MINBLOCKS;
    // bit 4 of the knob: the overlapped passes go to the high-priority side stream instead of the default-priority fill stream
    cudaStream_t const side = (overlapFill && !(g_fillOverlap & 16)) ? w->fillStream : w->sideStream;
    if (pipelined)
    {
        // the side stream starts where the main stream is now (placements done, list uploaded) with batch 0
        MMG_CUDA(cudaEventRecord(w->evSide[0], w->stream));
        MMG_CUDA(cudaStreamWaitEvent(side, w->evSide[0], 0));
        if (!overlapFill)
        {
            if (gather(0, w->sideStream)) return 1;
            MMG_CUDA(cudaEventRecord(w->evGather[0], w->sideStream));
        }
    }
    for (size_t b0 = 0; b0 < list.size(); b0 += kFillBatch)
    {
        const int m = (int)std::min<size_t>(kFillBatch, list.size() - b0);
        const int* dl = w->d_list + b0;
        const size_t b = b0 / kFillBatch;
        const int set = pipelined ? (int)(b & 1) : 0;
        if (overlapFill)
        {
            // side stream: lists + Prep records of batch b into set b & 1 (last read by the placement scan of batch b - 2), then the
            // batch's terrain / rock / lush passes (the rock and lush queues are only ever touched on this stream, in batch order).
            // Main stream: the placement scan + decorators of batch b once the side stream has finished the batch - while the side
            // stream is already in batch b + 1, whose chunks are different chunks.
            if (b >= 2) MMG_CUDA(cudaStreamWaitEvent(side, w->evScan[set], 0));
            if (gather(b, side)) return 1;
            if (launchFillTerrain(m, dl, (const int2*)w->d_origins, (const float*)w->d_height, (const float*)w->d_weights, (const float*)w->d_eroded,
                                  (const CaveLayer*)w->d_caves, w->d_blocks, w->d_rockQueue, w->d_lushQueue, w->d_lushCount, side, rockCtasPerSM))
                return 1;
            MMG_CUDA(cudaEventRecord(w->evGather[set], side));
        }
        else
        {
            if (pipelined && b + 1 < (size_t)nBatches)
            {
                // batch b + 1 is gathered and prepared on the side stream while this batch's terrain / rock passes run; its set was last read
                // by the placement scan of batch b - 1
                if (b >= 1) MMG_CUDA(cudaStreamWaitEvent(w->sideStream, w->evScan[set ^ 1], 0));
                if (gather(b + 1, w->sideStream)) return 1;
                MMG_CUDA(cudaEventRecord(w->evGather[set ^ 1], w->sideStream));
            }
            if (!pipelined && gather(b, w->stream)) return 1;
            if (launchFillTerrain(m, dl, (const int2*)w->d_origins, (const float*)w->d_height, (const float*)w->d_weights, (const float*)w->d_eroded,
                                  (const CaveLayer*)w->d_caves, w->d_blocks, w->d_rockQueue, w->d_lushQueue, w->d_lushCount, w->stream))
                return 1;
        }
        if (pipelined) MMG_CUDA(cudaStreamWaitEvent(w->stream, w->evGather[set], 0));
        if (launchFillFeatures(m, dl, (const int2*)w->d_origins, (const float*)w->d_height, (const float*)w->d_weights, (const CaveLayer*)w->d_caves,
                               gF[set], gCF[set], info[set], prepF[set], prepC[set], MAX_FEATURES, MAX_CAVE_FEATURES, w->d_blocks, w->stream))
            return 1;
        if (pipelined) MMG_CUDA(cudaEventRecord(w->evScan[set], w->stream));
        if (enc)
        {
            slots.resize(m);
            for (int i = 0; i < m; ++i) slots[i] = (int)hostSlot(list[b0 + i]);
            if (codecEncodeBatch(w, (int)(b0 / kFillBatch), m, dl, slots.data())) return 1;
        }
        else if (hostBlocks)
        {
            // stream the finished batch to the host while the next batch is being filled:
            // one copy per run of consecutive chunks whose host slots are consecutive too
            cudaEvent_t e = w->evBatch[(b0 / kFillBatch) & 1];
            MMG_CUDA(cudaEventRecord(e, w->stream));
            MMG_CUDA(cudaStreamWaitEvent(w->copyStream, e, 0));
            for (int i = 0; i < m;)
            {
                int j = i + 1;
                while (j < m && list[b0 + j] == list[b0 + j - 1] + 1 && hostSlot(list[b0 + j]) == hostSlot(list[b0 + j - 1]) + 1) ++j;
                const int c0 = list[b0 + i];
                MMG_CUDA(cudaMemcpyAsync(hostBlocks + hostSlot(c0) * 98304, w->d_blocks + (size_t)c0 * 98304, (size_t)(j - i) * 98304,
                                         cudaMemcpyDeviceToHost, w->copyStream));
                i = j;
            }
        }
    }
    if (enc && codecDeliver(w, nBatches, enc->bytes, enc)) return 1;
    for (int i : list) w->stage[i] = 6;
    return 0;
}

// every device buffer the stage runners would otherwise allocate at their first call, now (a streaming session reserves at set-up, as
// Terrain::initCuda does, terrain.cpp:154-185: cudaMalloc inside a tick costs the tick tens of milliseconds)
static int worldReserve(MmgenWorld* w)
{
    const std::vector<int> none;
    const std::vector<int2> noZones;
    w->reserving = true;
    const int rc = worldHeightfields(w, &none) || worldLayers(w, none) || worldErode(w, noZones) || worldCaves(w, none) || worldPlacements(w, none) ||
                   worldFill(w, none, nullptr, [](int) { return (size_t)0; });
    w->reserving = false;
    return rc;
}

static int worldGenerate(MmgenWorld* w, int stageMask, uint8_t* hostBlocks, EncodedOut* enc = nullptr)
{
    if (requireReady()) return 1;
    MMG_CUDA(cudaEventRecord(w->ev[12], w->stream));
    // host-delivery calls take their only input, the chunk origins, from host memory every time
    if (hostBlocks || enc) MMG_CUDA(cudaMemcpyAsync(w->d_origins, w->h_origins.data(), (size_t)w->n * sizeof(int2), cudaMemcpyHostToDevice, w->stream));
    const int nx = w->nx, nz = w->nz;
    if (stageMask & MMGEN_STAGE_HEIGHTFIELD)
    {
        MMG_CUDA(cudaEventRecord(w->ev[0], w->stream));
        if (worldHeightfields(w, nullptr)) return 1;
        MMG_CUDA(cudaEventRecord(w->ev[1], w->stream));
    }
    // A generate that runs layers, erosion and caves in one call overlaps them: the cave stage reads stage-1 products only
    // (heightfield, biome weights: chunk.cu:755-937), so it runs on `stream` while layers + erosion - 1 600 short dependent launches
    // that leave most of the GPU idle - run on the high-priority side stream; both join before the placements. The reference's
    // state machine orders them one after the other (terrain.cpp:587-960); the products are the same either way.
    constexpr int kOverlapMask = MMGEN_STAGE_LAYERS | MMGEN_STAGE_EROSION | MMGEN_STAGE_CAVES;
    const bool overlap = (stageMask & kOverlapMask) == kOverlapMask && !g_serialStages;
    struct StreamRoles      // whatever path leaves this function, the world's streams keep their roles
    {
        MmgenWorld* w; cudaStream_t main, side;
        ~StreamRoles() { w->stream = main; w->sideStream = side; }
    } roles{w, w->stream, w->sideStream};
    if (overlap)
    {
        MMG_CUDA(cudaEventRecord(w->evSide[0], w->stream));
        MMG_CUDA(cudaStreamWaitEvent(w->sideStream, w->evSide[0], 0));
        std::swap(w->stream, w->sideStream);      // the stage runners below enqueue on w->stream
    }
    if (stageMask & MMGEN_STAGE_LAYERS)
    {
        // chunks whose 3x3 neighbourhood lies inside the window (gatherHeightfield's condition)
        std::vector<int> list;
        for (int z = 1; z < nz - 1; ++z)
            for (int x = 1; x < nx - 1; ++x)
                if (w->stage[z * nx + x] >= 1) list.push_back(z * nx + x);
        MMG_CUDA(cudaEventRecord(w->ev[2], w->stream));
        if (worldLayers(w, list)) return 1;
        MMG_CUDA(cudaEventRecord(w->ev[3], w->stream));
    }
    if (stageMask & MMGEN_STAGE_EROSION)
    {
        MMG_CUDA(cudaEventRecord(w->ev[4], w->stream));
        w->erosionSweeps = 0;
        // zones are aligned to multiples of 12 chunks in WORLD chunk coordinates (terrain.cpp:259-262)
        auto floorDiv = [](int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); };
        std::vector<int2> corners;   // window-local corner of each erodable zone's 24x24-chunk window
        for (int zz = floorDiv(w->cz0, 12) * 12; zz < w->cz0 + nz; zz += 12)
            for (int zx = floorDiv(w->cx0, 12) * 12; zx < w->cx0 + nx; zx += 12)
            {
                const int lx0 = zx - 6 - w->cx0, lz0 = zz - 6 - w->cz0;
                if (lx0 < 0 || lz0 < 0 || lx0 + 24 > nx || lz0 + 24 > nz) continue;
                // with a target region only zones that meet target (+) 3 chunks are eroded
                if (w->hasTarget && (lx0 + 18 <= w->tx0 - 3 || lx0 + 6 >= w->tx0 + w->tnx + 3 ||
                                     lz0 + 18 <= w->tz0 - 3 || lz0 + 6 >= w->tz0 + w->tnz + 3)) continue;
                if (w->hasGlobal)
                {
                    // halo exchange: only zones holding a chunk whose placements this world computes itself
                    bool any = false;
                    for (int z = 6; z < 18 && !any; ++z)
                        for (int x = 6; x < 18 && !any; ++x) any = w->ownsPlacements(lx0 + x, lz0 + z);
                    if (!any) continue;
                }
                // a zone is eroded once: its 144 centre chunks move from stage 2 to 3 together (later stages are never demoted)
                if (w->stage[(lz0 + 6) * nx + lx0 + 6] >= 3) continue;
                bool ok = true;
                for (int z = 0; z < 24 && ok; ++z)
                    for (int x = 0; x < 24 && ok; ++x) ok = w->stage[(lz0 + z) * nx + lx0 + x] >= 2;
                if (ok) corners.push_back(make_int2(lx0, lz0));
            }
        std::vector<int> caved;
        if (overlap)
        {
            // the caves of every chunk that is eroded already or is about to be, queued on the main stream first
            std::vector<uint8_t> will(w->n, 0);
            for (const int2& c : corners)
                for (int z = 6; z < 18; ++z)
                    for (int x = 6; x < 18; ++x) will[(c.y + z) * nx + c.x + x] = 1;
            for (int i = 0; i < w->n; ++i)
                if ((w->stage[i] == 3 || (will[i] && w->stage[i] < 3)) && w->ownsPlacements(i % nx, i / nx)) caved.push_back(i);
            std::swap(w->stream, w->sideStream);      // main stream
            MMG_CUDA(cudaEventRecord(w->ev[6], w->stream));
            if (worldCaves(w, caved)) return 1;
            MMG_CUDA(cudaEventRecord(w->ev[7], w->stream));
            std::swap(w->stream, w->sideStream);      // side stream again
        }
        if (worldErode(w, corners)) return 1;
        MMG_CUDA(cudaEventRecord(w->ev[5], w->stream));
        if (overlap)
        {
            for (int i : caved) w->stage[i] = 4;      // worldErode marked its zones' chunks 3
            MMG_CUDA(cudaEventRecord(w->evSide[1], w->stream));
            std::swap(w->stream, w->sideStream);
            MMG_CUDA(cudaStreamWaitEvent(w->stream, w->evSide[1], 0));
        }
    }
    if ((stageMask & MMGEN_STAGE_CAVES) && !overlap)
    {
        std::vector<int> list;
        for (int i = 0; i < w->n; ++i)
            if (w->stage[i] == 3 && w->ownsPlacements(i % nx, i / nx)) list.push_back(i);
        MMG_CUDA(cudaEventRecord(w->ev[6], w->stream));
        if (worldCaves(w, list)) return 1;
        MMG_CUDA(cudaEventRecord(w->ev[7], w->stream));
    }
    if (stageMask & MMGEN_STAGE_FEATURES)
    {
        std::vector<int> list;
        for (int i = 0; i < w->n; ++i)
            if (w->stage[i] == 4) list.push_back(i);
        MMG_CUDA(cudaEventRecord(w->ev[8], w->stream));
        if (worldPlacements(w, list)) return 1;
        MMG_CUDA(cudaEventRecord(w->ev[9], w->stream));
    }
    if (stageMask & MMGEN_STAGE_FILL)
    {
        // chunks whose 7x7 neighbourhood has placements (gatherFeaturePlacements' condition)
        std::vector<int> list;
        for (int z = 3; z < nz - 3; ++z)
            for (int x = 3; x < nx - 3; ++x)
            {
                if (w->stage[z * nx + x] != 5 || !w->inTarget(x, z, 0)) continue;
                bool ok = true;
                for (int dz = -3; dz <= 3 && ok; ++dz)
                    for (int dx = -3; dx <= 3 && ok; ++dx) ok = w->stage[(z + dz) * nx + x + dx] >= 5;
                if (ok) list.push_back(z * nx + x);
            }
        MMG_CUDA(cudaEventRecord(w->ev[10], w->stream));
        // host slots: region raster order (a region row is contiguous in the window)
        auto slot = [w, nx](int c) -> size_t {
            return w->hasTarget ? (size_t)(c / nx - w->tz0) * w->tnx + (c % nx - w->tx0) : (size_t)c;
        };
        if (worldFill(w, list, hostBlocks, slot, enc)) return 1;
        MMG_CUDA(cudaEventRecord(w->ev[11], w->stream));
    }
    if (hostBlocks || enc)
    {
        // the whole-generate bracket includes the tail of the download
        MMG_CUDA(cudaEventRecord(w->evBatch[0], w->copyStream));
        MMG_CUDA(cudaStreamWaitEvent(w->stream, w->evBatch[0], 0));
    }
    MMG_CUDA(cudaEventRecord(w->ev[13], w->stream));
    if (hostBlocks || enc) MMG_CUDA(cudaStreamSynchronize(w->stream));
    return 0;
}

int mmgen_world_generate(MmgenWorld* w, int stageMask) { return worldGenerate(w, stageMask, nullptr); }

int mmgen_world_generate_to_host(MmgenWorld* w, int stageMask, uint8_t* out_blocks)
{
    if (!out_blocks)
    {
        g_lastError = "mmgen_world_generate_to_host: out_blocks is NULL";
        return 1;
    }
    return worldGenerate(w, stageMask, out_blocks);
}

int mmgen_world_generate_to_host_encoded(MmgenWorld* w, int stageMask, uint8_t* out_buf, size_t capBytes, uint64_t* out_index, size_t* out_bytes)
{
    if (!out_buf || !out_index || !out_bytes)
    {
        g_lastError = "mmgen_world_generate_to_host_encoded: null output pointer";
        return 1;
    }
    EncodedOut enc = {out_buf, capBytes, out_index, 0};
    *out_bytes = 0;
    const int rc = worldGenerate(w, stageMask, nullptr, &enc);
    *out_bytes = enc.bytes;
    return rc;
}

#include "mm_stream.inl"
#include "mm_codec.inl"

int mmgen_world_reset(MmgenWorld* w)
{
    MMG_CUDA(cudaStreamSynchronize(w->stream));
    std::fill(w->stage.begin(), w->stage.end(), 0);
    if (w->d_counts) MMG_CUDA(cudaMemsetAsync(w->d_counts, 0, (size_t)w->n * 2 * sizeof(int), w->stream));
    return 0;
}

int mmgen_world_rewind(MmgenWorld* w, int stage)
{
    if (stage < 0 || stage > 6) { g_lastError = "mmgen_world_rewind: stage must be 0..6"; return 1; }
    MMG_CUDA(cudaStreamSynchronize(w->stream));
    for (auto& s : w->stage) s = std::min<uint8_t>(s, (uint8_t)stage);
    return 0;
}

int mmgen_world_create_for_region(int rx0, int rz0, int rnx, int rnz, MmgenWorld** out)
{
    if (rnx <= 0 || rnz <= 0 || !out)
    {
        g_lastError = "mmgen_world_create_for_region: bad arguments";
        return 1;
    }
    // apron rule: fill R <= placements on R(+)3 <= erosion of every zone meeting R(+)3 <= layers on those zones (+)6
    // <= heightfields on one more ring of chunks (the 1-block slope border of the outermost layers)
    auto floorDiv = [](int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); };
    const int zx0 = floorDiv(rx0 - 3, 12) * 12, zx1 = floorDiv(rx0 + rnx + 2, 12) * 12 + 12;
    const int zz0 = floorDiv(rz0 - 3, 12) * 12, zz1 = floorDiv(rz0 + rnz + 2, 12) * 12 + 12;
    if (mmgen_world_create(zx0 - 7, zz0 - 7, zx1 - zx0 + 14, zz1 - zz0 + 14, out)) return 1;
    MmgenWorld* w = *out;
    w->hasTarget = true;
    w->tx0 = rx0 - w->cx0; w->tz0 = rz0 - w->cz0; w->tnx = rnx; w->tnz = rnz;
    return 0;
}

int mmgen_world_set_exchange_region(MmgenWorld* w, int gx0, int gz0, int gnx, int gnz)
{
    if (gnx <= 0 || gnz <= 0) { w->hasGlobal = false; return 0; }
    if (!w->hasTarget) { g_lastError = "mmgen_world_set_exchange_region: the world has no target region (mmgen_world_create_for_region)"; return 1; }
    w->hasGlobal = true;
    w->gx0 = gx0 - w->cx0; w->gz0 = gz0 - w->cz0; w->gnx = gnx; w->gnz = gnz;
    return 0;
}

// chunk indices (window raster) of a world-coordinate rectangle, validated; grows the exchange scratch
static int exchangeRect(MmgenWorld* w, int cx0, int cz0, int nx, int nz, std::vector<int>& idx, const char* who)
{
    const int lx0 = cx0 - w->cx0, lz0 = cz0 - w->cz0;
    if (nx <= 0 || nz <= 0 || lx0 < 0 || lz0 < 0 || lx0 + nx > w->nx || lz0 + nz > w->nz)
    {
        g_lastError = std::string(who) + ": rectangle outside the world's window";
        return 1;
    }
    idx.resize((size_t)nx * nz);
    for (int z = 0; z < nz; ++z)
        for (int x = 0; x < nx; ++x) idx[(size_t)z * nx + x] = (lz0 + z) * w->nx + lx0 + x;
    if (idx.size() > w->xCap)
    {
        cudaFree(w->d_xIdx); cudaFree(w->d_xOff);
        w->d_xIdx = nullptr; w->d_xOff = nullptr; w->xCap = 0;
        MMG_CUDA(cudaMalloc(&w->d_xIdx, idx.size() * sizeof(int)));
        MMG_CUDA(cudaMalloc(&w->d_xOff, idx.size() * sizeof(long long)));
        w->xCap = idx.size();
    }
    return 0;
}

int mmgen_world_pack_placements(MmgenWorld* w, int cx0, int cz0, int nx, int nz, void* d_buf, size_t capBytes, size_t* bytes)
{
    if (requireReady()) return 1;
    std::vector<int> idx;
    if (exchangeRect(w, cx0, cz0, nx, nz, idx, "mmgen_world_pack_placements")) return 1;
    const int n = (int)idx.size();
    for (int i : idx)
        if (w->stage[i] < 5) { g_lastError = "mmgen_world_pack_placements: a chunk of the rectangle has no placements yet"; return 1; }
    // the counts of the rectangle's rows are contiguous in the world's count plane
    std::vector<int> cnt((size_t)n * 2);
    MMG_CUDA(cudaMemcpy2DAsync(cnt.data(), (size_t)nx * 2 * sizeof(int), w->d_counts + (size_t)idx[0] * 2, (size_t)w->nx * 2 * sizeof(int),
                               (size_t)nx * 2 * sizeof(int), nz, cudaMemcpyDeviceToHost, w->stream));
    MMG_CUDA(cudaStreamSynchronize(w->stream));
    std::vector<long long> off(n);
    long long at = (long long)n * 2 * sizeof(int);
    for (int i = 0; i < n; ++i)
    {
        off[i] = at;
        at += (long long)cnt[2 * i] * sizeof(FeaturePlacement) + (long long)cnt[2 * i + 1] * sizeof(CaveFeaturePlacement);
    }
    if (bytes) *bytes = (size_t)at;
    if ((size_t)at > capBytes || !d_buf)
    {
        g_lastError = "mmgen_world_pack_placements: the message needs " + std::to_string(at) + " bytes, the buffer holds " + std::to_string(capBytes);
        return 2;
    }
    MMG_CUDA(cudaMemcpyAsync(w->d_xIdx, idx.data(), (size_t)n * sizeof(int), cudaMemcpyHostToDevice, w->stream));
    MMG_CUDA(cudaMemcpyAsync(w->d_xOff, off.data(), (size_t)n * sizeof(long long), cudaMemcpyHostToDevice, w->stream));
    MMG_LAUNCH(k_pack_placements, n, 256, 0, w->stream, (const int*)w->d_xIdx, (const long long*)w->d_xOff, (const FeaturePlacement*)w->d_features,
               (const CaveFeaturePlacement*)w->d_caveFeatures, (const int*)w->d_counts, (uint8_t*)d_buf);
    MMG_CUDA(cudaStreamSynchronize(w->stream));      // the caller hands d_buf to another stream (NCCL)
    return 0;
}

int mmgen_world_unpack_placements(MmgenWorld* w, int cx0, int cz0, int nx, int nz, const void* d_buf, size_t bytes)
{
    if (requireReady()) return 1;
    std::vector<int> idx;
    if (exchangeRect(w, cx0, cz0, nx, nz, idx, "mmgen_world_unpack_placements")) return 1;
    const int n = (int)idx.size();
    if (bytes < (size_t)n * 2 * sizeof(int)) { g_lastError = "mmgen_world_unpack_placements: message shorter than its header"; return 1; }
    if (!w->d_features) MMG_CUDA(cudaMalloc(&w->d_features, (size_t)w->n * kMaxOwnFeatures * sizeof(FeaturePlacement)));
    if (!w->d_caveFeatures) MMG_CUDA(cudaMalloc(&w->d_caveFeatures, (size_t)w->n * kMaxOwnCaveFeatures * sizeof(CaveFeaturePlacement)));
    if (!w->d_counts)
    {
        MMG_CUDA(cudaMalloc(&w->d_counts, (size_t)w->n * 2 * sizeof(int)));
        MMG_CUDA(cudaMemsetAsync(w->d_counts, 0, (size_t)w->n * 2 * sizeof(int), w->stream));
    }
    std::vector<int> cnt((size_t)n * 2);
    MMG_CUDA(cudaMemcpyAsync(cnt.data(), d_buf, cnt.size() * sizeof(int), cudaMemcpyDeviceToHost, w->stream));
    MMG_CUDA(cudaStreamSynchronize(w->stream));
    std::vector<long long> off(n);
    long long at = (long long)n * 2 * sizeof(int);
    for (int i = 0; i < n; ++i)
    {
        if (cnt[2 * i] < 0 || cnt[2 * i] > kMaxOwnFeatures || cnt[2 * i + 1] < 0 || cnt[2 * i + 1] > kMaxOwnCaveFeatures)
        {
            g_lastError = "mmgen_world_unpack_placements: corrupt header";
            return 1;
        }
        off[i] = at;
        at += (long long)cnt[2 * i] * sizeof(FeaturePlacement) + (long long)cnt[2 * i + 1] * sizeof(CaveFeaturePlacement);
    }
    if ((size_t)at != bytes) { g_lastError = "mmgen_world_unpack_placements: message length does not match its header"; return 1; }
    MMG_CUDA(cudaMemcpyAsync(w->d_xIdx, idx.data(), (size_t)n * sizeof(int), cudaMemcpyHostToDevice, w->stream));
    MMG_CUDA(cudaMemcpyAsync(w->d_xOff, off.data(), (size_t)n * sizeof(long long), cudaMemcpyHostToDevice, w->stream));
    MMG_LAUNCH(k_unpack_placements, n, 256, 0, w->stream, (const int*)w->d_xIdx, (const long long*)w->d_xOff, (const uint8_t*)d_buf, w->d_features,
               w->d_caveFeatures, w->d_counts);
    MMG_CUDA(cudaStreamSynchronize(w->stream));      // the scratch lists are reused by the next message
    for (int i : idx) w->stage[i] = std::max<uint8_t>(w->stage[i], 5);
    return 0;
}

int mmgen_chunk_costs(int n, const int32_t* origins, float* out_costs)
{
    if (requireReady()) return 1;
    MMG_BATCH_LOCK();
    if (n <= 0) return 0;
    Scratch* S = g_scratch;
    if (S[0].ensure((size_t)n * sizeof(int2)) || S[1].ensure((size_t)n * 256 * sizeof(float)) || S[2].ensure((size_t)n * NUM_BIOMES * 256 * sizeof(float)) ||
        S[3].ensure((size_t)n * 3 * sizeof(float)))
        return 1;
    MMG_CUDA(cudaMemcpyAsync(S[0].ptr, origins, (size_t)n * sizeof(int2), cudaMemcpyHostToDevice, g_stream));
    MMG_LAUNCH(k_heightfield, n, 256, kNoiseSmemBytes, g_stream, (const int*)nullptr, (const int2*)S[0].ptr, (float*)S[1].ptr, (float*)S[2].ptr);
    MMG_LAUNCH(k_chunk_cost, n, 256, 0, g_stream, (const float*)S[1].ptr, (const float*)S[2].ptr, (float*)S[3].ptr);
    MMG_CUDA(cudaMemcpyAsync(out_costs, S[3].ptr, (size_t)n * 3 * sizeof(float), cudaMemcpyDeviceToHost, g_stream));
    MMG_CUDA(cudaStreamSynchronize(g_stream));
    return 0;
}

int mmgen_world_window(MmgenWorld* w, int* out8)
{
    out8[0] = w->cx0; out8[1] = w->cz0; out8[2] = w->nx; out8[3] = w->nz;
    out8[4] = w->hasTarget ? w->cx0 + w->tx0 : w->cx0; out8[5] = w->hasTarget ? w->cz0 + w->tz0 : w->cz0;
    out8[6] = w->hasTarget ? w->tnx : w->nx; out8[7] = w->hasTarget ? w->tnz : w->nz;
    return 0;
}

int mmgen_world_total_ms(MmgenWorld* w, float* out)
{
    MMG_CUDA(cudaStreamSynchronize(w->stream));
    MMG_CUDA(cudaEventElapsedTime(out, w->ev[12], w->ev[13]));
    return 0;
}

int mmgen_world_download_region_blocks(MmgenWorld* w, uint8_t* out_blocks)
{
    MMG_CUDA(cudaStreamSynchronize(w->stream));
    if (!w->d_blocks)
    {
        g_lastError = "mmgen_world_download_region_blocks: nothing filled yet";
        return 1;
    }
    const int tx0 = w->hasTarget ? w->tx0 : 0, tz0 = w->hasTarget ? w->tz0 : 0;
    const int tnx = w->hasTarget ? w->tnx : w->nx, tnz = w->hasTarget ? w->tnz : w->nz;
    MMG_CUDA(cudaMemcpy2D(out_blocks, (size_t)tnx * 98304, w->d_blocks + ((size_t)tz0 * w->nx + tx0) * 98304, (size_t)w->nx * 98304,
                          (size_t)tnx * 98304, tnz, cudaMemcpyDeviceToHost));
    return 0;
}

int mmgen_world_device_ptrs(MmgenWorld* w, void** heightfield, void** biomeWeights, void** layers, void** caveLayers, void** blocks)
{
    if (heightfield) *heightfield = w->d_height;
    if (biomeWeights) *biomeWeights = w->d_weights;
    if (layers) *layers = w->d_eroded ? (void*)w->d_eroded : (void*)w->d_layers;
    if (caveLayers) *caveLayers = w->d_caves;
    if (blocks) *blocks = w->d_blocks;
    return 0;
}

// per-column FNV hashes of every filled chunk (list order), shared by the two checksums below
static int worldColumnHashes(MmgenWorld* w, std::vector<int>& list, std::vector<unsigned long long>& hs)
{
    list.clear();
    for (int i = 0; i < w->n; ++i)
        if (w->stage[i] == 6) list.push_back(i);
    hs.clear();
    if (list.empty()) return 0;
    const int m = (int)list.size();
    int* d_l = nullptr;
    unsigned long long* d_h = nullptr;
    MMG_CUDA(cudaMalloc(&d_l, (size_t)m * sizeof(int)));
    MMG_CUDA(cudaMalloc(&d_h, (size_t)m * 256 * sizeof(unsigned long long)));
    MMG_CUDA(cudaMemcpyAsync(d_l, list.data(), (size_t)m * sizeof(int), cudaMemcpyHostToDevice, w->stream));
    MMG_LAUNCH(k_column_hashes, (m * 256 + 255) / 256, 256, 0, w->stream, (const int*)d_l, m, (const uint8_t*)w->d_blocks, d_h);
    hs.resize((size_t)m * 256);
    MMG_CUDA(cudaMemcpyAsync(hs.data(), d_h, hs.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost, w->stream));
    MMG_CUDA(cudaStreamSynchronize(w->stream));
    cudaFree(d_l);
    cudaFree(d_h);
    return 0;
}

int mmgen_world_block_checksum(MmgenWorld* w, uint64_t* out)
{
    if (requireReady()) return 1;
    std::vector<int> list;
    std::vector<unsigned long long> hs;
    if (worldColumnHashes(w, list, hs)) return 1;
    uint64_t h = 14695981039346656037ull;
    for (unsigned long long v : hs)
        for (int b = 0; b < 8; ++b) { h ^= (v >> (8 * b)) & 0xff; h *= 1099511628211ull; }
    *out = h;
    return 0;
}

// hash of (chunk coordinates, block volume) of every filled chunk, list order = raster order of the filled chunks
static int worldChunkHashes(MmgenWorld* w, std::vector<int>& list, std::vector<uint64_t>& out)
{
    std::vector<unsigned long long> hs;
    if (worldColumnHashes(w, list, hs)) return 1;
    out.resize(list.size());
    for (size_t k = 0; k < list.size(); ++k)
    {
        const int cx = w->cx0 + list[k] % w->nx, cz = w->cz0 + list[k] / w->nx;
        uint64_t h = 14695981039346656037ull;
        auto mix = [&](unsigned long long v) { for (int b = 0; b < 8; ++b) { h ^= (v >> (8 * b)) & 0xff; h *= 1099511628211ull; } };
        mix((unsigned long long)(long long)cx);
        mix((unsigned long long)(long long)cz);
        for (int c = 0; c < 256; ++c) mix(hs[k * 256 + c]);
        out[k] = h;
    }
    return 0;
}

int mmgen_world_chunk_hash_sum(MmgenWorld* w, uint64_t* out)
{
    if (requireReady()) return 1;
    std::vector<int> list;
    std::vector<uint64_t> hs;
    if (worldChunkHashes(w, list, hs)) return 1;
    uint64_t total = 0;
    for (uint64_t h : hs) total += h;      // mod 2^64: independent of the order and of how the world is tiled
    *out = total;
    return 0;
}

int mmgen_world_chunk_hashes(MmgenWorld* w, int cap, int32_t* coords, uint64_t* hashes, int* n)
{
    if (requireReady()) return 1;
    std::vector<int> list;
    std::vector<uint64_t> hs;
    if (worldChunkHashes(w, list, hs)) return 1;
    const int m = std::min<int>(cap, (int)list.size());
    for (int k = 0; k < m; ++k)
    {
        coords[2 * k] = w->cx0 + list[k] % w->nx;
        coords[2 * k + 1] = w->cz0 + list[k] / w->nx;
        hashes[k] = hs[k];
    }
    if (n) *n = (int)list.size();
    return 0;
}

int mmgen_world_download_features(MmgenWorld* w, int maxPerChunk, MmgenFeaturePlacement* features,
                                  MmgenCaveFeaturePlacement* caveFeatures, int32_t* counts)
{
    MMG_CUDA(cudaStreamSynchronize(w->stream));
    if (!w->d_counts) { std::memset(counts, 0, (size_t)w->n * 2 * sizeof(int)); return 0; }
    MMG_CUDA(cudaMemcpy(counts, w->d_counts, (size_t)w->n * 2 * sizeof(int), cudaMemcpyDeviceToHost));
    for (int c = 0; c < w->n; ++c)
    {
        const int nf = std::min(counts[2 * c], maxPerChunk), nc = std::min(counts[2 * c + 1], maxPerChunk);
        if (nf) MMG_CUDA(cudaMemcpy(features + (size_t)c * maxPerChunk, w->d_features + (size_t)c * kMaxOwnFeatures, (size_t)nf * sizeof(FeaturePlacement), cudaMemcpyDeviceToHost));
        if (nc) MMG_CUDA(cudaMemcpy(caveFeatures + (size_t)c * maxPerChunk, w->d_caveFeatures + (size_t)c * kMaxOwnCaveFeatures, (size_t)nc * sizeof(CaveFeaturePlacement), cudaMemcpyDeviceToHost));
    }
    return 0;
}

int mmgen_world_erosion_sweeps(MmgenWorld* w, int* out)
{
    *out = w->erosionSweeps;
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
    for (int st = 1; st <= 6; ++st)
        if (cudaEventQuery(w->ev[2 * st - 1]) == cudaSuccess && cudaEventElapsedTime(&out7[st], w->ev[2 * st - 2], w->ev[2 * st - 1]) != cudaSuccess)
        {
            out7[st] = 0.f;
            cudaGetLastError();
        }
    return 0;
}

int mmgen_world_download(MmgenWorld* w, float* heightfield, float* biomeWeights, float* layers,
                         MmgenCaveLayer* caveLayers, uint8_t* blocks)
{
    MMG_CUDA(cudaStreamSynchronize(w->stream));
    if (heightfield && w->d_height) MMG_CUDA(cudaMemcpy(heightfield, w->d_height, (size_t)w->n * 256 * sizeof(float), cudaMemcpyDeviceToHost));
    if (biomeWeights && w->d_weights) MMG_CUDA(cudaMemcpy(biomeWeights, w->d_weights, (size_t)w->n * NUM_BIOMES * 256 * sizeof(float), cudaMemcpyDeviceToHost));
    // layers: eroded set where a chunk was eroded, else the S2 output
    if (layers && w->d_layers)
    {
        MMG_CUDA(cudaMemcpy(layers, w->d_layers, (size_t)w->n * NUM_MATERIALS * 256 * sizeof(float), cudaMemcpyDeviceToHost));
        if (w->d_eroded)
            for (int i = 0; i < w->n; ++i)
                if (w->stage[i] >= 3)
                    MMG_CUDA(cudaMemcpy(layers + (size_t)i * NUM_MATERIALS * 256, w->d_eroded + (size_t)i * NUM_MATERIALS * 256,
                                        NUM_MATERIALS * 256 * sizeof(float), cudaMemcpyDeviceToHost));
    }
    if (caveLayers && w->d_caves) MMG_CUDA(cudaMemcpy(caveLayers, w->d_caves, (size_t)w->n * 256 * MAX_CAVE_LAYERS * sizeof(CaveLayer), cudaMemcpyDeviceToHost));
    if (blocks && w->d_blocks) MMG_CUDA(cudaMemcpy(blocks, w->d_blocks, (size_t)w->n * 98304, cudaMemcpyDeviceToHost));
    return 0;
}

// see mm_featurefuncs.cuh: 0 = the reference as built on Linux (default, the parity target), 1 = the source-text reading
int mmgen_set_cave_grid_test(int honoured)
{
    if (requireReady()) return 1;
    const int v = honoured ? 1 : 0;
    MMG_CUDA(cudaMemcpyToSymbol(g_caveGridTestHonoured, &v, sizeof(v)));
    return 0;
}

// ------------------------------------------------------------------ measurement helpers
int mmgen_kernel_timing(int enable)
{
    if (requireReady()) return 1;
    MMG_CUDA(cudaDeviceSynchronize());
    g_kt.on = enable != 0;
    g_kt.used = 0;
    return 0;
}

int mmgen_kernel_times(int cap, float* out_ms, int32_t* out_launches, int* n)
{
    if (requireReady()) return 1;
    MMG_CUDA(cudaDeviceSynchronize());
    const int k = std::min<int>(cap, K_NUM);
    for (int i = 0; i < k; ++i) { out_ms[i] = 0.f; out_launches[i] = 0; }
    for (size_t e = 0; e < g_kt.used; ++e)
    {
        float ms = 0.f;
        MMG_CUDA(cudaEventElapsedTime(&ms, g_kt.pool[2 * e], g_kt.pool[2 * e + 1]));
        if (g_kt.slotOf[e] < k) { out_ms[g_kt.slotOf[e]] += ms; out_launches[g_kt.slotOf[e]] += g_kt.countOf[e]; }
    }
    g_kt.used = 0;
    if (n) *n = k;
    return 0;
}

int mmgen_work_counters(uint64_t* out32, int reset)
{
    if (requireReady()) return 1;
    MMG_CUDA(cudaDeviceSynchronize());
    static_assert(W_NUM == 32, "mmgen.h documents 32 counters");
    if (out32) MMG_CUDA(cudaMemcpyFromSymbol(out32, g_work, sizeof(unsigned long long) * W_NUM));
    if (reset)
    {
        static const unsigned long long zero[W_NUM] = {};
        MMG_CUDA(cudaMemcpyToSymbol(g_work, zero, sizeof(zero)));
    }
    return 0;
}

const char* mmgen_kernel_name(int slot) { return (slot >= 0 && slot < K_NUM) ? kKernelNames[slot] : ""; }

// self-test of the packed-fp32 noise routines (mm_arith.cuh): every pair evaluation against the two scalar evaluations it stands
// for, bit for bit, at n pseudo-random positions of the magnitudes the pipeline uses (offsets of thousands, steps of 1e-3 .. 1)
__global__ void k_selftest_packed_noise(int n, unsigned seed, unsigned long long* mismatches)
{
    noise_tab_stage();
    unsigned long long bad = 0ull;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
        Minstd rng = make_rng3(i, (int)(seed & 0x3ff), (int)(seed >> 10));
        const float sc = (i & 1) ? 6000.f : ((i & 2) ? 40.f : 2.f);
        const float ax = rng.u11() * sc, ay = rng.u11() * sc, az = rng.u11() * sc;
        const float bx = ax + 5923.45f, by = ay + 4129.42f, bz = az + 5790.48f;
        auto ne = [](float p, float q) { return __float_as_uint(p) != __float_as_uint(q); };
        {
            const f32x2 r = simplex3x2_raw<false>(f2_make(ax, bx), f2_make(ay, by), f2_make(az, bz));
            bad += ne(f2_lo(r), simplex3_raw<false>(ax, ay, az)) + ne(f2_hi(r), simplex3_raw<false>(bx, by, bz));
            const f32x2 t = simplex3x2_raw<true>(f2_make(ax, bx), f2_make(ay, by), f2_make(az, bz));
            bad += ne(f2_lo(t), simplex3_raw<true>(ax, ay, az)) + ne(f2_hi(t), simplex3_raw<true>(bx, by, bz));
        }
        {
            const f32x2 r = simplex2x2_raw<false>(f2_make(ax, bx), f2_make(az, bz));
            bad += ne(f2_lo(r), simplex2_raw<false>(ax, az)) + ne(f2_hi(r), simplex2_raw<false>(bx, bz));
            const f32x2 t = simplex2x2_raw<true>(f2_make(ax, bx), f2_make(az, bz));
            bad += ne(f2_lo(t), simplex2_raw<true>(ax, az)) + ne(f2_hi(t), simplex2_raw<true>(bx, bz));
        }
        {
            const float px = ax * 0.004f, py = ay * 0.004f, pz = az * 0.004f;
            bad += ne(fbm3_paired<5>(px, py, pz), fbm3<5>(px, py, pz)) + ne(fbm3_paired<4, true>(px, py, pz), fbm3<4, true>(px, py, pz)) +
                   ne(fbm3_paired<3>(px, py, pz), fbm3<3>(px, py, pz));
            float o1, o2, o3;
            fbm3_from3<5>(px, py, pz, &o1, &o2, &o3);
            bad += ne(o1, fbm3<5>(px, py, pz)) + ne(o2, fbm3<5>(px + 5923.45f, py + 4129.42f, pz + 5790.48f)) +
                   ne(o3, fbm3<5>(px + 1765.68f, py + 4704.36f, pz + 5692.12f));
            const f32x2 q = fbm2x2<4>(f2_make(px, bx * 0.01f), f2_make(pz, bz * 0.01f));
            bad += ne(f2_lo(q), fbm2<4>(px, pz)) + ne(f2_hi(q), fbm2<4>(bx * 0.01f, bz * 0.01f));
        }
    }
    if (bad) atomicAdd(mismatches, bad);
}

int mmgen_selftest_packed_noise(int n, uint32_t seed, uint64_t* out_mismatches)
{
    if (requireReady()) return 1;
    if (n <= 0 || !out_mismatches) { g_lastError = "mmgen_selftest_packed_noise: bad arguments"; return 1; }
    unsigned long long* d = nullptr;
    MMG_CUDA(cudaMalloc(&d, sizeof(unsigned long long)));
    MMG_CUDA(cudaMemsetAsync(d, 0, sizeof(unsigned long long), g_stream));
    MMG_LAUNCH(k_selftest_packed_noise, kNumSMs * 4, 256, kNoiseSmemBytes, g_stream, n, (unsigned)seed, d);
    unsigned long long h = 0ull;
    MMG_CUDA(cudaMemcpyAsync(&h, d, sizeof(h), cudaMemcpyDeviceToHost, g_stream));
    MMG_CUDA(cudaStreamSynchronize(g_stream));
    cudaFree(d);
    *out_mismatches = (uint64_t)h;
    return 0;
}

int mmgen_measure_fp32_peak(float* out_tflops)
{
    if (requireReady()) return 1;
    const int blocks = kNumSMs * 8, iters = 4096;
    float* d = nullptr;
    MMG_CUDA(cudaMalloc(&d, (size_t)blocks * 256 * sizeof(float)));
    cudaEvent_t a, b;
    MMG_CUDA(cudaEventCreate(&a));
    MMG_CUDA(cudaEventCreate(&b));
    float best = 0.f;
    for (int rep = 0; rep < 5; ++rep)
    {
        MMG_CUDA(cudaEventRecord(a, g_stream));
        MMG_LAUNCH(k_fp32_peak, blocks, 256, 0, g_stream, d, iters);
        MMG_CUDA(cudaEventRecord(b, g_stream));
        MMG_CUDA(cudaStreamSynchronize(g_stream));
        float ms = 0.f;
        MMG_CUDA(cudaEventElapsedTime(&ms, a, b));
        const double flop = 2.0 * 8 * 16 * (double)iters * blocks * 256;
        if (rep > 0) best = std::max(best, (float)(flop / (ms * 1e-3) / 1e12));
    }
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    cudaFree(d);
    *out_tflops = best;
    return 0;
}

int mmgen_set_serial_stages(int serial)
{
    g_serialStages = serial != 0;
    return 0;
}

int mmgen_set_fill_overlap(int mode)
{
    g_fillOverlap = mode < 0 ? 0 : mode;
    return 0;
}

int mmgen_set_rock_queue_per_chunk(int slots)
{
    g_rockQueuePerChunk = (slots <= 0 || slots > kRockQueuePerChunk) ? kRockQueuePerChunk : slots;
    return 0;
}

// ------------------------------------------------------------------ meshing (Chunk::createVBOs, chunk.cu:1781-2003)
int mmgen_world_mesh(MmgenWorld* w, int n, const int32_t* chunkCoords, int32_t* out_counts)
{
    if (requireReady()) return 1;
    if (n <= 0) return 0;
    if (!w->d_blocks) { g_lastError = "mmgen_world_mesh: nothing filled yet"; return 1; }
    std::vector<MeshChunk> list(n);
    for (int i = 0; i < n; ++i)
    {
        const int x = chunkCoords[2 * i] - w->cx0, z = chunkCoords[2 * i + 1] - w->cz0;
        if (x < 0 || z < 0 || x >= w->nx || z >= w->nz || w->stage[z * w->nx + x] != 6)
        {
            g_lastError = "mmgen_world_mesh: chunk is not filled";
            return 1;
        }
        auto filled = [&](int ax, int az) { return (ax >= 0 && az >= 0 && ax < w->nx && az < w->nz && w->stage[az * w->nx + ax] == 6) ? az * w->nx + ax : -1; };
        list[i].chunk = z * w->nx + x;
        list[i].nb[0] = filled(x, z + 1); list[i].nb[1] = filled(x + 1, z); list[i].nb[2] = filled(x, z - 1); list[i].nb[3] = filled(x - 1, z);
        list[i].origin = w->h_origins[z * w->nx + x];
    }
    if ((size_t)n > w->meshListCap)
    {
        cudaFree(w->d_meshList); cudaFree(w->d_meshColOff); cudaFree(w->d_meshTotals); cudaFree(w->d_meshBase);
        w->d_meshList = nullptr; w->d_meshColOff = nullptr; w->d_meshTotals = nullptr; w->d_meshBase = nullptr; w->meshListCap = 0;
        MMG_CUDA(cudaMalloc(&w->d_meshList, (size_t)n * sizeof(MeshChunk)));
        MMG_CUDA(cudaMalloc(&w->d_meshColOff, (size_t)n * 256 * sizeof(int)));
        MMG_CUDA(cudaMalloc(&w->d_meshTotals, (size_t)n * sizeof(int)));
        MMG_CUDA(cudaMalloc(&w->d_meshBase, (size_t)n * sizeof(long long)));
        w->meshListCap = (size_t)n;
    }
    cudaEvent_t e0 = w->evMesh[0], e1 = w->evMesh[1];
    MMG_CUDA(cudaEventRecord(e0, w->stream));
    MMG_CUDA(cudaMemcpyAsync(w->d_meshList, list.data(), (size_t)n * sizeof(MeshChunk), cudaMemcpyHostToDevice, w->stream));
    MMG_LAUNCH(k_mesh_count, n, 32 * kMeshWarps, 0, w->stream, (const MeshChunk*)w->d_meshList, (const uint8_t*)w->d_blocks, w->d_meshColOff, w->d_meshTotals);
    w->h_meshTotals.resize(n);
    MMG_CUDA(cudaMemcpyAsync(w->h_meshTotals.data(), w->d_meshTotals, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, w->stream));
    MMG_CUDA(cudaStreamSynchronize(w->stream));
    w->h_meshBase.assign(n + 1, 0);
    for (int i = 0; i < n; ++i) w->h_meshBase[i + 1] = w->h_meshBase[i] + w->h_meshTotals[i];
    const size_t need = (size_t)w->h_meshBase[n];
    if (need > w->meshVertCap)
    {
        cudaFree(w->d_meshVerts); cudaFree(w->d_meshIdx);
        w->d_meshVerts = nullptr; w->d_meshIdx = nullptr; w->meshVertCap = 0;
        const size_t cap = need + need / 4 + 1024;
        MMG_CUDA(cudaMalloc(&w->d_meshVerts, cap * sizeof(MeshVertex)));
        MMG_CUDA(cudaMalloc(&w->d_meshIdx, (cap / 4 + 1) * 6 * sizeof(uint32_t)));
        w->meshVertCap = cap;
    }
    MMG_CUDA(cudaMemcpyAsync(w->d_meshBase, w->h_meshBase.data(), (size_t)n * sizeof(long long), cudaMemcpyHostToDevice, w->stream));
    MMG_LAUNCH(k_mesh_emit, n, 32 * kMeshWarps, kMeshStripBytes, w->stream, (const MeshChunk*)w->d_meshList, (const uint8_t*)w->d_blocks, (const int*)w->d_meshColOff,
               (const long long*)w->d_meshBase, w->d_meshVerts, w->d_meshIdx);
    MMG_CUDA(cudaEventRecord(e1, w->stream));
    MMG_CUDA(cudaStreamSynchronize(w->stream));
    MMG_CUDA(cudaEventElapsedTime(&w->meshMs, e0, e1));
    w->h_meshCoords.assign(chunkCoords, chunkCoords + 2 * (size_t)n);
    if (out_counts)
        for (int i = 0; i < n; ++i)
        {
            out_counts[2 * i] = w->h_meshTotals[i];
            out_counts[2 * i + 1] = w->h_meshTotals[i] / 4 * 6;
        }
    return 0;
}

// OptixRenderer::buildChunkAccel (optixRenderer.cpp:223-368) uploads chunkPtr->verts / idx from host vectors (initFromVector,
// :229-230) and describes them to OptiX as one triangle array. The same description for the meshes that already sit in the
// device arena: nothing crosses PCIe, and one optixAccelBuild can take the whole batch.
int mmgen_world_mesh_gas_inputs(MmgenWorld* w, int cap, MmgenGasInput* out, int* n)
{
    const int m = (int)w->h_meshTotals.size();
    if (n) *n = m;
    for (int i = 0; i < m && i < cap; ++i)
    {
        MmgenGasInput g;
        g.vertexBuffer = (uint64_t)(uintptr_t)(w->d_meshVerts + w->h_meshBase[i]);
        g.numVertices = (uint32_t)w->h_meshTotals[i];
        g.vertexStrideInBytes = (uint32_t)sizeof(MeshVertex);
        g.indexBuffer = (uint64_t)(uintptr_t)(w->d_meshIdx + w->h_meshBase[i] / 4 * 6);
        g.numIndexTriplets = (uint32_t)(w->h_meshTotals[i] / 4 * 2);
        g.indexStrideInBytes = 12u;
        g.cx = w->h_meshCoords[2 * i];
        g.cz = w->h_meshCoords[2 * i + 1];
        out[i] = g;
    }
    return 0;
}

int mmgen_world_mesh_download(MmgenWorld* w, int i, MmgenVertex* out_verts, uint32_t* out_idx)
{
    if (i < 0 || i >= (int)w->h_meshTotals.size()) { g_lastError = "mmgen_world_mesh_download: no such chunk in the last mmgen_world_mesh call"; return 1; }
    const long long base = w->h_meshBase[i];
    const int nv = w->h_meshTotals[i];
    if (nv && out_verts) MMG_CUDA(cudaMemcpy(out_verts, w->d_meshVerts + base, (size_t)nv * sizeof(MeshVertex), cudaMemcpyDeviceToHost));
    if (nv && out_idx) MMG_CUDA(cudaMemcpy(out_idx, w->d_meshIdx + base / 4 * 6, (size_t)nv / 4 * 6 * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    return 0;
}

int mmgen_world_mesh_device_ptrs(MmgenWorld* w, int i, void** verts, void** idx, int* nVerts, int* nIdx)
{
    if (i < 0 || i >= (int)w->h_meshTotals.size()) { g_lastError = "mmgen_world_mesh_device_ptrs: no such chunk in the last mmgen_world_mesh call"; return 1; }
    if (verts) *verts = w->d_meshVerts + w->h_meshBase[i];
    if (idx) *idx = w->d_meshIdx + w->h_meshBase[i] / 4 * 6;
    if (nVerts) *nVerts = w->h_meshTotals[i];
    if (nIdx) *nIdx = w->h_meshTotals[i] / 4 * 6;
    return 0;
}

int mmgen_world_mesh_ms(MmgenWorld* w, float* out) { *out = w->meshMs; return 0; }

#ifdef MMG_FEATURE_STATS
// developer build only: surface feature types whose bit is 0 are not rasterised (results then differ from the reference)
extern "C" int mmgen_debug_feature_mask(unsigned mask)
{
    MMG_CUDA(cudaMemcpyToSymbol(g_debugFeatureMask, &mask, sizeof(mask)));
    return 0;
}
// developer build only (nvcc -DMMG_FEATURE_STATS, tools/feature_census.py): read and clear the rasteriser census
extern "C" int mmgen_debug_feature_stats(unsigned long long* out)
{
    MMG_CUDA(cudaDeviceSynchronize());
    MMG_CUDA(cudaMemcpyFromSymbol(out, g_featStats, sizeof(unsigned long long) * 64 * 4));
    static unsigned long long zero[64 * 4];
    MMG_CUDA(cudaMemcpyToSymbol(g_featStats, zero, sizeof(zero)));
    return 0;
}
// out[0] = voxels where huge_zero_mask's proof was wrong (must be 0), out[1] / out[2] = threshold voxels without / with proof,
// out[3] = voxels that went on to the warped specialCaveNoise, out[4] / out[5] = of those: decided by the threshold bounds /
// needing the exact threshold, out[6] = decided wrongly by the bounds (must be 0), out[7] / out[8] = Worley evaluations with
// cells outside / all cells inside the CTA's jitter table
extern "C" int mmgen_debug_huge_stats(unsigned long long* out)
{
    MMG_CUDA(cudaDeviceSynchronize());
    MMG_CUDA(cudaMemcpyFromSymbol(out, g_hugeMismatch, sizeof(unsigned long long)));
    MMG_CUDA(cudaMemcpyFromSymbol(out + 1, g_hugeVoxels, 2 * sizeof(unsigned long long)));
    MMG_CUDA(cudaMemcpyFromSymbol(out + 3, g_caveWarped, sizeof(unsigned long long)));
    MMG_CUDA(cudaMemcpyFromSymbol(out + 4, g_cavePending, 3 * sizeof(unsigned long long)));
    MMG_CUDA(cudaMemcpyFromSymbol(out + 7, g_caveTable, 2 * sizeof(unsigned long long)));
    return 0;
}
#endif
