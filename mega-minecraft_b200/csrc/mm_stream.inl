// Streaming scheduler: Terrain::tick (/root/reference/src/terrain/terrain.cpp:587-960) re-hosted on the
// device-resident world. Included by mmgen.cu (host code only; the C linkage of the API comes from mmgen.h).
//
// What is kept from the reference: the ChunkState machine (chunk.hpp:18-32), the spiral visiting order around
// the player (terrain.cpp:220-252, updateChunk :300-417), one FIFO queue per stage, the order in which a tick
// drains them (VBOs, fill, gather placements, placements, caves, erosion, layers, gather heightfield,
// heightfields), the action-time budget (terrain.cpp:67-82, refilled by deltaTime) and therefore the
// reference's batch caps per tick (166 heightfields, 100 layers, 62 caves, 62 fills, 1 zone), the neighbourhood
// conditions of gatherHeightfield (3x3) and gatherFeaturePlacements (7x7) (chunk.cu:53-147), the quadrant
// rule that decides which zones are re-tested for erosion (terrain.cpp:425-453) and the NEEDS_VBOS rule
// (terrain.cpp:568-585).
//
// What changes: a stage call moves nothing over PCIe (products stay in the world's HBM planes; the two gather
// steps are pure state transitions because the kernels read neighbours in place), the costs are a parameter
// (mmgen_stream_set_costs) so that a tick can be as large as the GPU likes, and a zone is eroded only when its
// whole 24x24-chunk gather window has layers. The reference erodes a zone as soon as the neighbour zones THAT
// EXIST are ready (terrain.cpp:488-523), which makes the eroded heights at the rim of the loaded area depend
// on the player's path; here a chunk's result is a pure function of its coordinates (the same as the batch
// world's, tests/test_gpu_parity.py::test_stream_*). The session window bounds the stream: spiral positions
// outside it are ignored.
namespace {

enum StreamState : uint8_t
{
    ST_EMPTY, ST_HAS_HEIGHTFIELD, ST_NEEDS_LAYERS, ST_HAS_LAYERS, ST_NEEDS_EROSION, ST_NEEDS_CAVES, ST_NEEDS_FEATURE_PLACEMENTS,
    ST_NEEDS_GATHER_FEATURE_PLACEMENTS, ST_READY_TO_FILL, ST_FILLED, ST_NEEDS_VBOS, ST_DRAWABLE
};
enum StreamCost
{
    COST_HEIGHTFIELD, COST_GATHER_HEIGHTFIELD, COST_LAYERS, COST_ERODE_ZONE, COST_CAVES, COST_PLACEMENTS, COST_GATHER_PLACEMENTS,
    COST_FILL, COST_VBOS, NUM_COSTS
};

}  // namespace

struct MmgenStream
{
    MmgenWorld* w = nullptr;
    int vbosGenRadius = 16, maxGenRadius = 16 + 24;              // terrain.cpp:64-65
    int maxActionTimePerFrame = 500, totalActionTimePerSecond = 60 * 500;   // terrain.cpp:69-70
    int cost[NUM_COSTS] = {3, 2, 5, 500, 8, 3, 5, 8, 500 / 3};  // terrain.cpp:72-80
    int actionTimeLeft = 0;
    int2 currentChunkPos = {0, 0}, lastChunkPos = {0, 0};
    bool needsUpdateChunks = true;
    std::vector<int2> spiral;
    std::vector<uint8_t> state, ready, exists;
    // zones of the window: zone grid aligned to multiples of 12 in world chunk coordinates
    int zx0 = 0, zz0 = 0, nzx = 0, nzz = 0;
    std::vector<uint8_t> zoneQueued, zoneInTry;
    std::vector<int> zonesToTry;                                 // insertion order (the reference's set is unordered)
    std::queue<int> zonesToErode;
    std::queue<int> qHeightfield, qGatherHeightfield, qLayers, qCaves, qPlacements, qGatherPlacements, qFill, qVbos;
    std::deque<int> newlyFilled;                                 // window indices, for mmgen_stream_take_filled
    cudaEvent_t ev[2] = {};
    uint64_t ticks = 0;
    bool meshing = false;
    std::vector<int32_t> lastMeshed;                              // (cx, cz) pairs of the chunks meshed by the last tick that meshed

    int idx(int cx, int cz) const { return (cz - w->cz0) * w->nx + (cx - w->cx0); }
    bool inWindow(int cx, int cz) const { return cx >= w->cx0 && cx < w->cx0 + w->nx && cz >= w->cz0 && cz < w->cz0 + w->nz; }
    void setState(int i, uint8_t s) { state[i] = s; ready[i] = 1; }     // Chunk::setState, chunk.cu:28-32
};

static int floorDiv12(int a) { return (a >= 0) ? a / 12 : -((-a + 11) / 12); }

// Terrain::generateSpiral (terrain.cpp:220-252) for a given outer radius
static void streamBuildSpiral(MmgenStream* s)
{
    s->spiral.clear();
    int x = 0, z = 0, d = 1, m = 1;
    for (;;)
    {
        while (2 * x * d < m) { s->spiral.push_back(make_int2(x, z)); x += d; }
        if (m > s->maxGenRadius * 2) return;
        while (2 * z * d < m) { s->spiral.push_back(make_int2(x, z)); z += d; }
        d = -d;
        ++m;
    }
}

// Terrain::updateChunk (terrain.cpp:300-417)
static void streamUpdateChunk(MmgenStream* s, int dx, int dz)
{
    const int cx = s->currentChunkPos.x + dx, cz = s->currentChunkPos.y + dz;
    if (!s->inWindow(cx, cz)) return;
    const int i = s->idx(cx, cz);
    if (!s->exists[i]) { s->exists[i] = 1; s->state[i] = ST_EMPTY; s->ready[i] = 1; }
    if (!s->ready[i]) return;
    switch (s->state[i])
    {
    case ST_EMPTY: s->ready[i] = 0; s->qHeightfield.push(i); return;
    case ST_HAS_HEIGHTFIELD: s->ready[i] = 0; s->qGatherHeightfield.push(i); return;
    case ST_NEEDS_LAYERS: s->ready[i] = 0; s->qLayers.push(i); return;
    case ST_NEEDS_CAVES: s->ready[i] = 0; s->qCaves.push(i); return;
    case ST_NEEDS_FEATURE_PLACEMENTS: s->ready[i] = 0; s->qPlacements.push(i); return;
    case ST_NEEDS_GATHER_FEATURE_PLACEMENTS: s->ready[i] = 0; s->qGatherPlacements.push(i); return;
    case ST_READY_TO_FILL: s->ready[i] = 0; s->qFill.push(i); return;
    default: break;
    }
    if (std::max(std::abs(dx), std::abs(dz)) > s->vbosGenRadius) return;
    if (s->state[i] == ST_NEEDS_VBOS) { s->ready[i] = 0; s->qVbos.push(i); }
}

// floodFillAndIterateNeighbors<2R+1 .. > (chunk.cu:53-147) without the data movement: every chunk within R of
// `i` that is in state `cur` and whose (2R+1)^2 neighbourhood is at `cur` or beyond moves to `next`
static void streamGather(MmgenStream* s, int i, int R, uint8_t cur, uint8_t next)
{
    // floodFill (chunk.cu:53-91): breadth-first from chunk i over the four edge neighbours, inside the (4R+1)^2 window around it,
    // through chunks at state `cur` or beyond only - a chunk behind a less advanced one is not found
    const int nx = s->w->nx, nz = s->w->nz;
    const int x0 = i % nx, z0 = i / nx, radius = 2 * R, D = 4 * R + 1;
    // (the reference marks a chunk visited when it is popped, so its queue holds duplicates; marking at push time reaches the
    // same set with every chunk queued once)
    uint8_t found[13 * 13] = {}, queued[13 * 13] = {};      // R <= 3
    int queue[13 * 13], head = 0, tail = 0;
    auto slot = [&](int x, int z) { return (z - z0 + radius) * D + (x - x0 + radius); };
    queue[tail++] = i;
    queued[slot(x0, z0)] = 1;
    while (head < tail)
    {
        const int p = queue[head++];
        const int px = p % nx, pz = p / nx;
        if (s->state[p] < cur) continue;
        found[slot(px, pz)] = 1;
        const int nb[4][2] = {{px, pz + 1}, {px + 1, pz}, {px, pz - 1}, {px - 1, pz}};
        for (const auto& n : nb)
        {
            if (n[0] < 0 || n[1] < 0 || n[0] >= nx || n[1] >= nz) continue;
            const int k = n[1] * nx + n[0];
            if (!s->exists[k] || abs(n[0] - x0) > radius || abs(n[1] - z0) > radius || queued[slot(n[0], n[1])]) continue;
            queued[slot(n[0], n[1])] = 1;
            queue[tail++] = k;
        }
    }
    // iterateNeighborChunks (chunk.cu:93-136): centres within R that are at `cur` and whose (2R+1)^2 neighbourhood was found
    for (int cz = z0 - R; cz <= z0 + R; ++cz)
        for (int cx = x0 - R; cx <= x0 + R; ++cx)
        {
            if (cx < 0 || cz < 0 || cx >= nx || cz >= nz || !found[slot(cx, cz)]) continue;
            const int c = cz * nx + cx;
            if (s->state[c] != cur) continue;
            bool ok = true;
            for (int oz = -R; oz <= R && ok; ++oz)
                for (int ox = -R; ox <= R && ok; ++ox)
                {
                    const int ax = cx + ox, az = cz + oz;
                    ok = ax >= 0 && az >= 0 && ax < nx && az < nz && found[slot(ax, az)];
                }
            if (ok) s->setState(c, next);
        }
}

// Terrain::addZonesToTryErosionSet (terrain.cpp:419-453): the chunk's zone and the three neighbour zones on the
// side of the quadrant the chunk lies in
static void streamAddZonesToTry(MmgenStream* s, int i)
{
    const int nx = s->w->nx;
    const int cx = s->w->cx0 + i % nx, cz = s->w->cz0 + i / nx;
    const int zx = floorDiv12(cx), zz = floorDiv12(cz);
    // own zone, then directions startDirIdx .. startDirIdx + 2 (clockwise from north) with the reference's start table: 4 / 6 for
    // the west half of the zone (south / north), 0 / 2 for the east half (terrain.cpp:436-444). For the east half these are not
    // the zones whose windows contain the chunk; the table is kept as it is (pinned to the real Terrain::tick through the model).
    static const int kDir[8][2] = {{0, 1}, {1, 1}, {1, 0}, {1, -1}, {0, -1}, {-1, -1}, {-1, 0}, {-1, 1}};
    const bool east = (cx - zx * 12) >= 6, north = (cz - zz * 12) >= 6;
    const int start = east ? (north ? 2 : 0) : (north ? 6 : 4);
    int cand[4][2] = {{zx, zz}, {0, 0}, {0, 0}, {0, 0}};
    for (int k = 0; k < 3; ++k) { cand[k + 1][0] = zx + kDir[(start + k) % 8][0]; cand[k + 1][1] = zz + kDir[(start + k) % 8][1]; }
    for (const auto& c : cand)
    {
        const int lx = c[0] - s->zx0, lz = c[1] - s->zz0;
        if (lx < 0 || lz < 0 || lx >= s->nzx || lz >= s->nzz) continue;
        const int z = lz * s->nzx + lx;
        if (s->zoneQueued[z] || s->zoneInTry[z]) continue;
        s->zoneInTry[z] = 1;
        s->zonesToTry.push_back(z);
    }
}

// Terrain::updateZones (terrain.cpp:525-566) with the full-window readiness rule (see the header comment)
static void streamUpdateZones(MmgenStream* s)
{
    const int nx = s->w->nx, nz = s->w->nz;
    for (int z : s->zonesToTry)
    {
        s->zoneInTry[z] = 0;
        const int lx0 = (s->zx0 + z % s->nzx) * 12 - 6 - s->w->cx0, lz0 = (s->zz0 + z / s->nzx) * 12 - 6 - s->w->cz0;
        if (lx0 < 0 || lz0 < 0 || lx0 + 24 > nx || lz0 + 24 > nz) continue;
        bool ok = true;
        for (int dz = 0; dz < 24 && ok; ++dz)
            for (int dx = 0; dx < 24 && ok; ++dx)
            {
                const int k = (lz0 + dz) * nx + lx0 + dx;
                ok = s->exists[k] && s->state[k] >= ST_HAS_LAYERS;
            }
        if (ok) { s->zonesToErode.push(z); s->zoneQueued[z] = 1; }
    }
    s->zonesToTry.clear();
}

// checkChunkAndNeighborsForNeedsVbos (terrain.cpp:568-585)
static void streamCheckNeedsVbos(MmgenStream* s, int i)
{
    const int nx = s->w->nx, nz = s->w->nz;
    if (i < 0) return;
    const int x = i % nx, z = i / nx;
    if (!s->exists[i] || s->state[i] < ST_FILLED) return;
    const int nb[4][2] = {{x + 1, z}, {x - 1, z}, {x, z + 1}, {x, z - 1}};
    for (const auto& n : nb)
    {
        if (n[0] < 0 || n[1] < 0 || n[0] >= nx || n[1] >= nz) return;
        const int k = n[1] * nx + n[0];
        if (!s->exists[k] || s->state[k] < ST_FILLED) return;
    }
    // the reference re-arms the state unconditionally (also for chunks that are already DRAWABLE); a chunk that
    // has been meshed is left alone here
    if (s->state[i] == ST_FILLED) s->setState(i, ST_NEEDS_VBOS);
}

#ifndef MMG_STREAM_MIN_FILL
#define MMG_STREAM_MIN_FILL 64
#endif
// the most chunks one tick can fill under the action-time budget: what the world sizes its fill scratch for
static void streamFillHint(MmgenStream* s)
{
    const long long perTick = s->cost[COST_FILL] > 0 ? (long long)s->maxActionTimePerFrame / s->cost[COST_FILL] : (long long)kFillBatch;
    s->w->fillHint = (size_t)std::min<long long>(std::max<long long>(perTick, MMG_STREAM_MIN_FILL), (long long)kFillBatch);
}

int mmgen_stream_create(int cx0, int cz0, int nx, int nz, MmgenStream** out)
{
    if (requireReady()) return 1;
    if (!out) { g_lastError = "mmgen_stream_create: out is NULL"; return 1; }
    MmgenStream* s = new MmgenStream();
    if (mmgen_world_create(cx0, cz0, nx, nz, &s->w)) { delete s; return 1; }
    s->state.assign(s->w->n, ST_EMPTY);
    s->ready.assign(s->w->n, 0);
    s->exists.assign(s->w->n, 0);
    s->zx0 = floorDiv12(cx0); s->zz0 = floorDiv12(cz0);
    s->nzx = floorDiv12(cx0 + nx - 1) - s->zx0 + 1; s->nzz = floorDiv12(cz0 + nz - 1) - s->zz0 + 1;
    s->zoneQueued.assign((size_t)s->nzx * s->nzz, 0);
    s->zoneInTry.assign((size_t)s->nzx * s->nzz, 0);
    for (auto& e : s->ev) MMG_CUDA(cudaEventCreate(&e));
    streamBuildSpiral(s);
    streamFillHint(s);
    if (worldReserve(s->w)) { mmgen_stream_destroy(s); return 1; }
    *out = s;
    return 0;
}

int mmgen_stream_destroy(MmgenStream* s)
{
    if (!s) return 0;
    for (auto& e : s->ev) if (e) cudaEventDestroy(e);
    mmgen_world_destroy(s->w);
    delete s;
    return 0;
}

int mmgen_stream_world(MmgenStream* s, MmgenWorld** out) { *out = s->w; return 0; }

int mmgen_stream_set_radii(MmgenStream* s, int vbosGenRadius, int maxGenRadius)
{
    if (vbosGenRadius < 0 || maxGenRadius < vbosGenRadius) { g_lastError = "mmgen_stream_set_radii: need 0 <= vbos <= max"; return 1; }
    s->vbosGenRadius = vbosGenRadius; s->maxGenRadius = maxGenRadius;
    streamBuildSpiral(s);
    s->needsUpdateChunks = true;
    return 0;
}

int mmgen_stream_set_costs(MmgenStream* s, const int32_t* costs9, int maxActionTimePerFrame, int totalActionTimePerSecond)
{
    if (maxActionTimePerFrame <= 0 || totalActionTimePerSecond <= 0) { g_lastError = "mmgen_stream_set_costs: budgets must be positive"; return 1; }
    if (costs9)
        for (int k = 0; k < NUM_COSTS; ++k)
        {
            if (costs9[k] < 0) { g_lastError = "mmgen_stream_set_costs: negative cost"; return 1; }
            s->cost[k] = costs9[k];
        }
    s->maxActionTimePerFrame = maxActionTimePerFrame; s->totalActionTimePerSecond = totalActionTimePerSecond;
    streamFillHint(s);
    return worldReserve(s->w);      // a larger budget needs larger fill scratch: now, not in a tick
}

// Terrain::setCurrentChunkPos with chunkPosFromPlayerPos (terrain.cpp:254-257, 1031-1034); block coordinates
int mmgen_stream_set_player(MmgenStream* s, float playerX, float playerZ)
{
    s->currentChunkPos = make_int2((int)floorf(playerX / 16.f), (int)floorf(playerZ / 16.f));
    return 0;
}

int mmgen_stream_tick(MmgenStream* s, float deltaTime, MmgenTickStats* out)
{
    if (requireReady()) return 1;
    MmgenWorld* w = s->w;
    MmgenTickStats st;
    std::memset(&st, 0, sizeof(st));
    if (s->currentChunkPos.x != s->lastChunkPos.x || s->currentChunkPos.y != s->lastChunkPos.y)
    {
        s->lastChunkPos = s->currentChunkPos;
        s->needsUpdateChunks = true;
    }
    if (s->needsUpdateChunks)
    {
        streamUpdateZones(s);
        for (const int2& d : s->spiral) streamUpdateChunk(s, d.x, d.y);
        s->needsUpdateChunks = false;
    }
    s->actionTimeLeft = (int)std::min<long long>((long long)s->actionTimeLeft + (long long)((double)s->totalActionTimePerSecond * deltaTime),
                                                 s->maxActionTimePerFrame);
    const int nx = w->nx;
    MMG_CUDA(cudaEventRecord(s->ev[0], w->stream));
    // createVBOs (terrain.cpp:638-655); buildChunkAccel (the OptiX hand-off) is outside this path
    {
        std::vector<int32_t> coords;
        std::vector<int> picked;
        int budget = s->actionTimeLeft;
        while (!s->qVbos.empty() && budget >= s->cost[COST_VBOS])
        {
            const int i = s->qVbos.front(); s->qVbos.pop();
            picked.push_back(i);
            budget -= s->cost[COST_VBOS];
            coords.push_back(w->cx0 + i % nx);
            coords.push_back(w->cz0 + i / nx);
        }
        // Chunk::createVBOs for the tick's chunks in one pass over the resident blocks (mm_mesh.cuh); the vertex / index
        // arrays stay in the world's device arena until the next tick (mmgen_world_mesh_device_ptrs / _download, chunk i =
        // i-th pair of mmgen_stream_last_meshed). The chunks become DRAWABLE only once their meshes exist: if the mesher
        // fails they go back to the head of the queue and the tick reports the error with the scheduler state unchanged.
        if (s->meshing && !coords.empty())
        {
            std::vector<int32_t> counts(coords.size());
            if (mmgen_world_mesh(w, (int)coords.size() / 2, coords.data(), counts.data()))
            {
                std::queue<int> q;
                for (int i : picked) q.push(i);
                while (!s->qVbos.empty()) { q.push(s->qVbos.front()); s->qVbos.pop(); }
                s->qVbos.swap(q);
                return 1;
            }
            for (size_t k = 0; k < counts.size(); k += 2) st.meshVertices += counts[k];
            s->lastMeshed = coords;
        }
        for (int i : picked) { s->state[i] = ST_DRAWABLE; s->ready[i] = 0; }
        if (!picked.empty()) s->needsUpdateChunks = true;
        s->actionTimeLeft = budget;
        st.vbos = (int)picked.size();
    }
    {
        std::vector<int> list;
        while (!s->qFill.empty() && s->actionTimeLeft >= s->cost[COST_FILL])
        {
            s->needsUpdateChunks = true;
            const int i = s->qFill.front(); s->qFill.pop();
            list.push_back(i);
            s->state[i] = ST_FILLED; s->ready[i] = 0;
            s->actionTimeLeft -= s->cost[COST_FILL];
        }
        if (worldFill(w, list, nullptr, [](int) -> size_t { return 0; })) return 1;
        st.filled = (int)list.size();
        for (int i : list)
        {
            s->newlyFilled.push_back(i);
            const int x = i % nx, z = i / nx;
            streamCheckNeedsVbos(s, i);
            streamCheckNeedsVbos(s, x + 1 < nx ? i + 1 : -1);
            streamCheckNeedsVbos(s, x > 0 ? i - 1 : -1);
            streamCheckNeedsVbos(s, z + 1 < w->nz ? i + nx : -1);
            streamCheckNeedsVbos(s, z > 0 ? i - nx : -1);
        }
    }
    while (!s->qGatherPlacements.empty() && s->actionTimeLeft >= s->cost[COST_GATHER_PLACEMENTS])
    {
        s->needsUpdateChunks = true;
        const int i = s->qGatherPlacements.front(); s->qGatherPlacements.pop();
        streamGather(s, i, 3, ST_NEEDS_GATHER_FEATURE_PLACEMENTS, ST_READY_TO_FILL);
        s->actionTimeLeft -= s->cost[COST_GATHER_PLACEMENTS];
        ++st.gatherPlacements;
    }
    {
        // the reference runs generateFeaturePlacements chunk by chunk on the CPU; here the tick's chunks are one launch
        std::vector<int> list;
        while (!s->qPlacements.empty() && s->actionTimeLeft >= s->cost[COST_PLACEMENTS])
        {
            s->needsUpdateChunks = true;
            const int i = s->qPlacements.front(); s->qPlacements.pop();
            list.push_back(i);
            s->setState(i, ST_NEEDS_GATHER_FEATURE_PLACEMENTS);
            s->actionTimeLeft -= s->cost[COST_PLACEMENTS];
        }
        if (worldPlacements(w, list)) return 1;
        st.placements = (int)list.size();
    }
    {
        std::vector<int> list;
        while (!s->qCaves.empty() && s->actionTimeLeft >= s->cost[COST_CAVES])
        {
            s->needsUpdateChunks = true;
            const int i = s->qCaves.front(); s->qCaves.pop();
            list.push_back(i);
            s->setState(i, ST_NEEDS_FEATURE_PLACEMENTS);
            s->actionTimeLeft -= s->cost[COST_CAVES];
        }
        if (worldCaves(w, list)) return 1;
        st.caves = (int)list.size();
    }
    {
        std::vector<int2> corners;
        while (!s->zonesToErode.empty() && s->actionTimeLeft >= s->cost[COST_ERODE_ZONE])
        {
            s->needsUpdateChunks = true;
            const int z = s->zonesToErode.front(); s->zonesToErode.pop();
            const int lx0 = (s->zx0 + z % s->nzx) * 12 - 6 - w->cx0, lz0 = (s->zz0 + z / s->nzx) * 12 - 6 - w->cz0;
            corners.push_back(make_int2(lx0, lz0));
            for (int dz = 6; dz < 18; ++dz)
                for (int dx = 6; dx < 18; ++dx) s->setState((lz0 + dz) * nx + lx0 + dx, ST_NEEDS_CAVES);
            s->actionTimeLeft -= s->cost[COST_ERODE_ZONE];
        }
        if (worldErode(w, corners)) return 1;      // all zones of the tick relax in lockstep (<= 32 per launch group)
        st.zonesEroded = (int)corners.size();
    }
    {
        std::vector<int> list;
        while (!s->qLayers.empty() && s->actionTimeLeft >= s->cost[COST_LAYERS])
        {
            s->needsUpdateChunks = true;
            const int i = s->qLayers.front(); s->qLayers.pop();
            list.push_back(i);
            s->setState(i, ST_HAS_LAYERS);
            streamAddZonesToTry(s, i);
            s->actionTimeLeft -= s->cost[COST_LAYERS];
        }
        if (worldLayers(w, list)) return 1;
        st.layers = (int)list.size();
    }
    while (!s->qGatherHeightfield.empty() && s->actionTimeLeft >= s->cost[COST_GATHER_HEIGHTFIELD])
    {
        s->needsUpdateChunks = true;
        const int i = s->qGatherHeightfield.front(); s->qGatherHeightfield.pop();
        streamGather(s, i, 1, ST_HAS_HEIGHTFIELD, ST_NEEDS_LAYERS);
        s->actionTimeLeft -= s->cost[COST_GATHER_HEIGHTFIELD];
        ++st.gatherHeightfields;
    }
    {
        std::vector<int> list;
        while (!s->qHeightfield.empty() && s->actionTimeLeft >= s->cost[COST_HEIGHTFIELD])
        {
            s->needsUpdateChunks = true;
            const int i = s->qHeightfield.front(); s->qHeightfield.pop();
            list.push_back(i);
            s->setState(i, ST_HAS_HEIGHTFIELD);
            s->actionTimeLeft -= s->cost[COST_HEIGHTFIELD];
        }
        if (worldHeightfields(w, &list)) return 1;
        st.heightfields = (int)list.size();
    }
    MMG_CUDA(cudaEventRecord(s->ev[1], w->stream));
    MMG_CUDA(cudaStreamSynchronize(w->stream));      // terrain.cpp:934-937
    MMG_CUDA(cudaEventElapsedTime(&st.deviceMs, s->ev[0], s->ev[1]));
    st.actionTimeLeft = s->actionTimeLeft;
    st.idle = !s->needsUpdateChunks && s->zonesToTry.empty() && s->zonesToErode.empty() && s->qHeightfield.empty() &&
              s->qGatherHeightfield.empty() && s->qLayers.empty() && s->qCaves.empty() && s->qPlacements.empty() &&
              s->qGatherPlacements.empty() && s->qFill.empty() && s->qVbos.empty();
    if (st.idle)
    {
        // Safety net, after everything the reference's rules can do is done: its quadrant table (streamAddZonesToTry) re-tests, for
        // chunks in the east half of a zone, zones whose windows do not contain the chunk, so the chunk that completes a zone's
        // window may never put that zone up for its test and the zone stays un-eroded (a hole in the world) until the player
        // moves. When the stream would fall idle, every zone that is not queued yet is put up once more.
        for (int z = 0; z < s->nzx * s->nzz; ++z)
        {
            if (s->zoneQueued[z]) continue;
            const int lx0 = (s->zx0 + z % s->nzx) * 12 - 6 - w->cx0, lz0 = (s->zz0 + z / s->nzx) * 12 - 6 - w->cz0;
            if (lx0 < 0 || lz0 < 0 || lx0 + 24 > w->nx || lz0 + 24 > w->nz) continue;
            bool ok = true;
            for (int dz = 0; dz < 24 && ok; ++dz)
                for (int dx = 0; dx < 24 && ok; ++dx)
                {
                    const int k = (lz0 + dz) * w->nx + lx0 + dx;
                    ok = s->exists[k] && s->state[k] >= ST_HAS_LAYERS;
                }
            if (!ok) continue;
            s->zoneInTry[z] = 1;
            s->zonesToTry.push_back(z);
            s->needsUpdateChunks = true;
            st.idle = 0;
        }
    }
    ++s->ticks;
    if (out) *out = st;
    return 0;
}

// reference ChunkState of every window chunk (raster order); chunks the spiral has not reached yet read EMPTY
int mmgen_stream_states(MmgenStream* s, uint8_t* out)
{
    std::memcpy(out, s->state.data(), s->state.size());
    return 0;
}

// chunk coordinates (cx, cz pairs) of chunks filled since the last call, in fill order; *n = pairs written
int mmgen_stream_take_filled(MmgenStream* s, int32_t* coords, int cap, int* n)
{
    int k = 0;
    while (k < cap && !s->newlyFilled.empty())
    {
        const int i = s->newlyFilled.front(); s->newlyFilled.pop_front();
        coords[2 * k] = s->w->cx0 + i % s->w->nx;
        coords[2 * k + 1] = s->w->cz0 + i / s->w->nx;
        ++k;
    }
    *n = k;
    return 0;
}

// block volume of one filled chunk into host memory (98 304 bytes, y fastest)
int mmgen_stream_download_chunk(MmgenStream* s, int cx, int cz, uint8_t* out_blocks)
{
    if (!s->inWindow(cx, cz) || s->state[s->idx(cx, cz)] < ST_FILLED || !s->w->d_blocks)
    {
        g_lastError = "mmgen_stream_download_chunk: chunk is not filled";
        return 1;
    }
    MMG_CUDA(cudaMemcpyAsync(out_blocks, s->w->d_blocks + (size_t)s->idx(cx, cz) * 98304, 98304, cudaMemcpyDeviceToHost, s->w->stream));
    MMG_CUDA(cudaStreamSynchronize(s->w->stream));
    return 0;
}

// createVBOs on the device for the chunks that leave the VBO queue (off by default: the state transition only)
int mmgen_stream_set_meshing(MmgenStream* s, int enable) { s->meshing = enable != 0; return 0; }

// (cx, cz) pairs of the chunks in the world's mesh arena, in arena order; *n = pairs written
int mmgen_stream_last_meshed(MmgenStream* s, int32_t* coords, int cap, int* n)
{
    const int k = std::min<int>(cap, (int)s->lastMeshed.size() / 2);
    std::memcpy(coords, s->lastMeshed.data(), (size_t)k * 2 * sizeof(int32_t));
    *n = k;
    return 0;
}
