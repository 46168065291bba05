// Host side of the MMCH1 chunk format (mm_codec.cuh): encoded delivery of a world's block volumes, the host decoder and
// region files. Included by mmgen.cu inside its extern "C" section's translation unit (like mm_stream.inl).

struct CodecState
{
    unsigned short* d_nRuns = nullptr;      // [kFillBatch][256]
    unsigned* d_sizes = nullptr;            // [kFillBatch]
    unsigned long long* d_offsets = nullptr;   // [kFillBatch]
    int* d_slots = nullptr;                 // [kFillBatch]
    unsigned long long* d_index = nullptr;  // [targets][2]
    unsigned long long* d_used = nullptr;   // arena bytes used
    unsigned long long* d_batchInfo = nullptr;   // [batches][2]
    unsigned long long* h_batchInfo = nullptr;   // pinned mirror
    uint8_t* d_arena = nullptr;
    size_t arenaCap = 0, indexCap = 0, batchCap = 0;
    std::vector<cudaEvent_t> ev;
};

static void codecFree(CodecState* c)
{
    if (!c) return;
    cudaFree(c->d_nRuns); cudaFree(c->d_sizes); cudaFree(c->d_offsets); cudaFree(c->d_slots); cudaFree(c->d_index);
    cudaFree(c->d_used); cudaFree(c->d_batchInfo); cudaFree(c->d_arena);
    if (c->h_batchInfo) cudaFreeHost(c->h_batchInfo);
    for (auto e : c->ev) cudaEventDestroy(e);
    delete c;
}

static int codecEnsure(MmgenWorld* w, size_t targets, size_t batches)
{
    if (!w->codec) w->codec = new CodecState();
    CodecState* c = w->codec;
    if (!c->d_nRuns)
    {
        MMG_CUDA(cudaMalloc(&c->d_nRuns, (size_t)kFillBatch * 256 * sizeof(unsigned short)));
        MMG_CUDA(cudaMalloc(&c->d_sizes, (size_t)kFillBatch * sizeof(unsigned)));
        MMG_CUDA(cudaMalloc(&c->d_offsets, (size_t)kFillBatch * sizeof(unsigned long long)));
        MMG_CUDA(cudaMalloc(&c->d_slots, (size_t)kFillBatch * sizeof(int)));
        MMG_CUDA(cudaMalloc(&c->d_used, sizeof(unsigned long long)));
    }
    if (targets > c->indexCap)
    {
        cudaFree(c->d_index); cudaFree(c->d_arena);
        c->d_index = nullptr; c->d_arena = nullptr; c->indexCap = 0; c->arenaCap = 0;
        MMG_CUDA(cudaMalloc(&c->d_index, targets * 2 * sizeof(unsigned long long)));
        // worst case (every chunk stored raw) so that the encoder can never run out of room
        MMG_CUDA(cudaMalloc(&c->d_arena, targets * (size_t)kChunkBytes));
        c->indexCap = targets; c->arenaCap = targets * (size_t)kChunkBytes;
    }
    if (batches > c->batchCap)
    {
        cudaFree(c->d_batchInfo);
        if (c->h_batchInfo) cudaFreeHost(c->h_batchInfo);
        c->d_batchInfo = nullptr; c->h_batchInfo = nullptr; c->batchCap = 0;
        MMG_CUDA(cudaMalloc(&c->d_batchInfo, batches * 2 * sizeof(unsigned long long)));
        MMG_CUDA(cudaMallocHost(&c->h_batchInfo, batches * 2 * sizeof(unsigned long long)));
        c->batchCap = batches;
    }
    while (c->ev.size() < batches)
    {
        cudaEvent_t e;
        MMG_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        c->ev.push_back(e);
    }
    return 0;
}

// encodes the filled chunks list[0..m) (device list dl, region slots h_slots) of fill batch b into the arena, in stream order
static int codecEncodeBatch(MmgenWorld* w, int b, int m, const int* dl, const int* h_slots)
{
    CodecState* c = w->codec;
    if (b == 0) MMG_CUDA(cudaMemsetAsync(c->d_used, 0, sizeof(unsigned long long), w->stream));
    MMG_CUDA(cudaMemcpyAsync(c->d_slots, h_slots, (size_t)m * sizeof(int), cudaMemcpyHostToDevice, w->stream));
    MMG_LAUNCH(k_encode_count, m, 256, 0, w->stream, dl, (const uint8_t*)w->d_blocks, c->d_nRuns, c->d_sizes);
    MMG_LAUNCH(k_encode_place, 1, kEncodePlaceThreads, 0, w->stream, m, (const int*)c->d_slots, (const unsigned*)c->d_sizes, c->d_used, c->d_offsets, c->d_index,
               c->d_batchInfo + 2 * (size_t)b);
    MMG_LAUNCH(k_encode_emit, m, 256, 0, w->stream, dl, (const uint8_t*)w->d_blocks, (const unsigned short*)c->d_nRuns, (const unsigned*)c->d_sizes,
               (const unsigned long long*)c->d_offsets, c->d_arena);
    MMG_CUDA(cudaMemcpyAsync(c->h_batchInfo + 2 * (size_t)b, c->d_batchInfo + 2 * (size_t)b, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, w->stream));
    MMG_CUDA(cudaEventRecord(c->ev[b], w->stream));
    return 0;
}

// after every batch has been queued: as each batch's encoder finishes, its arena segment goes to the host on the copy stream
// (overlapping the fill of the later batches); the index follows at the end. enc->bytes becomes the payload length.
static int codecDeliver(MmgenWorld* w, int batches, size_t targets, EncodedOut* enc)
{
    CodecState* c = w->codec;
    size_t total = 0;
    bool fits = true;
    for (int b = 0; b < batches; ++b)
    {
        MMG_CUDA(cudaEventSynchronize(c->ev[b]));
        const unsigned long long off = c->h_batchInfo[2 * b], n = c->h_batchInfo[2 * b + 1];
        total = (size_t)(off + n);
        if (total > enc->cap) { fits = false; continue; }      // keep counting: the caller learns the size it needs
        if (n) MMG_CUDA(cudaMemcpyAsync(enc->buf + off, c->d_arena + off, (size_t)n, cudaMemcpyDeviceToHost, w->copyStream));
    }
    enc->bytes = total;
    if (!fits)
    {
        MMG_CUDA(cudaStreamSynchronize(w->copyStream));
        g_lastError = "mmgen_world_generate_to_host_encoded: the encoded region needs " + std::to_string(total) + " bytes, the buffer holds " + std::to_string(enc->cap);
        return 2;
    }
    MMG_CUDA(cudaMemcpyAsync(enc->index, c->d_index, targets * 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, w->copyStream));
    return 0;
}

extern "C" int mmgen_decode_chunk(const uint8_t* enc, size_t nbytes, uint8_t* out_blocks)
{
    if (!enc || !out_blocks) { g_lastError = "mmgen_decode_chunk: null pointer"; return 1; }
    if (nbytes == (size_t)kChunkBytes) { std::memcpy(out_blocks, enc, kChunkBytes); return 0; }
    if (nbytes < (size_t)kCodecHeaderBytes) { g_lastError = "mmgen_decode_chunk: shorter than the header"; return 1; }
    const unsigned short* nRuns = reinterpret_cast<const unsigned short*>(enc);
    const uint8_t* p = enc + kCodecHeaderBytes;
    const uint8_t* end = enc + nbytes;
    for (int col = 0; col < 256; ++col)
    {
        uint8_t* o = out_blocks + (size_t)col * 384;
        int y = 0;
        for (int r = 0; r < nRuns[col]; ++r, p += 2)
        {
            if (p + 2 > end || p[1] == 0 || y + p[1] > 384) { g_lastError = "mmgen_decode_chunk: corrupt run list"; return 1; }
            std::memset(o + y, p[0], p[1]);
            y += p[1];
        }
        if (y != 384) { g_lastError = "mmgen_decode_chunk: a column's runs do not add up to 384"; return 1; }
    }
    return 0;
}

// ---- region files: "MMRG" u32 version=1, i32 rx0, rz0, rnx, rnz, u64 payloadBytes, then index u64[rnz*rnx][2] = {offset, bytes}
// relative to the payload, then the payload (MMCH1 chunks)
struct RegionHeader { char magic[4]; uint32_t version; int32_t rx0, rz0, rnx, rnz; uint64_t payloadBytes; };

extern "C" int mmgen_region_save(const char* path, int rx0, int rz0, int rnx, int rnz, const uint64_t* index, const uint8_t* payload, size_t payloadBytes)
{
    if (!path || !index || !payload || rnx <= 0 || rnz <= 0) { g_lastError = "mmgen_region_save: bad arguments"; return 1; }
    FILE* f = std::fopen(path, "wb");
    if (!f) { g_lastError = std::string("mmgen_region_save: cannot open ") + path; return 1; }
    RegionHeader h = {{'M', 'M', 'R', 'G'}, 1u, rx0, rz0, rnx, rnz, (uint64_t)payloadBytes};
    const size_t n = (size_t)rnx * rnz;
    bool ok = std::fwrite(&h, sizeof(h), 1, f) == 1 && std::fwrite(index, 16, n, f) == n && std::fwrite(payload, 1, payloadBytes, f) == payloadBytes;
    ok = (std::fclose(f) == 0) && ok;
    if (!ok) { g_lastError = std::string("mmgen_region_save: write failed: ") + path; return 1; }
    return 0;
}

struct MmgenRegionFile { FILE* f; RegionHeader h; std::vector<uint64_t> index; std::vector<uint8_t> buf; };

extern "C" int mmgen_region_open(const char* path, MmgenRegionFile** out, int32_t* out_rect4)
{
    FILE* f = std::fopen(path, "rb");
    if (!f) { g_lastError = std::string("mmgen_region_open: cannot open ") + path; return 1; }
    MmgenRegionFile* r = new MmgenRegionFile();
    r->f = f;
    bool ok = std::fread(&r->h, sizeof(r->h), 1, f) == 1 && std::memcmp(r->h.magic, "MMRG", 4) == 0 && r->h.version == 1 && r->h.rnx > 0 && r->h.rnz > 0;
    if (ok)
    {
        r->index.resize((size_t)r->h.rnx * r->h.rnz * 2);
        ok = std::fread(r->index.data(), 16, r->index.size() / 2, f) == r->index.size() / 2;
    }
    if (!ok)
    {
        std::fclose(f);
        delete r;
        g_lastError = std::string("mmgen_region_open: not a region file: ") + path;
        return 1;
    }
    if (out_rect4) { out_rect4[0] = r->h.rx0; out_rect4[1] = r->h.rz0; out_rect4[2] = r->h.rnx; out_rect4[3] = r->h.rnz; }
    *out = r;
    return 0;
}

extern "C" int mmgen_region_read_chunk(MmgenRegionFile* r, int cx, int cz, uint8_t* out_blocks)
{
    const int x = cx - r->h.rx0, z = cz - r->h.rz0;
    if (x < 0 || z < 0 || x >= r->h.rnx || z >= r->h.rnz) { g_lastError = "mmgen_region_read_chunk: chunk outside the file's region"; return 1; }
    const uint64_t off = r->index[2 * ((size_t)z * r->h.rnx + x)], n = r->index[2 * ((size_t)z * r->h.rnx + x) + 1];
    if (n == 0 || off + n > r->h.payloadBytes) { g_lastError = "mmgen_region_read_chunk: corrupt index"; return 1; }
    r->buf.resize(n);
    const long base = (long)(sizeof(RegionHeader) + r->index.size() * 8);
    if (std::fseek(r->f, base + (long)off, SEEK_SET) != 0 || std::fread(r->buf.data(), 1, n, r->f) != n)
    {
        g_lastError = "mmgen_region_read_chunk: read failed";
        return 1;
    }
    return mmgen_decode_chunk(r->buf.data(), n, out_blocks);
}

extern "C" int mmgen_region_close(MmgenRegionFile* r)
{
    if (!r) return 0;
    std::fclose(r->f);
    delete r;
    return 0;
}
