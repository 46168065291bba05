"""Parity of the CUDA path (called through the C ABI, include/mmgen.h) against
 (a) the golden vectors produced by the UNMODIFIED reference CUDA pipeline, and
 (b) the CPU oracle on other seeded windows (S1 on all 24 surface biomes, the whole pipeline on four of them;
     tests/test_reference_tour.py holds the whole pipeline to the reference's CUDA generator on all 24),
plus size-independent properties at larger sizes. Integer / byte outputs: bit-exact. Heights and
layers: bit-exact against the oracle; <= 1e-5 relative against the reference (north_star)."""
import numpy as np
import pytest

from conftest import same_placements, split_lists

pytestmark = pytest.mark.gpu
UNWRITTEN = 0x7FC0DEAD

# chunks where one surface biome has weight 1 over the whole chunk (found with the oracle's stage 1)
BIOME_CHUNKS = {0: (-2400, -864), 1: (-2400, -1968), 2: (-2400, -1344), 3: (-2400, 1536), 4: (-2400, -816), 5: (-2400, -912),
                6: (-2112, -1200), 7: (-2256, 1632), 8: (-2400, 672), 9: (-2400, -2352), 10: (-2352, -1776), 11: (-2400, 768),
                12: (-2400, -1776), 13: (-2400, 1008), 14: (-2400, 336), 15: (-2400, -1152), 16: (-2400, 1488), 17: (-2400, -2400),
                18: (-2400, -1536), 19: (-2400, -2256), 20: (-2352, 672), 21: (-2352, -576), 22: (-2400, -240), 23: (-2400, -2208)}


def origins_of(x0, z0, nx, nz):
    return np.array([[(x0 + x) * 16, (z0 + z) * 16] for z in range(nz) for x in range(nx)], np.int32)


# ------------------------------------------------------------------ against the reference's outputs
def test_batch_ops_vs_reference_golden(gen, golden):
    from oracle import oracle as orc
    g, nx, nz, origins = golden["g"], golden["nx"], golden["nz"], golden["origins"]
    h, w = gen.heightfields(origins)
    assert np.array_equal(h.view(np.uint32), g["heightfield"].view(np.uint32))
    assert np.array_equal(w.view(np.uint32), g["biome_weights"].view(np.uint32))
    h18 = orc.gather_h18(g["heightfield"], nx, nz)
    ring = g["ring_idx"]
    lay = gen.layers(origins[ring], np.stack([h18[int(i)] for i in ring]), g["biome_weights"][ring])
    written = g["ring_layers"].view(np.uint32) != UNWRITTEN
    assert np.array_equal(lay.view(np.uint32)[written], g["ring_layers"].view(np.uint32)[written])
    zone = g["zone_idx"]
    caves = gen.caves(origins[zone], g["heightfield"][zone], g["biome_weights"][zone])
    for f in ("start", "end", "bottomBiome", "topBiome"):
        assert np.array_equal(caves[f], g["cave_layers"][f]), f
    F, CF = gen.feature_placements(origins[zone], g["heightfield"][zone], g["biome_weights"][zone], g["zone_layers"], g["cave_layers"])
    rF, rCF = split_lists(g["features"], g["features_off"]), split_lists(g["cave_features"], g["cave_features_off"])
    assert all(same_placements(a, b) for a, b in zip(F, rF))
    assert all(same_placements(a, b[:4096]) for a, b in zip(CF, rCF))
    pos = {int(c): k for k, c in enumerate(zone)}
    lists = {int(c): rF[k] for k, c in enumerate(zone)}
    clists = {int(c): rCF[k] for k, c in enumerate(zone)}
    bidx = g["block_idx"]
    sel = np.array([pos[int(c)] for c in bidx])
    gf = [orc.gather_features(lists, int(c) % nx, int(c) // nx, nx) for c in bidx]
    gcf = [orc.gather_features(clists, int(c) % nx, int(c) // nx, nx) for c in bidx]
    blocks = gen.fill(origins[bidx], g["heightfield"][bidx], g["biome_weights"][bidx], g["zone_layers"][sel], g["cave_layers"][sel], gf, gcf)
    assert np.array_equal(blocks, g["blocks"])


def test_gather_features_op_vs_reference_golden(gen, golden):
    """mmgen_gather_features (Chunk::gatherFeaturePlacements, chunk.cu:1158-1196): the reference's own placement lists of the
    golden zone gathered in the reference's order; feeding the result to mmgen_fill must give the reference's blocks."""
    from oracle import oracle as orc
    g, nx, origins = golden["g"], golden["nx"], golden["origins"]
    zone, bidx = g["zone_idx"], g["block_idx"]
    rF, rCF = split_lists(g["features"], g["features_off"]), split_lists(g["cave_features"], g["cave_features_off"])
    off = gen.gather_offsets()
    assert [tuple(o) for o in off] == [tuple(o) for o in orc.GATHER_OFFSETS]
    pos = {int(c): k for k, c in enumerate(zone)}
    nb = np.array([[pos.get(int(c) + int(dx) + int(dz) * nx, -1) for dx, dz in off] for c in bidx], np.int32)
    assert (nb >= 0).all()                                     # the filled chunks' 7x7 neighbourhoods lie inside the zone
    gf, gcf, counts = gen.gather_features(nb, rF, [c[:4096] for c in rCF])
    lists = {int(c): rF[k] for k, c in enumerate(zone)}
    clists = {int(c): rCF[k][:4096] for k, c in enumerate(zone)}
    for k, c in enumerate(bidx):
        ef = orc.gather_features(lists, int(c) % nx, int(c) // nx, nx)
        ec = orc.gather_features(clists, int(c) % nx, int(c) // nx, nx)
        assert counts[k, 0] == len(ef) and counts[k, 1] == len(ec)
        assert same_placements(gf[k], ef[:2048]) and same_placements(gcf[k], ec[:4096])
    # an absent neighbour contributes nothing
    nb2 = nb[:1].copy()
    nb2[0, 5] = -1
    _, _, c2 = gen.gather_features(nb2, rF, [c[:4096] for c in rCF])
    assert c2[0, 0] == counts[0, 0] - len(rF[nb[0, 5]]) and c2[0, 1] == counts[0, 1] - min(len(rCF[nb[0, 5]]), 4096)
    sel = np.array([pos[int(c)] for c in bidx])
    blocks = gen.fill(origins[bidx], g["heightfield"][bidx], g["biome_weights"][bidx], g["zone_layers"][sel], g["cave_layers"][sel], gf, gcf)
    assert np.array_equal(blocks, g["blocks"])


def test_fill_rejects_counts_beyond_the_stride(gen, mm, golden):
    origins = golden["origins"][:1]
    z = np.zeros
    F = z((1, 4), mm.FeaturePlacement)
    CF = z((1, 4), mm.CaveFeaturePlacement)
    counts = np.array([[5, 0]], np.int32)
    import ctypes
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    rc = gen.L.mmgen_fill(1, p(origins), p(z((1, 256), np.float32)), p(z((1, 24, 256), np.float32)), p(z((1, 20, 256), np.float32)),
                          p(z((1, 256, 32), mm.CaveLayer)), p(F), p(CF), p(counts), 4, 4, p(z((1, 98304), np.uint8)))
    assert rc != 0 and b"placements but the strides" in gen.L.mmgen_last_error()


def test_cave_grid_test_switch_vs_oracle(gen, oracle, golden):
    """mmgen_set_cave_grid_test(1): the source-text reading of tryGenerateCaveFeaturePlacement (grid test honoured). The
    default (0) is the reference as built and is what every other test pins; both readings are checked against the oracle."""
    g, origins = golden["g"], golden["origins"]
    zone = g["zone_idx"][:24]
    args = (origins[zone], g["heightfield"][zone], g["biome_weights"][zone], g["zone_layers"][:24], g["cave_layers"][:24])
    try:
        gen._check(gen.L.mmgen_set_cave_grid_test(1))
        oracle.set_cave_grid_test(True)
        F1, CF1 = gen.feature_placements(*args)
        oF1, oCF1 = oracle.feature_placements(*args)
    finally:
        gen._check(gen.L.mmgen_set_cave_grid_test(0))
        oracle.set_cave_grid_test(False)
    F0, CF0 = gen.feature_placements(*args)
    assert all(same_placements(a, b) for a, b in zip(F1, oF1)) and all(same_placements(a, b) for a, b in zip(CF1, oCF1))
    n0, n1 = sum(len(c) for c in CF0), sum(len(c) for c in CF1)
    assert 0 < n1 < n0                                         # the honoured grid test thins the cave placements out
    rCF = split_lists(g["cave_features"], g["cave_features_off"])[:24]
    assert all(same_placements(a, b[:4096]) for a, b in zip(CF0, rCF))      # the default is the reference build's behaviour


def test_world_mode_vs_reference_golden(gen, mm, golden):
    g = golden["g"]
    world = gen.world(golden["x0"], golden["z0"], golden["nx"], golden["nz"])
    world.generate(mm.STAGE_ALL)
    assert np.array_equal(world.stages(), g["stage"])          # same reachability as the reference state machine
    d = world.download(heightfield=True, biome_weights=True, layers=True, cave_layers=True, blocks=True)
    assert np.array_equal(d["heightfield"].view(np.uint32), g["heightfield"].view(np.uint32))
    zone = g["zone_idx"]
    assert np.array_equal(d["layers"][zone][:, 10:].view(np.uint32), g["zone_layers"][:, 10:].view(np.uint32))   # eroded + backward layers
    for f in ("start", "end", "bottomBiome", "topBiome"):
        assert np.array_equal(d["cave_layers"][zone][f], g["cave_layers"][f]), f
    assert np.array_equal(d["blocks"][g["block_idx"]], g["blocks"])
    # erosion pad semantics: a second generate is idempotent (stages are pure functions of coordinates)
    c1 = world.block_checksum()
    world.generate(mm.STAGE_ALL)
    assert world.block_checksum() == c1
    world.close()


def test_region_world_vs_reference_golden(gen, mm, golden):
    """The apron rule: a world created for the 6x6 target region [3,9)^2 computes exactly the C2 window
    and delivers the same blocks, both by download and by the host-delivery call."""
    g = golden["g"]
    world = gen.region_world(3, 3, 6, 6)
    assert (world.cx0, world.cz0, world.nx, world.nz) == (-7, -7, 26, 26)
    world.generate(mm.STAGE_ALL)
    st = world.stages()
    assert (st == 6).sum() == 36 and (st >= 4).sum() == 144
    blocks = world.download_region_blocks()
    assert np.array_equal(blocks, g["blocks"])
    host = np.zeros_like(blocks)
    world.reset()
    assert world.stages().max() == 0
    world.generate_to_host(host.ctypes.data, mm.STAGE_ALL)
    assert np.array_equal(host, g["blocks"])
    assert world.total_ms() > 0
    world.close()


def test_tiles_are_bit_identical_to_one_world(gen, mm):
    """Multi-GPU tiling property on one GPU: the tiles of a region, generated independently, reproduce
    the region generated as one world (checksums of the block volumes and the bytes themselves)."""
    from mega_minecraft_b200 import tiling
    region = (-5, 20, 10, 8)
    whole = gen.region_world(*region)
    whole.generate(mm.STAGE_ALL)
    ref = whole.download_region_blocks().reshape(region[3], region[2], 16, 16, 384)
    ref_sum = whole.chunk_hash_sum()
    whole.close()
    total = 0
    for t in tiling.tiles(*region, 4):
        w = gen.region_world(*t)
        w.generate(mm.STAGE_ALL)
        b = w.download_region_blocks().reshape(t[3], t[2], 16, 16, 384)
        total = (total + w.chunk_hash_sum()) & 0xFFFFFFFFFFFFFFFF
        w.close()
        assert np.array_equal(b, ref[t[1] - region[1]:t[1] - region[1] + t[3], t[0] - region[0]:t[0] - region[0] + t[2]])
    assert total == ref_sum                         # the tiling-invariant world hash bench.py reports


def test_adapter_runs_reference_state_machine_on_new_kernels(gen, golden):
    """Drop-in proof: the reference's own Chunk/Zone state machine, CPU feature placement and gather
    (unmodified objects from oracle/_ref) with integration/chunk_adapter.cpp linked over its five batch entry points and
    the two CPU passes of stage 5 reproduces the reference's blocks bit for bit."""
    from oracle import refcuda
    if not refcuda.adapter_available():
        pytest.skip("oracle/_ref/libmmref_adapter.so not built (needs /root/reference at build time)")
    g = golden["g"]
    ra = refcuda.RefCuda(0, adapter=True)
    r = ra.generate(golden["x0"], golden["z0"], golden["nx"], golden["nz"], 6)
    # all seven overridden entry points ran: S1, S2, S3, S4, S6 and the two CPU passes of S5 (placements per chunk, gather per chunk)
    calls = [ra.L.mmadapter_calls(i) for i in range(7)]
    assert all(c > 0 for c in calls) and calls[5] >= 144 and calls[6] >= 36, calls
    assert np.array_equal(r["stage"], g["stage"])
    assert np.array_equal(r["heightfield"].view(np.uint32), g["heightfield"].view(np.uint32))
    assert np.array_equal(r["layers"][g["zone_idx"]][:, 10:].view(np.uint32), g["zone_layers"][:, 10:].view(np.uint32))
    for f in ("start", "end", "bottomBiome", "topBiome"):
        assert np.array_equal(r["cave_layers"][f], g["cave_layers"][f]), f
    rF = split_lists(g["features"], g["features_off"])
    assert all(same_placements(a, b) for a, b in zip(r["features"], rF))
    assert np.array_equal(r["block_idx"], g["block_idx"])
    assert np.array_equal(r["blocks"], g["blocks"])


# ------------------------------------------------------------------ against the oracle, other windows
@pytest.mark.parametrize("biome", sorted(BIOME_CHUNKS))
def test_stage1_all_biomes_vs_oracle(gen, oracle, biome):
    cx, cz = BIOME_CHUNKS[biome]
    origins = origins_of(cx - 2, cz - 2, 5, 5)
    h, w = gen.heightfields(origins)
    oh, ow = oracle.heightfields(origins)
    assert w[12, biome].min() == 1.0                            # the window really is that biome
    assert np.array_equal(w.view(np.uint32), ow.view(np.uint32))
    # the three biomes whose height function calls powf (ARCHIPELAGO, SPARSE_DESERT, MOUNTAINS; biomeFuncs.hpp:235,311,375):
    # powf goes through MUFU.RCP on the GPU (oracle/mm_devmath.h:dm_rcp_approx), last-bit differences allowed there
    tol_bits = 2 if biome in (1, 13, 23) else 0
    diff = np.abs(h.view(np.int32).astype(np.int64) - oh.view(np.int32))
    assert diff.max() <= tol_bits, int(diff.max())


@pytest.mark.parametrize("biome", [6, 12, 16, 18])
def test_full_pipeline_vs_oracle(gen, mm, oracle, biome):
    """Whole pipeline on a 26x26 window around another zone, device-resident world vs oracle stage by stage."""
    from oracle import oracle as orc
    cx, cz = BIOME_CHUNKS[biome]
    zx, zz = (cx // 12) * 12, (cz // 12) * 12
    x0, z0, nx, nz = zx - 7, zz - 7, 26, 26
    origins = origins_of(x0, z0, nx, nz)
    world = gen.world(x0, z0, nx, nz)
    world.generate(mm.STAGE_ALL)
    d = world.download(heightfield=True, biome_weights=True, layers=True, cave_layers=True, blocks=True)
    st = world.stages().ravel()
    assert (st == 6).sum() == 36 and (st >= 3).sum() == 144
    oh, ow = oracle.heightfields(origins)
    assert np.array_equal(d["biome_weights"].view(np.uint32), ow.view(np.uint32))
    assert np.abs(d["heightfield"].view(np.int32).astype(np.int64) - oh.view(np.int32)).max() <= 2
    # from here on feed the oracle the product's own upstream outputs so every stage is judged on identical inputs
    h, w = d["heightfield"], d["biome_weights"]
    h18 = orc.gather_h18(h, nx, nz)
    inner = sorted(h18.keys())
    ol = np.full((nx * nz, 20, 256), np.nan, np.float32)
    ol[inner] = oracle.layers(origins[inner], np.stack([h18[i] for i in inner]), w[inner])
    planes = orc.gather_zone(ol, h, nx, zx - 6 - x0, zz - 6 - z0)
    er, _ = oracle.erode_zone(planes)
    orc.scatter_zone(er, ol, nx, zx - 6 - x0, zz - 6 - z0)
    zone = np.nonzero(st >= 3)[0]
    assert np.array_equal(d["layers"][zone][:, 10:].view(np.uint32), ol[zone][:, 10:].view(np.uint32))
    sub = zone[::4]                                             # caves are the expensive part of the oracle
    oc = oracle.caves(origins[sub], h[sub], w[sub])
    for f in ("start", "end", "bottomBiome", "topBiome"):
        assert np.array_equal(d["cave_layers"][sub][f], oc[f]), f
    F, CF = world.download_features()
    oF, oCF = oracle.feature_placements(origins[zone], h[zone], w[zone], d["layers"][zone], d["cave_layers"][zone])
    assert all(same_placements(F[int(c)], a) for c, a in zip(zone, oF))
    assert all(same_placements(CF[int(c)], a[:4096]) for c, a in zip(zone, oCF))
    lists = {int(c): oF[k] for k, c in enumerate(zone)}
    clists = {int(c): oCF[k] for k, c in enumerate(zone)}
    filled = np.nonzero(st == 6)[0][::5]
    gf = [orc.gather_features(lists, int(c) % nx, int(c) // nx, nx) for c in filled]
    gcf = [orc.gather_features(clists, int(c) % nx, int(c) // nx, nx) for c in filled]
    ob = oracle.fill(origins[filled], h[filled], w[filled], d["layers"][filled], d["cave_layers"][filled], gf, gcf)
    assert np.array_equal(d["blocks"][filled], ob)
    world.close()


# ------------------------------------------------------------------ edge cases and properties
def test_empty_and_ragged_inputs(gen, mm):
    h, w = gen.heightfields(np.zeros((0, 2), np.int32))
    assert h.shape == (0, 256) and w.shape == (0, 24, 256)
    # a window too small for any zone: S1/S2 only, nothing filled, no error
    world = gen.world(0, 0, 5, 3)
    world.generate(mm.STAGE_ALL)
    st = world.stages()
    assert st.max() == 2 and (st == 2).sum() == 3 and st[0, 0] == 1
    world.close()
    # negative coordinates and a non-square window that holds exactly one erodable zone
    world = gen.world(-19, 5, 26, 26)
    world.generate(mm.STAGE_ALL)
    st = world.stages()
    assert (st >= 3).sum() == 144 and (st == 6).sum() == 36
    world.close()


def test_batch_results_independent_of_batching(gen):
    """Chunks are pure functions of their coordinates: any batch split gives the same bytes."""
    origins = origins_of(100, -40, 6, 4)
    h, w = gen.heightfields(origins)
    perm = np.random.default_rng(3).permutation(len(origins))
    h2, w2 = gen.heightfields(origins[perm])
    assert np.array_equal(h[perm], h2) and np.array_equal(w[perm], w2)
    h3, _ = gen.heightfields(origins[:5])
    assert np.array_equal(h[:5], h3)
    caves = gen.caves(origins[:4], h[:4], w[:4])
    caves2 = gen.caves(origins[2:4], h[2:4], w[2:4])
    assert caves[2:4].tobytes() == caves2.tobytes()


def test_cave_layer_invariants_large(gen):
    """C4-style volume (8x8 chunks here): structural invariants of the cave-layer encoding."""
    origins = origins_of(0, 0, 8, 8)
    h, w = gen.heightfields(origins)
    c = gen.caves(origins, h, w)
    start, end = c["start"].astype(np.int64), c["end"].astype(np.int64)
    used = start != 384
    assert (end[used] > start[used]).all()                       # non-empty air runs
    assert ((start[:, :, 1:] > end[:, :, :-1]) | ~used[:, :, 1:]).all()      # sorted, disjoint
    assert (used[:, :, 1:] <= used[:, :, :-1]).all()             # compacted to the front
    last = used.sum(axis=2) - 1
    cols = np.take_along_axis(end, np.maximum(last, 0)[..., None], axis=2)[..., 0]
    assert (cols[last >= 0] == 384).all()                        # the top run is open sky
    assert (c["bottomBiome"] <= 4).all() and (c["topBiome"] <= 4).all()


def test_erosion_properties(gen, oracle):
    """Relaxation only raises layer starts, never above the layer's end, and is idempotent."""
    rng = np.random.default_rng(5)
    base = rng.uniform(100, 140, (384, 384)).astype(np.float32)
    planes = np.empty((9, 384, 384), np.float32)
    planes[8] = base + 12
    for l in range(8):
        planes[l] = base + l * 1.5 * rng.uniform(0, 1, (384, 384)).astype(np.float32)
    planes[:8] = np.minimum.accumulate(planes[::-1], axis=0)[::-1][:8]      # monotone stack
    out, sweeps = gen.erode_zone(planes)
    ref, _ = oracle.erode_zone(planes)
    assert np.array_equal(out.view(np.uint32), ref[:8].view(np.uint32))
    assert (out >= planes[:8] - 1e-3).all()
    again, _ = gen.erode_zone(np.concatenate([out, planes[8:9]]))
    assert np.abs(again - out).max() < 1e-3


def test_c4_cave_and_fill_stress_32x32(gen, mm, oracle):
    """BASELINE.json config 4: 32x32 chunks [0,32)^2 of full 16x384x16 volumes. Size-independent properties over all
    1024 chunks, the batch operators against the resident world on the same chunks (two host paths, one kernel set),
    and the oracle's caves + fill on a sample fed with the product's own upstream outputs."""
    from oracle import oracle as orc
    world = gen.region_world(0, 0, 32, 32)
    world.generate(mm.STAGE_ALL)
    st = world.stages()
    assert (st == 6).sum() == 1024
    d = world.download(heightfield=True, biome_weights=True, layers=True, cave_layers=True, blocks=True)
    nx = world.nx
    filled = np.nonzero(st.ravel() == 6)[0]
    blocks = d["blocks"][filled]
    h = d["heightfield"][filled]
    # y = 0 is bedrock everywhere (chunk.cu:1207-1210); nothing but air above the highest ground + the tallest feature
    bedrock = blocks[0, 0, 0, 0]
    assert (blocks[:, :, :, 0] == bedrock).all()
    top = max(int(np.floor(d["heightfield"].max())), 128) + 1 + 120      # tallest feature bound, featurePlacement.hpp via c_featureHeightBounds
    assert (blocks[:, :, :, min(top + 1, 384):] == 0).all()
    # cave layers: non-empty, sorted, disjoint runs in every column of the region
    cl = d["cave_layers"][filled]
    start, end = cl["start"].astype(np.int64), cl["end"].astype(np.int64)
    used = start != 384
    assert (end[used] > start[used]).all() and ((start[:, :, 1:] > end[:, :, :-1]) | ~used[:, :, 1:]).all()
    # idempotence and the two checksums
    c1, s1 = world.block_checksum(), world.chunk_hash_sum()
    world.reset()
    world.generate(mm.STAGE_ALL)
    assert (world.block_checksum(), world.chunk_hash_sum()) == (c1, s1)
    # batch operators on a sample: same bytes as the resident world
    origins = origins_of(world.cx0, world.cz0, world.nx, world.nz)
    sample = filled[:: 97][:8]
    hb, wb = gen.heightfields(origins[sample])
    assert np.array_equal(hb, d["heightfield"][sample]) and np.array_equal(wb, d["biome_weights"][sample])
    cb = gen.caves(origins[sample], hb, wb)
    assert cb.tobytes() == d["cave_layers"][sample].tobytes()
    F, CF = world.download_features()
    placed = {int(c): F[int(c)] for c in np.nonzero(st.ravel() >= 5)[0]}
    cplaced = {int(c): CF[int(c)] for c in np.nonzero(st.ravel() >= 5)[0]}
    gf = [orc.gather_features(placed, int(c) % nx, int(c) // nx, nx) for c in sample]
    gcf = [orc.gather_features(cplaced, int(c) % nx, int(c) // nx, nx) for c in sample]
    bb = gen.fill(origins[sample], hb, wb, d["layers"][sample], cb, gf, gcf)
    assert np.array_equal(bb, d["blocks"][sample])
    # oracle on the same sample (caves on half of it: the expensive part of the oracle)
    oc = oracle.caves(origins[sample[:4]], hb[:4], wb[:4])
    for f in ("start", "end", "bottomBiome", "topBiome"):
        assert np.array_equal(cb[:4][f], oc[f]), f
    ob = oracle.fill(origins[sample], hb, wb, d["layers"][sample], cb, gf, gcf)
    assert np.array_equal(bb, ob)
    world.close()


def test_packed_noise_equals_scalar_noise(gen):
    """The noise kernels evaluate simplex samples two at a time on the packed fp32 instructions of sm_100 (mm_arith.cuh). Every
    pair routine must return, bit for bit, what the two scalar evaluations return - 4 M positions x 21 comparisons each."""
    for seed in (1, 77):
        assert gen.selftest_packed_noise(1 << 22, seed) == 0


def test_stage_overlap_is_result_neutral(gen, mm, golden):
    """A full generate runs layers + erosion on a side stream while the caves (stage-1 inputs only) run on the main stream. Serial
    and overlapped runs must give the same world - the reference's blocks for the golden window, and identical per-chunk hashes,
    eroded layers and cave layers for a region that spans several erosion zones."""
    hashes = {}
    for serial in (1, 0):
        try:
            gen.L.mmgen_set_serial_stages(serial)
            world = gen.region_world(3, 3, 6, 6)
            world.generate(mm.STAGE_ALL)
            assert np.array_equal(world.download_region_blocks(), golden["g"]["blocks"])
            world.close()
            world = gen.region_world(-20, 5, 40, 30)
            world.generate(mm.STAGE_ALL)
            d = world.download(layers=True, cave_layers=True)
            hashes[serial] = (world.chunk_hashes(), d["layers"].copy(), d["cave_layers"].copy(), world.stages().copy())
            world.close()
        finally:
            gen.L.mmgen_set_serial_stages(0)
    assert hashes[0][0] == hashes[1][0] and len(hashes[0][0]) == 40 * 30
    assert np.array_equal(hashes[0][3], hashes[1][3])
    done = hashes[0][3].ravel() >= 4      # eroded and caved (chunks outside that hold no product)
    assert done.sum() >= 40 * 30
    assert np.array_equal(hashes[0][1][done], hashes[1][1][done]) and np.array_equal(hashes[0][2][done], hashes[1][2][done])


def test_fill_overlap_is_result_neutral(gen, mm, golden):
    """A fill that spans several 2048-chunk batches can run the terrain / rock / lush passes of batch b + 1 on a second stream while the
    placement scan of batch b runs (mmgen_set_fill_overlap). Every mode must give the same chunks: identical per-chunk hashes over a
    region of four batches, twice per mode (the second run reuses the first run's buffers and events), and the reference's blocks for
    the golden window."""
    ref = None
    try:
        for mode in (0, 8, 4, 16 + 8):
            gen.set_fill_overlap(mode)
            world = gen.region_world(-30, 40, 96, 72)
            for rep in range(2):
                world.reset()
                world.generate(mm.STAGE_ALL)
                hs = world.chunk_hashes()
                assert len(hs) == 96 * 72
                if ref is None:
                    ref = hs
                assert hs == ref, "fill overlap mode %d changes the world (run %d)" % (mode, rep)
            world.close()
            world = gen.region_world(3, 3, 6, 6)
            world.generate(mm.STAGE_ALL)
            assert np.array_equal(world.download_region_blocks(), golden["g"]["blocks"])
            world.close()
    finally:
        gen.set_fill_overlap(mm.FILL_OVERLAP_DEFAULT)


def test_rock_queue_overflow_is_result_neutral(gen, mm, golden):
    """k_fill_terrain queues rock voxels for the dense kernel k_fill_rock; voxels that do not fit in the queue are
    finished in place. With a queue far too small the blocks must still equal the reference's."""
    try:
        gen.L.mmgen_set_rock_queue_per_chunk(1500)
        world = gen.region_world(3, 3, 6, 6)
        world.generate(mm.STAGE_ALL)
        assert np.array_equal(world.download_region_blocks(), golden["g"]["blocks"])
        world.close()
    finally:
        gen.L.mmgen_set_rock_queue_per_chunk(0)
