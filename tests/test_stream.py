"""The streaming scheduler (Terrain::tick re-hosted on the device-resident world, mm_stream.inl): BASELINE.json's
config 3 (64x64-chunk streaming region at the player-render-distance load pattern).

CPU: the scheduler model (oracle/terrain_model.py) keeps the reference's batch caps and reaches the expected region.
GPU: the C++ scheduler agrees with the model tick by tick, and the chunks it fills - in spiral order, 62 at a time -
are bit-identical to the same chunks generated as one batch world."""
import numpy as np
import pytest

from oracle import terrain_model as tm

DT = 1.0 / 32.0      # exactly representable: the float (C ABI) and double (model) budgets refill identically


def expected_filled_range(cx0, nx, pc, R):
    """Chunks [lo, hi) along one axis that a stream confined to [cx0, cx0+nx) with the player at chunk pc and
    generation radius R can fill: existing chunks, minus the layer ring, eroded zones only, minus 3 for placements."""
    a, b = max(cx0, pc - R), min(cx0 + nx, pc + R + 1)          # chunks that exist
    a, b = a + 1, b - 1                                          # with layers
    zones = [z for z in range(a // 12 - 1, b // 12 + 2) if 12 * z - 6 >= a and 12 * z + 18 <= b]
    if not zones:
        return 0, 0
    return 12 * zones[0] + 3, 12 * zones[-1] + 12 - 3


def test_spiral_matches_reference_shape():
    s = tm.spiral(40)
    assert s[0] == (0, 0) and len(set(s)) == len(s)
    assert set(s) >= {(x, z) for x in range(-40, 41) for z in range(-40, 41)}       # covers the whole generation square
    r = [max(abs(x), abs(z)) for x, z in s]
    assert all(b >= a - 1 for a, b in zip(r, r[1:]))                               # rings grow outwards
    assert tm.spiral(16)[:9] == s[:9]


def test_model_keeps_reference_batch_caps_and_reaches_the_region():
    m = tm.TerrainModel(-41, -41, 82, 82)
    log = m.run_until_idle()
    assert log[-1]["idle"]
    assert max(t["heightfields"] for t in log) <= 166 and max(t["layers"] for t in log) <= 100
    assert max(t["caves"] for t in log) <= 62 and max(t["filled"] for t in log) <= 62 and max(t["zonesEroded"] for t in log) <= 1
    # an erosion takes the whole frame budget: nothing else runs in that tick before it (terrain.cpp:79, 790-812)
    for t in log:
        if t["zonesEroded"]:
            assert t["filled"] == t["caves"] == t["placements"] == t["gatherPlacements"] == t["vbos"] == 0
    lo, hi = expected_filled_range(-41, 82, 0, 40)
    assert (lo, hi) == (-21, 21)
    assert m.filled() == sorted((x, z) for x in range(lo, hi) for z in range(lo, hi))
    # chunks are filled nearest-first: the first filled chunk is within a zone of the player
    assert max(abs(m.filled_order[0][0]), abs(m.filled_order[0][1])) <= 12


KEYS = ("heightfields", "gatherHeightfields", "layers", "zonesEroded", "caves", "placements", "gatherPlacements", "filled", "vbos")


@pytest.mark.parametrize("moves,window", [([(0, 0)], (-41, -41, 82, 82)), ([(0, 0), (24, 0)], (-41, -41, 106, 82)), ([(-7, 5)], (-48, -36, 82, 82))])
def test_model_equals_the_real_terrain_tick(moves, window):
    """Pins the scheduler model to the reference's OWN scheduler: the unmodified terrain.cpp (oracle/_ref/libmmref_terrain.so,
    built by oracle/Makefile with g++ and one force-included compatibility header) is run headless with the generation kernels
    switched off - its batch sizes, orders and budgets do not depend on the data - and must agree with oracle/terrain_model.py
    tick by tick in all nine per-stage counts and in the order chunks are filled, for a standing player, a walking player
    and a player away from the origin. The one thing taken FROM the run is the order in which the reference erodes zones that
    become ready in the same update: it iterates an unordered_set<Zone*> there (heap-address order)."""
    from oracle import refterrain
    if not refterrain.available():
        pytest.skip("oracle/_ref/libmmref_terrain.so not built (needs /root/reference at build time)")
    ref = refterrain.run_session(moves, DT)
    m = tm.TerrainModel(*window, zone_order=[tuple(z) for z in ref["eroded"]])
    log = []
    for i, mv in enumerate(moves):
        m.set_player_chunk(*mv)
        seg = m.run_until_idle(DT)
        # the reference run keeps ticking a few idle ticks after each move (that is how the driver detects idleness)
        end = ref["segments"][i + 1] if i + 1 < len(moves) else len(ref["ticks"])
        log += [[t[k] for k in KEYS] for t in seg]
        log += [[0] * 9] * (end - len(log))
    assert len(log) == len(ref["ticks"])
    bad = [i for i, (a, b) in enumerate(zip(log, ref["ticks"])) if a != b]
    assert not bad, "first tick that differs: %d model %s reference %s" % (bad[0], log[bad[0]], ref["ticks"][bad[0]])
    assert [list(c) for c in m.filled_order] == ref["filled"]                      # the order chunks are filled in
    assert sum(t[7] for t in ref["ticks"]) == len(m.filled()) > 1500
    assert max(t[0] for t in ref["ticks"]) == 166 and max(t[2] for t in ref["ticks"]) == 100 and max(t[7] for t in ref["ticks"]) == 62


def test_model_player_walk_extends_the_filled_region():
    m = tm.TerrainModel(-41, -41, 106, 82)
    m.run_until_idle()
    n0 = len(m.filled())
    m.set_player_chunk(24, 0)
    log = m.run_until_idle()
    assert sum(t["filled"] for t in log) == len(m.filled()) - n0 > 0
    lo, hi = expected_filled_range(-41, 106, 24, 40)
    assert {c[0] for c in m.filled()} == set(range(-21, hi))


@pytest.mark.gpu
@pytest.mark.parametrize("costs", ["reference", "b200"])
def test_stream_matches_model_and_batch_world(gen, mm, costs):
    """C3 in small: session window of 58x58 chunks, generation radius 28 (the reference uses 40), player at the origin."""
    win, R = (-29, -29, 58, 58), 28
    kw = {}
    if costs == "b200":      # a frame budget large enough that every queue drains in one tick
        kw = dict(max_per_frame=1 << 24, per_second=1 << 30)
    t = mm.Terrain(gen, *win)
    t.set_radii(8, R)
    if kw:
        t.set_costs(mm.REFERENCE_COSTS, kw["max_per_frame"], kw["per_second"])
    model = tm.TerrainModel(*win, vbos_gen_radius=8, max_gen_radius=R, **kw)
    for k in range(100000):
        a, b = t.tick(DT).as_dict(), model.tick(DT)
        for f in b:
            assert a[f] == b[f], (k, f, a, b)
        if b["idle"]:
            break
    if costs == "b200":
        assert k < 40
    filled = t.take_filled(1 << 16)
    assert [tuple(c) for c in filled] == model.filled_order
    states = t.states()
    for (cx, cz), s in model.state.items():
        assert states[cz - win[1], cx - win[0]] == s
    lo, hi = expected_filled_range(win[0], win[2], 0, R)
    assert len(filled) == (hi - lo) ** 2 > 0
    # the same chunks as one batch world: tiling-invariant hash of (coordinates, blocks) and a direct comparison of a few chunks
    world = gen.region_world(lo, lo, hi - lo, hi - lo)
    world.generate(mm.STAGE_ALL)
    assert world.chunk_hash_sum() == t.chunk_hash_sum()
    ref = world.download_region_blocks().reshape(hi - lo, hi - lo, 16, 16, 384)
    for cx, cz in ((lo, lo), (0, 0), (hi - 1, lo + 2), (-3, hi - 1)):
        assert np.array_equal(t.download_chunk(cx, cz), ref[cz - lo, cx - lo])
    world.close()
    t.close()


@pytest.mark.gpu
def test_stream_player_walk(gen, mm):
    """The player walks two zones east in four steps while the stream is still working; everything that ends up filled
    equals the batch world, and chunks are never generated twice."""
    win, R = (-29, -29, 82, 58), 28
    t = mm.Terrain(gen, *win)
    t.set_radii(8, R)
    t.set_costs(mm.REFERENCE_COSTS, 4000, 60 * 4000)      # 8x the reference frame budget, same relative costs
    model = tm.TerrainModel(*win, vbos_gen_radius=8, max_gen_radius=R, max_per_frame=4000, per_second=60 * 4000)
    total_hf = 0
    for step in range(5):
        t.setCurrentChunkPos(6 * step, 0)
        model.set_player_chunk(6 * step, 0)
        for k in range(40 if step < 4 else 100000):
            a, b = t.tick(DT).as_dict(), model.tick(DT)
            assert all(a[f] == b[f] for f in b), (step, k, a, b)
            total_hf += a["heightfields"]
            if b["idle"]:
                break
    assert b["idle"]
    assert total_hf == len(model.state)                  # one heightfield launch slot per chunk that exists
    xs = sorted({c[0] for c in model.filled()})
    zs = sorted({c[1] for c in model.filled()})
    assert len(model.filled()) == len(xs) * len(zs)      # a rectangle
    world = gen.region_world(xs[0], zs[0], len(xs), len(zs))
    world.generate(mm.STAGE_ALL)
    assert world.chunk_hash_sum() == t.chunk_hash_sum()
    world.close()
    t.close()


@pytest.mark.gpu
def test_stream_errors(gen, mm):
    t = mm.Terrain(gen, 0, 0, 30, 30)
    with pytest.raises(mm.MmgenError):
        t.set_radii(8, 4)
    with pytest.raises(mm.MmgenError):
        t.set_costs(mm.REFERENCE_COSTS, 0, 1)
    with pytest.raises(mm.MmgenError):
        t.download_chunk(3, 3)          # not filled yet
    assert t.tick(0.0).as_dict()["heightfields"] == 0     # no budget, nothing happens
    t.close()


@pytest.mark.gpu
def test_stream_meshes_drawable_chunks(gen, mm):
    """With meshing on, the VBO stage runs createVBOs on the device for the chunks it makes DRAWABLE; the vertex totals
    equal a direct mesh of the same chunks afterwards (all four neighbours are filled by then, as in the reference)."""
    win, R = (-29, -29, 58, 58), 28
    t = mm.Terrain(gen, *win)
    t.set_radii(6, R)
    t.set_meshing(True)
    t.set_costs(mm.REFERENCE_COSTS, 4000, 60 * 4000)
    log = t.run_until_idle(DT)
    states = t.states()
    drawable = np.argwhere(states == mm.chunkgen.CHUNK_DRAWABLE)
    assert len(drawable) == sum(s["vbos"] for s in log) == 13 * 13
    total = sum(s["meshVertices"] for s in log)
    assert total > 0
    # the same chunks meshed in one call on a batch world of the same region
    lo, hi = expected_filled_range(win[0], win[2], 0, R)
    world = gen.region_world(lo, lo, hi - lo, hi - lo)
    world.generate(mm.STAGE_ALL)
    coords = np.array([[win[0] + x, win[1] + z] for z, x in drawable], np.int32)
    counts = world.mesh(coords, download=False)
    assert int(counts[:, 0].sum()) == total
    world.close()
    t.close()
