"""Halo exchange of placement lists between tiles (SURVEY.md 8(e) option B).

CPU: the exchange plan (which rectangles cross between which ranks) and the predictor-based cuts. GPU: four tiles on ONE GPU
exchange their strips through mmgen_world_pack_placements / mmgen_world_unpack_placements (the same calls bench.py puts NCCL
send / recv between) and must reproduce the region generated as one world bit for bit."""
import numpy as np
import pytest


@pytest.fixture(scope="module")
def pkg(mm):
    from mega_minecraft_b200 import sharding, tiling
    return tiling, sharding


@pytest.mark.parametrize("n", [2, 4, 8])
def test_exchange_plan_covers_every_foreign_ring_chunk_once(pkg, n):
    tiling, sharding = pkg
    region = (0, 0, 96, 64)
    tiles = tiling.tiles(*region, n)
    owner = {}
    for r, (x0, z0, nx, nz) in enumerate(tiles):
        for z in range(z0, z0 + nz):
            for x in range(x0, x0 + nx):
                owner[(x, z)] = r
    for r, (x0, z0, nx, nz) in enumerate(tiles):
        plan = sharding.exchange_plan(tiles, r)
        got = {}
        for peer, send, recv in plan:
            # what I receive from the peer is what the peer's plan sends to me
            back = [p for p in sharding.exchange_plan(tiles, peer) if p[0] == r]
            assert len(back) == 1 and back[0][1] == recv and back[0][2] == send
            for z in range(recv[1], recv[1] + recv[3]):
                for x in range(recv[0], recv[0] + recv[2]):
                    assert owner[(x, z)] == peer and (x, z) not in got
                    got[(x, z)] = peer
        ring = {(x, z) for z in range(z0 - 3, z0 + nz + 3) for x in range(x0 - 3, x0 + nx + 3)
                if (x, z) in owner and owner[(x, z)] != r}
        assert set(got) == ring                      # every ring chunk owned by somebody else arrives exactly once


def test_cut_by_cost_equalises_predicted_cost(pkg):
    _, sharding = pkg
    rng = np.random.default_rng(1)
    cost = 1.0 + rng.random((256, 256)) + np.linspace(0, 2, 256)[None, :]        # denser towards +x
    bal = sharding.Balancer((0, 0, 256, 256), 8)
    tiles = bal.cut_by_cost(cost)
    cover = np.zeros((256, 256), np.int32)
    sums = []
    for x0, z0, nx, nz in tiles:
        cover[z0:z0 + nz, x0:x0 + nx] += 1
        sums.append(cost[z0:z0 + nz, x0:x0 + nx].sum())
    assert (cover == 1).all()
    assert max(sums) / (sum(sums) / 8) < 1.02
    equal = [cost[z0:z0 + nz, x0:x0 + nx].sum() for x0, z0, nx, nz in sharding.Balancer((0, 0, 256, 256), 8).tiles()]
    assert max(equal) / (sum(equal) / 8) > 1.1      # the equal tiles it started from are not balanced


@pytest.mark.gpu
def test_exchanged_tiles_are_bit_identical_to_one_world(gen, mm, pkg):
    import torch
    tiling, sharding = pkg
    region = (-20, 37, 30, 22)                       # negative coordinates, not zone aligned
    whole = gen.region_world(*region)
    whole.generate(mm.STAGE_ALL)
    ref = whole.download_region_blocks().reshape(region[3], region[2], 16, 16, 384)
    ref_sum = whole.chunk_hash_sum()
    whole.close()
    tiles = tiling.tiles(*region, 4)
    worlds = [gen.region_world(*t) for t in tiles]
    try:
        caved = []
        for w in worlds:
            w.set_exchange_region(*region)
            w.generate(mm.STAGE_ALL & ~mm.STAGE_FILL)
            st = w.stages()
            caved.append(int((st >= 4).sum()))
            assert int((st == 6).sum()) == 0
        # stages 4 + 5a ran on the own tile and on ring chunks outside the region only
        recompute = [(t[2] + 6) * (t[3] + 6) for t in tiles]
        assert all(c < r for c, r in zip(caved, recompute))
        buf = torch.empty(64 << 20, dtype=torch.uint8, device="cuda")
        moved = 0
        for r, w in enumerate(worlds):
            for peer, send, recv in sharding.exchange_plan(tiles, r):
                n, ok = w.pack_placements(send, buf.data_ptr(), buf.numel())
                assert ok and n >= send[2] * send[3] * 8
                back = [p for p in sharding.exchange_plan(tiles, peer) if p[0] == r][0]
                worlds[peer].unpack_placements(back[2], buf.data_ptr(), n)
                moved += n
        assert moved > 0
        # a buffer that is too small reports the size needed instead of writing
        p0 = sharding.exchange_plan(tiles, 0)[0]
        need, ok = worlds[0].pack_placements(p0[1], buf.data_ptr(), 16)
        assert not ok and need > 16
        total = 0
        for t, w in zip(tiles, worlds):
            w.generate(mm.STAGE_FILL)
            assert int((w.stages() == 6).sum()) == t[2] * t[3]
            b = w.download_region_blocks().reshape(t[3], t[2], 16, 16, 384)
            assert np.array_equal(b, ref[t[1] - region[1]:t[1] - region[1] + t[3], t[0] - region[0]:t[0] - region[0] + t[2]])
            total = (total + w.chunk_hash_sum()) & 0xFFFFFFFFFFFFFFFF
        assert total == ref_sum
    finally:
        for w in worlds:
            w.close()


@pytest.mark.gpu
def test_chunk_costs_match_the_heightfield(gen):
    origins = np.array([[x * 16, z * 16] for z in range(-2, 2) for x in range(40, 44)], np.int32)
    h, w = gen.heightfields(origins)
    c = gen.chunk_costs(origins)
    hi = np.floor(h).astype(np.int64)
    assert np.allclose(c[:, 0], np.maximum(hi, 128).sum(axis=1), rtol=1e-6)
    assert np.allclose(c[:, 1], np.clip(hi, 1, 383).sum(axis=1), rtol=1e-6)
    assert np.allclose(c[:, 2], (1.0 - w[:, :8, :].sum(axis=1)).sum(axis=1), rtol=1e-4, atol=1e-2)
