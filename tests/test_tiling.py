"""Host-side logic of the multi-GPU path (no GPU): tiling, apron rule, and the cross-rank reductions over
gloo with world_size 2."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


@pytest.fixture(scope="module")
def pkg(mm):
    from mega_minecraft_b200 import sharding, tiling
    return tiling, sharding


@pytest.mark.parametrize("n", [1, 2, 4, 8])
@pytest.mark.parametrize("region", [(0, 0, 256, 256), (-5, 20, 10, 8), (-100, -37, 64, 96)])
@pytest.mark.parametrize("align", [1, 12])
def test_tiles_partition_the_region(pkg, n, region, align):
    tiling, _ = pkg
    if region[2] < 16 and align == 12:
        pytest.skip("region smaller than the alignment grid")
    ts = tiling.tiles(*region, n, align=align)
    assert len(ts) == n
    cover = np.zeros((region[3], region[2]), np.int32)
    for (x0, z0, nx, nz) in ts:
        assert nx > 0 and nz > 0
        cover[z0 - region[1]:z0 - region[1] + nz, x0 - region[0]:x0 - region[0] + nx] += 1
    assert (cover == 1).all()                      # every chunk in exactly one tile
    if align > 1 and n > 1:
        for (x0, z0, nx, nz) in ts:                # interior cuts sit on the zone grid
            assert x0 == region[0] or x0 % align == 0
            assert z0 == region[1] or z0 % align == 0


def test_grid_for_is_pinned(pkg):
    tiling, _ = pkg
    assert [tiling.grid_for(n) for n in (1, 2, 3, 4, 6, 8)] == [(1, 1), (2, 1), (1, 3), (2, 2), (2, 3), (4, 2)]
    # rank order is row-major: the first `columns` tiles share z0
    ts = tiling.tiles(0, 0, 256, 256, 8)
    assert len({t[1] for t in ts[:4]}) == 1 and ts[4][1] > ts[0][1]


def test_apron_rule(pkg):
    tiling, _ = pkg
    # C2: filling chunks [3,9)^2 needs exactly the reference's C2 window [-7,19)^2 (one zone + pad + layer ring)
    assert tiling.apron_window(3, 3, 6, 6) == (-7, -7, 26, 26)
    c = tiling.stage_chunk_counts(3, 3, 6, 6)
    assert c == {"S1": 676, "S2": 576, "S3_zones": 1, "S4": 144, "S5": 144, "S6": 36}
    # C5 on one GPU
    c = tiling.stage_chunk_counts(0, 0, 256, 256)
    assert c["S6"] == 65536 and c["S3_zones"] == 23 * 23 and c["S4"] == 262 * 262 and c["S1"] == 290 * 290
    # negative coordinates: zones are floor-aligned
    assert tiling.apron_window(-13, -1, 1, 1) == (-24 - 7, -12 - 7, 24 + 14, 24 + 14)


def test_combine_checksums_is_order_sensitive(pkg):
    _, sharding = pkg
    a = sharding.combine_checksums([1, 2, 3])
    assert a != sharding.combine_checksums([3, 2, 1]) and a == sharding.combine_checksums([1, 2, 3])
    assert sharding.reduce_scalar(3.5) == 3.5 and sharding.gather_u64(2 ** 63 + 5) == [2 ** 63 + 5]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world_size, port, region, out):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import mmgen_loader
    mmgen_loader.load()
    from mega_minecraft_b200 import sharding
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    tile = sharding.rank_tile(region, rank, world_size)
    # stand-in for the per-tile block checksum: a function of the tile's chunk coordinates only
    xs, zs = np.meshgrid(np.arange(tile[0], tile[0] + tile[2]), np.arange(tile[1], tile[1] + tile[3]))
    check = int((xs.astype(np.int64) * 73856093 ^ zs.astype(np.int64) * 19349663).sum()) & 0xFFFFFFFFFFFFFFFF | (1 << 63)
    checks = sharding.gather_u64(check)
    slowest = sharding.reduce_scalar(10.0 + rank, "max")
    chunks = sharding.reduce_scalar(tile[2] * tile[3], "sum")

    class FakeGen:      # stage-1 cost features as a function of the chunk origin: the strips must reassemble to the one-rank map
        def chunk_costs(self, origins):
            o = origins.astype(np.float32)
            return np.stack([o[:, 0] + 1000.0, o[:, 1] * 2.0, o[:, 0] - o[:, 1]], axis=1).astype(np.float32)

    cost = sharding.chunk_cost_map(FakeGen(), region, rank, world_size, weights=(1.0, 2.0, 3.0, 0.5))
    tiles = sharding.Balancer(region, world_size).cut_by_cost(np.abs(cost) + 1.0)
    if rank == 0:
        out.put((checks, slowest, chunks, cost, tiles))
    else:
        out.put(tiles)
    dist.destroy_process_group()


def test_two_rank_reductions_over_gloo(pkg):
    tiling, sharding = pkg
    region = (-8, 4, 24, 36)
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, region, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    got = [q.get(), q.get()]
    (checks, slowest, chunks, cost, tiles0), tiles1 = (got[0], got[1]) if len(got[0]) == 5 else (got[1], got[0])
    assert slowest == 11.0 and chunks == region[2] * region[3]

    class FakeGen:
        def chunk_costs(self, origins):
            o = origins.astype(np.float32)
            return np.stack([o[:, 0] + 1000.0, o[:, 1] * 2.0, o[:, 0] - o[:, 1]], axis=1).astype(np.float32)

    single = sharding.chunk_cost_map(FakeGen(), region, 0, 1, weights=(1.0, 2.0, 3.0, 0.5))
    assert cost.shape == (region[3], region[2]) and np.array_equal(cost, single)      # the all-gathered strips = the one-rank map
    assert tiles0 == tiles1                                                            # every rank derives the same cuts
    expect = []
    for t in tiling.tiles(*region, 2):
        xs, zs = np.meshgrid(np.arange(t[0], t[0] + t[2]), np.arange(t[1], t[1] + t[3]))
        expect.append(int((xs.astype(np.int64) * 73856093 ^ zs.astype(np.int64) * 19349663).sum()) & 0xFFFFFFFFFFFFFFFF | (1 << 63))
    assert checks == expect                                           # 64-bit values survive the gather, in rank order
    assert sharding.combine_checksums(checks) == sharding.combine_checksums(expect)


def test_feedback_balancer_converges_and_partitions(pkg):
    """Synthetic cost density with a strong gradient: a few feedback rounds bring the tiles to equal cost, and the
    tiles always partition the region."""
    _, sharding = pkg
    region = (0, 0, 256, 256)
    zz, xx = np.meshgrid(np.arange(256), np.arange(256), indexing="ij")
    density = 1.0 + 0.5 * (zz < 128) + 0.3 * np.sin(xx / 40.0) + 0.2 * (xx > 200)      # cost per chunk

    def cost(t):
        return float(density[t[1]:t[1] + t[3], t[0]:t[0] + t[2]].sum())

    for n in (2, 4, 8):
        b = sharding.Balancer(region, n)
        first = None
        for _ in range(4):
            ts = b.tiles()
            cover = np.zeros((256, 256), np.int32)
            for (x0, z0, nx, nz) in ts:
                assert nx > 0 and nz > 0
                cover[z0:z0 + nz, x0:x0 + nx] += 1
            assert (cover == 1).all()
            imb = b.update([cost(t) for t in ts])
            first = imb if first is None else first
        final = [cost(t) for t in b.tiles()]
        assert max(final) / (sum(final) / len(final)) < 1.03 <= first


def test_cost_map_samples_every_second_chunk_and_cuts_balance(pkg):
    """chunk_cost_map evaluates the stage-1 features at every `stride`-th chunk (the sampled chunk's value stands for its stride x
    stride block) and Balancer.cut_by_cost places the cuts of the 4 x 2 grid so that every tile carries about the same predicted
    cost: for a smooth synthetic cost field the heaviest tile stays within a few per cent of the mean, where equal tiles do not."""
    tiling, sharding = pkg
    region = (-7, 3, 101, 64)
    seen = []

    class FakeGen:
        def chunk_costs(self, origins):
            seen.append(origins.copy())
            x, z = origins[:, 0].astype(np.float64) / 16.0, origins[:, 1].astype(np.float64) / 16.0
            f = 1.0 + 0.8 * np.sin(x / 23.0) * np.cos(z / 17.0) + 0.004 * (x + z)
            return np.stack([f * 1e4, f * 2e4, 256.0 * np.ones_like(f)], axis=1).astype(np.float32)

    w = (1.0, 0.5, 0.25, 0.1)
    full = sharding.chunk_cost_map(FakeGen(), region, weights=w, stride=1)
    half = sharding.chunk_cost_map(FakeGen(), region, weights=w, stride=2)
    assert full.shape == half.shape == (region[3], region[2])
    assert len(seen[0]) == region[2] * region[3] and len(seen[1]) == 51 * 32      # a quarter of the evaluations
    assert np.array_equal(half[::2, ::2], full[::2, ::2])                          # sampled chunks carry their own value ...
    assert np.array_equal(half[1::2, 1::2], full[0:-1:2, 0:-1:2][:half[1::2, 1::2].shape[0], :half[1::2, 1::2].shape[1]])   # ... and lend it to their block
    for cost in (full, half):
        tiles = sharding.Balancer(region, 8).cut_by_cost(cost)
        cover = np.zeros((region[3], region[2]), np.int32)
        loads = []
        for (x0, z0, nx, nz) in tiles:
            cover[z0 - region[1]:z0 - region[1] + nz, x0 - region[0]:x0 - region[0] + nx] += 1
            loads.append(full[z0 - region[1]:z0 - region[1] + nz, x0 - region[0]:x0 - region[0] + nx].sum())
        assert (cover == 1).all()
        assert max(loads) / (sum(loads) / 8) < 1.06
    equal = [full[z0 - region[1]:z0 - region[1] + nz, x0 - region[0]:x0 - region[0] + nx].sum() for (x0, z0, nx, nz) in tiling.tiles(*region, 8)]
    assert max(equal) / (sum(equal) / 8) > 1.10
