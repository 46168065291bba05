"""The placement-culling tables of the CUDA path (csrc/mm_tables.cuh: c_featureReach, c_caveFeatureReach,
c_caveFeatureBand) against the rasterisers themselves: the oracle's place_feature / place_cave_feature are
brute-forced over a box of voxels around random placements of every type, and no filled voxel may lie
outside the reach / band the product culls with. (The tables are derived from each rasteriser's own first
rejection test; this is the independent check.)"""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TABLES = open(os.path.join(ROOT, "mega-minecraft_b200", "csrc", "mm_tables.cuh")).read()


def _table(name):
    m = re.search(name + r"\[[^\]]*\](?:\[[^\]]*\])?\s*=\s*\{(.*?)\};", TABLES, re.S)
    body = m.group(1).replace("kReachAll", str(1 << 20))
    return [int(v) for v in re.findall(r"-?\d+", body)]


FEATURE_REACH = _table("c_featureReach")
CAVE_REACH = _table("c_caveFeatureReach")
CAVE_BAND = np.array(_table("c_caveFeatureBand")).reshape(-1, 4)
HB = np.array(_table("c_featureHeightBounds")).reshape(-1, 2)
CHB = np.array(_table("c_caveFeatureHeightBounds")).reshape(-1, 2)


def _extent(oracle, cave, placements, radius, ylo, yhi):
    p = np.ascontiguousarray(placements, np.int32)
    out = np.zeros((len(p), 7), np.int32)
    ylo, yhi = np.ascontiguousarray(ylo, np.int32), np.ascontiguousarray(yhi, np.int32)
    oracle.L.mmo_feature_extent(int(cave), len(p), p.ctypes.data_as(ctypes.c_void_p), int(radius), ylo.ctypes.data_as(ctypes.c_void_p),
                                yhi.ctypes.data_as(ctypes.c_void_p), out.ctypes.data_as(ctypes.c_void_p), oracle.nthreads)
    return out


def test_tables_have_one_entry_per_type():
    assert len(FEATURE_REACH) == 21 and len(CAVE_REACH) == 10 and CAVE_BAND.shape == (10, 4)
    assert HB.shape == (21, 2) and CHB.shape == (10, 2)


@pytest.mark.parametrize("feature", range(1, 21))
def test_surface_feature_reach(oracle, feature):
    rng = np.random.default_rng(100 + feature)
    n = 6 if FEATURE_REACH[feature] > 30 else 16
    # sea-level dependent types (coral, kelp, iceberg) only rasterise when placed deep enough
    ys = rng.integers(60, 100, n) if feature in (2, 3, 4) else rng.integers(70, 170, n)
    pl = np.stack([np.full(n, feature), rng.integers(-3000, 3000, n), ys, rng.integers(-3000, 3000, n), np.zeros(n, np.int64)], axis=1)
    reach = FEATURE_REACH[feature]
    box = min(reach + 6, 80)
    out = _extent(oracle, 0, pl, box, np.maximum(pl[:, 2] + HB[feature, 0] - 3, 0), np.minimum(pl[:, 2] + HB[feature, 1] + 3, 383))
    assert out[:, 0].sum() > 0, "no placement of this type rasterised: the check would be vacuous"
    assert out[:, 1].max() <= reach and out[:, 2].max() <= reach


@pytest.mark.parametrize("feature", range(1, 10))
def test_cave_feature_reach_and_band(oracle, feature):
    rng = np.random.default_rng(200 + feature)
    n = 24
    lh = rng.integers(1, 60, n)
    lh[:4] = (1, 2, 3, 120)
    pl = np.stack([np.full(n, feature), rng.integers(-3000, 3000, n), rng.integers(5, 120, n), rng.integers(-3000, 3000, n), lh], axis=1)
    reach = CAVE_REACH[feature]
    # the reference only offers voxels inside [y + lo, y + layerHeight + hi] to the rasteriser (chunk.cu:1486-1491)
    ylo = np.maximum(pl[:, 2] + CHB[feature, 0], 0)
    yhi = np.minimum(pl[:, 2] + lh + CHB[feature, 1], 383)
    out = _extent(oracle, 1, pl, reach + 6, ylo, yhi)
    hit = out[:, 0] > 0
    assert hit.any(), "no placement of this type rasterised: the check would be vacuous"
    assert out[:, 1].max() <= reach and out[:, 2].max() <= reach
    a, a_ceil, b, b_ceil = CAVE_BAND[feature]
    lo_seen = np.where(a_ceil, out[:, 5], out[:, 3])[hit]      # relative to the ceiling or to the floor
    hi_seen = np.where(b_ceil, out[:, 6], out[:, 4])[hit]
    assert lo_seen.min() >= a and hi_seen.max() <= b
