"""Known answers away from the golden window (GPU): three regions, all six stages through the device-resident world, held
to the REFERENCE's block volumes.

tests/golden/region_hashes.{json,npz} were made by tools/region_hashes.py on a B200: the unmodified reference CUDA pipeline
(oracle/_ref) generated each region, and its block volumes were hashed per chunk with the function the product computes on
the device (mmgen_world_chunk_hashes: FNV-1a per column, then over (cx, cz, 256 column hashes)). The regions pin what the
golden window cannot: far-away coordinates (hash arguments of 1e7 and more, the Payne-Hanek path of sinf), other biomes,
batches of a few thousand chunks. The product must reproduce every chunk hash except the chunks that held a block flip
when the pins were made - those flips are listed in the json with coordinates, both block IDs and both sides' layer starts
(3 of 2.3e8 voxels and 4 of 1.0e8). All seven sit on the first row / column of a zone whose neighbour the reference driver
eroded earlier: there the reference's eroded heights depend on the zone order (tests/test_reference_tour.py::
test_reference_erosion_depends_on_zone_order_along_seams) and the product, which erodes every zone as
if it were the first, is a few hundredths of a block away. The product must not differ anywhere else."""
import json
import os

import numpy as np
import pytest

from conftest import ROOT

PINS = os.path.join(ROOT, "tests", "golden", "region_hashes.json")
REGIONS = [(0, 0, 48, 48), (-300, 500, 32, 32), (4000, -4000, 24, 36)]


@pytest.mark.gpu
@pytest.mark.parametrize("region", REGIONS)
def test_region_chunk_hashes_match_the_reference(gen, mm, region):
    pins = json.load(open(PINS))["regions"]["%d,%d,%d,%d" % region]
    ref_hashes = np.load(os.path.join(ROOT, "tests", "golden", "region_hashes.npz"))["%d,%d,%d,%d" % region]
    rx0, rz0, rnx, rnz = region
    w = gen.region_world(*region)
    try:
        w.generate(mm.STAGE_ALL)
        w.sync()
        own = w.chunk_hashes()
    finally:
        w.close()
    assert len(own) == rnx * rnz == len(ref_hashes)
    differing = {(rx0 + k % rnx, rz0 + k // rnx) for k in range(rnx * rnz) if own[(rx0 + k % rnx, rz0 + k // rnx)] != int(ref_hashes[k])}
    known = {tuple(f["chunk"]) for f in pins["product_flips_when_pinned"]}
    assert differing <= known, "chunks that differ from the reference and held no known flip: %s" % sorted(differing - known)
    assert len(pins["product_flips_when_pinned"]) / pins["voxels"] < 1e-6
