"""Shared fixtures. `-m "not gpu"`: oracle vs reference-generated golden vectors, host logic, ABI surface.
`-m gpu`: the CUDA path (through the C ABI) against the oracle and the golden vectors."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def golden():
    """Reference-CUDA outputs for the window chunks [-7,19)^2 (tools/make_golden.py)."""
    g = np.load(os.path.join(ROOT, "tests", "golden", "c2_window.npz"))
    x0, z0, nx, nz = (int(v) for v in g["window"])
    origins = np.array([[(x0 + x) * 16, (z0 + z) * 16] for z in range(nz) for x in range(nx)], np.int32)
    return {"g": g, "x0": x0, "z0": z0, "nx": nx, "nz": nz, "origins": origins}


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as orc
    orc.build()
    return orc.Oracle()


@pytest.fixture(scope="session")
def mm():
    import mmgen_loader
    return mmgen_loader.load()


@pytest.fixture(scope="session")
def gen(mm):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    mm.build()
    return mm.ChunkGen(0)


def split_lists(arr, off):
    return [arr[off[i]:off[i + 1]] for i in range(len(off) - 1)]


def same_placements(a, b):
    """Field-wise equality of placement records (padding bytes are not part of the value)."""
    if len(a) != len(b):
        return False
    for f in a.dtype.names:
        if f.startswith("pad"):
            continue
        if not np.array_equal(a[f], b[f]):
            return False
    return True
