"""The C-ABI library: it loads, exports every symbol include/mmgen.h declares, and fails loudly
(no CPU fallback) when there is no GPU. No compute calls here."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib(mm):
    mm.build()
    return ctypes.CDLL(mm.lib_path())


def test_exports_every_declared_symbol(lib):
    header = open(os.path.join(ROOT, "include", "mmgen.h")).read()
    names = sorted(set(re.findall(r"\b(mmgen_[a-z_0-9]+)\s*\(", header)))
    assert len(names) >= 20
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_wire_struct_sizes(mm):
    assert mm.CaveLayer.itemsize == 12 and mm.FeaturePlacement.itemsize == 20 and mm.CaveFeaturePlacement.itemsize == 24
    assert mm.FeaturePlacement.fields["x"][1] == 4 and mm.FeaturePlacement.fields["canReplaceBlocks"][1] == 16
    assert mm.CaveFeaturePlacement.fields["layerHeight"][1] == 16 and mm.CaveFeaturePlacement.fields["canReplaceBlocks"][1] == 20


def test_no_cpu_fallback_without_gpu(mm, lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib.mmgen_last_error.restype = ctypes.c_char_p
    assert lib.mmgen_init(0) != 0
    assert b"no CPU fallback" in lib.mmgen_last_error() or b"CUDA" in lib.mmgen_last_error()
    import numpy as np
    out = np.zeros(256, np.float32)
    origins = np.zeros(2, np.int32)
    rc = lib.mmgen_heightfields(1, origins.ctypes.data_as(ctypes.c_void_p), out.ctypes.data_as(ctypes.c_void_p), None)
    assert rc != 0
    with pytest.raises(mm.MmgenError):
        mm.ChunkGen(0)


def test_product_does_not_reference_oracle():
    pkg = os.path.join(ROOT, "mega-minecraft_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".cu", ".cuh", ".h", ".py", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle/" not in text.replace("oracle/tools", "") or f.endswith(".py") is False or "import oracle" not in text, f
                assert "#include \"../../oracle" not in text and "from oracle" not in text and "import oracle" not in text, f


def test_scheduling_knob_defaults_agree(mm):
    """The Python binding restores `FILL_OVERLAP_DEFAULT` after tests and bench runs that change mmgen_set_fill_overlap: it has to be
    the mode the library starts in (csrc/mmgen.cu), and every knob of the header has a binding method."""
    src = open(os.path.join(ROOT, "mega-minecraft_b200", "csrc", "mmgen.cu")).read()
    m = re.search(r"static int g_fillOverlap = (\d+);", src)
    assert m and int(m.group(1)) == mm.FILL_OVERLAP_DEFAULT
    for method in ("set_fill_overlap", "set_serial_stages"):
        assert callable(getattr(mm.ChunkGen, method))
