"""Chunk wire / on-disk format MMCH1 (include/mmgen.h): known-answer decode, corrupt input, region files (CPU: the decoder and
the region file code are host C and need no GPU); device-side encoder and encoded delivery against the reference's golden
blocks (GPU)."""
import numpy as np
import pytest


def np_encode(blocks):
    """Independent restatement of the format in numpy: uint16 nRuns[256], then {block, length <= 255} runs per column, zero
    padded to a multiple of 16 bytes; raw when that would not be shorter."""
    cols = np.ascontiguousarray(blocks, np.uint8).reshape(256, 384)
    counts, runs = np.zeros(256, "<u2"), []
    for c in range(256):
        col = cols[c]
        starts = np.flatnonzero(np.concatenate([[True], col[1:] != col[:-1]]))
        ends = np.concatenate([starts[1:], [384]])
        n = 0
        for s, e in zip(starts, ends):
            length = int(e - s)
            while length > 0:
                take = min(length, 255)
                runs.append((int(col[s]), take))
                length -= take
                n += 1
        counts[c] = n
    body = counts.tobytes() + bytes(b for r in runs for b in r)
    body += b"\0" * (-len(body) % 16)
    return body if len(body) < 98304 else cols.tobytes()


def synthetic_chunk(seed):
    rng = np.random.default_rng(seed)
    b = np.zeros((256, 384), np.uint8)
    for c in range(256):
        h = int(rng.integers(60, 300))
        b[c, 0] = 1
        b[c, 1:h] = rng.choice([57, 73, 60, 5], p=[0.7, 0.1, 0.1, 0.1]) if c % 3 else 57
        b[c, h // 2:h // 2 + int(rng.integers(0, 9))] = 0
        b[c, h:h + 1] = 20
    return b.reshape(16, 16, 384)


def test_decode_known_answer_and_long_runs(mm):
    blocks = synthetic_chunk(1)
    blocks[3, 4, :] = 0                                  # one column that is a single 384-long run: split 255 + 129
    enc = np_encode(blocks)
    assert len(enc) % 16 == 0 and len(enc) < 20000
    assert np.array_equal(mm.decode_chunk(enc), blocks)
    n_runs = np.frombuffer(enc[:512], "<u2")
    assert n_runs[4 + 16 * 3] == 2
    noisy = np.random.default_rng(2).integers(0, 200, (16, 16, 384), dtype=np.uint8)      # incompressible: stored raw
    enc = np_encode(noisy)
    assert len(enc) == 98304 and np.array_equal(mm.decode_chunk(enc), noisy)


def test_decode_rejects_corrupt_input(mm):
    enc = bytearray(np_encode(synthetic_chunk(3)))
    with pytest.raises(mm.MmgenError):
        mm.decode_chunk(bytes(enc[:300]))                                   # shorter than the header
    bad = bytearray(enc)
    bad[512 + 1] = 0                                                        # a zero-length run
    with pytest.raises(mm.MmgenError):
        mm.decode_chunk(bytes(bad))
    bad = bytearray(enc)
    bad[0] = (bad[0] + 1) & 0xFF                                            # one run too many in column 0
    with pytest.raises(mm.MmgenError):
        mm.decode_chunk(bytes(bad))
    with pytest.raises(mm.MmgenError):
        mm.decode_chunk(bytes(enc[:len(enc) // 2]))                         # truncated run list


def test_region_file_round_trip(mm, tmp_path):
    region = (-3, 7, 3, 2)
    chunks = [synthetic_chunk(10 + i) for i in range(6)]
    encs = [np_encode(c) for c in chunks]
    index = np.zeros((6, 2), np.uint64)
    off = 0
    for i, e in enumerate(encs):
        index[i] = (off, len(e))
        off += len(e)
    payload = np.frombuffer(b"".join(encs), np.uint8)
    path = tmp_path / "r.mmrg"
    mm.save_region(path, region, index, payload)
    f = mm.RegionFile(path)
    try:
        assert f.region == region
        for i in (5, 0, 3):                                                 # random access
            cx, cz = region[0] + i % 3, region[1] + i // 3
            assert np.array_equal(f.read_chunk(cx, cz), chunks[i])
        with pytest.raises(mm.MmgenError):
            f.read_chunk(region[0] + 3, region[1])
    finally:
        f.close()
    with pytest.raises(mm.MmgenError):
        mm.RegionFile(tmp_path / "missing.mmrg")
    (tmp_path / "junk.mmrg").write_bytes(b"not a region file at all, definitely not")
    with pytest.raises(mm.MmgenError):
        mm.RegionFile(tmp_path / "junk.mmrg")


@pytest.mark.gpu
def test_encoded_delivery_equals_reference_golden(gen, mm, golden, tmp_path):
    g = golden["g"]
    world = gen.region_world(3, 3, 6, 6)
    try:
        buf = np.zeros(36 * 98304, np.uint8)
        index = np.zeros((36, 2), np.uint64)
        n = world.generate_to_host_encoded(buf.ctypes.data, buf.size, index)
        assert 36 * 512 < n < 36 * 98304 // 4                               # generated terrain compresses well
        assert int(index[:, 1].sum()) == n and (index[:, 0] % 16 == 0).all()
        for i in range(36):
            off, size = int(index[i, 0]), int(index[i, 1])
            enc = buf[off:off + size]
            assert np.array_equal(mm.decode_chunk(enc), g["blocks"][i]), i
            assert enc.tobytes() == np_encode(g["blocks"][i])                # the device encoder emits exactly the format
        # a buffer that is too small: error that names the size needed, nothing written beyond it
        world.reset()
        small = np.zeros(4096, np.uint8)
        with pytest.raises(mm.MmgenError, match=str(n)):
            world.generate_to_host_encoded(small.ctypes.data, small.size, index)
        # region file from the delivery, random access back
        path = tmp_path / "c2.mmrg"
        mm.save_region(path, (3, 3, 6, 6), index, buf[:n])
        f = mm.RegionFile(path)
        assert np.array_equal(f.read_chunk(8, 3), g["blocks"][5]) and np.array_equal(f.read_chunk(3, 8), g["blocks"][30])
        f.close()
    finally:
        world.close()
