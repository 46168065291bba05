"""Meshing (Chunk::createVBOs on the device, mm_mesh.cuh) against the reference's own createVBOs:
 - CPU: the committed fixture tests/golden/c2_mesh.npz is what oracle/_ref produces today (where it is built);
 - GPU: the kernels reproduce the reference's vertex and index arrays byte for byte, null neighbours included."""
import hashlib

import numpy as np
import pytest

from oracle import refcuda


@pytest.fixture(scope="module")
def mesh_golden():
    import os
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "c2_mesh.npz"))


def neighbours_in_region(blocks_by_coord, cx, cz):
    return [blocks_by_coord.get((cx + dx, cz + dz)) for dx, dz in ((0, 1), (1, 0), (0, -1), (-1, 0))]


def test_fixture_is_the_reference_output(golden, mesh_golden):
    if not refcuda.available():
        pytest.skip("oracle/_ref/libmmref_cuda.so not built (needs /root/reference at build time)")
    g = golden["g"]
    by_coord = {(golden["x0"] + int(i) % golden["nx"], golden["z0"] + int(i) // golden["nx"]): g["blocks"][k] for k, i in enumerate(g["block_idx"])}
    assert refcuda.VERTEX.itemsize == 40
    for k in (0, 7, 14, 35):
        cx, cz = (int(v) for v in mesh_golden["coords"][k])
        v, ix = refcuda.mesh_chunk(cx, cz, by_coord[(cx, cz)], neighbours_in_region(by_coord, cx, cz))
        assert (len(v), len(ix)) == tuple(mesh_golden["counts"][k])
        assert hashlib.sha1(v.tobytes()).hexdigest() == str(mesh_golden["verts_sha1"][k])
        assert hashlib.sha1(ix.tobytes()).hexdigest() == str(mesh_golden["idx_sha1"][k])
    # structure of the stored interior chunk: quads, 6 indices per 4 vertices, indices inside the chunk
    v, ix = mesh_golden["full_verts"], mesh_golden["full_idx"]
    assert len(v) % 4 == 0 and len(ix) == len(v) // 4 * 6 and ix.max() == len(v) - 1
    assert np.array_equal(ix.reshape(-1, 6) - ix.reshape(-1, 6)[:, :1], np.tile([0, 1, 2, 0, 2, 3], (len(ix) // 6, 1)))


@pytest.mark.gpu
def test_mesh_matches_reference_createvbos(gen, mm, golden, mesh_golden):
    world = gen.region_world(3, 3, 6, 6)
    world.generate(mm.STAGE_ALL)
    assert np.array_equal(world.download_region_blocks(), golden["g"]["blocks"])
    coords = mesh_golden["coords"]
    meshes = world.mesh(coords)
    assert world.mesh_ms() > 0
    for k, (v, ix) in enumerate(meshes):
        assert (len(v), len(ix)) == tuple(mesh_golden["counts"][k]), (k, coords[k])
        if tuple(coords[k]) == tuple(mesh_golden["full_coord"]):       # readable diff first
            ref = mesh_golden["full_verts"]
            for f in ("pos", "nor", "uv", "m"):
                assert np.array_equal(v[f], ref[f]), f
            assert np.array_equal(ix, mesh_golden["full_idx"])
        assert hashlib.sha1(v.tobytes()).hexdigest() == str(mesh_golden["verts_sha1"][k]), (k, coords[k])
        assert hashlib.sha1(ix.tobytes()).hexdigest() == str(mesh_golden["idx_sha1"][k]), (k, coords[k])
    # any subset / order of chunks gives the same per-chunk arrays (arena offsets do not leak into the indices)
    sub = world.mesh(coords[[20, 3]])
    assert sub[0][0].tobytes() == meshes[20][0].tobytes() and sub[1][1].tobytes() == meshes[3][1].tobytes()
    # hand-off to the path tracer (OptixRenderer::buildChunkAccel, optixRenderer.cpp:223-368): the triangle-array description
    # of every meshed chunk points into the device arena - no host vectors, no upload
    meshes = world.mesh(coords)
    gas = world.mesh_gas_inputs()
    assert gas.dtype.itemsize == 40 and len(gas) == len(coords)
    assert np.array_equal(np.stack([gas["cx"], gas["cz"]], axis=1), coords)
    assert (gas["vertexStrideInBytes"] == 40).all() and (gas["indexStrideInBytes"] == 12).all()
    assert np.array_equal(gas["numVertices"], mesh_golden["counts"][:, 0]) and np.array_equal(gas["numIndexTriplets"] * 3, mesh_golden["counts"][:, 1])
    # chunks are packed back to back in the arena, vertices 40 bytes apart, 6 indices per 4 vertices
    assert np.array_equal(np.diff(gas["vertexBuffer"].astype(np.int64)), gas["numVertices"][:-1].astype(np.int64) * 40)
    assert np.array_equal(np.diff(gas["indexBuffer"].astype(np.int64)), gas["numIndexTriplets"][:-1].astype(np.int64) * 12)
    assert (gas["vertexBuffer"] % 8 == 0).all() and (gas["indexBuffer"] % 4 == 0).all()
    import ctypes
    for k in (0, 7, 35):                                    # the same addresses mmgen_world_mesh_device_ptrs / _download use
        pv, pi, nv, ni = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_int(0), ctypes.c_int(0)
        gen._check(gen.L.mmgen_world_mesh_device_ptrs(world.h, k, ctypes.byref(pv), ctypes.byref(pi), ctypes.byref(nv), ctypes.byref(ni)))
        assert (pv.value, pi.value, nv.value, ni.value) == (int(gas["vertexBuffer"][k]), int(gas["indexBuffer"][k]), int(gas["numVertices"][k]),
                                                            3 * int(gas["numIndexTriplets"][k]))
    with pytest.raises(mm.MmgenError):
        world.mesh(np.array([[100, 100]], np.int32))
    world.close()


@pytest.mark.gpu
def test_mesh_other_biomes_vs_reference_mesher(gen, mm):
    """Block volumes from other parts of the world (plants, water, crystals, leaves) meshed by both sides."""
    if not refcuda.available():
        pytest.skip("oracle/_ref/libmmref_cuda.so not built")
    for rx, rz in ((-150, -54), (-150, 96)):
        world = gen.region_world(rx, rz, 3, 3)
        world.generate(mm.STAGE_ALL)
        blocks = world.download_region_blocks().reshape(3, 3, 16, 16, 384)
        by_coord = {(rx + x, rz + z): blocks[z, x] for z in range(3) for x in range(3)}
        coords = np.array([[rx + 1, rz + 1], [rx, rz + 2]], np.int32)
        for (cx, cz), (v, ix) in zip(coords, world.mesh(coords)):
            rv, rix = refcuda.mesh_chunk(int(cx), int(cz), by_coord[(int(cx), int(cz))], neighbours_in_region(by_coord, int(cx), int(cz)))
            assert len(v) == len(rv)
            for f in ("pos", "nor", "uv", "m"):
                assert np.array_equal(v[f], rv[f]), (cx, cz, f)
            assert np.array_equal(ix, rix)
        world.close()
