"""Host model (no GPU) of the tile-level fixed-point shortcut of k_erode_sweep (mm_stage23.cuh).

The kernel sweeps a zone tile by tile (32 x 32 cells, ping-pong planes) and RETURNS AT ONCE from a tile whose 3 x 3 tile
neighbourhood the previous sweep left unchanged, leaving whatever the output planes already hold. The claim is that the planes
then hold exactly what a full Jacobi sweep would have written. This file restates the sweep (one cell rule, chunk.cu:542-590
in its race-free Jacobi reading) and the kernel's flag protocol in numpy and checks the claim bit for bit against sweeping
every tile every time, over several layers with the carried heights of the first sweep, on terrain with steep and flat parts.
The CUDA kernel itself is held to the oracle by tests/test_gpu_parity.py; this test holds the ARGUMENT."""
import numpy as np

T = 32                      # tile side
SQRT2 = np.float32(1.41421356237309504880168872420)
DIRS = [(0, 1), (1, 1), (1, 0), (1, -1), (0, -1), (-1, -1), (-1, 0), (-1, 1)]      # c_dirVecs2d: odd entries are diagonals


def sweep_cells(s_in, e_up, acc_in, rep, is_first):
    """One Jacobi sweep of the whole plane: returns (out_s, acc_out, changed mask). float32 throughout."""
    a = acc_in if is_first else np.zeros_like(acc_in)
    s = s_in + a
    e = e_up + a
    sp = np.pad(s, 1, mode="edge")      # clamp-to-edge halo, chunk.cu:545
    ep = np.pad(e, 1, mode="edge")
    n = s.shape[0]
    ns = s.copy()
    max_t = e - s
    rep_diag = np.float32(rep * SQRT2)
    for d, (dx, dz) in enumerate(DIRS):
        sj = sp[1 + dz:1 + dz + n, 1 + dx:1 + dx + n]
        ej = ep[1 + dz:1 + dz + n, 1 + dx:1 + dx + n]
        ns = np.maximum(ns, sj - (rep_diag if d & 1 else np.float32(rep)))
        max_t = np.maximum(max_t, ej - sj)
    ns = np.minimum(ns, e)
    active = max_t > 0
    changed = active & (ns != s)
    out_s = np.where(active, ns, s_in)
    acc_out = np.where(changed, (ns - s) + acc_in, acc_in)
    return out_s.astype(np.float32), acc_out.astype(np.float32), changed


def erode(layers, reps, skip):
    """Erodes `layers` (top layer last) the way erodeZonesDevice drives k_erode_sweep: layers from the top down, groups of 8
    sweeps until a group's last sweep changes nothing, ping-pong planes, the first sweep of a layer folds in the carried heights.
    skip=True applies the kernel's tile protocol: tile flags in three rotating rows, `force` for the first two sweeps of a layer,
    a skipped tile leaves both of its output planes untouched."""
    n = layers[0].shape[0]
    nt = n // T
    planes = [l.copy() for l in layers] + [np.full((n, n), np.float32(1e9))]      # plane above the top layer: never binding
    scratch = np.full((n, n), np.float32(np.nan))                                  # a skipped tile must never expose this
    acc = [np.zeros((n, n), np.float32), np.full((n, n), np.float32(np.nan))]
    acc_in = 0
    flags = np.zeros((3, nt, nt), bool)
    sweep_no = 0
    executed = 0
    for layer in range(len(layers) - 1, -1, -1):
        cur, other = planes[layer], scratch
        layer_sweeps = 0
        first = True
        converged = False
        while not converged:
            last_changed = False
            for _ in range(8):
                force = layer_sweeps < 2
                row_prev, row_cur, row_next = flags[(sweep_no + 2) % 3], flags[sweep_no % 3], flags[(sweep_no + 1) % 3]
                row_next[:] = False
                full_s, full_acc, changed = sweep_cells(cur, planes[layer + 1], acc[acc_in], reps[layer], first)
                out_s, out_acc = other, acc[1 - acc_in]
                any_changed = False
                for tz in range(nt):
                    for tx in range(nt):
                        if skip and not force and not row_prev[max(tz - 1, 0):tz + 2, max(tx - 1, 0):tx + 2].any():
                            continue      # the CTA returns at once
                        executed += 1
                        sl = (slice(tz * T, tz * T + T), slice(tx * T, tx * T + T))
                        out_s[sl] = full_s[sl]
                        out_acc[sl] = full_acc[sl]
                        if changed[sl].any():
                            row_cur[tz, tx] = True
                            any_changed = True
                cur, other = other, cur
                acc_in = 1 - acc_in
                first = False
                layer_sweeps += 1
                sweep_no += 1
                last_changed = any_changed
            converged = not last_changed
        # an even number of sweeps per group: the result is back in the layer's own plane
        assert cur is planes[layer]
        scratch = other
    return planes[:len(layers)], acc[acc_in], executed


def terrain(n, seed):
    rng = np.random.default_rng(seed)
    x, z = np.meshgrid(np.arange(n, dtype=np.float32), np.arange(n, dtype=np.float32))
    base = (40 + 25 * np.sin(x / 17) * np.cos(z / 23)).astype(np.float32)
    base[n // 3:n // 3 + 6, :] += np.float32(30)                       # a cliff: long relaxation, local activity
    base[:T, :T] = np.float32(50)                                      # a flat tile: quiet from the start
    l0 = base
    l1 = (l0 + rng.random((n, n), dtype=np.float32) * 6).astype(np.float32)
    l2 = (l1 + rng.random((n, n), dtype=np.float32) * 3 + np.float32(0.5)).astype(np.float32)
    return [l0, l1, l2]


def test_tile_skip_gives_the_planes_of_full_sweeps():
    reps = [np.float32(1.0), np.float32(0.839099586), np.float32(0.577350318)]
    for seed in (1, 2):
        layers = terrain(4 * T, seed)
        full_planes, full_acc, full_exec = erode(layers, reps, skip=False)
        skip_planes, skip_acc, skip_exec = erode(layers, reps, skip=True)
        for a, b in zip(full_planes, skip_planes):
            assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
        assert np.array_equal(full_acc.view(np.uint32), skip_acc.view(np.uint32))
        assert not np.isnan(skip_acc).any() and not any(np.isnan(p).any() for p in skip_planes)
        assert skip_exec < full_exec      # the shortcut did skip tiles
