"""Chunk-coordinate tiling of a world over the GPUs of one box.

Every stage of the generation path is a pure function of world coordinates plus neighbour data within
a bounded radius, so a world region splits into rectangular tiles that are generated independently:
each GPU builds the apron its tile needs (`mmgen_world_create_for_region`) and keeps its block
volumes. No data-path collective is needed for the recompute variant; the only cross-rank traffic
is the equality checksum and the timing reduction.
"""


def grid_for(n_ranks):
    """(columns, rows) = (cuts along x, cuts along z) of the tile grid: as square as possible, columns >= rows
    (1, 2, 4, 8 ranks -> 1x1, 2x1, 2x2, 4x2). Rank order is row-major (z-major): rank = row * columns + column."""
    if n_ranks < 1:
        raise ValueError("n_ranks must be >= 1")
    gx = 1
    while gx * gx * 2 <= n_ranks and n_ranks % (gx * 2) == 0:
        gx *= 2
    while n_ranks % gx:
        gx -= 1
    return gx, n_ranks // gx


def split_points(start, length, parts, align=1):
    """parts+1 cut positions over [start, start+length), interior cuts rounded to a multiple of `align`
    (in world chunk coordinates) when that leaves every part non-empty."""
    cuts = [start]
    for i in range(1, parts):
        c = start + (length * i) // parts
        if align > 1:
            a = int(round(c / align)) * align
            if cuts[-1] < a < start + length:
                c = a
        cuts.append(c)
    cuts.append(start + length)
    if any(b <= a for a, b in zip(cuts, cuts[1:])):
        raise ValueError("region too small for %d parts" % parts)
    return cuts


def tiles(rx0, rz0, rnx, rnz, n_ranks, align=1):
    """List of n_ranks tiles (x0, z0, nx, nz) in rank order (z-major) that partition the region exactly."""
    gx, gz = grid_for(n_ranks)
    xs, zs = split_points(rx0, rnx, gx, align), split_points(rz0, rnz, gz, align)
    return [(xs[i], zs[j], xs[i + 1] - xs[i], zs[j + 1] - zs[j]) for j in range(gz) for i in range(gx)]


def apron_window(x0, z0, nx, nz):
    """Window (cx0, cz0, wnx, wnz) that mmgen_world_create_for_region allocates for a tile: zones
    (12-chunk aligned) meeting tile (+) 3 chunks, plus 6 chunks of erosion pad, plus the ring of chunks
    whose heightfield borders the outermost layers."""
    zx0 = ((x0 - 3) // 12) * 12
    zx1 = ((x0 + nx + 2) // 12) * 12 + 12
    zz0 = ((z0 - 3) // 12) * 12
    zz1 = ((z0 + nz + 2) // 12) * 12 + 12
    return zx0 - 7, zz0 - 7, zx1 - zx0 + 14, zz1 - zz0 + 14


def stage_chunk_counts(x0, z0, nx, nz):
    """Chunks each stage touches to fill a tile: dict S1..S6 (S3 in zones)."""
    cx0, cz0, wnx, wnz = apron_window(x0, z0, nx, nz)
    zones = ((wnx - 14) // 12) * ((wnz - 14) // 12)
    return {"S1": wnx * wnz, "S2": (wnx - 2) * (wnz - 2), "S3_zones": zones, "S4": (nx + 6) * (nz + 6), "S5": (nx + 6) * (nz + 6),
            "S6": nx * nz}
