"""One process per GPU: which tile a rank generates and the few cross-rank reductions the path needs.

The generation path has no data-path exchange in the recompute variant (option A): a
rank builds the apron of its own tile. What crosses ranks is bookkeeping only - the device time of a
step (MAX), work counters (SUM) and the per-tile block checksums (gather) - so the same code runs over
NCCL on the GPUs and over gloo in the CPU tests.
"""
import numpy as np
import torch
import torch.distributed as dist

from . import tiling


def rank_tile(region, rank, world_size, align=1):
    """(x0, z0, nx, nz) of the tile `rank` fills when `region` = (rx0, rz0, rnx, rnz) is split over world_size ranks."""
    return tiling.tiles(region[0], region[1], region[2], region[3], world_size, align)[rank]


def _dev():
    return torch.device("cuda", torch.cuda.current_device()) if dist.is_initialized() and dist.get_backend() == "nccl" else torch.device("cpu")


def reduce_scalar(x, op="max"):
    """MAX / SUM of a python float over all ranks (identity when torch.distributed is not initialised)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(x)
    t = torch.tensor([float(x)], dtype=torch.float64, device=_dev())
    dist.all_reduce(t, op=dist.ReduceOp.MAX if op == "max" else dist.ReduceOp.SUM)
    return float(t.item())


def gather_u64(value):
    """All ranks' 64-bit values (e.g. tile block checksums) in rank order, on every rank."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return [int(value)]
    # int64 carries the bit pattern; split in two 32-bit halves to stay clear of signed overflow
    t = torch.tensor([int(value) >> 32, int(value) & 0xFFFFFFFF], dtype=torch.int64, device=_dev())
    out = [torch.zeros_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return [(int(o[0].item()) << 32) | int(o[1].item()) for o in out]


def combine_checksums(values):
    """Order-sensitive 64-bit FNV-1a over the per-tile checksums (rank order = tile order)."""
    h = 14695981039346656037
    for v in values:
        for b in range(8):
            h ^= (int(v) >> (8 * b)) & 0xFF
            h = (h * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return h


def gather_floats(value):
    """All ranks' float values in rank order, on every rank."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return [float(value)]
    t = torch.tensor([float(value)], dtype=torch.float64, device=_dev())
    out = [torch.zeros_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return [float(o.item()) for o in out]


class Balancer:
    """Feedback load balancing of the tiling.

    The cost of generating a chunk varies over the world (terrain height, cave and feature density; measured
    +-15 % between the 8 equal tiles of the 256x256 bench world), and it cannot be predicted from the cheap
    stages. What can be done is what a streaming generator does anyway: measure. Every rank reports the device
    time of its tile; the cost density (ms per chunk) of each tile is spread over its chunks and the cuts of
    the gx x gz grid are moved so that every tile carries the same predicted cost - first the cuts between
    the rows (z), then the cuts inside every row (x). Tiles stay rectangles that partition the region
    exactly, rank order stays row-major, and every rank computes the same cuts from the same gathered times.
    """

    def __init__(self, region, world_size):
        self.region = tuple(region)
        self.gx, self.gz = tiling.grid_for(world_size)
        rx0, rz0, rnx, rnz = self.region
        self.zs = tiling.split_points(rz0, rnz, self.gz)
        self.xs = [tiling.split_points(rx0, rnx, self.gx) for _ in range(self.gz)]

    def tiles(self):
        return [(self.xs[j][i], self.zs[j], self.xs[j][i + 1] - self.xs[j][i], self.zs[j + 1] - self.zs[j])
                for j in range(self.gz) for i in range(self.gx)]

    @staticmethod
    def _cuts(cost, parts, start):
        """Cut positions over a 1-D cost profile so that each part carries the same cumulative cost (>= 1 cell per part)."""
        cum = np.concatenate([[0.0], np.cumsum(cost)])
        cuts = [0]
        for k in range(1, parts):
            target = cum[-1] * k / parts
            c = int(np.searchsorted(cum, target))
            # nearest of the two neighbouring positions
            if c > 0 and abs(cum[c - 1] - target) <= abs(cum[min(c, len(cum) - 1)] - target):
                c -= 1
            c = max(c, cuts[-1] + 1)
            c = min(c, len(cost) - (parts - k))
            cuts.append(c)
        cuts.append(len(cost))
        return [start + c for c in cuts]

    def update(self, times):
        """times: per-rank cost of the current tiles (rank order). Moves the cuts; returns the predicted imbalance before the move."""
        rx0, rz0, rnx, rnz = self.region
        dens = np.zeros((rnz, rnx))
        for (x0, z0, nx, nz), t in zip(self.tiles(), times):
            dens[z0 - rz0:z0 - rz0 + nz, x0 - rx0:x0 - rx0 + nx] = float(t) / (nx * nz)
        self.zs = self._cuts(dens.sum(axis=1), self.gz, rz0)
        self.xs = [self._cuts(dens[self.zs[j] - rz0:self.zs[j + 1] - rz0].sum(axis=0), self.gx, rx0) for j in range(self.gz)]
        return max(times) / (sum(times) / len(times))
