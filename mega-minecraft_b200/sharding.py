"""One process per GPU: which tile a rank generates and the few cross-rank reductions the path needs.

The generation path has no data-path exchange in the recompute variant (option A): a
rank builds the apron of its own tile. What crosses ranks is bookkeeping only - the device time of a
step (MAX), work counters (SUM) and the per-tile block checksums (gather) - so the same code runs over
NCCL on the GPUs and over gloo in the CPU tests.
"""
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
