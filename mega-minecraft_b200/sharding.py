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


def _grow(t, g):
    return (t[0] - g, t[1] - g, t[2] + 2 * g, t[3] + 2 * g)


def _intersect(a, b):
    x0, z0 = max(a[0], b[0]), max(a[1], b[1])
    x1, z1 = min(a[0] + a[2], b[0] + b[2]), min(a[1] + a[3], b[1] + b[3])
    return (x0, z0, x1 - x0, z1 - z0) if x1 > x0 and z1 > z0 else None


def exchange_plan(tiles, rank, halo=3):
    """Halo exchange of placement lists (SURVEY.md 8(e) option B): [(peer, send_rect, recv_rect)] for `rank`. A rank fills
    its tile from the placements of tile (+) 3 chunks (gather order chunk.cu:1158-1187); the part of that ring that lies in
    a peer's tile is computed there and sent: send_rect = my tile n (peer tile (+) halo), recv_rect = peer tile n (my tile
    (+) halo). Both are non-empty for the same peers, so every pair exchanges exactly one message each way."""
    me = tiles[rank]
    plan = []
    for peer, t in enumerate(tiles):
        if peer == rank:
            continue
        send, recv = _intersect(me, _grow(t, halo)), _intersect(t, _grow(me, halo))
        assert (send is None) == (recv is None)
        if send is not None:
            plan.append((peer, send, recv))
    return plan


class HaloExchange:
    """Moves the packed placement lists between the ranks' device-resident worlds with NCCL send / recv over NVLink (two
    rounds: message lengths, then payloads), one device buffer pair per peer."""

    def __init__(self, tiles, rank, bytes_per_chunk=48 * 1024):
        self.plan = exchange_plan(tiles, rank)
        dev = _dev()
        self.send = {p: torch.empty(s[2] * s[3] * bytes_per_chunk, dtype=torch.uint8, device=dev) for p, s, _ in self.plan}
        self.recv = {p: torch.empty(r[2] * r[3] * bytes_per_chunk, dtype=torch.uint8, device=dev) for p, _, r in self.plan}
        self.slen = {p: torch.zeros(1, dtype=torch.int64, device=dev) for p, _, _ in self.plan}
        self.rlen = {p: torch.zeros(1, dtype=torch.int64, device=dev) for p, _, _ in self.plan}
        self.bytes_sent = 0

    def run(self, world):
        """world: the rank's World after stages 1-5 on its own tile. Returns the bytes this rank sent."""
        if not self.plan:
            return 0
        sizes = {}
        for peer, srect, _ in self.plan:
            n, ok = world.pack_placements(srect, self.send[peer].data_ptr(), self.send[peer].numel())
            if not ok:                                   # rare: a denser strip than the buffer was sized for
                self.send[peer] = torch.empty(n + n // 4, dtype=torch.uint8, device=self.send[peer].device)
                n, ok = world.pack_placements(srect, self.send[peer].data_ptr(), self.send[peer].numel())
                assert ok
            sizes[peer] = n
            self.slen[peer].fill_(n)
        ops = []
        for peer, _, _ in self.plan:
            ops.append(dist.P2POp(dist.isend, self.slen[peer], peer))
            ops.append(dist.P2POp(dist.irecv, self.rlen[peer], peer))
        for r in dist.batch_isend_irecv(ops):
            r.wait()
        rsizes = {peer: int(self.rlen[peer].item()) for peer, _, _ in self.plan}
        for peer, n in rsizes.items():
            if n > self.recv[peer].numel():
                self.recv[peer] = torch.empty(n + n // 4, dtype=torch.uint8, device=self.recv[peer].device)
        ops = []
        for peer, _, _ in self.plan:
            ops.append(dist.P2POp(dist.isend, self.send[peer][:sizes[peer]], peer))
            ops.append(dist.P2POp(dist.irecv, self.recv[peer][:rsizes[peer]], peer))
        for r in dist.batch_isend_irecv(ops):
            r.wait()
        torch.cuda.synchronize()
        for peer, _, rrect in self.plan:
            world.unpack_placements(rrect, self.recv[peer].data_ptr(), rsizes[peer])
        self.bytes_sent = sum(sizes.values())
        return self.bytes_sent


# Relative cost of a chunk from its stage-1 features (ChunkGen.chunk_costs: cave voxels, fill voxels, land columns), fitted to
# measured per-tile device times of the 256x256 bench world on a B200 (tools/fit_cost_model.py; only ratios matter)
# (25 tiles of three tilings, 49-85 ms each: 1.4 % rms, 4.6 % worst-case error; by area alone 12.6 % / 22.4 %; profiles/r02_cost_model_v2.txt.
# Refitted after the kernels of DESIGN.md 5 items 21-25: the cave stage got cheaper relative to the fill, its voxel count is almost
# collinear with the fill's and the non-negative fit gives it no weight of its own.)
COST_WEIGHTS = (0.0, 8.842e-4, 2.056e-3, 4.005e-3)      # ms: x cave voxels / 1e4, x fill voxels / 1e4, x land columns / 256, per chunk


_cost_buffers = {}      # (shape, device) -> tensors reused by chunk_cost_map: it runs inside every timed step


def chunk_cost_map(gen, region, rank=0, world_size=1, weights=COST_WEIGHTS, stride=2):
    """(rnz, rnx) predicted cost per chunk of `region`, from stage 1 alone: every rank evaluates the features of a strip of
    rows (mmgen_chunk_costs), the strips are all-gathered, every rank ends up with the same map. The features are taken at every
    `stride`-th chunk in x and z and repeated over the chunks in between: terrain height and land fraction vary over hundreds of
    blocks, a quarter of the stage-1 evaluations place the cuts as well as all of them."""
    rx0, rz0, rnx, rnz = region
    snx, snz = -(-rnx // stride), -(-rnz // stride)      # sampled grid
    rows = tiling.split_points(0, snz, world_size) if snz >= world_size else [0] + [snz] * world_size
    z0, z1 = rows[rank], rows[rank + 1]
    zz, xx = np.meshgrid(rz0 + stride * np.arange(z0, z1, dtype=np.int32), rx0 + stride * np.arange(snx, dtype=np.int32), indexing="ij")
    origins = np.ascontiguousarray(np.stack([xx.ravel() * 16, zz.ravel() * 16], axis=1), np.int32)
    f = gen.chunk_costs(origins).reshape(z1 - z0, snx, 3) if len(origins) else np.zeros((0, snx, 3), np.float32)
    if dist.is_initialized() and dist.get_world_size() > 1:
        most = max(b - a for a, b in zip(rows, rows[1:]))
        dev = _dev()
        key = (most, snx, world_size, str(dev))
        if key not in _cost_buffers:
            pin = dev.type == "cuda"
            _cost_buffers[key] = (torch.zeros((most, snx, 3), dtype=torch.float32, pin_memory=pin),
                                  torch.empty((most, snx, 3), dtype=torch.float32, device=dev),
                                  torch.empty((world_size * most, snx, 3), dtype=torch.float32, device=dev),
                                  torch.empty((world_size * most, snx, 3), dtype=torch.float32, pin_memory=pin))
        h_in, d_in, d_out, h_out = _cost_buffers[key]
        h_in[:z1 - z0] = torch.from_numpy(f)
        d_in.copy_(h_in, non_blocking=True)
        dist.all_gather_into_tensor(d_out, d_in)          # one collective, one read-back
        h_out.copy_(d_out)
        out = h_out.numpy().reshape(world_size, most, snx, 3)
        f = np.concatenate([out[r, :b - a] for r, (a, b) in enumerate(zip(rows, rows[1:]))], axis=0)
    c = weights[0] * f[:, :, 0] / 1e4 + weights[1] * f[:, :, 1] / 1e4 + weights[2] * f[:, :, 2] / 256.0 + weights[3]
    if stride > 1:
        c = np.repeat(np.repeat(c, stride, axis=0), stride, axis=1)[:rnz, :rnx]
    return c


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

    def cut_by_cost(self, cost):
        """cost: (rnz, rnx) predicted cost per chunk (chunk_cost_map). Moves the cuts so that every tile carries the same
        predicted cost - rows first, then the cuts inside every row - without any rehearsal pass."""
        rx0, rz0, rnx, rnz = self.region
        cost = np.asarray(cost, np.float64)
        assert cost.shape == (rnz, rnx)
        self.zs = self._cuts(cost.sum(axis=1), self.gz, rz0)
        self.xs = [self._cuts(cost[self.zs[j] - rz0:self.zs[j + 1] - rz0].sum(axis=0), self.gx, rx0) for j in range(self.gz)]
        return self.tiles()

    def update(self, times):
        """times: per-rank cost of the current tiles (rank order). Moves the cuts; returns the predicted imbalance before the move."""
        rx0, rz0, rnx, rnz = self.region
        dens = np.zeros((rnz, rnx))
        for (x0, z0, nx, nz), t in zip(self.tiles(), times):
            dens[z0 - rz0:z0 - rz0 + nz, x0 - rx0:x0 - rx0 + nx] = float(t) / (nx * nz)
        self.zs = self._cuts(dens.sum(axis=1), self.gz, rz0)
        self.xs = [self._cuts(dens[self.zs[j] - rz0:self.zs[j + 1] - rz0].sum(axis=0), self.gx, rx0) for j in range(self.gz)]
        return max(times) / (sum(times) / len(times))
