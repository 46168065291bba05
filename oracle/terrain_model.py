"""TEST INFRASTRUCTURE - NOT PRODUCT CODE.

Pure-Python restatement of the reference's chunk scheduler for the generation path:
Terrain::generateSpiral / updateChunk / updateChunks / addZonesToTryErosionSet / updateZones / tick
(/root/reference/src/terrain/terrain.cpp:220-252, 300-417, 419-453, 525-566, 587-960) and the two
neighbourhood gathers (chunk.cu:53-147), with no data: only ChunkStates, queues and the action-time budget.
It is the checker for the streaming scheduler in mega-minecraft_b200/csrc/mm_stream.inl: per-tick batch
sizes and the final set of filled chunks must agree tick by tick (tests/test_stream.py).

One rule differs from the reference on purpose and is the same rule the product documents: a zone is
queued for erosion when its WHOLE 24x24-chunk gather window has layers (the reference only waits for the
neighbour zones that already exist, terrain.cpp:488-523), and chunks outside the session window never exist.

PINNED against the executable reference: oracle/_ref/libmmref_terrain.so is the UNMODIFIED terrain.cpp (g++ with one
force-included compatibility header, oracle/Makefile target `terrain`) driven headless by oracle/refterrain_driver.cpp;
tests/test_stream.py::test_model_equals_the_real_terrain_tick holds this model to it tick by tick (batch sizes of all nine
stages, fill order) for a standing and a walking player. The data is pinned separately: whatever order the scheduler fills
chunks in, their blocks equal the batch world's, which is held bit-exact to the reference's output.
"""
from collections import deque

(EMPTY, HAS_HEIGHTFIELD, NEEDS_LAYERS, HAS_LAYERS, NEEDS_EROSION, NEEDS_CAVES, NEEDS_FEATURE_PLACEMENTS,
 NEEDS_GATHER_FEATURE_PLACEMENTS, READY_TO_FILL, FILLED, NEEDS_VBOS, DRAWABLE) = range(12)

# terrain.cpp:69-80
MAX_ACTION_TIME_PER_FRAME = 500
TOTAL_ACTION_TIME_PER_SECOND = 60 * MAX_ACTION_TIME_PER_FRAME
COSTS = dict(heightfield=3, gather_heightfield=2, layers=5, erode=MAX_ACTION_TIME_PER_FRAME, caves=8, placements=3,
             gather_placements=5, fill=8, vbos=MAX_ACTION_TIME_PER_FRAME // 3)


def spiral(max_gen_radius):
    """terrain.cpp:220-252"""
    out = []
    x = z = 0
    d = m = 1
    while True:
        while 2 * x * d < m:
            out.append((x, z))
            x += d
        if m > max_gen_radius * 2:
            return out
        while 2 * z * d < m:
            out.append((x, z))
            z += d
        d = -d
        m += 1


class TerrainModel:
    def __init__(self, cx0, cz0, nx, nz, vbos_gen_radius=16, max_gen_radius=40, costs=None,
                 max_per_frame=MAX_ACTION_TIME_PER_FRAME, per_second=TOTAL_ACTION_TIME_PER_SECOND, zone_order=None):
        # zone_order: [(zone x, zone z) in zone units] - when several zones become ready in the same update, queue them in this
        # order. The reference iterates an unordered_set<Zone*> there (terrain.cpp:543-563), i.e. in heap-address order; replaying
        # the order a real run took makes the model comparable with that run tick by tick. Default: insertion order.
        self.zone_rank = {z: i for i, z in enumerate(zone_order)} if zone_order is not None else None
        self.cx0, self.cz0, self.nx, self.nz = cx0, cz0, nx, nz
        self.vbos_r, self.max_r = vbos_gen_radius, max_gen_radius
        self.cost = dict(COSTS if costs is None else costs)
        self.max_per_frame, self.per_second = max_per_frame, per_second
        self.spiral = spiral(max_gen_radius)
        self.state = {}            # (cx, cz) -> state, only chunks that exist
        self.ready = {}
        self.cur = (0, 0)
        self.last = (0, 0)
        self.needs_update = True
        self.left = 0
        self.q = {k: deque() for k in ("heightfield", "gather_heightfield", "layers", "caves", "placements", "gather_placements", "fill", "vbos")}
        self.zones_to_try = []     # insertion order
        self.zones_queued = set()
        self.zones_to_erode = deque()
        self.filled_order = []

    # -- helpers
    def in_window(self, c):
        return self.cx0 <= c[0] < self.cx0 + self.nx and self.cz0 <= c[1] < self.cz0 + self.nz

    def set_state(self, c, s):
        self.state[c] = s
        self.ready[c] = True

    def set_player_chunk(self, cx, cz):
        self.cur = (cx, cz)

    def update_chunk(self, dx, dz):
        c = (self.cur[0] + dx, self.cur[1] + dz)
        if not self.in_window(c):
            return
        if c not in self.state:
            self.state[c] = EMPTY
            self.ready[c] = True
        if not self.ready[c]:
            return
        s = self.state[c]
        route = {EMPTY: "heightfield", HAS_HEIGHTFIELD: "gather_heightfield", NEEDS_LAYERS: "layers", NEEDS_CAVES: "caves",
                 NEEDS_FEATURE_PLACEMENTS: "placements", NEEDS_GATHER_FEATURE_PLACEMENTS: "gather_placements", READY_TO_FILL: "fill"}
        if s in route:
            self.ready[c] = False
            self.q[route[s]].append(c)
            return
        if max(abs(dx), abs(dz)) > self.vbos_r:
            return
        if s == NEEDS_VBOS:
            self.ready[c] = False
            self.q["vbos"].append(c)

    def gather(self, c, R, cur, nxt):
        """floodFillAndIterateNeighbors<4R+1> (chunk.cu:53-147). floodFill: breadth-first from c over the four edge
        neighbours, inside the (4R+1)^2 window around c, through chunks at state `cur` or beyond only - a chunk behind a
        less advanced one is not found. iterateNeighborChunks: centres within R of c that are at `cur` and whose whole
        (2R+1)^2 neighbourhood was found move to `nxt`."""
        radius = 2 * R
        found, visited, queue = set(), set(), deque([c])
        while queue:
            p = queue.popleft()
            visited.add(p)
            if self.state.get(p, -1) < cur:
                continue
            found.add(p)
            for n in ((p[0], p[1] + 1), (p[0] + 1, p[1]), (p[0], p[1] - 1), (p[0] - 1, p[1])):
                if n not in self.state or n in visited or max(abs(n[0] - c[0]), abs(n[1] - c[1])) > radius:
                    continue
                queue.append(n)
        for cz in range(c[1] - R, c[1] + R + 1):
            for cx in range(c[0] - R, c[0] + R + 1):
                if (cx, cz) not in found or self.state[(cx, cz)] != cur:
                    continue
                if all((cx + ox, cz + oz) in found for oz in range(-R, R + 1) for ox in range(-R, R + 1)):
                    self.set_state((cx, cz), nxt)

    # Terrain::addZonesToTryErosionSet (terrain.cpp:431-453): the chunk's own zone plus three neighbour zones chosen by the
    # quadrant of the zone the chunk lies in - start direction 4 / 6 for the west half (south / north), 0 / 2 for the east half
    # (south / north), then three consecutive directions clockwise from north. (For the east half that is NOT the three zones
    # whose gather windows contain the chunk - the table is what the reference does, and the scheduler is pinned to it.)
    QUADRANT_ZONES = {(False, False): ((0, -1), (-1, -1), (-1, 0)), (False, True): ((-1, 0), (-1, 1), (0, 1)),
                      (True, False): ((0, 1), (1, 1), (1, 0)), (True, True): ((1, 0), (1, -1), (0, -1))}

    def add_zones_to_try(self, c):
        zx, zz = c[0] // 12, c[1] // 12
        east, north = c[0] - 12 * zx >= 6, c[1] - 12 * zz >= 6
        for z in ((zx, zz),) + tuple((zx + dx, zz + dz) for dx, dz in self.QUADRANT_ZONES[(east, north)]):
            if z in self.zones_queued or z in self.zones_to_try:
                continue
            self.zones_to_try.append(z)

    def update_zones(self):
        if self.zone_rank is not None:
            self.zones_to_try.sort(key=lambda z: self.zone_rank.get(z, 1 << 30))
        for z in self.zones_to_try:
            x0, z0 = 12 * z[0] - 6, 12 * z[1] - 6
            if all(self.state.get((x0 + dx, z0 + dz), -1) >= HAS_LAYERS for dz in range(24) for dx in range(24)):
                self.zones_to_erode.append(z)
                self.zones_queued.add(z)
        self.zones_to_try = []

    def check_needs_vbos(self, c):
        if self.state.get(c, -1) < FILLED:
            return
        for n in ((c[0] + 1, c[1]), (c[0] - 1, c[1]), (c[0], c[1] + 1), (c[0], c[1] - 1)):
            if self.state.get(n, -1) < FILLED:
                return
        if self.state[c] == FILLED:
            self.set_state(c, NEEDS_VBOS)

    # -- terrain.cpp:587-960
    def tick(self, dt):
        st = dict(heightfields=0, gatherHeightfields=0, layers=0, zonesEroded=0, caves=0, placements=0, gatherPlacements=0, filled=0, vbos=0)
        if self.cur != self.last:
            self.last = self.cur
            self.needs_update = True
        if self.needs_update:
            self.update_zones()
            for dx, dz in self.spiral:
                self.update_chunk(dx, dz)
            self.needs_update = False
        self.left = min(self.left + int(self.per_second * dt), self.max_per_frame)

        def drain(name):
            out = []
            while self.q[name] and self.left >= self.cost[name]:
                self.needs_update = True
                out.append(self.q[name].popleft())
                self.left -= self.cost[name]
            return out

        for c in drain("vbos"):
            self.state[c] = DRAWABLE
            self.ready[c] = False
            st["vbos"] += 1
        filled = drain("fill")
        for c in filled:
            self.state[c] = FILLED
            self.ready[c] = False
        for c in filled:
            self.filled_order.append(c)
            self.check_needs_vbos(c)
            for n in ((c[0] + 1, c[1]), (c[0] - 1, c[1]), (c[0], c[1] + 1), (c[0], c[1] - 1)):
                self.check_needs_vbos(n)
        st["filled"] = len(filled)
        for c in drain("gather_placements"):
            self.gather(c, 3, NEEDS_GATHER_FEATURE_PLACEMENTS, READY_TO_FILL)
            st["gatherPlacements"] += 1
        for c in drain("placements"):
            self.set_state(c, NEEDS_GATHER_FEATURE_PLACEMENTS)
            st["placements"] += 1
        for c in drain("caves"):
            self.set_state(c, NEEDS_FEATURE_PLACEMENTS)
            st["caves"] += 1
        while self.zones_to_erode and self.left >= self.cost["erode"]:
            self.needs_update = True
            z = self.zones_to_erode.popleft()
            for dz in range(12):
                for dx in range(12):
                    self.set_state((12 * z[0] + dx, 12 * z[1] + dz), NEEDS_CAVES)
            self.left -= self.cost["erode"]
            st["zonesEroded"] += 1
        for c in drain("layers"):
            self.set_state(c, HAS_LAYERS)
            self.add_zones_to_try(c)
            st["layers"] += 1
        for c in drain("gather_heightfield"):
            self.gather(c, 1, HAS_HEIGHTFIELD, NEEDS_LAYERS)
            st["gatherHeightfields"] += 1
        for c in drain("heightfield"):
            self.set_state(c, HAS_HEIGHTFIELD)
            st["heightfields"] += 1
        st["actionTimeLeft"] = self.left
        st["idle"] = int(not self.needs_update and not self.zones_to_try and not self.zones_to_erode and not any(self.q.values()))
        if st["idle"]:
            # the product's safety net (mm_stream.inl): zones whose window is complete but which the reference's quadrant table
            # never put up for their test are tried when the stream would otherwise fall idle (never happens in the runs that
            # are compared with the real terrain.cpp: there the reference itself finishes every zone)
            zx0, zz0 = self.cx0 // 12, self.cz0 // 12
            for zz in range(zz0, (self.cz0 + self.nz - 1) // 12 + 1):
                for zx in range(zx0, (self.cx0 + self.nx - 1) // 12 + 1):
                    z = (zx, zz)
                    if z in self.zones_queued:
                        continue
                    x0, z0 = 12 * zx - 6, 12 * zz - 6
                    if all(self.state.get((x0 + dx, z0 + dz), -1) >= HAS_LAYERS for dz in range(24) for dx in range(24)):
                        self.zones_to_try.append(z)
                        self.needs_update = True
                        st["idle"] = 0
        return st

    def run_until_idle(self, dt=1.0 / 60.0, max_ticks=100000):
        log = []
        for _ in range(max_ticks):
            st = self.tick(dt)
            log.append(st)
            if st["idle"]:
                break
        return log

    def filled(self):
        return sorted(c for c, s in self.state.items() if s >= FILLED)
