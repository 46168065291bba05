#!/usr/bin/env python3
"""Developer tool (GPU box): BASELINE.json config 3 - a 64x64-chunk region streamed around a standing player at the
reference's load pattern (spiral order, per-stage queues, action-time budget), through mmgen_stream_*.
Prints one JSON line per cost profile: ticks, chunks filled, device ms, wall ms, chunks/s.
usage: tools/stream_bench.py [radius=51]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mmgen_loader  # noqa: E402

mm = mmgen_loader.load()
R = int(sys.argv[1]) if len(sys.argv) > 1 else 51
gen = mm.ChunkGen(0)
profiles = {
    "reference costs (500 per frame: <=166 heightfields / 100 layers / 62 caves / 62 fills / 1 zone per tick)": (500, 60 * 500),
    "8x frame budget": (4000, 60 * 4000),
    "unbounded frame budget (every queue drains each tick)": (1 << 24, 1 << 30),
}
for name, (cap, rate) in profiles.items():
    for rep in range(2):      # first pass allocates
        t = mm.Terrain(gen, -R - 1, -R - 1, 2 * R + 2, 2 * R + 2)
        t.set_radii(16, R)
        t.set_costs(mm.REFERENCE_COSTS, cap, rate)
        t0 = time.perf_counter()
        log = t.run_until_idle(1.0 / 32.0)
        wall = time.perf_counter() - t0
        filled = sum(s["filled"] for s in log)
        dev = sum(s["deviceMs"] for s in log)
        h = t.chunk_hash_sum()
        t.close()
    print(json.dumps({"profile": name, "gen_radius": R, "ticks": len(log), "chunks_filled": filled, "device_ms": round(dev, 2),
                      "wall_ms": round(1e3 * wall, 2), "chunks_per_s_wall": round(filled / wall, 1), "hash": "%016x" % h,
                      "max_batch": {k: max(s[k] for s in log) for k in ("heightfields", "layers", "caves", "filled", "zonesEroded")}}))
