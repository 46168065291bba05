#!/usr/bin/env python3
"""Developer tool (GPU box): per-kernel device times of the streaming scheduler (8x and unbounded frame budget), to see which
kernel a change of the streaming numbers comes from. usage: python tools/stream_diag.py"""
import os, sys, time
sys.path.insert(0, os.getcwd())
import mmgen_loader
mm = mmgen_loader.load()
R = 51
gen = mm.ChunkGen(0)
for name, (cap, rate) in {"8x": (4000, 60 * 4000), "unbounded": (1 << 24, 1 << 30)}.items():
    for rep in range(2):
        t = mm.Terrain(gen, -R - 1, -R - 1, 2 * R + 2, 2 * R + 2)
        t.set_radii(16, R)
        t.set_costs(mm.REFERENCE_COSTS, cap, rate)
        if rep == 1: gen.kernel_timing(True)
        t0 = time.perf_counter()
        log = t.run_until_idle(1.0 / 32.0)
        wall = time.perf_counter() - t0
        if rep == 1:
            kt = gen.kernel_times(); gen.kernel_timing(False)
            print(name, "wall %.1f ms" % (wall * 1e3), {k: (round(v[0], 2), v[1]) for k, v in kt.items() if v[1]})
        t.close()
