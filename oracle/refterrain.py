"""TEST INFRASTRUCTURE. Runs the reference's own chunk scheduler - the UNMODIFIED terrain.cpp in oracle/_ref/libmmref_terrain.so
(oracle/Makefile target `terrain`, oracle/refterrain_driver.cpp) - headless and returns what it did tick by tick.

    python -m oracle.refterrain '{"moves": [[0, 0], [24, 0]], "dt": 0.03125, "skip_kernels": true}'

prints one JSON object: ticks = per tick the nine MmgenTickStats counts (heightfields, gatherHeightfields, layers, zonesEroded,
caves, placements, gatherPlacements, filled, vbos), filled = chunk coordinates in fill order, eroded = zone coordinates (zone
units) in erosion order, built = chunk coordinates in the order they reached the VBO stage, segments = tick index at which each player move was made. It runs in its own process: the reference
keeps its state in file-static buffers and pointer-keyed hash sets. With skip_kernels the generation kernels are not launched
(no GPU needed): batch sizes, orders and budgets do not depend on the data."""
import ctypes
import json
import os
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, "_ref", "libmmref_terrain.so")


def available():
    return os.path.exists(LIB)


def run_session(moves, dt=1.0 / 32.0, skip_kernels=True, max_ticks=20000, timeout=600):
    """moves: player chunk positions; the session ticks until idle after each. Returns the dict described above."""
    spec = json.dumps({"moves": [list(m) for m in moves], "dt": dt, "skip_kernels": bool(skip_kernels), "max_ticks": max_ticks})
    out = subprocess.run([sys.executable, os.path.abspath(__file__), spec], capture_output=True, text=True, timeout=timeout)
    if out.returncode != 0 or not out.stdout.strip():
        raise RuntimeError("reference scheduler run failed (rc %d): %s" % (out.returncode, out.stderr[-2000:]))
    return json.loads(out.stdout.strip().splitlines()[-1])


def _main(spec):
    L = ctypes.CDLL(LIB)
    if L.mmrt_create(0, 1 if spec["skip_kernels"] else 0) != 0:
        raise SystemExit("mmrt_create failed")
    out9 = (ctypes.c_int * 9)()
    buf = (ctypes.c_int * 40000)()
    ticks, filled, eroded, built, segments = [], [], [], [], []
    for cx, cz in spec["moves"]:
        L.mmrt_set_player_chunk(int(cx), int(cz))
        segments.append(len(ticks))
        quiet = 0
        while quiet < 3 and len(ticks) < spec["max_ticks"]:
            L.mmrt_tick(ctypes.c_float(spec["dt"]), out9)
            ticks.append(list(out9))
            n = L.mmrt_take_eroded(buf, 20000)
            eroded += [[buf[2 * k] // 12, buf[2 * k + 1] // 12] for k in range(n)]
            n = L.mmrt_take_filled(buf, 20000)
            filled += [[buf[2 * k], buf[2 * k + 1]] for k in range(n)]
            n = L.mmrt_take_built(buf, 20000)
            built += [[buf[2 * k], buf[2 * k + 1]] for k in range(n)]
            quiet = quiet + 1 if sum(ticks[-1]) == 0 else 0
    print(json.dumps({"ticks": ticks, "filled": filled, "eroded": eroded, "built": built, "segments": segments}), flush=True)
    os._exit(0)      # the reference's teardown (file-static CUDA buffers) is not part of what is being checked


if __name__ == "__main__":
    _main(json.loads(sys.argv[1]))
