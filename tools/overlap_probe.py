#!/usr/bin/env python3
"""Developer tool (GPU box): do k_caves (FP32-issue-bound) and the fill kernels (latency-bound) gain from running
concurrently on two streams? World A runs S4 while world B runs S6; compare with each alone."""
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mmgen_loader  # noqa: E402

mm = mmgen_loader.load()
gen = mm.ChunkGen(0)
S = int(sys.argv[1]) if len(sys.argv) > 1 else 96
A = gen.region_world(0, 0, S, S)
B = gen.region_world(400, 0, S, S)
PRE_A = mm.STAGE_HEIGHTFIELD | mm.STAGE_LAYERS | mm.STAGE_EROSION
PRE_B = PRE_A | mm.STAGE_CAVES | mm.STAGE_FEATURES


def prep():
    A.reset(); B.reset()
    A.generate(PRE_A); B.generate(PRE_B)
    A.sync(); B.sync()


def run_a():
    A.generate(mm.STAGE_CAVES); A.sync()


def run_b():
    B.generate(mm.STAGE_FILL); B.sync()


for rep in range(3):
    prep(); t0 = time.perf_counter(); run_a(); ta = time.perf_counter() - t0
    prep(); t0 = time.perf_counter(); run_b(); tb = time.perf_counter() - t0
    prep()
    th = [threading.Thread(target=run_a), threading.Thread(target=run_b)]
    t0 = time.perf_counter()
    for t in th: t.start()
    for t in th: t.join()
    tc = time.perf_counter() - t0
    print("caves alone %.1f ms, fill alone %.1f ms, sum %.1f ms, concurrent %.1f ms (%.1f %% of the sum)" % (1e3 * ta, 1e3 * tb, 1e3 * (ta + tb), 1e3 * tc, 100 * tc / (ta + tb)))
