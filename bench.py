#!/usr/bin/env python3
"""Headless bench of the chunk-generation path: chunks/s for the full six-stage generation of a world
region (BASELINE.json metric), its roofline reading and two measured baselines.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--world S]
      own arm, BASELINE config 5: the S x S-chunk world [0, S)^2 (default 256). One process per GPU (torchrun for N > 1); the
      region is split into N chunk-coordinate tiles whose cuts come from a stage-1 cost predictor (every rank evaluates the
      features of a strip, all-gather, same cuts everywhere - repeated inside every timed step); each rank fills its tile in
      its own device-resident world => total work fixed => "scaling": "strong". The 3-chunk placement ring around a tile is
      either recomputed or exchanged over NCCL send / recv; both are measured before the timed steps and the faster is kept
      (--halo).
      `value`  = target chunks / device time of mmgen_world_generate (everything stays in HBM; CUDA events, MAX over ranks).
      `e2e`    = the same through mmgen_world_generate_to_host: chunk origins come from host memory and the raw block
                 volumes are delivered into pinned HOST memory inside the timed region.
      `e2e_encoded` = the same with the volumes run-length coded on the device (wire format MMCH1) - an extra, not the headline.
      `roofline`, `stages` = per-stage algorithmic FLOPs / bytes over device time; S1-S3 and the placement scan from
                 device work counters; ncu pipe figures of the committed captures next to the canonical-FLOP rates. Layers + erosion
                 run on a side stream concurrently with the caves (config.stage_overlap): stage times overlap.
  python bench.py --config c4      BASELINE config 4: cave + fill stress on 32x32 chunks, S4 and S6 timed in isolation.
  python bench.py --config c3      BASELINE config 3: 64x64-chunk streaming region at the reference's tick pattern (mmgen_stream_*).
  python bench.py --impl reference [...]
      the UNMODIFIED reference pipeline (chunk.cu built by oracle/Makefile into oracle/_ref) on the
      same GPU with its own batch caps, pinned staging buffers and CPU stages, on a bounded sample.
  python bench.py --impl reference-cpu [...]
      the CPU restatement (oracle/) on all host cores, on a bounded sample.

oracle/ is only executed here for the cpu_baseline / reference legs, never on the measured product path.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "chunks_per_sec_full_6stage_generation"
UNIT = "chunks/s"
# canonical algorithmic costs, SURVEY.md 8(d)
F_S2, F_S3, F_W3CELL, F_SIN = 140, 330, 60, 20
FLOP_CAVE_VOXEL = 23 * F_S3 + 27 * F_W3CELL + 81 * F_SIN          # one evaluated voxel of shouldGenerateCaveAtBlock
FLOP_CAVE_BIOME = 11 * F_S3 + 12 * F_S2                           # one getCaveBiome
# what k_caves executes of FLOP_CAVE_VOXEL after its exact early-outs (threshold bounds, huge-caves proof, tabulated cell hashes),
# counted by the census build (profiles/r02_census_v2.txt: warped noise + Worley at 77.2 % of the algorithmic voxels, fbmA at 19.3 % of
# those, the huge-caves term at 0.4 %, cell hashes tabulated)
EXECUTED_OVER_ALGORITHMIC_CAVES = 0.48
# dram__bytes_read.sum + dram__bytes_write.sum of one k_caves launch of 4096 chunks (ncu --set full, profiles/r02_k_caves_final.txt):
# 26.8 MB read + 407.9 MB written, against 107 528 algorithmic bytes per chunk (98 304 of them the CaveLayer output)
CAVES_DRAM_BYTES_PER_CHUNK = (26.836736e6 + 407.892992e6) / 4096
BYTES_FILL_CHUNK = 242688                                         # S6 compulsory I/O per chunk (without feature lists)
BYTES_CAVES_CHUNK = 107528
BYTES_S1_CHUNK, BYTES_S2_CHUNK = 25608, 46360
BYTES_S3_CELL_SWEEP = 12                                          # 2 planes read + 1 written per cell and sweep (1 769 472 B per 384^2 sweep)
# SURVEY.md 8(d) static call counts of getHeight per active biome, in canonical FLOPs: simplex2 140, Worley2 = 9 cells x 45 + 18 sinf x 20
F_W2 = 9 * 45 + 18 * F_SIN
FLOP_BIOME_HEIGHT = [5 * F_S2] * 24
for _b, _c in {1: 9 * F_S2, 14: 6 * F_S2, 16: 6 * F_S2, 23: 15 * F_S2, 8: 16 * F_S2 + 2 * F_W2, 9: 15 * F_S2 + F_W2, 13: 6 * F_S2 + F_W2,
               15: 9 * F_S2 + 2 * F_W2, 19: 6 * F_S2 + F_W2 + 3 * F_SIN}.items():
    FLOP_BIOME_HEIGHT[_b] = _c
FLOP_BIOME_NOISE = 11 * F_S2                                      # getBiomeNoise per column
FLOP_FBM5 = 5 * F_S2                                              # one stratified layer's thickness noise (S2)
# issue-slot utilisation of the S6 kernels from the committed ncu captures (the placement scan has no noise-primitive FLOP model:
# SURVEY.md 8(d) counts weights, smoothsteps and layer / list logic as 0 FLOP)
NCU_ISSUE = {"k_caves": (83.7, "profiles/r02_k_caves_final.txt"),
             "k_fill_rock": (70.6, "profiles/r02_k_fill_rock_final.txt (captured before the kernel's second regrouping level, DESIGN.md 5 item 30: "
                                   "the kernel is 7 % faster since and its later noise calls run on fuller warps than the 27.7 lanes listed)"),
             "k_fill_terrain": (80.8, "profiles/r02_k_fill_terrain_final.txt"), "k_fill_features": (75.6, "profiles/r02_k_fill_features_final.txt"),
             "k_erode_sweep": (38.3, "profiles/r02_k_erode_sweep_final.txt")}
# the same captures: FMA-pipe and ALU-pipe cycles active (% of peak), active lanes per executed warp instruction
NCU_PIPES = {"k_caves": (47.6, 58.5, 30.3), "k_fill_rock": (43.0, 49.6, 27.7), "k_fill_terrain": (15.3, 65.8, 23.1),
             "k_fill_features": (19.2, 59.0, 25.5), "k_erode_sweep": (10.6, 21.3, 29.9)}


def peaks():
    p = {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0, "src": "fallback"}
    f = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(f):
        try:
            j = json.load(open(f))
            p.update(hbm_gbs=float(j["hbm_gbs"]), sm_max_mhz=float(j.get("sm_max_mhz", 1965.0)), src="measured")
        except Exception:
            pass
    # no FP32 figure is measured by the driver: nominal CUDA-core peak = 148 SMs x 128 lanes x 2 FLOP x max clock
    p["fp32_tflops"] = 148 * 128 * 2 * p["sm_max_mhz"] * 1e6 / 1e12
    return p


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            if len(r) < 7:
                continue
            try:
                sm.append(float(r[0])); mx = float(r[1])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def prefer_gpu_numa_node(index):
    """Best effort: make this process allocate host memory on the NUMA node the GPU hangs off, before the pinned delivery buffer
    is created - at 8 GPUs the raw block volumes (6.4 GB per step) otherwise all land on the node the container's CPUs belong to
    and half the GPUs write across the socket link. Uses set_mempolicy(MPOL_PREFERRED) through libc; returns what happened."""
    try:
        import ctypes
        out = subprocess.run(["nvidia-smi", "-i", str(index), "--query-gpu=pci.bus_id", "--format=csv,noheader"], capture_output=True, text=True, timeout=20)
        bus = out.stdout.strip().lower()
        if bus.startswith("00000000:"):
            bus = bus[4:]                              # sysfs uses a 4-digit PCI domain
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bus).read().strip())
        nodes = sorted(int(d[4:]) for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit())
        if node < 0 or node not in nodes:
            return {"gpu_numa_node": node, "nodes": nodes, "policy": "unchanged (no NUMA information)"}
        libc = ctypes.CDLL("libc.so.6", use_errno=True)
        mask = ctypes.c_ulong(1 << node)
        rc = libc.syscall(238, 1, ctypes.byref(mask), 64)      # SYS_set_mempolicy (x86-64), MPOL_PREFERRED
        return {"gpu_numa_node": node, "nodes": nodes, "policy": "MPOL_PREFERRED" if rc == 0 else "unchanged (set_mempolicy errno %d)" % ctypes.get_errno()}
    except Exception as e:      # noqa: BLE001
        return {"policy": "unchanged (%s)" % type(e).__name__}


# ------------------------------------------------------------------------------------------ CPU port
def cpu_stage_rates(nthreads, cave_chunks=None, fill_chunks=None):
    """Times the oracle port stage by stage on one 26x26-chunk window (zone (0,0) + pad + ring, the C2
    window) and returns per-stage rates in chunks/s (zones/s for S3) on `nthreads` host threads."""
    from oracle import oracle as orc
    orc.build()
    o = orc.Oracle(nthreads)
    x0, z0, nx, nz = -7, -7, 26, 26
    origins = np.array([[(x0 + x) * 16, (z0 + z) * 16] for z in range(nz) for x in range(nx)], np.int32)
    t = {}
    t0 = time.perf_counter(); h, w = o.heightfields(origins); t["S1"] = (time.perf_counter() - t0, nx * nz)
    h18 = orc.gather_h18(h, nx, nz)
    inner = sorted(h18.keys())
    t0 = time.perf_counter()
    lay = np.zeros((nx * nz, 20, 256), np.float32)
    lay[inner] = o.layers(origins[inner], np.stack([h18[i] for i in inner]), w[inner])
    t["S2"] = (time.perf_counter() - t0, len(inner))
    planes = orc.gather_zone(lay, h, nx, 1, 1)
    t0 = time.perf_counter(); er, _ = o.erode_zone(planes); dt = time.perf_counter() - t0
    t["S3_zones"] = (dt / nthreads, 1)            # single-threaded per zone; zones are independent => one zone per thread
    orc.scatter_zone(er, lay, nx, 1, 1)
    zone = np.array([z * nx + x for z in range(7, 19) for x in range(7, 19)])
    sub = zone[:cave_chunks] if cave_chunks else zone
    t0 = time.perf_counter(); cl_sub = o.caves(origins[sub], h[sub], w[sub]); t["S4"] = (time.perf_counter() - t0, len(sub))
    t0 = time.perf_counter(); F, CF = o.feature_placements(origins[sub], h[sub], w[sub], lay[sub], cl_sub); t["S5"] = (time.perf_counter() - t0, len(sub))
    # fill: own placements only unless the whole zone was caved (the cost is dominated by the per-voxel noise)
    k = fill_chunks or min(len(sub), nthreads)
    fsel = sub[:k]
    gf = [F[i] for i in range(k)]
    gcf = [CF[i] for i in range(k)]
    t0 = time.perf_counter(); o.fill(origins[fsel], h[fsel], w[fsel], lay[fsel], cl_sub[:k], gf, gcf); t["S6"] = (time.perf_counter() - t0, k)
    return {s: n / max(dt, 1e-9) for s, (dt, n) in t.items()}, {s: {"seconds": round(dt, 3), "units": n} for s, (dt, n) in t.items()}


def cpu_whole_job_rate(rates, counts):
    """chunks/s the CPU port would reach on the bench workload: time = sum over stages of units / rate."""
    tt = sum(counts[s] / rates[s] for s in ("S1", "S2", "S3_zones", "S4", "S5", "S6"))
    return counts["S6"] / tt


# ------------------------------------------------------------------------------------------ arms
def run_reference_cuda(args):
    from oracle import refcuda
    if not refcuda.available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libmmref_cuda.so is not built (needs /root/reference at build time)"}))
        return
    r = refcuda.RefCuda(0)
    Z = args.ref_zones                                       # Z x Z zones + 6 chunks of pad + the layer ring
    x0 = z0 = -12 * (Z // 2) - 7
    n = 12 * Z + 14
    times, filled = [], 0
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        rc = r.L.mmref_generate(x0, z0, n, n, 6, 0)
        dt = time.perf_counter() - t0
        if rc != 0:
            raise RuntimeError("mmref_generate failed: %d" % rc)
        if i >= args.warmup:
            times.append(dt)
        filled = sum(1 for k in range(n * n) if r.L.mmref_stage(k) == 6)
    stage_ms = [r.L.mmref_stage_ms(s) for s in range(8)]
    total = sum(times)
    val = filled * len(times) / total
    sample = "%dx%d-chunk window (%d chunks filled per step; every stage on everything the reference state machine reaches), " \
             "reference batch caps 166/100/62/62, pinned staging, 1 host thread + its CUDA kernels on GPU 0" % (n, n, filled)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic (the world is a pure function of chunk coordinates; no dataset exists)",
        "config": {"workload": "full 6-stage generation, reference CUDA pipeline (unmodified chunk.cu for sm_100) incl. its CPU stages", "sample": sample,
                   "same_config_note": "a bounded sample, not the 256x256 world: %d chunks filled per step; chunks/s is per filled chunk and includes the "
                                       "apron (S1-S3 on the whole window are < 2 %% of the reference's step)" % filled},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": 1, "kind": "reference", "sample": sample,
                         "note": "the reference has no CPU generator: its own implementation of the path is CUDA kernels driven by one host thread"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "stage_wall_ms": {"S1": stage_ms[1], "S2": stage_ms[2], "S3": stage_ms[3], "S4": stage_ms[4], "S5a": stage_ms[5], "S5b": stage_ms[6], "S6": stage_ms[7]},
    }))


def run_reference_cpu(args):
    import mmgen_loader
    mmgen_loader.load()
    from mega_minecraft_b200 import tiling
    nthreads = os.cpu_count() or 1
    counts = tiling.stage_chunk_counts(0, 0, args.world, args.world)
    vals = []
    for i in range(max(1, min(args.steps, 2))):
        rates, detail = cpu_stage_rates(nthreads)
        vals.append(cpu_whole_job_rate(rates, counts))
    val = float(np.mean(vals))
    sample = "oracle port, one 26x26-chunk window per step timed stage by stage, extrapolated with the workload's per-stage chunk counts"
    print(json.dumps({"impl": "reference-cpu", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": 0, "higher_is_better": True,
                      "config": {"workload": "%dx%d-chunk world, full 6-stage generation" % (args.world, args.world)},
                      "cpu_baseline": {"value": val, "unit": UNIT, "cores": nthreads, "kind": "port", "sample": sample},
                      "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "stage_detail": detail}))


def cheap_stage_rooflines(wc, steps, kernels, fp32_peak, hbm_peak):
    """Rooflines of S1 / S2 / S3 from the library's work counters (mmgen_work_counters), per step of this rank."""
    wc = [float(v) / steps for v in wc]
    ms = lambda k: kernels.get(k, {"ms_per_step": 0.0})["ms_per_step"]
    f1 = sum(wc[b] * FLOP_BIOME_HEIGHT[b] for b in range(24)) + wc[24] * FLOP_BIOME_NOISE
    f2 = wc[25] * FLOP_FBM5
    b3 = wc[27] * 1024 * BYTES_S3_CELL_SWEEP
    s1 = {"algorithmic_flop": f1, "fp32_tflops": f1 / 1e12 / max(ms("k_heightfield") / 1e3, 1e-9), "columns": wc[24],
          "active_biomes_per_column": sum(wc[:24]) / max(wc[24], 1.0), "hbm_gbs": wc[24] / 256 * BYTES_S1_CHUNK / 1e9 / max(ms("k_heightfield") / 1e3, 1e-9)}
    s2 = {"algorithmic_flop": f2, "fp32_tflops": f2 / 1e12 / max(ms("k_layers") / 1e3, 1e-9), "fbm5_per_column": wc[25] / max(wc[26], 1.0),
          "hbm_gbs": wc[26] / 256 * BYTES_S2_CHUNK / 1e9 / max(ms("k_layers") / 1e3, 1e-9)}
    s3 = {"tiles_swept": wc[27], "tiles_launched": wc[27] + wc[28], "algorithmic_bytes": b3,
          "plane_traffic_gbs": b3 / 1e9 / max(ms("k_erode_sweep") / 1e3, 1e-9)}
    for d in (s1, s2):
        d["fp32_frac"] = d["fp32_tflops"] / fp32_peak
        d["hbm_frac"] = d["hbm_gbs"] / hbm_peak
        d["bound"] = "fp32"
    s3["frac_of_hbm_peak"] = s3["plane_traffic_gbs"] / hbm_peak
    s3["bound"] = "L2 bandwidth + launch latency (the planes of a 32-zone batch, 75 MB, stay in the 126 MB L2; HBM peak is the only measured " \
                  "bandwidth figure, so the fraction is against it)"
    return s1, s2, s3


def run_c4(args):
    """BASELINE.json config 4: cave / cave-biome + chunk-fill stress on [0,32)^2 - S4 and S6 timed in isolation with S1-S3 (and S5
    for the fill) precomputed and resident."""
    import torch
    import mmgen_loader
    mm = mmgen_loader.load()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the generation path has no CPU fallback")
    gen = mm.ChunkGen(0)
    pk = peaks()
    world = gen.region_world(0, 0, 32, 32)
    world.generate(mm.STAGE_ALL)
    world.sync()
    n_target = 1024
    st = world.stages().ravel()
    n_caved = int((st >= 4).sum())
    hgt = np.floor(world.download(heightfield=True)["heightfield"]).astype(np.int64)
    cave_voxels = int(np.maximum(hgt[st >= 4], 128).sum())
    fill_voxels = int(np.clip(hgt[st == 6], 1, 383).sum())
    fp32 = min(gen.measure_fp32_peak(), pk["fp32_tflops"])
    sampler = ClockSampler(0)
    t4 = t6 = 0.0
    for i in range(args.warmup + args.steps):
        if i == args.warmup:
            torch.cuda.synchronize()
            gen.kernel_timing(True)
            sampler.start()
            l0 = gen.launch_count()
        world.rewind(3)
        world.generate(mm.STAGE_CAVES)
        a = float(world.stage_ms()[4])
        world.generate(mm.STAGE_FEATURES)
        world.generate(mm.STAGE_FILL)
        b = float(world.stage_ms()[6])
        if i >= args.warmup:
            t4 += a
            t6 += b
    torch.cuda.synchronize()
    clocks = sampler.stop()
    kt = gen.kernel_times()
    gen.kernel_timing(False)
    launches = gen.launch_count() - l0
    t4 /= args.steps
    t6 /= args.steps
    assert int((world.stages() == 6).sum()) == n_target
    kernels = {k: {"ms_per_step": v[0] / args.steps, "launches_per_step": v[1] / args.steps} for k, v in kt.items() if v[1]}
    s4 = {"ms": t4, "chunks": n_caved, "chunks_per_s": n_caved / (t4 / 1e3), "fp32_tflops": cave_voxels * FLOP_CAVE_VOXEL / 1e12 / (t4 / 1e3),
          "hbm_gbs": n_caved * BYTES_CAVES_CHUNK / 1e9 / (t4 / 1e3), "executed_over_algorithmic": EXECUTED_OVER_ALGORITHMIC_CAVES}
    s6 = {"ms": t6, "chunks": n_target, "chunks_per_s": n_target / (t6 / 1e3), "fp32_tflops": fill_voxels * FLOP_CAVE_BIOME / 1e12 / (t6 / 1e3),
          "hbm_gbs": n_target * BYTES_FILL_CHUNK / 1e9 / (t6 / 1e3)}
    for d in (s4, s6):
        d["fp32_frac"] = d["fp32_tflops"] / fp32
        d["hbm_frac"] = d["hbm_gbs"] / pk["hbm_gbs"]
    s4["executed_fp32_frac"] = s4["fp32_frac"] * EXECUTED_OVER_ALGORITHMIC_CAVES
    dom = "k_caves"
    out = {"metric": "c4_cave_and_fill_stress_chunks_per_sec", "value": n_target / ((t4 + t6) / 1e3), "unit": UNIT, "n_gpus": 1, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": t4 + t6, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
           "data": "synthetic (the world is a pure function of chunk coordinates; no dataset exists)",
           "config": {"workload": "BASELINE config 4: 32x32 chunks [0,32)^2 of full 16x384x16 volumes; S4 (caves + cave biomes, on the region (+) 3 chunks "
                                  "= %d chunks) and S6 (fill + decorators, 1024 chunks) each timed in isolation by CUDA events, S1-S3 and S5 resident" % n_caved,
                      "l2": "S4 writes %.0f MB and S6 %.0f MB per step: beyond the 126 MB L2 only for S6; no flush (inputs of a step were written by the "
                            "previous stage of the same step)" % (n_caved * 98304 / 1e6, n_target * 98304 / 1e6)},
           "stages": {"S4": s4, "S6": s6}, "kernels": kernels, "gpu_launches": int(launches), "clocks": clocks,
           "roofline": {"bound": "fp32", "kernel": dom, "achieved": s4["fp32_tflops"], "peak": fp32, "unit": "TFLOP/s", "frac": s4["fp32_frac"],
                        "traffic": CAVES_DRAM_BYTES_PER_CHUNK * n_caved, "executed": {"frac": s4["executed_fp32_frac"]},
                        "note": "S6 as HBM: %.1f GB/s of compulsory bytes = %.4f of the measured %.0f GB/s; it is bound by per-voxel noise and the placement "
                                "scan, not by HBM" % (s6["hbm_gbs"], s6["hbm_frac"], pk["hbm_gbs"])},
           "e2e": None, "cpu_baseline": None}
    print(json.dumps(out), flush=True)
    world.close()


def run_c3(args):
    """BASELINE.json config 3: a 64x64-chunk region streamed around a standing player at the reference's load pattern (spiral order,
    per-stage FIFO queues, action-time budget: Terrain::tick, terrain.cpp:587-960) through mmgen_stream_*."""
    import torch
    import mmgen_loader
    mm = mmgen_loader.load()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the generation path has no CPU fallback")
    gen = mm.ChunkGen(0)
    R = 51                                     # generation radius: fills the 66x66 chunks around the player
    profiles = {"reference": (500, 60 * 500), "budget_x8": (4000, 60 * 4000), "unbounded": (1 << 24, 1 << 30)}
    res = {}
    sampler = ClockSampler(0)
    launches = 0
    for name, (cap, rate) in profiles.items():
        runs = []
        for rep in range(args.warmup + args.steps):
            if name == "reference" and rep == args.warmup:
                sampler.start()
            l0 = gen.launch_count()
            t = mm.Terrain(gen, -R - 1, -R - 1, 2 * R + 2, 2 * R + 2)
            t.set_radii(16, R)
            t.set_costs(mm.REFERENCE_COSTS, cap, rate)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            log = t.run_until_idle(1.0 / 32.0)
            torch.cuda.synchronize()
            wall = time.perf_counter() - t0
            filled = sum(s["filled"] for s in log)
            dev = sum(s["deviceMs"] for s in log)
            h = t.chunk_hash_sum()
            t.close()
            if rep >= args.warmup:
                runs.append((wall, dev, filled, len(log)))
                if name == "reference":
                    launches += gen.launch_count() - l0
        if name == "reference":
            clocks = sampler.stop()
        # one more session, untimed, with CUDA event pairs around every hot-kernel launch: where the device time of the ticks goes
        gen.kernel_timing(True)
        t = mm.Terrain(gen, -R - 1, -R - 1, 2 * R + 2, 2 * R + 2)
        t.set_radii(16, R)
        t.set_costs(mm.REFERENCE_COSTS, cap, rate)
        t.run_until_idle(1.0 / 32.0)
        torch.cuda.synchronize()
        t.close()
        ktimes = {k: {"ms": round(v[0], 3), "launches": int(v[1])} for k, v in gen.kernel_times().items() if v[1]}
        gen.kernel_timing(False)
        wall = sum(r[0] for r in runs) / len(runs)
        dev = sum(r[1] for r in runs) / len(runs)
        res[name] = {"ticks": runs[0][3], "chunks_filled": runs[0][2], "wall_ms": 1e3 * wall, "device_ms": dev, "chunks_per_s": runs[0][2] / wall,
                     "frame_budget": cap, "hash": "%016x" % h, "kernels": ktimes,
                     "max_batch": {k: max(s[k] for s in log) for k in ("heightfields", "layers", "caves", "filled", "zonesEroded")}}
    r = res["reference"]
    out = {"metric": "c3_streaming_chunks_per_sec", "value": r["chunks_per_s"], "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": r["wall_ms"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
           "data": "synthetic (the world is a pure function of chunk coordinates; no dataset exists)",
           "config": {"workload": "BASELINE config 3: 66x66 chunks streamed around a standing player (generation radius 51) by the re-hosted Terrain::tick "
                                  "at the reference's action-time costs (<= 166 heightfields / 100 layers / 62 caves / 62 fills / 1 zone per tick); "
                                  "a step = one whole session from an empty world to idle, wall clock incl. every per-tick synchronisation",
                      "timing": "wall clock around the session (the scheduler is host code; device time of the ticks' launches is in device_ms)"},
           "profiles": res, "gpu_launches": int(launches), "clocks": clocks, "e2e": None, "cpu_baseline": None,
           "roofline": None}
    print(json.dumps(out), flush=True)


def run_own(args):
    import torch
    import torch.distributed as dist
    import mmgen_loader
    mm = mmgen_loader.load()
    from mega_minecraft_b200 import tiling

    rank = int(os.environ.get("RANK", "0"))
    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # libraries (NCCL's version banner) write to fd 1: send everything but the final JSON line to stderr
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the generation path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world_size > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        torch.cuda.synchronize()
        if world_size > 1:
            dist.barrier()
        torch.cuda.synchronize()

    from mega_minecraft_b200 import sharding

    def max_over_ranks(x):
        return sharding.reduce_scalar(x, "max")

    def sum_over_ranks(x):
        return sharding.reduce_scalar(x, "sum")

    S = args.world
    gen = mm.ChunkGen(local_rank)
    fill_overlap = mm.FILL_OVERLAP_DEFAULT if args.fill_overlap is None else args.fill_overlap
    gen.set_fill_overlap(fill_overlap)
    region = (0, 0, S, S)
    # Tiling (N > 1): the cuts are placed by a cost PREDICTOR that needs stage 1 only - every rank evaluates the stage-1 cost
    # features of a strip of the region (mmgen_chunk_costs, < 1 ms), the strips are all-gathered and every rank derives the same
    # cuts (sharding.chunk_cost_map / Balancer.cut_by_cost). The prediction is repeated inside every timed step and its wall time
    # is part of the step. --balance-rounds > 0 adds the round-1 feedback passes (untimed full generations) on top.
    bal = sharding.Balancer(region, world_size)
    balance_log = []
    predictor_ms = 0.0
    if world_size > 1 and not args.equal_tiles:
        t0 = time.perf_counter()
        bal.cut_by_cost(sharding.chunk_cost_map(gen, region, rank, world_size))
        predictor_ms = 1e3 * (time.perf_counter() - t0)
    rounds = args.balance_rounds if world_size > 1 else 0
    world = None
    for it in range(rounds + 1):
        tile = bal.tiles()[rank]
        world = gen.region_world(*tile)
        if it == rounds:
            break
        world.generate(mm.STAGE_ALL)            # allocations
        world.reset()
        world.generate(mm.STAGE_ALL)
        times = sharding.gather_floats(world.total_ms())
        balance_log.append({"tiles": [list(t) for t in bal.tiles()], "rank_ms": [round(t, 2) for t in times]})
        bal.update(times)
        world.close()
    tiles_final = bal.tiles()
    n_target = tile[2] * tile[3]
    numa = prefer_gpu_numa_node(local_rank) if world_size > 1 else {"policy": "unchanged (one GPU)"}
    host = torch.empty(n_target * 98304, dtype=torch.uint8, pin_memory=True)
    exchange = sharding.HaloExchange(tiles_final, rank) if world_size > 1 else None

    def predict():
        """The in-step part of the balancing: stage-1 cost features of this rank's strip, all-gather, cuts. Returns wall ms."""
        if world_size == 1 or args.equal_tiles or rounds:
            return 0.0
        t0 = time.perf_counter()
        b2 = sharding.Balancer(region, world_size)
        b2.cut_by_cost(sharding.chunk_cost_map(gen, region, rank, world_size))
        assert b2.tiles() == tiles_final, "the predictor is deterministic: the tiles of the resident worlds stay valid"
        return 1e3 * (time.perf_counter() - t0)

    def step(variant, to_host=False):
        """One generation of this rank's tile. Returns (ms, exchange ms): device time of the generate call(s) (CUDA events on the
        world's stream) + wall time of the predictor and, for the exchange variant, of pack / NCCL send-recv / unpack."""
        ms = predict()
        step.predict_ms += ms
        world.reset()
        if variant == "exchange":
            world.generate(mm.STAGE_ALL & ~mm.STAGE_FILL)
            ms += world.total_ms()
            t0 = time.perf_counter()
            exchange.run(world)
            x = 1e3 * (time.perf_counter() - t0)
            deliver(mm.STAGE_FILL, to_host)
            return ms + x + world.total_ms(), x
        deliver(mm.STAGE_ALL, to_host)
        return ms + world.total_ms(), 0.0

    enc_index = torch.empty(n_target * 2, dtype=torch.int64, pin_memory=True)
    enc_bytes = [0]

    def deliver(mask, to_host):
        if to_host == "encoded":      # run-length coded on the device (format MMCH1), payload + index into pinned host memory
            enc_bytes[0] = world.generate_to_host_encoded(host.data_ptr(), host.numel(), enc_index.numpy().view(np.uint64).reshape(-1, 2), mask)
        elif to_host:
            world.generate_to_host(host.data_ptr(), mask)
        else:
            world.generate(mask)

    step.predict_ms = 0.0

    def set_variant(variant):
        if variant == "exchange":
            world.set_exchange_region(*region)
        else:
            world.set_exchange_region(0, 0, 0, 0)

    # ---- which halo variant is faster on this box? (north_star: "recomputed redundantly or exchanged ... whichever measures faster")
    variants = {}
    if world_size > 1 and args.halo == "auto":
        for v in ("recompute", "exchange"):
            set_variant(v)
            step(v)                               # allocations / NCCL channels
            barrier()
            tms = 0.0
            for _ in range(2):
                tms += step(v)[0]
            variants[v] = max_over_ranks(tms / 2)
            barrier()
        variant = min(variants, key=variants.get)
    else:
        variant = args.halo if (world_size > 1 and args.halo != "auto") else "recompute"
    set_variant(variant)
    launches0 = gen.launch_count()

    for i in range(args.warmup):
        if i == args.warmup - 1:
            gen.kernel_timing(True)              # the last warm-up step creates the CUDA events the per-kernel timing of the timed steps reuses
        step(variant)
    gen.kernel_times()
    gen.kernel_timing(False)
    world.sync()
    assert int((world.stages() == 6).sum()) == n_target, "not every target chunk was filled"
    checksum = world.block_checksum()
    hash_sum = world.chunk_hash_sum()

    # ---- device-resident leg
    sampler = ClockSampler(local_rank)
    stage_ms = np.zeros(7)
    fp32_measured = gen.measure_fp32_peak()          # FFMA microbenchmark on this GPU, right before the timed region
    gen.kernel_timing(True)                          # CUDA event pairs around every hot-kernel launch, on the world's stream
    gen.work_counters(reset=True)                    # S1 / S2 / S3 work counters of the timed steps
    sampler.start()                                  # forks nvidia-smi: tens of ms, different on every rank - before the barrier
    barrier()
    l0 = gen.launch_count()
    t0 = time.perf_counter()
    dev_ms = 0.0
    xch_ms = 0.0
    step.predict_ms = 0.0
    ktimes = {}
    t_step = t_book = 0.0
    for _ in range(args.steps):
        ta = time.perf_counter()
        a, x = step(variant)
        tb = time.perf_counter()
        dev_ms += a
        xch_ms += x
        stage_ms += world.stage_ms()
        for k, v in gen.kernel_times().items():      # {kernel: (device ms, launches)} of this step; frees the event pairs for the next one
            ktimes[k] = (ktimes.get(k, (0.0, 0))[0] + v[0], ktimes.get(k, (0.0, 0))[1] + v[1])
        t_step += tb - ta
        t_book += time.perf_counter() - tb
    rank_ms = [t / args.steps for t in sharding.gather_floats(dev_ms)]
    predict_ms_per_step = step.predict_ms / args.steps
    # where a rank's wall time goes per step: inside step() (predictor + generate + exchange, host view) and the bookkeeping between steps
    rank_wall = [[round(1e3 * v / args.steps, 2) for v in sharding.gather_floats(q)] for q in (t_step, t_book)]
    rank_predict = [round(v / args.steps, 2) for v in sharding.gather_floats(step.predict_ms)]
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop()
    wcount = gen.work_counters(reset=True)
    gen.kernel_timing(False)
    launches = gen.launch_count() - l0
    dev_s = max_over_ranks(dev_ms / 1e3)
    wall_s = max_over_ranks(wall)
    total_chunks = S * S
    value = total_chunks * args.steps / dev_s

    # ---- S6 kernel times without the cross-batch overlap (one extra untimed step): with mmgen_set_fill_overlap on, the event pairs of the
    # timed steps bracket kernels that share the SMs with another stream's kernels (or wait behind them), so their elapsed times overlap
    ktimes_s6 = {}
    if fill_overlap:
        gen.set_fill_overlap(0)
        gen.kernel_timing(True)
        step(variant)
        ktimes_s6 = {k: v[0] for k, v in gen.kernel_times().items() if v[1]}
        gen.kernel_timing(False)
        gen.set_fill_overlap(fill_overlap)

    # ---- end-to-end leg: origins from host memory, block volumes into pinned host memory
    step(variant, to_host=True)
    barrier()
    t0 = time.perf_counter()
    e2e_ms = 0.0
    for _ in range(args.steps):
        e2e_ms += step(variant, to_host=True)[0]
    barrier()
    e2e_wall = max_over_ranks(time.perf_counter() - t0)
    e2e_dev = max_over_ranks(e2e_ms / 1e3)
    e2e_value = total_chunks * args.steps / max(e2e_wall, e2e_dev)
    host_sum = int(host.view(torch.int64).sum().item()) if rank == 0 else 0
    sample_slots = np.unique(np.linspace(0, n_target - 1, 64).astype(np.int64))       # raw volumes kept to check the encoded delivery against
    raw_sample = host.numpy().reshape(-1)[: n_target * 98304].reshape(n_target, 98304)[sample_slots].copy()

    # ---- the same with the block volumes delivered in the library's wire format (MMCH1 run-length code, decoded on the host by
    # mmgen_decode_chunk): an extra figure, not the headline e2e - the reference's contract is the raw volume
    step(variant, to_host="encoded")
    barrier()
    t0 = time.perf_counter()
    enc_ms = 0.0
    for _ in range(args.steps):
        enc_ms += step(variant, to_host="encoded")[0]
    barrier()
    enc_wall = max_over_ranks(time.perf_counter() - t0)
    enc_dev = max_over_ranks(enc_ms / 1e3)
    enc_value = total_chunks * args.steps / max(enc_wall, enc_dev)
    enc_total = sum_over_ranks(enc_bytes[0])
    idx = enc_index.numpy().view(np.uint64).reshape(-1, 2)      # every rank: 64 chunks of its tile decode to the raw delivery's bytes
    for k, slot in enumerate(sample_slots):
        blk = mm.decode_chunk(host.numpy()[int(idx[slot, 0]):int(idx[slot, 0] + idx[slot, 1])])
        assert np.array_equal(blk.reshape(-1), raw_sample[k]), "encoded delivery differs from the raw delivery at slot %d" % slot
    h2d = sum_over_ranks(world.n * 8)
    d2h = sum_over_ranks(n_target * 98304)

    # ---- roofline of the dominant kernel (algorithmic FLOPs from the heightfield)
    stage_ms /= args.steps
    hgt = world.download(heightfield=True)["heightfield"]
    st = world.stages().ravel()
    hi = np.floor(hgt).astype(np.int64)
    cave_voxels = int(np.maximum(hi[st >= 4], 128).sum())                       # 0 < y <= max(floor(h), 128)
    fill_voxels = int(np.clip(hi[st == 6], 1, 383).sum())                       # 0 < y <= h: getCaveBiome per voxel
    pk = peaks()
    flop_s4 = cave_voxels * FLOP_CAVE_VOXEL
    flop_s6 = fill_voxels * FLOP_CAVE_BIOME
    ach4 = sum_over_ranks(flop_s4) / 1e12 / max(max_over_ranks(stage_ms[4] / 1e3), 1e-9)
    ach6 = sum_over_ranks(flop_s6) / 1e12 / max(max_over_ranks(stage_ms[6] / 1e3), 1e-9)
    hbm6 = sum_over_ranks(n_target * BYTES_FILL_CHUNK) / 1e9 / max(max_over_ranks(stage_ms[6] / 1e3), 1e-9)
    # per-kernel device time per step (this rank), from the event pairs recorded during the timed steps
    kernels = {k: {"ms_per_step": v[0] / args.steps, "launches_per_step": v[1] / args.steps} for k, v in ktimes.items() if v[1]}
    # dominant kernel = the single kernel with the most device time; its algorithmic FLOPs are defined for k_caves (every
    # evaluated voxel of shouldGenerateCaveAtBlock) and k_fill_terrain (getCaveBiome per voxel 0 < y <= h, chunk.cu:1243-1370)
    flops_of = {"k_caves": flop_s4, "k_fill_terrain": flop_s6}
    dom = max(kernels, key=lambda k: kernels[k]["ms_per_step"]) if kernels else "k_caves"
    roof_kernel = dom if dom in flops_of else max(flops_of, key=lambda k: kernels.get(k, {"ms_per_step": 0})["ms_per_step"])
    rk = kernels.get(roof_kernel, {"ms_per_step": float("nan"), "launches_per_step": 1})
    rk_launches = max(rk["launches_per_step"], 1)
    avg_launch_ms = rk["ms_per_step"] / rk_launches
    ach = sum_over_ranks(flops_of[roof_kernel]) / world_size / rk_launches / 1e12 / max(avg_launch_ms / 1e3, 1e-12)      # per GPU
    fp32_all = max_over_ranks(fp32_measured)
    fp32_peak = min(fp32_all, pk["fp32_tflops"]) if fp32_all > 0 else pk["fp32_tflops"]
    roof = {"bound": "fp32", "kernel": roof_kernel, "achieved": ach, "peak": fp32_peak, "unit": "TFLOP/s", "frac": ach / fp32_peak,
            "traffic": (CAVES_DRAM_BYTES_PER_CHUNK * int((st >= 4).sum()) / rk_launches if roof_kernel == "k_caves" else None),
            "traffic_src": "ncu capture of one 4096-chunk k_caves launch (profiles/r02_k_caves_final.txt), scaled to this run's chunks per launch; "
                           "algorithmic bytes per chunk: 107528",
            "avg_launch_ms": avg_launch_ms, "launches_per_step": rk_launches,
            "algorithmic_flop_per_launch": flops_of[roof_kernel] / rk_launches,
            "executed": ({"flop_ratio_to_algorithmic": EXECUTED_OVER_ALGORITHMIC_CAVES, "achieved": ach * EXECUTED_OVER_ALGORITHMIC_CAVES,
                          "frac": ach * EXECUTED_OVER_ALGORITHMIC_CAVES / fp32_peak,
                          "src": "census build (profiles/r02_census_v2.txt): evaluations proved unnecessary are not executed. Both figures are work "
                                 "rates in SURVEY.md 8(d)'s canonical FLOPs (simplex3 = 330 FLOP, a hashed Worley cell 60 + 3 x 20), not pipe "
                                 "occupancy: this implementation reads the lattice hash and the gradients from tables and retires two fp32 "
                                 "operations per packed instruction, so a canonical FLOP costs it less than a FLOP. What the hardware saw is in `ncu`."}
                         if roof_kernel == "k_caves" else None),
            "ncu": {"issue_slot_utilisation_pct": NCU_ISSUE[roof_kernel][0], "fma_pipe_pct": NCU_PIPES[roof_kernel][0],
                    "alu_pipe_pct": NCU_PIPES[roof_kernel][1], "active_lanes_of_32": NCU_PIPES[roof_kernel][2], "src": NCU_ISSUE[roof_kernel][1]}
                   if roof_kernel in NCU_ISSUE else None,
            "peak_src": "FFMA microbenchmark run by this process on the same GPU just before the timed region (mmgen_measure_fp32_peak: %.1f TFLOP/s; "
                        "nominal 148 SM x 128 lanes x 2 x %.0f MHz = %.1f); MEASURED_PEAKS.json carries HBM and bf16 tensor figures only" % (
                            fp32_measured, pk["sm_max_mhz"], pk["fp32_tflops"]),
            "dominant_kernel_by_time": dom,
            "note": "no stage is a dense contraction and none is HBM-bound (S6 moves %.0f GB/s of compulsory bytes), so the bound is the FP32 pipe; "
                    "algorithmic FLOPs = noise-primitive calls of the reference algorithm x canonical cost (SURVEY.md 8d), evaluations this "
                    "implementation proves unnecessary still count; k_fill_features (placement rasterisation) has no FLOP model and is reported by time" % hbm6}
    r1, r2, r3 = cheap_stage_rooflines(wcount, args.steps, kernels, fp32_peak, pk["hbm_gbs"])      # this rank's counters and kernel times
    s6_serial_ms = {k: ktimes_s6.get(k, kernels[k]["ms_per_step"]) for k in kernels}      # = the timed steps' figures when the overlap is off
    if ktimes_s6:
        # the S6 kernels' entries carry the time the kernel takes by itself (extra step, overlap off); what the event pair of the timed
        # steps spans - the kernel sharing the SMs with the other stream's kernels, or queued behind them - is kept next to it
        for k in ("k_gather_features", "k_fill_terrain", "k_fill_rock", "k_fill_lush", "k_prepare_placements", "k_fill_features", "k_decorators"):
            if k in kernels and k in ktimes_s6:
                kernels[k] = {"ms_per_step": ktimes_s6[k], "launches_per_step": kernels[k]["launches_per_step"],
                              "elapsed_ms_in_timed_steps": kernels[k]["ms_per_step"], "src": "extra untimed step with mmgen_set_fill_overlap(0)"}
    s6_sum = sum(s6_serial_ms.get(k, 0.0) for k in ("k_fill_terrain", "k_fill_rock", "k_fill_lush", "k_prepare_placements", "k_fill_features", "k_decorators"))
    s6k = {k: {"ms": s6_serial_ms[k], "share_of_S6": s6_serial_ms[k] / max(s6_sum, 1e-9),
               "issue_slot_utilisation_pct": NCU_ISSUE[k][0], "fma_pipe_pct": NCU_PIPES[k][0], "alu_pipe_pct": NCU_PIPES[k][1],
               "active_lanes_of_32": NCU_PIPES[k][2], "src": NCU_ISSUE[k][1]}
           for k in ("k_fill_terrain", "k_fill_rock", "k_fill_features") if k in kernels}
    if "k_fill_features" in s6k:
        # work model of the placement scan from the kernel's own counters (mmgen_work_counters [29..31], this rank, per step)
        placements, pairs, rast = (float(wcount[i]) / args.steps for i in (29, 30, 31))
        kms = max(s6k["k_fill_features"]["ms"], 1e-9)
        s6k["k_fill_features"].update({
            "algorithmic_placement_tests": placements * 98304.0,      # the reference tests every gathered placement at every voxel of the chunk
            "pairs_examined": pairs, "pairs_rasterised": rast, "examined_over_algorithmic": pairs / max(placements * 98304.0, 1.0),
            "algorithmic_tests_per_s": placements * 98304.0 / (kms * 1e-3), "rasterised_pairs_per_s": rast / (kms * 1e-3),
            "note": "no FLOP model (SURVEY.md 8(d): integer / shared-memory work): algorithmic work = gathered placements x 98 304 voxels as the "
                    "reference scans them; pairs_examined = (column, y) pairs inside the placements' clipped boxes, pairs_rasterised = the ones "
                    "that were neither claimed by an earlier placement nor excluded by the air test"})
    stages = {
        "S1": dict(r1, ms=float(stage_ms[1])), "S2": dict(r2, ms=float(stage_ms[2])),
        "S3": dict(r3, ms=float(stage_ms[3]), sweeps=world.erosion_sweeps(), issue_slot_utilisation_pct=NCU_ISSUE["k_erode_sweep"][0],
                   note="runs on the side stream concurrently with S4 (config.stage_overlap): its kernel and stage times are stretched by the sharing "
                        "and are not what the stage costs the step - serial it takes 27 ms per 256x256 world (profiles/r02_stage_overlap.txt), "
                        "overlapped the step grows by about 12 ms over a step without erosion"),
        "S4": {"ms": float(stage_ms[4]), "bound": "fp32", "fp32_tflops": ach4, "fp32_frac": ach4 / (fp32_peak * world_size),
               "executed_fp32_frac": ach4 / (fp32_peak * world_size) * EXECUTED_OVER_ALGORITHMIC_CAVES,
               "hbm_gbs": sum_over_ranks(int((st >= 4).sum()) * BYTES_CAVES_CHUNK) / 1e9 / max(max_over_ranks(stage_ms[4] / 1e3), 1e-9),
               "issue_slot_utilisation_pct": NCU_ISSUE["k_caves"][0]},
        "S5": {"ms": float(stage_ms[5]), "bound": "latency (integer / RNG walk per column; time only, SURVEY.md 8(d))"},
        "S6": {"ms": float(stage_ms[6]), "bound": "fp32 / issue (per-voxel noise + placement scan), not HBM", "fp32_tflops": ach6,
               "fp32_frac": ach6 / (fp32_peak * world_size), "hbm_gbs": hbm6, "hbm_frac": hbm6 / (pk["hbm_gbs"] * world_size),
               "hbm_peak_src": pk["src"], "kernels": s6k,
               "note": "algorithmic FLOPs = getCaveBiome (11 simplex3 + 12 simplex2) per voxel 0 < y <= h as the reference evaluates it; the placement "
                       "scan (k_fill_features) is integer / shared-memory work that SURVEY.md 8(d) counts as 0 FLOP: it is reported by time and by "
                       "the issue-slot utilisation of its ncu capture"},
    }
    counts = tiling.stage_chunk_counts(*tile)
    checks = sharding.gather_u64(checksum)

    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world_size, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dev_s / args.steps, "wall_ms_per_step": 1e3 * wall_s / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic (the world is a pure function of chunk coordinates; no dataset exists)",
        "config": {"workload": "%dx%d-chunk world [0,%d)^2, full 6-stage generation (S1 heightfield/biomes, S2 layers, S3 erosion, S4 caves, "
                               "S5 feature placement+gather, S6 fill+decorators)" % (S, S, S),
                   "tiling": ("one tile (the whole region)" if world_size == 1 else "%d chunk-coordinate tiles; halo variant '%s'%s; cuts %s%s" % (
                       world_size, variant,
                       " (placement lists of the 3-chunk ring exchanged with NCCL send/recv: %d bytes sent by rank 0 per step, %.2f ms per step incl. "
                       "pack / unpack; stages 4 + 5a run on the own tile only)" % (exchange.bytes_sent, xch_ms / args.steps) if variant == "exchange"
                       else " (apron recomputed per tile, no data-path collective)",
                       "equal" if args.equal_tiles else "from the stage-1 cost predictor, recomputed inside every timed step (%.2f ms of wall time per step, part of "
                       "the step time; %.1f ms at set-up incl. allocations)" % (predict_ms_per_step, predictor_ms),
                       "; then moved by feedback from %d untimed passes (per-rank device time)" % rounds if rounds else "")),
                   "halo_variants_ms_per_step": {k: round(v, 2) for k, v in variants.items()},
                   "tiles": [list(t) for t in tiles_final], "chunks_touched_rank0": counts,
                   "stage_overlap": "layers + erosion (S2, S3) run on a high-priority side stream while the caves (S4, stage-1 inputs only) run on the "
                                    "main stream; they join before S5. stages.S3.ms / S4.ms are elapsed times of overlapping stages and do not add up to "
                                    "ms_per_step; the kernel times under `kernels` and the S4 roofline are measured with both streams sharing the SMs. "
                                    "S6 (mmgen_set_fill_overlap %d): %s" % (fill_overlap, "the terrain / rock / lush passes of fill batch b + 1 run on a second stream "
                                    "while the placement scan + decorators of batch b run on the main stream (different chunks' volumes); the S6 entries of "
                                    "`kernels` give ms_per_step from one extra untimed step with the overlap off (each kernel by itself) and, as "
                                    "elapsed_ms_in_timed_steps, what their event pairs span in the timed steps, where they share the SMs with the other "
                                    "stream's kernels or queue behind them" if fill_overlap else "fill passes in sequence"),
                   "l2": "working set per step (>= %.1f GB written) far exceeds the 126 MB L2; no flush needed" % (n_target * 98304 / 1e9)},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "ms_per_step": 1e3 * max(e2e_wall, e2e_dev) / args.steps, "host_checksum": host_sum,
                "d2h_gbs": d2h / 1e9 / (max(e2e_wall, e2e_dev) / args.steps), "host_memory_rank0": numa},
        "e2e_encoded": {"value": enc_value, "unit": UNIT, "d2h_bytes_per_step": int(enc_total), "compression": d2h / max(enc_total, 1),
                        "ms_per_step": 1e3 * max(enc_wall, enc_dev) / args.steps,
                        "note": "block volumes run-length coded on the device (wire format MMCH1, include/mmgen.h) and delivered as payload + index; "
                                "decoded on the host by mmgen_decode_chunk"},
        "gpu_launches": int(sum_over_ranks(launches)), "rank_ms": [round(t, 2) for t in rank_ms], "balance_passes": balance_log,
        "rank_wall_ms": {"step": rank_wall[0], "bookkeeping_between_steps": rank_wall[1], "predictor_incl_wait_for_slowest_rank": rank_predict},
        "roofline": roof, "kernels": kernels, "stages": stages, "clocks": clocks, "block_checksums": [("%016x" % c) for c in checks], "world_hash": "%016x" % (sum(sharding.gather_u64(hash_sum)) & 0xFFFFFFFFFFFFFFFF),
    }
    if rank == 0 and not args.no_cpu and world_size == 1:
        nthreads = os.cpu_count() or 1
        rates, detail = cpu_stage_rates(nthreads, cave_chunks=args.cpu_cave_chunks)
        out["cpu_baseline"] = {"value": cpu_whole_job_rate(rates, tiling.stage_chunk_counts(0, 0, S, S)), "unit": UNIT, "cores": nthreads, "kind": "port",
                               "sample": "oracle port (C++ -O2, no fast-math) on %d host threads: S1 on 676, S2 on 576 chunks, S3 on 1 zone, S4+S5 on %d, "
                                         "S6 on %d chunks of the C2 window; whole-job rate extrapolated with this workload's per-stage chunk counts. "
                                         "A thin extrapolation, and it FLATTERS the CPU: S6 is timed with each chunk's own placements only, not the "
                                         "gathered 49-list scan (up to 6144 placements per voxel) that dominates the reference's fill"
                                         % (nthreads, detail["S4"]["units"], detail["S6"]["units"]),
                               "stage_rates": {k: round(v, 2) for k, v in rates.items()}}
    elif rank == 0:
        out["cpu_baseline"] = None
    if rank == 0:
        sys.stdout.flush()
        os.dup2(real_stdout, 1)          # the JSON line is the only thing this arm writes to stdout
        print(json.dumps(out), flush=True)
    world.close()
    if world_size > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference", "reference-cpu"])
    ap.add_argument("--world", type=int, default=256, help="side of the target region in chunks")
    ap.add_argument("--balance-rounds", type=int, default=0, help="extra feedback passes (untimed full generations) that move the tile cuts (N > 1)")
    ap.add_argument("--equal-tiles", action="store_true", help="equal tiles instead of the stage-1 cost predictor (N > 1)")
    ap.add_argument("--halo", default="auto", choices=["auto", "recompute", "exchange"],
                    help="N > 1: recompute the 3-chunk placement ring per tile, exchange it over NCCL, or measure both and keep the faster (auto)")
    ap.add_argument("--ref-zones", type=int, default=3, help="the reference CUDA arm generates ZxZ erosion zones (+ apron) per step")
    ap.add_argument("--cpu-cave-chunks", type=int, default=0, help="bound the CPU baseline's S4 sample (0 = the whole zone)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--fill-overlap", type=int, default=None, help="mmgen_set_fill_overlap mode (default: the library's)")
    ap.add_argument("--config", default="c5", choices=["c5", "c3", "c4"],
                    help="c5 (default): the 256x256-chunk world the metric is quoted on; c3: 64x64 streaming region at the reference's tick pattern; "
                         "c4: cave + fill stress on 32x32 chunks, S4 and S6 timed in isolation (BASELINE.json configs 3 and 4; one GPU)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    if args.impl != "b200":
        if rank != 0:
            return
        if args.impl == "reference":
            run_reference_cuda(args)
        else:
            run_reference_cpu(args)
        return
    if args.config != "c5":
        if rank != 0:
            return
        (run_c3 if args.config == "c3" else run_c4)(args)
        return
    run_own(args)


if __name__ == "__main__":
    main()
