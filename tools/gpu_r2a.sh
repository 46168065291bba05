# round 2, call A (GPU box): full GPU suite incl. the reference tour, reference-pinned region hashes, default vs split-features bench
OUT=gpurun_out/r2a; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q --durations=15 > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -25 $OUT/pytest.log
cp gpurun_out/parity_tour.json $OUT/ 2>/dev/null
timeout 600 python tools/region_hashes.py --write > $OUT/region_hashes.log 2>&1; echo "hashes rc=$?"; cat $OUT/region_hashes.log | tail -5
cp tests/golden/region_hashes.json $OUT/ 2>/dev/null
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu > $OUT/bench_default.json 2> $OUT/bench_default.err; echo "bench rc=$?"
MMGEN_LIB=$PWD/mega-minecraft_b200/libmmgen_split.so timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu > $OUT/bench_split.json 2> $OUT/bench_split.err; echo "split rc=$?"
python - <<P
import json
for n in ("default", "split"):
    try:
        j = json.loads(open("$OUT/bench_%s.json" % n).read().strip().splitlines()[-1])
        print(n, round(j["value"]), j["world_hash"], round(j["e2e"]["value"]), {k: round(v["ms_per_step"], 1) for k, v in j["kernels"].items()})
    except Exception as e:
        print(n, "failed", e)
P
