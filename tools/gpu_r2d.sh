# round 2, call D (GPU box): suite after the rafflesia FMA fix + erosion order test, reference pins, census of k_fill_features
OUT=gpurun_out/r2d; mkdir -p $OUT
timeout 900 python tools/region_hashes.py --write > $OUT/region_hashes.log 2>&1; echo "hashes rc=$?"; grep -c "'x'" $OUT/region_hashes.log
timeout 1500 python -m pytest tests -m gpu -q --durations=5 > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -12 $OUT/pytest.log
cp gpurun_out/parity_tour.json $OUT/ 2>/dev/null
MMGEN_LIB=$PWD/mega-minecraft_b200/libmmgen_stats.so timeout 600 python tools/feature_census.py 128 > $OUT/census.txt 2>&1; echo "census rc=$?"; head -40 $OUT/census.txt
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; tail -3 $OUT/bench.err
python - <<P
import json
j = json.loads(open("$OUT/bench.json").read().strip().splitlines()[-1])
print(round(j["value"]), j["world_hash"], round(j["e2e"]["value"]), {k: round(v["ms_per_step"], 1) for k, v in j["kernels"].items()})
P
