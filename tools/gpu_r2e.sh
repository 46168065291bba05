# round 2, call E (GPU box): suite with the erosion deviation test, bench lines c5 (short) / c4 / c3
OUT=gpurun_out/r2e; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q --durations=5 > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -12 $OUT/pytest.log
cp gpurun_out/parity_tour.json $OUT/ 2>/dev/null
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; tail -3 $OUT/bench.err
timeout 600 python bench.py --config c4 > $OUT/bench_c4.json 2> $OUT/bench_c4.err; echo "c4 rc=$?"; tail -3 $OUT/bench_c4.err
timeout 600 python bench.py --config c3 --steps 2 --warmup 1 > $OUT/bench_c3.json 2> $OUT/bench_c3.err; echo "c3 rc=$?"; tail -3 $OUT/bench_c3.err
python - <<P
import json
j = json.loads(open("$OUT/bench.json").read().strip().splitlines()[-1])
print(round(j["value"]), j["world_hash"], round(j["e2e"]["value"]))
print(json.dumps(j["stages"])[:3000])
j = json.loads(open("$OUT/bench_c4.json").read().strip().splitlines()[-1]); print("c4", round(j["value"]), json.dumps(j["stages"]))
j = json.loads(open("$OUT/bench_c3.json").read().strip().splitlines()[-1]); print("c3", round(j["value"]), {k: (v["ticks"], round(v["wall_ms"], 1), round(v["chunks_per_s"])) for k, v in j["profiles"].items()})
P
