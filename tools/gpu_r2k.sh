# round 2, call K (GPU box): suite after the scheduler pin / list ring / GAS inputs, c3 and a short c5
OUT=gpurun_out/r2k; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q --durations=5 > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -8 $OUT/pytest.log
timeout 600 python bench.py --config c3 --steps 2 --warmup 1 > $OUT/bench_c3.json 2> $OUT/bench_c3.err; echo "c3 rc=$?"; tail -3 $OUT/bench_c3.err
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; tail -3 $OUT/bench.err
python - <<P
import json
j = json.loads(open("$OUT/bench_c3.json").read().strip().splitlines()[-1]); print("c3", round(j["value"]), {k: (v["ticks"], round(v["wall_ms"], 1), round(v["device_ms"], 1), round(v["chunks_per_s"])) for k, v in j["profiles"].items()})
j = json.loads(open("$OUT/bench.json").read().strip().splitlines()[-1])
print(round(j["value"]), j["world_hash"], round(j["e2e"]["value"]), round(j["e2e_encoded"]["value"]), j["ms_per_step"])
P
