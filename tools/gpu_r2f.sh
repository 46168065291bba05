# round 2, call F (GPU box): suite (zone-order test, reference pins installed), bench after shape culling
OUT=gpurun_out/r2f; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q --durations=5 > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -8 $OUT/pytest.log
cp gpurun_out/parity_tour.json $OUT/ 2>/dev/null
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; tail -3 $OUT/bench.err
python - <<P
import json
j = json.loads(open("$OUT/bench.json").read().strip().splitlines()[-1])
print(round(j["value"]), j["world_hash"], round(j["e2e"]["value"]), {k: round(v["ms_per_step"], 1) for k, v in j["kernels"].items()})
P
