# round 2, call J (GPU box): suite (codec), default bench with the encoded e2e leg
OUT=gpurun_out/r2j; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q --durations=5 > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -8 $OUT/pytest.log
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; tail -5 $OUT/bench.err
python - <<P
import json
j = json.loads(open("$OUT/bench.json").read().strip().splitlines()[-1])
print(round(j["value"]), j["world_hash"], round(j["e2e"]["value"]), j["e2e_encoded"], j["ms_per_step"], {k: round(v["ms_per_step"], 1) for k, v in j["kernels"].items()})
P
