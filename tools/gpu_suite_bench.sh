# round 2, (GPU box): full GPU suite + default bench (no CPU leg) after the packed-fp32 / column-group work
OUT=gpurun_out/suite; mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest.log
timeout 900 python bench.py --no-cpu > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; tail -c 300 $OUT/bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/suite/bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['world_hash'], d['e2e']['value'], d['e2e_encoded']['value'], {k:round(v['ms_per_step'],1) for k,v in d['kernels'].items()})
PY
