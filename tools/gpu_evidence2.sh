# round 2, second session (GPU box): the record at the final commit - GPU suite, default bench (with the CPU leg), reference arm, smoke
OUT=gpurun_out/evidence2; mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $OUT/pytest.log
timeout 600 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; tail -c 300 $OUT/bench.err
timeout 400 python bench.py --impl reference > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "ref rc=$?"
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $OUT/smoke.log
python - <<'PY'
import json
for f in ('bench','bench_ref'):
    try:
        d=json.loads(open('gpurun_out/evidence2/%s.json'%f).read().strip().splitlines()[-1])
        print(f, d.get('value'), d.get('ms_per_step'), (d.get('e2e') or {}).get('value'), (d.get('cpu_baseline') or {}).get('value'), d.get('world_hash'))
    except Exception as e: print(f, 'ERR', e)
PY
