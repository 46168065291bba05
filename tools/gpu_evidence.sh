# round 2, (GPU box): the round's evidence run - GPU suite, default bench (with the CPU leg), reference arm, BASELINE configs 3 and 4, smoke, mesh
OUT=gpurun_out/evidence; mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $OUT/pytest.log
timeout 1200 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; tail -c 300 $OUT/bench.err
timeout 900 python bench.py --impl reference > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "ref rc=$?"
timeout 600 python bench.py --config c3 > $OUT/bench_c3.json 2> $OUT/bench_c3.err; echo "c3 rc=$?"
timeout 600 python bench.py --config c4 > $OUT/bench_c4.json 2> $OUT/bench_c4.err; echo "c4 rc=$?"
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $OUT/smoke.log
timeout 300 python tools/mesh_bench.py > $OUT/mesh.json 2>&1; echo "mesh rc=$?"
python - <<'PY'
import json
for f in ('bench','bench_ref','bench_c3','bench_c4'):
    try:
        d=json.loads(open('gpurun_out/evidence/%s.json'%f).read().strip().splitlines()[-1])
        print(f, d.get('value'), d.get('ms_per_step'), (d.get('e2e') or {}).get('value'), (d.get('cpu_baseline') or {}).get('value'))
    except Exception as e: print(f, 'ERR', e)
PY
