# round 2, call 3L (GPU box): BASELINE config 3 per-kernel times with the current library; stream tests
OUT=gpurun_out/r3l; mkdir -p $OUT
timeout 600 python bench.py --config c3 --steps 2 --warmup 1 > $OUT/c3_cur2.json 2> $OUT/c3_cur2.err; echo "rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r3l/c3_cur2.json').read().strip().splitlines()[-1])
for n,p in d['profiles'].items(): print(n, round(p['wall_ms'],1), round(p['device_ms'],1), p['ticks'], round(p['chunks_per_s']))
PY
timeout 900 python -m pytest tests/test_stream.py tests/test_gpu_parity.py tests/test_mesh.py -m gpu -q -x 2>&1 | tail -2
