# round 2, call 3M (GPU box): BASELINE config 3 on ONE box - round-2 start library, current library, current with a 512-chunk scratch floor
OUT=gpurun_out/r3m; mkdir -p $OUT
for rep in 1 2; do for v in base cur; do
  L=$PWD/mega-minecraft_b200/libmmgen.so; [ $v = base ] && L=$PWD/mega-minecraft_b200/libmmgen_base.so; [ $v = sf512 ] && L=$PWD/mega-minecraft_b200/libmmgen_sf512.so
  MMGEN_LIB=$L timeout 600 python bench.py --config c3 --steps 3 --warmup 1 > $OUT/c3_$v.json 2> $OUT/c3_$v.err
  python - $v <<'PY'
import json,sys
d=json.loads(open('gpurun_out/r3m/c3_%s.json'%sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], ' '.join('%s wall %.1f dev %.1f'%(n,p['wall_ms'],p['device_ms']) for n,p in d['profiles'].items()))
PY
done; done
timeout 900 python -m pytest tests/test_stream.py tests/test_mesh.py tests/test_exchange.py -m gpu -q -x 2>&1 | tail -2
