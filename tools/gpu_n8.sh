# usage: bash tools/gpu_n8.sh <N> <tag>   (GPU box with N GPUs, via gpurun --gpus N): the driver's launch line for the scaling bench
N=${1:-8}; TAG=${2:-n$N}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err; echo "rc=$?"
tail -c 400 $OUT/bench_n$N.err
python - $OUT/bench_n$N.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['world_hash'], 'e2e', d['e2e']['value'], 'enc', d['e2e_encoded']['value'], 'rank_ms', d.get('rank_ms'), d['config'].get('tiling'), d['config'].get('halo_variants_ms_per_step'))
PY
