# round 2, call I (2-GPU box): bench at N=2 with the predictor cuts, both halo variants measured (auto)
OUT=gpurun_out/r2i; mkdir -p $OUT
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 3 --warmup 3 > $OUT/bench_n2.json 2> $OUT/bench_n2.err; echo "n2 rc=$?"; tail -5 $OUT/bench_n2.err
python - <<P
import json
j = json.loads(open("$OUT/bench_n2.json").read().strip().splitlines()[-1])
print(round(j["value"]), j["world_hash"], round(j["e2e"]["value"]), j["rank_ms"], j["config"]["tiling"], j["config"]["halo_variants_ms_per_step"], j["config"]["tiles"])
P
