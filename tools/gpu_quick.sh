# usage: bash tools/gpu_quick.sh <tag> [census]   (on the GPU box, via gpurun): GPU tests + a short bench (+ the rasteriser census)
TAG=${1:-quick}
OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $OUT/pytest.log
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"
python - <<P
import json
j = json.loads(open("$OUT/bench.json").read().strip().splitlines()[-1])
print(j["value"], j["world_hash"], j.get("e2e", {}).get("value"))
print({k: round(v["ms_per_step"], 1) for k, v in j["kernels"].items()})
P
if [ "$2" = census ]; then MMGEN_LIB=$PWD/mega-minecraft_b200/libmmgen_stats.so python tools/feature_census.py 128 2>&1 | tail -30; fi
