# usage: bash tools/gpu_scale.sh <tag> <N> [bench args]   (on an N-GPU box via gpurun --gpus N)
TAG=${1:-scale}; N=${2:-2}; shift; shift
OUT=gpurun_out/$TAG; mkdir -p $OUT
if [ "$N" = "1" ]; then
  timeout 600 python bench.py --gpus 1 "$@" > $OUT/bench_n1.json 2> $OUT/bench_n1.err; echo "n1 rc=$?"
else
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N "$@" > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err; echo "n$N rc=$?"
fi
tail -c 400 $OUT/bench_n$N.err
