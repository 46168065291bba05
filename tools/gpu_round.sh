# usage: bash tools/gpu_round.sh <tag> <world-side> [bench args...]   (on the GPU box, via gpurun)
TAG=${1:-round}; SIDE=${2:-96}; shift; shift
OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest.log
timeout 900 python bench.py --world $SIDE "$@" > $OUT/bench.log 2>&1; echo "bench rc=$?"; tail -c 3500 $OUT/bench.log
