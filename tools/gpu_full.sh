# usage: bash tools/gpu_full.sh <tag>   (on the GPU box, via gpurun): tests, default bench, reference arm, launch list
TAG=${1:-full}
OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $OUT/pytest.log
timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; tail -c 600 $OUT/bench.err
timeout 900 python bench.py --impl reference > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "ref rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches.csv python bench.py --world 48 --steps 1 --warmup 1 --no-cpu > $OUT/ncu_bench.log 2>&1; echo "ncu rc=$?"
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $OUT/smoke.log
