# usage: bash tools/gpu_final.sh <tag>   (on the GPU box, via gpurun): everything the round report cites, in one call
TAG=${1:-final}
OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $OUT/pytest.log
timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; tail -c 300 $OUT/bench.err
timeout 900 python bench.py --impl reference > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "ref rc=$?"
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $OUT/smoke.log
timeout 300 python tools/stream_bench.py > $OUT/stream.jsonl 2>&1; echo "stream rc=$?"
timeout 300 python tools/mesh_bench.py > $OUT/mesh.json 2>&1; echo "mesh rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches.csv python bench.py --world 48 --steps 1 --warmup 1 --no-cpu > $OUT/ncu_bench.log 2>&1; echo "ncu rc=$?"
# one launch from the middle of a 128x128-chunk world (launch 2 of 5 for the cave kernels, 9 of 32 for the fill kernels)
for K in k_caves:2 k_fill_features:9 k_fill_rock:9 k_fill_terrain:9 k_erode_sweep:300; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:${K%%:*} -s ${K##*:} -c 1 -f -o $OUT/${K%%:*} python tools/profile_driver.py 128 1 > $OUT/ncu_${K%%:*}.log 2>&1
done
ls $OUT
