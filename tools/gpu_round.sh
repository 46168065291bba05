mkdir -p gpurun_out/r1a
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r1a/smi.txt; nproc >> gpurun_out/r1a/smi.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r1a/pytest.log 2>&1; echo "pytest rc=$?" 
tail -5 gpurun_out/r1a/pytest.log
timeout 300 python bench.py --world 96 --steps 2 --warmup 1 --no-cpu > gpurun_out/r1a/bench96.log 2>&1; echo "bench96 rc=$?"; tail -c 3000 gpurun_out/r1a/bench96.log
timeout 900 python bench.py --steps 2 --warmup 1 > gpurun_out/r1a/bench256.log 2>&1; echo "bench256 rc=$?"; tail -c 4000 gpurun_out/r1a/bench256.log
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r1a/benchref.log 2>&1; echo "benchref rc=$?"; tail -c 2000 gpurun_out/r1a/benchref.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r1a/launches.csv python bench.py --world 24 --steps 1 --warmup 1 --no-cpu > gpurun_out/r1a/ncu_bench.log 2>&1; echo "ncu rc=$?"
