# round 2, call W (GPU box): ncu source-level capture of k_caves (fused phases, 5 blocks)
OUT=gpurun_out/r2w; mkdir -p $OUT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_caves -s 2 -c 1 -f -o $OUT/k_caves python tools/profile_driver.py 128 1 > $OUT/ncu_k_caves.log 2>&1
ls $OUT
