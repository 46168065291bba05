# round 2, call Y (GPU box): ncu source-level captures of k_caves (16 columns), k_fill_terrain, k_fill_features, k_fill_rock
OUT=gpurun_out/r2y; mkdir -p $OUT
for K in k_caves:2 k_fill_terrain:9 k_fill_features:9 k_fill_rock:9; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:${K%%:*} -s ${K##*:} -c 1 -f -o $OUT/${K%%:*} python tools/profile_driver.py 128 1 > $OUT/ncu_${K%%:*}.log 2>&1
done
ls $OUT
