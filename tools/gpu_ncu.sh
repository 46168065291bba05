# round 2, (GPU box): final ncu evidence - launch list of the bench command, --set full captures of the five hot kernels, census
OUT=gpurun_out/ncu; mkdir -p $OUT
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches.csv python bench.py --world 48 --steps 1 --warmup 1 --no-cpu > $OUT/ncu_bench.log 2>&1; echo "ncu launch list rc=$?"
# one launch from the middle of a 128x128-chunk world (launch 2 of 5 for the cave kernel, 4 of 8 for the fill kernels)
for K in k_caves:2 k_fill_features:4 k_fill_rock:4 k_fill_terrain:4 k_erode_sweep:300; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:${K%%:*} -s ${K##*:} -c 1 -f -o $OUT/${K%%:*} python tools/profile_driver.py 128 1 > $OUT/ncu_${K%%:*}.log 2>&1
done
MMGEN_LIB=$PWD/mega-minecraft_b200/libmmgen_stats.so timeout 600 python tools/feature_census.py 128 > $OUT/census.txt 2>&1; tail -4 $OUT/census.txt
ls $OUT
