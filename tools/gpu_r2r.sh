# round 2, call R (GPU box): ncu --set full of k_caves / k_fill_rock with the packed-fp32 build (source-level stall samples)
OUT=gpurun_out/r2r; mkdir -p $OUT
for K in k_caves:2 k_fill_rock:9; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:${K%%:*} -s ${K##*:} -c 1 -f -o $OUT/${K%%:*} python tools/profile_driver.py 128 1 > $OUT/ncu_${K%%:*}.log 2>&1
done
ls -la $OUT
