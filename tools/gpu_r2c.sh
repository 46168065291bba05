# round 2, call C (GPU box): reference pins with flip diagnostics, full GPU suite, ncu capture of the new k_fill_features
OUT=gpurun_out/r2c; mkdir -p $OUT
timeout 900 python tools/region_hashes.py --write > $OUT/region_hashes.log 2>&1; echo "hashes rc=$?"; grep -c "'x'" $OUT/region_hashes.log
timeout 1500 python -m pytest tests -m gpu -q --durations=5 > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -12 $OUT/pytest.log
cp gpurun_out/parity_tour.json $OUT/ 2>/dev/null
for K in k_fill_features:9; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:${K%%:*} -s ${K##*:} -c 1 -f -o $OUT/${K%%:*} python tools/profile_driver.py 128 1 > $OUT/ncu_${K%%:*}.log 2>&1; echo "ncu rc=$?"
done
ls -la $OUT
