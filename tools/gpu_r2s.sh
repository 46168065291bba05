# round 2, call S (GPU box): table staging by one bulk copy per CTA (cp.async.bulk + mbarrier) on top of the packed-fp32 noise
OUT=gpurun_out/r2s; mkdir -p $OUT
for v in base q9r8; do MMGEN_LIB=$PWD/mega-minecraft_b200/libmmgen_$v.so python tools/variant_time.py 128 2>&1 | tail -1; done | tee $OUT/variants.txt
python tools/variant_time.py 128 2>&1 | tail -1 | tee -a $OUT/variants.txt
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_region_hashes.py tests/test_reference_tour.py -m gpu -q -x 2>&1 | tail -2
