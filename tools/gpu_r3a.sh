# round 2, call 3A (GPU box): k_fill_terrain per chunk row (16 columns, 256 threads) against the per-column kernel; parity
OUT=gpurun_out/r3a; mkdir -p $OUT
for v in tcol trow3 trow4; do MMGEN_LIB=$PWD/mega-minecraft_b200/libmmgen_$v.so python tools/variant_time.py 128 k_fill_terrain k_fill_rock k_fill_features 2>&1 | tail -1; done | tee $OUT/variants.txt
python tools/variant_time.py 128 k_fill_terrain k_fill_rock k_fill_features 2>&1 | tail -1 | tee -a $OUT/variants.txt
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_region_hashes.py tests/test_reference_tour.py -m gpu -q -x 2>&1 | tail -2
