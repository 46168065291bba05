# round 2, call 3C (GPU box): k_fill_rock occupancy variants with the packed noise; parity of the default build
OUT=gpurun_out/r3c; mkdir -p $OUT
python tools/variant_time.py 128 k_fill_terrain k_fill_rock 2>&1 | tail -1 | tee $OUT/variants.txt
for v in r9 r10 r12; do MMGEN_LIB=$PWD/mega-minecraft_b200/libmmgen_$v.so python tools/variant_time.py 128 k_fill_terrain k_fill_rock 2>&1 | tail -1; done | tee -a $OUT/variants.txt
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_region_hashes.py tests/test_reference_tour.py -m gpu -q -x 2>&1 | tail -2
