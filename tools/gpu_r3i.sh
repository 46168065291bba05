# round 2, call 3I (GPU box): k_fill_rock as two kernels (near / bulk voxels), occupancy targets r<near>b<bulk>
OUT=gpurun_out/r3i; mkdir -p $OUT
python tools/variant_time.py 128 k_fill_rock 2>&1 | tail -1 | tee $OUT/variants.txt
for v in rone r6b8 r8b9 r8b10 r8b12 r10b10; do MMGEN_LIB=$PWD/mega-minecraft_b200/libmmgen_$v.so python tools/variant_time.py 128 k_fill_rock 2>&1 | tail -1; done | tee -a $OUT/variants.txt
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_region_hashes.py -m gpu -q -x 2>&1 | tail -2
