# round 2, call P (GPU box): packed fp32 (FFMA2) noise pairs in k_caves / k_fill_rock / k_cave_biomes / k_fill_lush - base vs packed at
# three occupancy targets (caves / rock min blocks), parity
OUT=gpurun_out/r2p; mkdir -p $OUT
for v in base p9 q10r8 q9r7 q8r6; do MMGEN_LIB=$PWD/mega-minecraft_b200/libmmgen_$v.so python tools/variant_time.py 128 2>&1 | tail -1; done | tee $OUT/variants2.txt
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_region_hashes.py tests/test_reference_tour.py -m gpu -q -x 2>&1 | tail -2
