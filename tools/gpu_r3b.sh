# round 2, call 3B (GPU box): k_fill_terrain_rows thread-count / occupancy variants
OUT=gpurun_out/r3b; mkdir -p $OUT
for v in t320m5 t384m4 t384m5 t512m4; do MMGEN_LIB=$PWD/mega-minecraft_b200/libmmgen_$v.so python tools/variant_time.py 128 k_fill_terrain 2>&1 | tail -1; done | tee $OUT/variants.txt
MMGEN_LIB=$PWD/mega-minecraft_b200/libmmgen_t384m5.so timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -2
