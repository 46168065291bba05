# round 2, call M (GPU box): k_fill_terrain with 1 / 2 / 4 columns per CTA: time + golden parity of the multi-column builds
OUT=gpurun_out/r2m; mkdir -p $OUT
for v in 1 2 4; do MMGEN_LIB=$PWD/mega-minecraft_b200/libmmgen_tc$v.so python tools/variant_time.py 128 k_fill_terrain k_fill_rock k_fill_features 2>&1 | tail -1; done | tee $OUT/variants.txt
for v in 2 4; do MMGEN_LIB=$PWD/mega-minecraft_b200/libmmgen_tc$v.so timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_region_hashes.py -m gpu -q -x 2>&1 | tail -2; done
