# round 2, call 3G (GPU box): several erosion tiles per CTA
OUT=gpurun_out/r3g; mkdir -p $OUT
python tools/variant_time.py 128 k_erode_sweep 2>&1 | tail -1 | tee $OUT/variants.txt
for v in e2 e3 e4 e6; do MMGEN_LIB=$PWD/mega-minecraft_b200/libmmgen_$v.so python tools/variant_time.py 128 k_erode_sweep 2>&1 | tail -1; done | tee -a $OUT/variants.txt
MMGEN_LIB=$PWD/mega-minecraft_b200/libmmgen_e4.so timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_region_hashes.py -m gpu -q -x 2>&1 | tail -2
