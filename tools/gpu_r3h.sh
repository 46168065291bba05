# round 2, call 3H (GPU box): chunks per fill batch 1024 / 2048 / 4096 (512 before); codec + parity with 2048
OUT=gpurun_out/r3h; mkdir -p $OUT
for v in fb1024 fb2048 fb4096; do MMGEN_LIB=$PWD/mega-minecraft_b200/libmmgen_$v.so python tools/variant_time.py 128 2>&1 | tail -1; done | tee $OUT/variants2.txt
MMGEN_LIB=$PWD/mega-minecraft_b200/libmmgen_fb2048.so timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_codec.py tests/test_region_hashes.py -m gpu -q -x 2>&1 | tail -2
