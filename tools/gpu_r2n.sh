# round 2, call N (GPU box): base vs vector stores in k_fill_terrain vs noise tables read from global memory
OUT=gpurun_out/r2n; mkdir -p $OUT
for v in base vstore nglobal; do MMGEN_LIB=$PWD/mega-minecraft_b200/libmmgen_$v.so python tools/variant_time.py 128 2>&1 | tail -1; done | tee $OUT/variants.txt
MMGEN_LIB=$PWD/mega-minecraft_b200/libmmgen_nglobal.so timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_region_hashes.py -m gpu -q -x 2>&1 | tail -2
