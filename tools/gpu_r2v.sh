# round 2, call V (GPU box): k_caves with the a-priori jitter table (noise and Worley in one phase), min blocks 4 / 5, census of the table hit rate, parity
OUT=gpurun_out/r2v; mkdir -p $OUT
for v in c2m5 f4 f5; do MMGEN_LIB=$PWD/mega-minecraft_b200/libmmgen_$v.so python tools/variant_time.py 128 k_caves 2>&1 | tail -1; done | tee $OUT/variants.txt
MMGEN_LIB=$PWD/mega-minecraft_b200/libmmgen_stats.so timeout 600 python tools/feature_census.py 128 > $OUT/census.txt 2>&1; tail -5 $OUT/census.txt
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_region_hashes.py tests/test_reference_tour.py -m gpu -q -x 2>&1 | tail -2
