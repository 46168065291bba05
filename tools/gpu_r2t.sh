# round 2, call T (GPU box): k_caves with two columns per 288-thread CTA and a survivor list (min blocks 3 / 4 / 5) against the committed kernel
OUT=gpurun_out/r2t; mkdir -p $OUT
for v in base c2m3 c2m4 c2m5; do MMGEN_LIB=$PWD/mega-minecraft_b200/libmmgen_$v.so python tools/variant_time.py 128 2>&1 | tail -1; done | tee $OUT/variants.txt
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_region_hashes.py tests/test_reference_tour.py -m gpu -q -x 2>&1 | tail -2
