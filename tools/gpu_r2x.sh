# round 2, call X (GPU box): k_caves with C columns per CTA and a survivor-list loop: (columns, threads, min blocks) variants, parity
OUT=gpurun_out/r2x; mkdir -p $OUT
for v in g8t256m5 h8t256m5 h16t256m4 h16t256m5 h16t320m4 h16t384m3 h16t512m2 h32t256m4 h32t512m2; do MMGEN_LIB=$PWD/mega-minecraft_b200/libmmgen_$v.so python tools/variant_time.py 128 k_caves 2>&1 | tail -1; done | tee $OUT/variants2.txt
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_region_hashes.py tests/test_reference_tour.py -m gpu -q -x 2>&1 | tail -2
