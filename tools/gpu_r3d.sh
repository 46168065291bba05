# round 2, call 3D (GPU box): k_fill_features thread-count / occupancy variants
OUT=gpurun_out/r3d; mkdir -p $OUT
for v in f384m2 f448m2 f512m2 f576m2 f768m1 f896m1 f1024m1; do MMGEN_LIB=$PWD/mega-minecraft_b200/libmmgen_$v.so python tools/variant_time.py 128 k_fill_features 2>&1 | tail -1; done | tee $OUT/variants3.txt
