# round 2, call G (GPU box): k_fill_features with the shape-culling variants (MMG_SHAPE 0..4) + ncu of variant 1
OUT=gpurun_out/r2g; mkdir -p $OUT
for v in 0 1 2 3 4; do MMGEN_LIB=$PWD/mega-minecraft_b200/libmmgen_shape$v.so python tools/variant_time.py 128 k_fill_features k_fill_terrain k_fill_rock 2>&1 | tail -1; done | tee $OUT/variants.txt
MMGEN_LIB=$PWD/mega-minecraft_b200/libmmgen_shape1.so timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fill_features -s 9 -c 1 -f -o $OUT/k_fill_features_shape1 python tools/profile_driver.py 128 1 > $OUT/ncu1.log 2>&1; echo "ncu rc=$?"
