# round 2, call U (GPU box): two-column k_caves, min blocks 5 / 6; ncu source-level capture of the 5-block build
OUT=gpurun_out/r2u; mkdir -p $OUT
for v in c2m5 c2m6; do MMGEN_LIB=$PWD/mega-minecraft_b200/libmmgen_$v.so python tools/variant_time.py 128 k_caves 2>&1 | tail -1; done | tee $OUT/variants.txt
MMGEN_LIB=$PWD/mega-minecraft_b200/libmmgen_c2m5.so timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_caves -s 2 -c 1 -f -o $OUT/k_caves python tools/profile_driver.py 128 1 > $OUT/ncu_k_caves.log 2>&1
ls $OUT
