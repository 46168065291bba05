# round 2, call H (GPU box): k_fill_features queue depth (MMG_QCAP) x shape culling (MMG_SHAPE)
OUT=gpurun_out/r2h; mkdir -p $OUT
for f in mega-minecraft_b200/libmmgen_s*.so; do MMGEN_LIB=$PWD/$f python tools/variant_time.py 128 k_fill_features 2>&1 | tail -1; done | tee $OUT/variants.txt
