for v in cur cb8k cb16k; do L=$PWD/mega-minecraft_b200/libmmgen.so; [ $v != cur ] && L=$PWD/mega-minecraft_b200/libmmgen_$v.so
MMGEN_LIB=$L python - <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import mmgen_loader
mm = mmgen_loader.load()
gen = mm.ChunkGen(0)
w = gen.region_world(0, 0, 256, 256)
best = 1e9
for rep in range(3):
    w.reset(); w.generate(mm.STAGE_ALL); w.sync(); best = min(best, w.total_ms())
print(os.path.basename(os.environ['MMGEN_LIB']), round(best, 2), '%016x' % w.chunk_hash_sum())
PY
done
