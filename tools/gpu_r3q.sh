# round 2, call 3Q (GPU box): erosion CTA shape (rows of 32 threads per tile CTA) now that erosion shares the SMs with the cave kernel
OUT=gpurun_out/r3q; mkdir -p $OUT
python - <<'PY' | tee gpurun_out/r3q/erode_rows.txt
import sys, os, subprocess
for lib in ('libmmgen.so', 'libmmgen_er2.so', 'libmmgen_er4.so', 'libmmgen_er16.so'):
    code = '''
import sys, os
sys.path.insert(0, os.getcwd())
import mmgen_loader
mm = mmgen_loader.load()
gen = mm.ChunkGen(0)
w = gen.region_world(0, 0, 256, 256)
for serial in (1, 0):
    gen.set_serial_stages(serial)
    w.reset(); w.generate(mm.STAGE_ALL); w.sync()
    w.reset(); w.generate(mm.STAGE_ALL); w.sync()
    print("%s", "serial" if serial else "overlap", round(w.total_ms(), 2), [round(float(v), 1) for v in w.stage_ms()], "%%016x" %% w.chunk_hash_sum())
''' % lib
    env = dict(os.environ, MMGEN_LIB=os.path.join(os.getcwd(), 'mega-minecraft_b200', lib))
    print(subprocess.run([sys.executable, '-c', code], env=env, capture_output=True, text=True).stdout, end='')
PY
