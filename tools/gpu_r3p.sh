# round 2, call 3P (GPU box): layers + erosion on a side stream while the caves run - serial against overlapped, parity
OUT=gpurun_out/r3p; mkdir -p $OUT
python - <<'PY' | tee gpurun_out/r3p/overlap.txt
import sys, os
sys.path.insert(0, os.getcwd())
import mmgen_loader
mm = mmgen_loader.load()
gen = mm.ChunkGen(0)
for S in (128, 256):
    w = gen.region_world(0, 0, S, S)
    for serial in (1, 0, 1, 0):
        gen.L.mmgen_set_serial_stages(serial)
        w.reset(); w.generate(mm.STAGE_ALL); w.sync()
        w.reset(); w.generate(mm.STAGE_ALL); w.sync()
        print(S, 'serial' if serial else 'overlap', round(w.total_ms(), 2), [round(float(v), 1) for v in w.stage_ms()], '%016x' % w.chunk_hash_sum())
    w.close()
gen.L.mmgen_set_serial_stages(0)
PY
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_region_hashes.py tests/test_exchange.py tests/test_reference_tour.py -m gpu -q -x 2>&1 | tail -2
