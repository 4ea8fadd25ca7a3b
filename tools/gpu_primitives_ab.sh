mkdir -p gpurun_out/r2h
O=gpurun_out/r2h
timeout 900 python -m pytest tests/test_gpu_primitives.py tests/test_gpu_parity.py tests/test_gpu_merging.py -q 2>&1 | tail -6 | tee $O/pytest_gpu.txt
timeout 400 python bench.py --no-also --no-cpu-baseline --steps 20 > $O/batch_own.json 2> $O/err1
AM3D_OPTIONS=own_primitives=0 timeout 400 python bench.py --no-also --no-cpu-baseline --steps 20 > $O/batch_cub.json 2> $O/err2
timeout 300 python bench.py --scaling weak --no-also --no-cpu-baseline --steps 20 > $O/b512_own.json 2> $O/err3
AM3D_OPTIONS=own_primitives=0 timeout 300 python bench.py --scaling weak --no-also --no-cpu-baseline --steps 20 > $O/b512_cub.json 2> $O/err4
for f in batch_own batch_cub b512_own b512_cub; do python - <<PY
import json
d=json.loads([l for l in open("$O/${f}.json") if l.startswith("{")][-1])
print("$f", round(d["ms_per_step"],3), "ms/step  sweep", round(d["roofline"]["sweep_ms_per_step"],3), "e2e ms", round(d["e2e"]["ms_per_step"],3), "launches", d["gpu_launches"], "detect", round(d["phase_ms_last_step"]["detection"],3))
PY
done
