mkdir -p gpurun_out/r2g
O=gpurun_out/r2g
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 | tee $O/pytest_gpu.txt
timeout 400 python bench.py --no-also --no-cpu-baseline --steps 20 > $O/batch_tail.json 2> $O/err3
AM3D_OPTIONS=pgs_tail_fusion=0 timeout 400 python bench.py --no-also --no-cpu-baseline --steps 20 > $O/batch_notail.json 2> $O/err4
for f in batch_tail batch_notail; do python - <<PY
import json
d=json.loads([l for l in open("$O/${f}.json") if l.startswith("{")][-1])
print("$f", round(d["ms_per_step"],3), "ms/step  sweep", round(d["roofline"]["sweep_ms_per_step"],3), "e2e ms", round(d["e2e"]["ms_per_step"],3), "launches", d["gpu_launches"], d["phase_ms_last_step"])
PY
done
