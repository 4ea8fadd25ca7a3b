mkdir -p gpurun_out/r2e
O=gpurun_out/r2e
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 | tee $O/pytest_gpu.txt
timeout 400 python bench.py --no-also --no-cpu-baseline --steps 20 > $O/batch.json 2> $O/err3
timeout 300 python bench.py --workload stack --merging 0 --no-cpu-baseline --steps 20 > $O/stack.json 2> $O/err7
for f in batch stack; do python - <<PY
import json
d=json.loads([l for l in open("$O/${f}.json") if l.startswith("{")][-1])
print("$f", round(d["ms_per_step"],3), "ms/step  e2e ms", round(d["e2e"]["ms_per_step"],3), "value", round(d["value"]), "e2e", round(d["e2e"]["value"]))
PY
done
