mkdir -p gpurun_out/r2i
O=gpurun_out/r2i
for n in 2048 1024 512; do
AM3D_OPTIONS=pgs_clusters=0,pgs_persistent=0 timeout 300 python bench.py --size $n --no-also --no-cpu-baseline --steps 20 > $O/b${n}_phase.json 2> $O/err2
for v in phase; do python - <<PY
import json
d=json.loads([l for l in open("$O/b${n}_${v}.json") if l.startswith("{")][-1])
print("$n $v", round(d["ms_per_step"],3), "ms/step  sweep", round(d["roofline"]["sweep_ms_per_step"],3), d["roofline"]["kernel"], "e2e ms", round(d["e2e"]["ms_per_step"],3), "launches", d["gpu_launches"])
PY
done; done
