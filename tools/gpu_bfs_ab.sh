mkdir -p gpurun_out/r2i
timeout 600 python -m pytest tests/test_gpu_merging.py tests/test_gpu_scale.py tests/test_gpu_scenes.py tests/test_gpu_hooks.py -q 2>&1 | tail -4
python tools/batch_steps.py 512 168 > gpurun_out/r2i/steps512_bfs.txt 2>&1
python tools/batch_steps.py 512 168 scene_bfs=0 > gpurun_out/r2i/steps512_nobfs.txt 2>&1
python tools/batch_steps.py 4096 168 > gpurun_out/r2i/steps4096_bfs.txt 2>&1
python tools/batch_steps.py 4096 168 scene_bfs=0 > gpurun_out/r2i/steps4096_nobfs.txt 2>&1
for f in steps512_bfs steps512_nobfs steps4096_bfs steps4096_nobfs; do python - <<PY
import re
t=[float(re.search(r"total\s+([\d.]+)",l).group(1)) for l in open("gpurun_out/r2i/$f.txt") if l.startswith("step")]
u=[float(re.search(r"upd\s+([\d.]+)",l).group(1)) for l in open("gpurun_out/r2i/$f.txt") if l.startswith("step")]
print("$f", "mean total", round(sum(t[6:46])/40,3), "max upd", max(u), "mean upd", round(sum(u[6:46])/40,3))
PY
done
