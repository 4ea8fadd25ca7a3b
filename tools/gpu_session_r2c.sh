# fast-row A/B: parity tests, then funnel / batch / batch512 / stack with and without the branch-free rows
mkdir -p gpurun_out/r2c
O=gpurun_out/r2c
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 | tee $O/pytest_gpu.txt
timeout 300 python bench.py --workload funnel --steps 3 --warmup 3 --no-cpu-baseline > $O/funnel_fast.json 2> $O/err1
AM3D_OPTIONS=pgs_fast_rows=0 timeout 300 python bench.py --workload funnel --steps 3 --warmup 3 --no-cpu-baseline > $O/funnel_plain.json 2> $O/err2
timeout 400 python bench.py --no-also --no-cpu-baseline --steps 20 > $O/batch_fast.json 2> $O/err3
AM3D_OPTIONS=pgs_fast_rows=0 timeout 400 python bench.py --no-also --no-cpu-baseline --steps 20 > $O/batch_plain.json 2> $O/err4
timeout 300 python bench.py --scaling weak --no-also --no-cpu-baseline --steps 20 > $O/b512_fast.json 2> $O/err5
AM3D_OPTIONS=pgs_fast_rows=0 timeout 300 python bench.py --scaling weak --no-also --no-cpu-baseline --steps 20 > $O/b512_plain.json 2> $O/err6
timeout 300 python bench.py --workload stack --merging 0 --no-cpu-baseline --steps 20 > $O/stack_fast.json 2> $O/err7
AM3D_OPTIONS=pgs_fast_rows=0 timeout 300 python bench.py --workload stack --merging 0 --no-cpu-baseline --steps 20 > $O/stack_plain.json 2> $O/err8
for f in funnel batch b512 stack; do for v in fast plain; do python - <<PY
import json
d=json.loads([l for l in open("$O/${f}_${v}.json") if l.startswith("{")][-1])
print("$f $v", round(d["ms_per_step"],3), "ms/step  sweep", round(d["roofline"]["sweep_ms_per_step"],3), "value", round(d["value"]), "e2e", round(d["e2e"]["value"]))
PY
done; done
