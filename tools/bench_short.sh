# usage: tools/bench_short.sh [bench args...]: one bench line, condensed
python bench.py "$@" --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(d['config']['workload'], 'value %.3g'%d['value'], 'e2e %.3g'%d['e2e']['value'], 'ms/step %.2f'%d['ms_per_step'], {k:round(v,2) for k,v in d['phase_ms_last_step'].items()}, 'pgs launch ms %.2f'%d['roofline']['avg_launch_ms'], 'frac %.3f'%d['roofline']['frac'], 'launches', d['gpu_launches'], 'wall %.2f'%d['wall_ms_per_step'], 'coll', d['collections_last_step'], 'series', d['step_ms_series'])
"
