mkdir -p gpurun_out/r2
timeout 600 python tools/funnel_run.py 100 70 0.6 gpurun_out/r2/funnel100.jsonl 55e6 2>&1 | tail -12 | cut -c1-700
nvidia-smi --query-gpu=memory.used,memory.total --format=csv | tail -1
