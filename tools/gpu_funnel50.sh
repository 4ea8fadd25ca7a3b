mkdir -p gpurun_out/r2
timeout 900 python tools/funnel_run.py 50 120 0.6 gpurun_out/r2/funnel50.jsonl 2>&1 | tail -16
