mkdir -p gpurun_out/r2f
O=gpurun_out/r2f
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 | tee $O/pytest_gpu.txt
timeout 300 python tools/funnel_steps.py 20 44 2>&1 | tail -6 | tee $O/funnel_steps.txt
