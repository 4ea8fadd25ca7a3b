mkdir -p gpurun_out/r2c
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 | tee gpurun_out/r2c/pytest_gpu.txt
