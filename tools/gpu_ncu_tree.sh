mkdir -p gpurun_out/r2f
export AM3D_CUDA_PROFILER=1
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,launch__grid_size --clock-control none -k regex:"k_narrow|k_tree" --csv --log-file gpurun_out/r2f/launches_tree.csv python bench.py --workload funnel --steps 1 --warmup 5 --no-cpu-baseline > gpurun_out/r2f/ncu_tree.log 2>&1
grep -v "^==" gpurun_out/r2f/launches_tree.csv | cut -d, -f5,12- | tail -20
