mkdir -p gpurun_out/r2d
export AM3D_CUDA_PROFILER=1
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2d/launches_funnel_fast.csv python bench.py --workload funnel --steps 1 --warmup 5 --no-cpu-baseline > gpurun_out/r2d/ncu_launches_funnel_fast.log 2>&1
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"k_pgs_giant" -s 130 -c 2 -f -o gpurun_out/r2d/full_giant python bench.py --workload funnel --steps 1 --warmup 5 --no-cpu-baseline > gpurun_out/r2d/ncu_full_giant.log 2>&1
