mkdir -p gpurun_out/r2
O=gpurun_out/r2
./tools/micro/fp64_latency > $O/fp64_latency.txt 2>&1
timeout 400 python bench.py --workload funnel --steps 3 --warmup 3 --no-cpu-baseline > $O/bench_funnel20.json 2> $O/err4
export AM3D_CUDA_PROFILER=1
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_funnel.csv python bench.py --workload funnel --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_launches_funnel.log 2>&1
tail -c 1500 $O/bench_funnel20.json
