mkdir -p gpurun_out/r1
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 400 python bench.py > gpurun_out/r1/bench_batch.json 2> gpurun_out/r1/bench_batch.err; tail -c 400 gpurun_out/r1/bench_batch.json
timeout 300 python bench.py --impl reference --steps 40 --warmup 3 > gpurun_out/r1/bench_reference_batch.json 2>&1
timeout 300 python bench.py --workload stack --merging 0 > gpurun_out/r1/bench_stack_m0.json 2> gpurun_out/r1/err1
timeout 300 python bench.py --workload stack --merging 1 --no-cpu-baseline > gpurun_out/r1/bench_stack_m1.json 2> gpurun_out/r1/err2
timeout 300 python bench.py --workload pile --merging 1 --no-cpu-baseline > gpurun_out/r1/bench_pile_m1.json 2> gpurun_out/r1/err3
timeout 300 python bench.py --workload funnel --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r1/bench_funnel20.json 2> gpurun_out/r1/err4
export AM3D_CUDA_PROFILER=1
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r1/launches_batch.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r1/ncu_launches.log 2>&1
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r1/launches_stack.csv python bench.py --workload stack --merging 0 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r1/ncu_launches_stack.log 2>&1
timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"k_pgs|k_narrow|k_assemble|k_warm_start|k_contact_set|k_pairs_grid|k_bpc_accumulate" -c 14 -f -o gpurun_out/r1/full_batch python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r1/ncu_full_batch.log 2>&1
timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"k_pgs|k_narrow_box|k_pairs_grid" -c 4 -f -o gpurun_out/r1/full_stack python bench.py --workload stack --merging 0 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r1/ncu_full_stack.log 2>&1
ls -la gpurun_out/r1
