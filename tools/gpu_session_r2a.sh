# Round-2 session, part A (essentials first): tests, smoke, default bench, reference arm, launch list + full capture of the default.
mkdir -p gpurun_out/r2
O=gpurun_out/r2
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -5 | tee $O/pytest_gpu.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 | tee $O/smoke.txt
timeout 600 python bench.py > $O/bench_batch.json 2> $O/bench_batch.err; tail -c 300 $O/bench_batch.json
timeout 300 python bench.py --impl reference --steps 40 --warmup 3 > $O/bench_reference_batch.json 2>&1
timeout 400 python bench.py --scaling weak --no-also --no-cpu-baseline > $O/bench_batch512.json 2> $O/err512
timeout 300 python bench.py --workload stack --merging 0 --no-cpu-baseline > $O/bench_stack_m0.json 2> $O/err1
export AM3D_CUDA_PROFILER=1
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_batch.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-also > $O/ncu_launches.log 2>&1
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"k_pgs_color|k_pgs_cluster|k_pgs_persistent|k_warm_start|k_assemble|k_narrow_box|k_contact_set" -c 34 -f -o $O/full_batch python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-also > $O/ncu_full_batch.log 2>&1
ls -la $O
