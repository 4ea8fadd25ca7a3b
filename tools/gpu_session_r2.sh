# Round-2 measurement session on one B200 box (run through gpurun): tests, smoke, bench lines, ncu launch lists and
# --set full captures.  Outputs land in gpurun_out/r2/; tools/make_profiles_r2.py turns them into profiles/r2_*.
mkdir -p gpurun_out/r2
O=gpurun_out/r2
if [ "$1" != "ncu-only" ]; then
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3 | tee $O/pytest_gpu.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 | tee $O/smoke.txt
timeout 600 python bench.py > $O/bench_batch.json 2> $O/bench_batch.err; tail -c 300 $O/bench_batch.json
timeout 300 python bench.py --impl reference --steps 40 --warmup 3 > $O/bench_reference_batch.json 2>&1
timeout 400 python bench.py --scaling weak --no-also --no-cpu-baseline > $O/bench_batch512.json 2> $O/err512
timeout 300 python bench.py --workload stack --merging 0 --no-cpu-baseline > $O/bench_stack_m0.json 2> $O/err1
timeout 300 python bench.py --workload stack --merging 1 --no-cpu-baseline > $O/bench_stack_m1.json 2> $O/err2
timeout 300 python bench.py --workload pile --merging 1 --no-cpu-baseline > $O/bench_pile_m1.json 2> $O/err3
timeout 400 python bench.py --workload funnel --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_funnel20.json 2> $O/err4
fi
export AM3D_CUDA_PROFILER=1
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_batch.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-also > $O/ncu_launches.log 2>&1
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_batch512.csv python bench.py --scaling weak --steps 2 --warmup 3 --no-cpu-baseline --no-also > $O/ncu_launches512.log 2>&1
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_stack.csv python bench.py --workload stack --merging 0 --steps 2 --warmup 3 --no-cpu-baseline > $O/ncu_launches_stack.log 2>&1
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"k_pgs_color|k_pgs_tail|k_warm_start|k_warm_apply|k_assemble|k_narrow_box|k_contact_set" -c 30 -f -o $O/full_batch python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-also > $O/ncu_full_batch.log 2>&1
ncu -i $O/full_batch.ncu-rep --page raw --csv > $O/full_batch_raw.csv 2>/dev/null; rm -f $O/full_batch.ncu-rep   # (gpurun_out travels back only below 64 MiB)
timeout 400 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"k_pgs_cluster|k_pgs_persistent|k_bfs_layers|k_color_coop" -c 6 -f -o $O/full_batch512 python bench.py --scaling weak --steps 1 --warmup 3 --no-cpu-baseline --no-also > $O/ncu_full_batch512.log 2>&1
ncu -i $O/full_batch512.ncu-rep --page raw --csv > $O/full_batch512_raw.csv 2>/dev/null; rm -f $O/full_batch512.ncu-rep   # (gpurun_out travels back only below 64 MiB)
timeout 400 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"k_pgs_color|k_pgs_persistent|k_narrow_box|k_pairs_grid" -c 4 -f -o $O/full_stack python bench.py --workload stack --merging 0 --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_full_stack.log 2>&1
ncu -i $O/full_stack.ncu-rep --page raw --csv > $O/full_stack_raw.csv 2>/dev/null; rm -f $O/full_stack.ncu-rep   # (gpurun_out travels back only below 64 MiB)
timeout 400 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"k_pgs_giant|k_tree_tasks|k_narrow_tree" -s 150 -c 6 -f -o $O/full_funnel python bench.py --workload funnel --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_full_funnel.log 2>&1
ncu -i $O/full_funnel.ncu-rep --page raw --csv > $O/full_funnel_raw.csv 2>/dev/null; rm -f $O/full_funnel.ncu-rep   # (gpurun_out travels back only below 64 MiB)
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -k regex:"k_narrow|k_tree" --csv --log-file $O/launches_funnel_tree.csv python bench.py --workload funnel --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_launches_funnel_tree.log 2>&1
ls -la $O
