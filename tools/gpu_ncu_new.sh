mkdir -p gpurun_out/r2
O=gpurun_out/r2
export AM3D_CUDA_PROFILER=1
timeout 400 ncu --profile-from-start off --set full --clock-control none -k regex:"k_rs_scatter|k_rs_hist|k_scan_apply|k_pgs_tail|k_export_ints" -c 14 -f -o $O/full_new python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-also > $O/ncu_full_new.log 2>&1
ncu -i $O/full_new.ncu-rep --page raw --csv > $O/full_new_raw.csv 2>/dev/null; rm -f $O/full_new.ncu-rep
timeout 300 ncu --profile-from-start off --set full --clock-control none -k regex:"k_tree_tasks|k_narrow_tree" -c 4 -f -o $O/full_tree python bench.py --workload funnel --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_full_tree.log 2>&1
ncu -i $O/full_tree.ncu-rep --page raw --csv > $O/full_tree_raw.csv 2>/dev/null; rm -f $O/full_tree.ncu-rep
ls -la $O
