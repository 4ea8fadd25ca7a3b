# bench lines only (no profiler): refreshes gpurun_out/r1/bench_*.json
mkdir -p gpurun_out/r1
timeout 300 python bench.py > gpurun_out/r1/bench_batch.json 2> gpurun_out/r1/bench_batch.err
timeout 200 python bench.py --impl reference --steps 40 --warmup 3 > gpurun_out/r1/bench_reference_batch.json 2>&1
timeout 200 python bench.py --workload stack --merging 0 > gpurun_out/r1/bench_stack_m0.json 2> gpurun_out/r1/err1
timeout 200 python bench.py --workload stack --merging 1 --no-cpu-baseline > gpurun_out/r1/bench_stack_m1.json 2> gpurun_out/r1/err2
timeout 200 python bench.py --workload pile --merging 1 --no-cpu-baseline > gpurun_out/r1/bench_pile_m1.json 2> gpurun_out/r1/err3
timeout 120 python bench.py --workload funnel --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r1/bench_funnel20.json 2> gpurun_out/r1/err4
tail -c 300 gpurun_out/r1/bench_batch.json
