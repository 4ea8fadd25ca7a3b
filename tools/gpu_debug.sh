mkdir -p gpurun_out/r2f
python tools/debug_torsos.py > gpurun_out/r2f/torsos_split.txt 2>&1
sed -n '/^step 107/,/^step 112/p' gpurun_out/r2f/torsos_split.txt | cut -c1-700
