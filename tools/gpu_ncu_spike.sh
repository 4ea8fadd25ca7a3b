mkdir -p gpurun_out/r2i
AM3D_PROFILE_STEP=150 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2i/launches_spike512.csv python tools/batch_steps.py 512 152 > gpurun_out/r2i/ncu_spike.log 2>&1
AM3D_PROFILE_STEP=152 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2i/launches_calm512.csv python tools/batch_steps.py 512 154 > gpurun_out/r2i/ncu_calm.log 2>&1
tail -3 gpurun_out/r2i/ncu_spike.log | cut -c1-200
