mkdir -p gpurun_out/r2
timeout 600 python profiles/ab_modes.py 4,3 3 3 > gpurun_out/r2/ab31.log 2>&1
tail -4 gpurun_out/r2/ab31.log
