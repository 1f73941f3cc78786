mkdir -p gpurun_out/r2
timeout 600 python profiles/ab_e2e.py 8 > gpurun_out/r2/ab_e2e.log 2>&1
tail -3 gpurun_out/r2/ab_e2e.log
