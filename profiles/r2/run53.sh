mkdir -p gpurun_out/r2
timeout 600 python profiles/e2e_profile.py 60 > gpurun_out/r2/e2e_profile.log 2>&1
head -75 gpurun_out/r2/e2e_profile.log
