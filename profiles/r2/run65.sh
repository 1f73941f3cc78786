mkdir -p gpurun_out/r2
timeout 600 python profiles/host_profile_many.py 24 45 > gpurun_out/r2/host_profile_many.log 2>&1
head -70 gpurun_out/r2/host_profile_many.log
