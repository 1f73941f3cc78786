mkdir -p gpurun_out/r2
timeout 600 python profiles/stall_probe.py 20 > gpurun_out/r2/stall_probe.log 2>&1
tail -24 gpurun_out/r2/stall_probe.log
