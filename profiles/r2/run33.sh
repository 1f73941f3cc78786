mkdir -p gpurun_out/r2
(timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -k tp_conv) > gpurun_out/r2/t_k33.log 2>&1
tail -2 gpurun_out/r2/t_k33.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2/launches_rb.csv python profiles/run_profile.py 2 > gpurun_out/r2/launches_rb.log 2>&1
