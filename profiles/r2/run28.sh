mkdir -p gpurun_out/r2
export CB200_ACCUM_MODE=4
(timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -k tp_conv) > gpurun_out/r2/t_k28.log 2>&1
tail -12 gpurun_out/r2/t_k28.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2/launches_sus.csv python profiles/run_profile.py 2 > gpurun_out/r2/launches_sus.log 2>&1
