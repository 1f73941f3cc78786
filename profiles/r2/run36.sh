mkdir -p gpurun_out/r2
(timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -k tp_conv) > gpurun_out/r2/t_k36.log 2>&1
tail -3 gpurun_out/r2/t_k36.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2/launches_v5.csv python profiles/run_profile.py 2 > gpurun_out/r2/launches_v5.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,sm__icc_request_hit_rate.pct,sm__icc_requests.sum,gcc__cache_requests_type_instruction.sum,gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed,smsp__inst_executed.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:tp_accumulate -s 19 -c 1 --csv --log-file gpurun_out/r2/icc_v5.csv python profiles/run_profile.py 2 > gpurun_out/r2/icc_v5.log 2>&1
