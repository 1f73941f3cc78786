mkdir -p gpurun_out/r2
(timeout 200 python -m pytest tests/test_gpu_kernels.py -x -q -k tp_conv) > gpurun_out/r2/t_k38.log 2>&1
tail -3 gpurun_out/r2/t_k38.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2/launches_rw.csv python profiles/run_profile.py 2 > gpurun_out/r2/launches_rw.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum,sm__icc_request_hit_rate.pct,gcc__cache_requests_type_instruction.sum,smsp__inst_executed.sum,smsp__issue_active.max.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:tp_accumulate -s 19 -c 1 --csv --log-file gpurun_out/r2/icc_rw.csv python profiles/run_profile.py 2 > gpurun_out/r2/icc_rw.log 2>&1
