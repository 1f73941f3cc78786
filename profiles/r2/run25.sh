mkdir -p gpurun_out/r2
(timeout 300 python -m pytest tests/test_gpu_fullsize.py -x -q -k "blockwise or bench_batch_score") > gpurun_out/r2/t_k26.log 2>&1
tail -15 gpurun_out/r2/t_k26.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2/launches_tt4.csv python profiles/run_profile.py 2 > gpurun_out/r2/launches_tt4.log 2>&1
tail -3 gpurun_out/r2/launches_tt4.log
