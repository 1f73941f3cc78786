mkdir -p gpurun_out/r2
(timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_kernels.py -q -m gpu -x) > gpurun_out/r2/t_67.log 2>&1
tail -5 gpurun_out/r2/t_67.log
timeout 600 python profiles/host_profile_many.py 24 14 > gpurun_out/r2/host_profile_many2.log 2>&1
head -24 gpurun_out/r2/host_profile_many2.log
timeout 900 python profiles/run_configs.py --config 4 > gpurun_out/r2/config4_n1.json 2> gpurun_out/r2/config4_n1.err
cat gpurun_out/r2/config4_n1.json; tail -2 gpurun_out/r2/config4_n1.err
