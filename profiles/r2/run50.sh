mkdir -p gpurun_out/r2
(timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -q -m gpu -x) > gpurun_out/r2/t_50.log 2>&1
tail -4 gpurun_out/r2/t_50.log
bash profiles/r2/run47.sh
