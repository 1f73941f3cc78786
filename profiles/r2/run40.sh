mkdir -p gpurun_out/r2
(timeout 600 python -m pytest tests/test_gpu_model.py -q -m gpu -k "removes_a_whole or two_gpus") > gpurun_out/r2/t_40.log 2>&1
tail -40 gpurun_out/r2/t_40.log
