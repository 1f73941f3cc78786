mkdir -p gpurun_out/r2
(timeout 600 python -m pytest tests/test_gpu_model.py -q -m gpu -x -k "multi_batch") > gpurun_out/r2/t_63.log 2>&1
tail -30 gpurun_out/r2/t_63.log
