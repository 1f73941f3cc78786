mkdir -p gpurun_out/r2
(timeout 600 python -m pytest tests/test_gpu_model.py -q -m gpu -x) > gpurun_out/r2/t_72.log 2>&1
tail -2 gpurun_out/r2/t_72.log
bash profiles/r2/run71.sh
