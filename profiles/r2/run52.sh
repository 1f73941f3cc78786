mkdir -p gpurun_out/r2
(timeout 1500 python -m pytest tests -q -m gpu -x) > gpurun_out/r2/t_52.log 2>&1
tail -4 gpurun_out/r2/t_52.log
bash profiles/r2/run47.sh
