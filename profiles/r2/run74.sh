mkdir -p gpurun_out/r2
(timeout 900 python -m pytest tests -q -m gpu -x) > gpurun_out/r2/t_final3.log 2>&1
tail -3 gpurun_out/r2/t_final3.log
