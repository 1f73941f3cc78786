mkdir -p gpurun_out/r2
(timeout 900 python -m pytest tests/test_gpu_model.py -x -q -m gpu -k "reuse or reproducible or golden or sampling") > gpurun_out/r2/t_m34.log 2>&1
tail -15 gpurun_out/r2/t_m34.log
CB200_GRAPH_CACHE=0 timeout 600 python profiles/ab_modes.py 4 2 3 > gpurun_out/r2/ab34_nocache.log 2>&1; tail -2 gpurun_out/r2/ab34_nocache.log
CB200_GRAPH_CACHE=2 timeout 600 python profiles/ab_modes.py 4 2 3 > gpurun_out/r2/ab34_cache.log 2>&1; tail -2 gpurun_out/r2/ab34_cache.log
