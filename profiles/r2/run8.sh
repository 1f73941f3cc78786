mkdir -p gpurun_out/r2
export CB200_ACCUM_MODE=4
(time timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q --durations=5) > gpurun_out/r2/t_k8.log 2>&1
tail -15 gpurun_out/r2/t_k8.log
if grep -q "passed" gpurun_out/r2/t_k8.log && ! grep -q "failed" gpurun_out/r2/t_k8.log; then
(time timeout 900 python -m pytest tests -x -q -m gpu -k "test_gpu_model or bench_batch or blockwise" ) > gpurun_out/r2/t_all8.log 2>&1
tail -6 gpurun_out/r2/t_all8.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-config3 > gpurun_out/r2/bench_run8.json 2> gpurun_out/r2/bench_run8.err
cut -c1-200 gpurun_out/r2/bench_run8.json; tail -3 gpurun_out/r2/bench_run8.err
fi
