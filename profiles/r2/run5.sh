mkdir -p gpurun_out/r2
(time timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q --durations=5) > gpurun_out/r2/t_k5.log 2>&1
tail -12 gpurun_out/r2/t_k5.log
(time timeout 900 python -m pytest tests -x -q -m gpu -k "not test_gpu_kernels and (not fullsize or bench_batch_sampling or blockwise)" ) > gpurun_out/r2/t_all5.log 2>&1
tail -6 gpurun_out/r2/t_all5.log
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2/bench_run5.json 2> gpurun_out/r2/bench_run5.err
cut -c1-200 gpurun_out/r2/bench_run5.json; tail -3 gpurun_out/r2/bench_run5.err
CB200_CUDA_GRAPH=0 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-config3 > gpurun_out/r2/bench_run5_nograph.json 2> gpurun_out/r2/bench_run5_nograph.err
cut -c1-200 gpurun_out/r2/bench_run5_nograph.json
CB200_CUDA_GRAPH=0 CB200_ACCUM_MODE=2 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-config3 > gpurun_out/r2/bench_run5_mode2.json 2> gpurun_out/r2/bench_run5_mode2.err
cut -c1-200 gpurun_out/r2/bench_run5_mode2.json
