mkdir -p gpurun_out/r2
(time python -m pytest tests -x -q -m gpu -k "not fullsize or bench_batch_sampling or get_model" --durations=5) > gpurun_out/r2/t_all4.log 2>&1
tail -12 gpurun_out/r2/t_all4.log
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2/bench_run4.json 2> gpurun_out/r2/bench_run4.err
cut -c1-200 gpurun_out/r2/bench_run4.json; tail -3 gpurun_out/r2/bench_run4.err
CB200_CUDA_GRAPH=0 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-config3 > gpurun_out/r2/bench_run4_nograph.json 2> gpurun_out/r2/bench_run4_nograph.err
cut -c1-200 gpurun_out/r2/bench_run4_nograph.json
