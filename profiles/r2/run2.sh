mkdir -p gpurun_out/r2
(time python -m pytest tests/test_gpu_fullsize.py -x -q -k "config5 or config1 or get_model or tor_bond or blockwise" --durations=10) > gpurun_out/r2/t_fullsize2.log 2>&1
tail -4 gpurun_out/r2/t_fullsize2.log
python bench.py --steps 3 --warmup 3 > gpurun_out/r2/bench_run2.json 2> gpurun_out/r2/bench_run2.err
cut -c1-400 gpurun_out/r2/bench_run2.json; tail -5 gpurun_out/r2/bench_run2.err
