mkdir -p gpurun_out/r2
(time timeout 900 python -m pytest tests -x -q -m gpu -k "test_gpu_model or bench_batch_sampling or get_model" ) > gpurun_out/r2/t_all7.log 2>&1
tail -6 gpurun_out/r2/t_all7.log
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2/bench_run7.json 2> gpurun_out/r2/bench_run7.err
cut -c1-200 gpurun_out/r2/bench_run7.json; tail -3 gpurun_out/r2/bench_run7.err
CB200_CUDA_GRAPH=0 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2/bench_run7_nograph.json 2> gpurun_out/r2/bench_run7_nograph.err
cut -c1-200 gpurun_out/r2/bench_run7_nograph.json
