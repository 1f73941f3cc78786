mkdir -p gpurun_out/r2
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2/smi.txt
(time python -m pytest tests/test_gpu_fullsize.py -x -q --durations=20) > gpurun_out/r2/t_fullsize.log 2>&1
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2/bench_base.json 2> gpurun_out/r2/bench_base.err
CB200_WORKSPACE_MB=64 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2/bench_ws64.json 2> gpurun_out/r2/bench_ws64.err
CB200_WORKSPACE_MB=256 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2/bench_ws256.json 2> gpurun_out/r2/bench_ws256.err
tail -5 gpurun_out/r2/t_fullsize.log; cat gpurun_out/r2/bench_base.json | cut -c1-300
