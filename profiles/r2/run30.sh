mkdir -p gpurun_out/r2
(timeout 1500 python -m pytest tests -m gpu -x -q) > gpurun_out/r2/t_full30.log 2>&1
tail -5 gpurun_out/r2/t_full30.log
for m in 4 3; do
CB200_ACCUM_MODE=$m timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-config3 > gpurun_out/r2/bench_run30_m$m.json 2> gpurun_out/r2/bench_run30_m$m.err
cut -c1-160 gpurun_out/r2/bench_run30_m$m.json; tail -2 gpurun_out/r2/bench_run30_m$m.err
done
