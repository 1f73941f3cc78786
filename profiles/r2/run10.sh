mkdir -p gpurun_out/r2
export CB200_ACCUM_MODE=4
(time timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q) > gpurun_out/r2/t_k10.log 2>&1
tail -4 gpurun_out/r2/t_k10.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-config3 > gpurun_out/r2/bench_run10.json 2> gpurun_out/r2/bench_run10.err
cut -c1-200 gpurun_out/r2/bench_run10.json; tail -3 gpurun_out/r2/bench_run10.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2/launches_m4b.csv python profiles/run_profile.py 2 > gpurun_out/r2/launches_m4b.log 2>&1
unset CB200_ACCUM_MODE
timeout 600 python profiles/e2e_profile.py 40 > gpurun_out/r2/e2e_profile.log 2>&1
