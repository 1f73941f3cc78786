mkdir -p gpurun_out/r2
export CB200_ACCUM_MODE=4
(time timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q) > gpurun_out/r2/t_k11.log 2>&1
tail -4 gpurun_out/r2/t_k11.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-config3 > gpurun_out/r2/bench_run11.json 2> gpurun_out/r2/bench_run11.err
cut -c1-200 gpurun_out/r2/bench_run11.json; tail -3 gpurun_out/r2/bench_run11.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2/launches_m4c.csv python profiles/run_profile.py 2 > gpurun_out/r2/launches_m4c.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tp_accumulate -s 19 -c 1 -f -o gpurun_out/r2/k3_m4c python profiles/run_profile.py 2 > gpurun_out/r2/k3_m4c.log 2>&1
