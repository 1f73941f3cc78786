mkdir -p gpurun_out/r2
export CB200_ACCUM_MODE=4
(timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q) > gpurun_out/r2/t_k17.log 2>&1
tail -3 gpurun_out/r2/t_k17.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-config3 > gpurun_out/r2/bench_run17.json 2> gpurun_out/r2/bench_run17.err
cut -c1-200 gpurun_out/r2/bench_run17.json; tail -3 gpurun_out/r2/bench_run17.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2/launches_m4f.csv python profiles/run_profile.py 2 > gpurun_out/r2/launches_m4d.log 2>&1
CB200_EXTRA_NVCC_FLAGS="-DCB_PHASE_TIMING" python -c "from confidence_bootstrapping_b200 import build as b; print(b.build_library(force=True))"
timeout 600 python profiles/phase_timing_ws.py > gpurun_out/r2/phase_ws4.log 2>&1
python -c "from confidence_bootstrapping_b200 import build as b; print(b.build_library(force=True))"
