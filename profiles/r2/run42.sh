mkdir -p gpurun_out/r2
timeout 600 python profiles/run_configs.py --config 5 > gpurun_out/r2/config5_n1.json 2> gpurun_out/r2/config5_n1.err
cat gpurun_out/r2/config5_n1.json; tail -2 gpurun_out/r2/config5_n1.err
timeout 900 python profiles/run_configs.py --config 4 > gpurun_out/r2/config4_n1.json 2> gpurun_out/r2/config4_n1.err
cat gpurun_out/r2/config4_n1.json; tail -2 gpurun_out/r2/config4_n1.err
timeout 1200 python bench.py > gpurun_out/r2/bench_n1.json 2> gpurun_out/r2/bench_n1.err
cat gpurun_out/r2/bench_n1.json; tail -3 gpurun_out/r2/bench_n1.err
