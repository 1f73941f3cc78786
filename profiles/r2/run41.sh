mkdir -p gpurun_out/r2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29541 bench.py --gpus 8 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2/bench_n8.json 2> gpurun_out/r2/bench_n8.err
cut -c1-200 gpurun_out/r2/bench_n8.json; tail -2 gpurun_out/r2/bench_n8.err
timeout 600 $TR --master-port 29542 profiles/run_configs.py --config 5 > gpurun_out/r2/config5_n8.json 2> gpurun_out/r2/config5_n8.err
cat gpurun_out/r2/config5_n8.json; tail -2 gpurun_out/r2/config5_n8.err
timeout 900 $TR --master-port 29543 profiles/run_configs.py --config 4 > gpurun_out/r2/config4_n8.json 2> gpurun_out/r2/config4_n8.err
cat gpurun_out/r2/config4_n8.json; tail -2 gpurun_out/r2/config4_n8.err
