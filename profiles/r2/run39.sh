mkdir -p gpurun_out/r2
(timeout 600 python -m pytest tests/test_gpu_model.py -x -q -m gpu -k "two_gpus") > gpurun_out/r2/t_2gpu.log 2>&1
tail -3 gpurun_out/r2/t_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2/bench_n2.json 2> gpurun_out/r2/bench_n2.err
cut -c1-300 gpurun_out/r2/bench_n2.json; tail -3 gpurun_out/r2/bench_n2.err
