mkdir -p gpurun_out/r2
(timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_kernels.py -q -m gpu -x) > gpurun_out/r2/t_68.log 2>&1
tail -3 gpurun_out/r2/t_68.log
timeout 600 python profiles/host_profile_many.py 24 12 > gpurun_out/r2/host_profile_many3.log 2>&1
head -22 gpurun_out/r2/host_profile_many3.log
timeout 900 python profiles/run_configs.py --config 4 > gpurun_out/r2/config4_n1.json 2> gpurun_out/r2/config4_n1.err
python -c "
import json
j=json.loads(open('gpurun_out/r2/config4_n1.json').read().strip().splitlines()[-1]); print('config4', j['poses_per_s'], j['wall_s'])"
tail -2 gpurun_out/r2/config4_n1.err
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2/bench_n1e.json 2> gpurun_out/r2/bench_n1e.err
python -c "
import json
j=json.loads(open('gpurun_out/r2/bench_n1e.json').read().strip().splitlines()[-1]); print(j['value'], j['e2e']['value'], j['config3']['poses_per_s'])"
