mkdir -p gpurun_out/r2
(timeout 600 python -m pytest tests/test_gpu_model.py -q -m gpu -x -k "sampling_many or multi_batch or reuse or golden") > gpurun_out/r2/t_64.log 2>&1
tail -30 gpurun_out/r2/t_64.log
timeout 900 python profiles/run_configs.py --config 4 > gpurun_out/r2/config4_n1.json 2> gpurun_out/r2/config4_n1.err
cat gpurun_out/r2/config4_n1.json; tail -2 gpurun_out/r2/config4_n1.err
timeout 900 python bench.py --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/r2/bench_n1d.json 2> gpurun_out/r2/bench_n1d.err
python - <<'PY'
import json
j=json.loads(open('gpurun_out/r2/bench_n1d.json').read().strip().splitlines()[-1])
print(j['value'], j['e2e']['value'], j.get('config3',{}).get('poses_per_s'))
PY
tail -2 gpurun_out/r2/bench_n1d.err
