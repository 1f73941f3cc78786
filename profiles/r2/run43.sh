mkdir -p gpurun_out/r2
(timeout 600 python -m pytest tests/test_gpu_model.py -q -m gpu -k "reuse or reproducible") > gpurun_out/r2/t_43.log 2>&1
tail -3 gpurun_out/r2/t_43.log
timeout 900 python profiles/run_configs.py --config 4 > gpurun_out/r2/config4_n1.json 2> gpurun_out/r2/config4_n1.err
cat gpurun_out/r2/config4_n1.json; tail -2 gpurun_out/r2/config4_n1.err
timeout 900 python bench.py --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/r2/bench_n1b.json 2> gpurun_out/r2/bench_n1b.err
python - <<'PY'
import json
j=json.loads(open('gpurun_out/r2/bench_n1b.json').read().strip().splitlines()[-1])
print(j['value'], j['e2e']['value'], j.get('config3'))
PY
tail -2 gpurun_out/r2/bench_n1b.err
