mkdir -p gpurun_out/r2
timeout 1500 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-config3 > gpurun_out/r2/bench_s20.json 2> gpurun_out/r2/bench_s20.err
python - <<'PY'
import json
j=json.loads(open('gpurun_out/r2/bench_s20.json').read().strip().splitlines()[-1])
print(j['value'], j['e2e']['value']); print(j['ms_per_step_each']); print(j['e2e_ms_each'])
PY
tail -2 gpurun_out/r2/bench_s20.err
