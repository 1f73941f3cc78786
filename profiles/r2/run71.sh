mkdir -p gpurun_out/r2
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-config3 > gpurun_out/r2/bench_s20_last.json 2> gpurun_out/r2/bench_s20_last.err
python - <<'PY'
import json
j=json.loads(open('gpurun_out/r2/bench_s20_last.json').read().strip().splitlines()[-1])
print(round(j['value'],1), round(j['e2e']['value'],1)); print(j['ms_per_step_each']); print(j['e2e_ms_each'])
PY
tail -1 gpurun_out/r2/bench_s20_last.err
