mkdir -p gpurun_out/r2
for R in 1 2; do
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-config3 > gpurun_out/r2/bench_s20_r$R.json 2> gpurun_out/r2/bench_s20_r$R.err
python - <<PY
import json
j=json.loads(open('gpurun_out/r2/bench_s20_r$R.json').read().strip().splitlines()[-1])
print("run $R", round(j["value"],1), round(j["e2e"]["value"],1), j["allocator"]); print(j['ms_per_step_each']); print(j['e2e_ms_each'])
PY
tail -1 gpurun_out/r2/bench_s20_r$R.err
done
