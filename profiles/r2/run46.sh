mkdir -p gpurun_out/r2
T0=$(date +%s)
timeout 1500 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2/bench_driver_form.json 2> gpurun_out/r2/bench_driver_form.err
T1=$(date +%s); echo "cb200 arm wall $((T1-T0)) s"
python - <<'PY'
import json
j=json.loads(open('gpurun_out/r2/bench_driver_form.json').read().strip().splitlines()[-1])
print(j['value'], j['e2e']['value'], j['roofline']['frac'], j['cpu_baseline']['value'], j.get('config3',{}).get('poses_per_s'), j['clocks'])
PY
tail -2 gpurun_out/r2/bench_driver_form.err
timeout 1500 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2/bench_ref_driver_form.json 2> gpurun_out/r2/bench_ref_driver_form.err
T2=$(date +%s); echo "reference arm wall $((T2-T1)) s"
cut -c1-400 gpurun_out/r2/bench_ref_driver_form.json; tail -2 gpurun_out/r2/bench_ref_driver_form.err
