mkdir -p gpurun_out/r2
(timeout 1500 python -m pytest tests -q -m gpu -x) > gpurun_out/r2/t_final2.log 2>&1
tail -3 gpurun_out/r2/t_final2.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 1500 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2/bench_n1_final2.json 2> gpurun_out/r2/bench_n1_final2.err
python - <<'PY'
import json
j=json.loads(open('gpurun_out/r2/bench_n1_final2.json').read().strip().splitlines()[-1])
print(j['value'], j['e2e']['value'], j['roofline']['frac'], j['roofline']['avg_launch_ms'], j['cpu_baseline']['value'], j.get('config3',{}).get('poses_per_s'), j['clocks'], j['allocator'])
print(j['ms_per_step_each']); print(j['e2e_ms_each'])
PY
tail -2 gpurun_out/r2/bench_n1_final2.err
