mkdir -p gpurun_out/r2m
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2m/pytest.txt 2>&1; tail -2 gpurun_out/r2m/pytest.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2m/smoke.txt 2>&1; tail -1 gpurun_out/r2m/smoke.txt
python bench.py --steps 20 --warmup 3 > gpurun_out/r2m/default.json 2> gpurun_out/r2m/default.err
python - <<'PY'
import json
j=json.loads(open('gpurun_out/r2m/default.json').read().strip().splitlines()[-1])
print(j['value'], j['ms_per_step'], j['e2e']['ms_per_step'], j['e2e']['transport'], j['parity_vs_cpu_sample'], j['roofline']['traffic'], (j['roofline'].get('int_issue') or {}).get('frac'), j['cpu_baseline']['cores'], j['config4_stream']['value'], j['gpu_launches'])
PY
