mkdir -p gpurun_out/r2e
SR_NO_BUILD=1 timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r2e/pytest.txt 2>&1; tail -4 gpurun_out/r2e/pytest.txt
python bench.py --steps 50 --warmup 5 > gpurun_out/r2e/default.json 2> gpurun_out/r2e/default.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2e/reference.json 2> gpurun_out/r2e/reference.err
python - <<'PY'
import json
j=json.loads(open('gpurun_out/r2e/default.json').read().strip().splitlines()[-1])
print(j['value'], j['ms_per_step'], j['kernel_ms'], j['e2e']['ms_per_step'], j['e2e']['transport'], j['parity_vs_cpu_sample'], j['roofline']['other_kernels'])
PY
