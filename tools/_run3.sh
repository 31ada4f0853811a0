mkdir -p gpurun_out/r2c
NCU="ncu --set full --clock-control none --import-source on"
# DTW static vs dynamic at configs[2] shape (smaller batch keeps the replays short)
$NCU -k regex:dtw_kernel -c 1 -o gpurun_out/r2c/prof_dtw_static python bench.py --workload dtw --batch 16384 --steps 1 --warmup 1 --dtw-variant 0 > gpurun_out/r2c/ncu_dtw0.log 2>&1
$NCU -k regex:dtw_dyn_kernel -c 1 -o gpurun_out/r2c/prof_dtw_dyn python bench.py --workload dtw --batch 16384 --steps 1 --warmup 1 --dtw-variant 1 > gpurun_out/r2c/ncu_dtw1.log 2>&1
# the recognise step: every kernel once (vad, mfcc, dtw) after warm-up
$NCU -k regex:"vad_kernel|mfcc_kernel|dtw_kernel" -s 9 -c 3 -o gpurun_out/r2c/prof_step python bench.py --steps 1 --warmup 3 --no-cpu --no-stream > gpurun_out/r2c/ncu_step.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r2c/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-stream > gpurun_out/r2c/launch_run.log 2>&1
ls -la gpurun_out/r2c
SR_NO_BUILD=1 timeout 600 python -m pytest tests -q -m gpu -k "torchrun or streaming" > gpurun_out/r2c/pytest.txt 2>&1; tail -3 gpurun_out/r2c/pytest.txt
python bench.py --workload stream --templates 20 > gpurun_out/r2c/stream.json 2> gpurun_out/r2c/stream.err
python bench.py --steps 20 --warmup 3 > gpurun_out/r2c/default.json 2> gpurun_out/r2c/default.err
python - <<'PY'
import json
j=json.loads(open('gpurun_out/r2c/stream.json').read().strip().splitlines()[-1])
print({k:{a:b for a,b in v.items() if 'latency' in a or 'largest' in a} for k,v in j['results'].items()})
j=json.loads(open('gpurun_out/r2c/default.json').read().strip().splitlines()[-1])
print(j['value'], j['ms_per_step'], j['e2e']['ms_per_step'], j['e2e']['transport'])
PY
